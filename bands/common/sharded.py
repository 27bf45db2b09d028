"""Frame sharding of a band's video loop over the GPUs of one box.

`<band>.py --gpus N` re-launches itself as N worker processes (rank r on GPU devices[r], torch.distributed over
127.0.0.1: NCCL when every rank has its own GPU -- the encoded frames travel GPU -> NVLink -> GPU of the writer rank -- and
gloo on a CPU-only host or when ranks share a GPU, which is how the tests run it).  Every rank owns a contiguous range of frames
(prisma_b200.shard.frame_range; flow bands read one halo frame), runs the band's normal loop over it and hands every
ordered output (video frames, per-frame scalars) to an `OrderedStreams` sink:

  * rank 0 writes its own frames straight into the band's VideoWriters (its range comes first), then receives the
    buffered frames of rank 1, 2, ... in order and appends them -- one video / csv per band, identical to the
    single-process run;
  * the other ranks keep their encoded frames in host memory until their range is done and send them in chunks
    (torch.distributed.send; u8 payload, any dtype viewed as bytes).

There is no data-path collective and nothing goes through the file system (round 1 wrote raw frames to disk).  Per-frame
FILE outputs (<sub>/%05d.png, .npy, .flo) are written by the rank that computed them: their names carry the frame index.
"""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from prisma_b200.shard import frame_range  # noqa: E402

ENV_RANK, ENV_WORLD, ENV_PORT, ENV_DEVS = "PRISMA_SHARD_RANK", "PRISMA_SHARD_WORLD", "PRISMA_SHARD_PORT", "PRISMA_SHARD_DEVICES"
CHUNK_BYTES = 64 << 20


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def strip_flags(argv, flags=("--gpus", "--device-list", "--device")):
    """argv without the sharding flags (both `--flag value` and `--flag=value`)."""
    out, skip = [], False
    for x in argv:
        if skip:
            skip = False
            continue
        if x in flags:
            skip = True
            continue
        if any(x.startswith(f + "=") for f in flags):
            continue
        out.append(x)
    return out


def launch(script, argv, gpus, devices=None):
    """Parent side of `--gpus N`: N workers of `script` (same arguments minus the sharding flags); returns when all are
    done, raises if one failed.  Worker r runs on GPU devices[r % len(devices)]."""
    devices = devices or list(range(gpus))
    port = free_port()
    procs = []
    for r in range(gpus):
        env = dict(os.environ)
        env.update({ENV_RANK: str(r), ENV_WORLD: str(gpus), ENV_PORT: str(port), ENV_DEVS: ",".join(str(d) for d in devices)})
        cmd = [sys.executable, script] + strip_flags(list(argv)) + ["--device", str(devices[r % len(devices)])]
        procs.append(subprocess.Popen(cmd, env=env))
    rcs = [p.wait() for p in procs]
    if any(rcs):
        raise RuntimeError("frame-range workers failed: exit codes %s" % rcs)


class ShardContext:
    """What a band's loop needs to know about its place in the job.  world == 1: a plain single-process run."""

    def __init__(self, rank=0, world=1, dist=None, cuda=False, device=0):
        self.rank, self.world, self.dist, self.cuda, self.device = rank, world, dist, cuda, device

    @staticmethod
    def from_env(device=0, backend=None):
        if ENV_RANK not in os.environ:
            return ShardContext()
        rank, world, port = int(os.environ[ENV_RANK]), int(os.environ[ENV_WORLD]), int(os.environ[ENV_PORT])
        import torch
        import torch.distributed as dist
        devs = [d for d in os.environ.get(ENV_DEVS, "").split(",") if d != ""]
        distinct = len(set(devs)) == len(devs) and len(devs) >= world  # NCCL refuses two ranks on one GPU (tests do that)
        cuda = (torch.cuda.is_available() and distinct) if backend is None else backend == "nccl"
        if cuda:
            torch.cuda.set_device(device)
        dist.init_process_group("nccl" if cuda else "gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                                **({"device_id": torch.device("cuda", device)} if cuda else {}))
        return ShardContext(rank, world, dist, cuda, device)

    def frames(self, total, halo=0):
        """(start, stop, first): this rank owns frames [start, stop) and must read from `first` (= start - halo, clipped)."""
        return frame_range(self.rank, self.world, total, halo)

    def is_writer(self):
        return self.rank == 0

    # ---- point-to-point transport of byte blocks (NCCL: through device memory; gloo: host tensors)
    def _send(self, arr, dst):
        import torch
        flat = np.ascontiguousarray(arr).reshape(-1).view(np.uint8)
        for o in range(0, flat.size, CHUNK_BYTES):
            t = torch.from_numpy(flat[o:o + CHUNK_BYTES])
            self.dist.send(t.cuda(self.device) if self.cuda else t, dst)

    def _recv(self, nbytes, src):
        import torch
        out = np.empty(nbytes, np.uint8)
        for o in range(0, nbytes, CHUNK_BYTES):
            n = min(CHUNK_BYTES, nbytes - o)
            t = torch.empty(n, dtype=torch.uint8, device=("cuda:%d" % self.device) if self.cuda else "cpu")
            self.dist.recv(t, src)
            out[o:o + n] = t.cpu().numpy()
        return out

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


class OrderedStreams:
    """Frame-ordered outputs of a sharded loop.  `writers`: name -> zero-argument factory of an object with .write(frame) /
    .close() (created on the writer rank only, lazily, so absent streams cost nothing)."""

    def __init__(self, ctx, writers):
        self.ctx, self.factories = ctx, writers
        self.open_writers = {}
        self.buf = {}       # rank > 0: name -> list of frames
        self.rows = []      # per-frame scalar tuples of this rank, in frame order

    def _writer(self, name):
        if name not in self.open_writers:
            self.open_writers[name] = self.factories[name]()
        return self.open_writers[name]

    def write(self, name, frame):
        if self.ctx.is_writer():
            self._writer(name).write(frame)
        else:
            self.buf.setdefault(name, []).append(np.ascontiguousarray(frame))

    def scalars(self, *row):
        self.rows.append(tuple(float(v) for v in row))

    def finish(self):
        """Returns the full scalar table (frame order) on the writer rank, None elsewhere; closes the writers."""
        ctx = self.ctx
        names = sorted(self.factories)
        table = list(self.rows)
        if ctx.world > 1:
            if ctx.is_writer():
                for src in range(1, ctx.world):
                    hdr = ctx._recv(8 * (2 + 5 * len(names)), src).view(np.int64)
                    nrows, ncols = int(hdr[0]), int(hdr[1])
                    if nrows:
                        table.extend(tuple(r) for r in ctx._recv(nrows * ncols * 8, src).view(np.float64).reshape(nrows, ncols).tolist())
                    for i, name in enumerate(names):
                        n, h, w, c, isz = (int(v) for v in hdr[2 + 5 * i: 7 + 5 * i])
                        if n == 0:
                            continue
                        per = h * w * c * isz
                        step = max(1, CHUNK_BYTES // per)
                        for o in range(0, n, step):
                            k = min(step, n - o)
                            block = ctx._recv(k * per, src).reshape(k, h, w, c * isz)
                            for j in range(k):
                                fr = block[j] if isz == 1 else block[j].view(np.uint16 if isz == 2 else np.float32)
                                self._writer(name).write(fr.reshape(h, w, c))
            else:
                hdr = np.zeros(2 + 5 * len(names), np.int64)
                hdr[0], hdr[1] = len(self.rows), (len(self.rows[0]) if self.rows else 0)
                for i, name in enumerate(names):
                    fr = self.buf.get(name, [])
                    if fr:
                        a = fr[0] if fr[0].ndim == 3 else fr[0][..., None]
                        hdr[2 + 5 * i: 7 + 5 * i] = (len(fr), a.shape[0], a.shape[1], a.shape[2], a.dtype.itemsize)
                ctx._send(hdr, 0)
                if self.rows:
                    ctx._send(np.asarray(self.rows, np.float64), 0)
                for i, name in enumerate(names):
                    fr = self.buf.get(name, [])
                    if not fr:
                        continue
                    per = fr[0].nbytes
                    step = max(1, CHUNK_BYTES // per)
                    for o in range(0, len(fr), step):
                        ctx._send(np.stack(fr[o:o + step]), 0)
                self.buf.clear()
        for w in self.open_writers.values():
            w.close()
        return table if ctx.is_writer() else None
