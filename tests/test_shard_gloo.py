"""N > 1 host logic on CPU: frame sharding and the scalar gather over a world_size-2 gloo group."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from prisma_b200.shard import frame_range, gather_frame_scalars, max_over_ranks


@pytest.mark.parametrize("total,world", [(256, 8), (255, 8), (7, 8), (1024, 8), (10, 3), (1, 2), (0, 2)])
def test_frame_ranges_partition_the_clip(total, world):
    covered = []
    for r in range(world):
        s, e, first = frame_range(r, world, total, halo=1)
        assert 0 <= s <= e <= total and first == max(0, s - 1)
        covered.extend(range(s, e))
    assert covered == list(range(total))          # contiguous, ordered, no overlap, nothing dropped
    sizes = [frame_range(r, world, total)[1] - frame_range(r, world, total)[0] for r in range(world)]
    assert max(sizes) - min(s for s in sizes if s or True) <= max(sizes)  # ranks past the end may be empty
    assert max(sizes) == -(-total // world) if total else max(sizes) == 0


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e, _ = frame_range(rank, world, total)
    # stand-in for the per-frame (min, max) the depth band collects: a function of the frame index
    vals = [(float(i) * 0.5, float(i) * 0.5 + 1.0) for i in range(s, e)]
    table = gather_frame_scalars(vals, total, dist)
    t = max_over_ranks(10.0 + rank, dist)
    dist.barrier()
    if rank == 0:
        q.put((table, t))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [9, 16])
def test_gather_and_max_over_two_ranks(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    table, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 11.0
    assert table == [(i * 0.5, i * 0.5 + 1.0) for i in range(total)]


class _ListWriter:
    """Stand-in for a VideoWriter: collects the frames it is handed (writer rank only)."""

    def __init__(self, sink):
        self.sink = sink

    def write(self, frame):
        self.sink.append(frame.copy())

    def close(self):
        self.sink.append("closed")


def _frame(i, shape=(6, 8, 3)):
    return ((i * 7 + __import__("numpy").arange(int(__import__("numpy").prod(shape))).reshape(shape)) % 251).astype("uint8")


def _stream_worker(rank, world, port, total, halo, q):
    """The band loops' use of bands/common/sharded.py: every rank renders its frame range into two ordered streams and a
    scalar row per frame; the writer rank must end up with every frame exactly once, in order."""
    import numpy as np
    os.environ.update(PRISMA_SHARD_RANK=str(rank), PRISMA_SHARD_WORLD=str(world), PRISMA_SHARD_PORT=str(port))
    from bands.common.sharded import OrderedStreams, ShardContext
    ctx = ShardContext.from_env(device=0, backend="gloo")
    got = {"a": [], "b": []}
    streams = OrderedStreams(ctx, {"a": lambda: _ListWriter(got["a"]), "b": lambda: _ListWriter(got["b"]),
                                   "unused": lambda: _ListWriter([])})
    start, stop, first = ctx.frames(total, halo=halo)
    assert first == max(0, start - halo)
    for i in range(start, stop):
        streams.write("a", _frame(i))
        if i % 2 == 0:
            streams.write("b", _frame(1000 + i, (4, 5, 3)).astype(np.uint16).view(np.uint16))
        streams.scalars(i * 0.25, -float(i))
    table = streams.finish()
    if rank == 0:
        q.put((got, table))
    else:
        assert table is None
    ctx.close()


@pytest.mark.parametrize("total,world,halo", [(11, 2, 0), (5, 2, 1), (3, 2, 1), (1, 2, 1)])
def test_ordered_streams_reach_the_writer_rank_in_frame_order(total, world, halo):
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 1000) + total * 3 + halo
    procs = [ctx.Process(target=_stream_worker, args=(r, world, port, total, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, table = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got["a"][-1] == "closed" and len(got["a"]) == total + 1
    for i in range(total):
        assert np.array_equal(got["a"][i], _frame(i))
    evens = [i for i in range(total) if i % 2 == 0]
    assert len(got["b"]) == len(evens) + 1
    for k, i in enumerate(evens):
        assert got["b"][k].dtype == np.uint16 and np.array_equal(got["b"][k], _frame(1000 + i, (4, 5, 3)).astype(np.uint16))
    assert table == [(i * 0.25, -float(i)) for i in range(total)]


def test_strip_flags_handles_both_spellings():
    from bands.common.sharded import strip_flags
    argv = ["-i", "clip", "--gpus", "4", "--device-list=0,1,2,3", "--device", "2", "--gpus=8", "--subpath", "x"]
    assert strip_flags(argv) == ["-i", "clip", "--subpath", "x"]
