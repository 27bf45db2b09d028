"""Frame sharding across the GPUs of one box: frames of a clip are independent units (depth, mask)
or depend on (i-1, i) only (flow), so rank r of R owns a contiguous range and no data-path collective exists.  The only
communication is the barrier / max-over-ranks of the timings and the gather of per-frame scalars ((min, max), max
displacement) to the rank that writes the .csv files -- torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
import math


def frame_range(rank, world, total, halo=0):
    """Rank r of R: frames [start, stop) = [r*ceil(T/R), min(T, (r+1)*ceil(T/R))); `first` = start - halo clipped at 0 is
    the first frame the rank must READ (flow needs frame start-1: halo=1)."""
    per = math.ceil(total / world) if world > 0 else total
    start = min(total, rank * per)
    stop = min(total, (rank + 1) * per)
    return start, stop, max(0, start - halo)


def max_over_ranks(x, dist=None, device=None):
    """max over ranks of a host scalar (the bench's timing rule: max over ranks, never the mean)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_frame_scalars(values, total, dist=None, device=None):
    """Per-frame scalars of this rank's range (list of k-tuples, in frame order) -> on every rank the full [total][k] table in
    frame order (all_gather of equal-size padded blocks; the writer rank saves the .csv files)."""
    import torch
    k = len(values[0]) if values else 1
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [tuple(float(c) for c in v) for v in values]
    world, rank = dist.get_world_size(), dist.get_rank()
    per = math.ceil(total / world)
    kk = torch.tensor([k], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(kk, op=dist.ReduceOp.MAX)
    k = int(kk.item())
    block = torch.full((per, k), float("nan"), dtype=torch.float64, device=device or "cpu")
    for i, v in enumerate(values):
        block[i, :len(v)] = torch.tensor([float(c) for c in v], dtype=torch.float64)
    out = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(out, block)
    table = []
    for r in range(world):
        s, e, _ = frame_range(r, world, total)
        table.extend(tuple(row) for row in out[r][: e - s].tolist())
    return table
