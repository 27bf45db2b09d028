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
