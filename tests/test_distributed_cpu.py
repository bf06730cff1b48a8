"""World-size-2 gloo tests (CPU) of the multi-process plumbing bench.py uses for N > 1: shard assignment by
rank (clouds are independent: no data-path collective), barrier, max-over-ranks timing, rank-0 reporting."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from se3conv3d_b200 import shard
    from se3conv3d_b200 import workloads as wl
    # every rank owns its own clouds (weak scaling): different seeds give different, independent shards
    pts, batch = wl.synthetic_bodies(2, 500, seed=shard.shard_seed(rank))
    ids = shard.cloud_ids(rank, world, clouds_per_rank=2)
    local_ms = 10.0 + 5.0 * rank                      # rank 1 is the slow one
    ms = shard.max_over_ranks(local_ms, device=torch.device("cpu"))
    total_points = shard.sum_over_ranks(pts.shape[0], device=torch.device("cpu"))
    shard.barrier(device=torch.device("cpu"))
    out.put((rank, ids, float(pts.sum()), ms, total_points, shard.is_reporter(rank)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharding_and_max_over_ranks_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, ids0, s0, ms0, tot0, rep0), (r1, ids1, s1, ms1, tot1, rep1) = res
    assert ids0 == [0, 1] and ids1 == [2, 3]            # disjoint global cloud ids
    assert s0 != s1                                      # different data per shard
    assert ms0 == ms1 == 15.0                            # the step time is the max over ranks
    assert tot0 == tot1 == 2000                          # value = units of ALL ranks / that time
    assert rep0 and not rep1                             # rank 0 alone prints the JSON line


def test_single_process_helpers_do_not_need_a_process_group():
    from se3conv3d_b200 import shard
    assert shard.max_over_ranks(3.5, device=torch.device("cpu")) == 3.5
    assert shard.sum_over_ranks(7, device=torch.device("cpu")) == 7
    assert shard.cloud_ids(0, 1, 4) == [0, 1, 2, 3]
    shard.barrier(device=torch.device("cpu"))
