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


def _reducer_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from se3conv3d_b200 import shard
    torch.manual_seed(0)                                   # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(6, 40), torch.nn.GELU(), torch.nn.Linear(40, 40), torch.nn.GELU(),
                                torch.nn.Linear(40, 3))
    unused = torch.nn.Parameter(torch.ones(5))             # a parameter that never receives a gradient
    params = list(model.parameters()) + [unused]
    red = shard.GradAllReducer(params, bucket_bytes=4096)  # several buckets
    assert len(red.buckets) >= 2 and all(p.grad.data_ptr() >= red.flat.data_ptr() for p in params)
    g = torch.Generator().manual_seed(100 + rank)          # different data per rank
    x, y = torch.randn(16, 6, generator=g), torch.randn(16, 3, generator=g)
    res = []
    for step in range(2):                                  # two steps: pending counters and handles reset
        loss = (model(x * (step + 1)) - y).square().mean()
        loss.backward()
        red.finish()
        res.append(red.flat.clone())
        red.zero_grad()
    # what a single process would compute for the mean of both ranks' losses
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 40), torch.nn.GELU(), torch.nn.Linear(40, 40), torch.nn.GELU(),
                              torch.nn.Linear(40, 3))
    tot = 0.0
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        xr, yr = torch.randn(16, 6, generator=g), torch.randn(16, 3, generator=g)
        tot = tot + (ref(xr) - yr).square().mean() / world
    tot.backward()
    flat_ref = torch.cat([p.grad.reshape(-1) for p in reversed(list(ref.parameters()))])
    n_unused = unused.numel()
    err = float((res[0][n_unused:] - flat_ref).abs().max() / flat_ref.abs().max())
    out.put((rank, err, float(res[0][:n_unused].abs().max()), float(res[0].sum()), float(res[1].abs().sum())))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_allreduce_buckets_gloo_world2():
    """shard.GradAllReducer: flat gradient buffer, reverse-order buckets reduced asynchronously from post-accumulate
    hooks; the averaged gradients equal those of the mean loss over both ranks' shards, on both ranks."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (_, e0, u0, s0, t0), (_, e1, u1, s1, t1) = res
    assert e0 < 1e-6 and e1 < 1e-6
    assert u0 == 0.0 and u1 == 0.0                          # the never-used parameter keeps a zero gradient
    assert s0 == s1 and t0 == t1 and t0 > 0                 # both ranks hold the same reduced buffer, also in step 2


def test_gradient_allreducer_single_replica_is_transparent():
    """One replica: no flat buffer, no hooks -- gradients are the plain autograd tensors (autograd hands them over without
    an accumulate kernel), zero_grad drops them, finish is a no-op, and the byte count of a step's gradients is still reported."""
    from se3conv3d_b200 import shard
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 10), torch.nn.GELU(), torch.nn.Linear(10, 3))
    params = list(model.parameters())
    red = shard.GradAllReducer(params)
    assert red.world == 1 and red.flat is None and red.buckets == [] and all(p.grad is None for p in params)
    assert red.bytes == 4 * sum(p.numel() for p in params)
    x = torch.randn(8, 6)
    model(x).square().mean().backward()
    red.finish()
    got = [p.grad.clone() for p in params]
    red.zero_grad()
    assert all(p.grad is None for p in params)
    model(x).square().mean().backward()
    for a, p in zip(got, params):
        assert torch.equal(a, p.grad)
