"""Multi-GPU plumbing of the hot path (one process per GPU, launched by torchrun).

The path shards by cloud: grid keys carry the batch id, ball-query ranges are clamped to the batch's key span
and k-NN stops at batch boundaries (custom_ops/ball_query/grid_utils.cuh:92, find_ranges_grid_ds.cu:103-104,
knn_query/knn_query.cu:60,97), so no edge ever crosses two clouds and there is NO data-path collective.  Every rank
builds the hierarchy and runs the convolutions of its own clouds; the only cross-rank operations are the barrier
and the max-over-ranks of the device time around the timed region (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_seed(rank):
    """Seed of the synthetic clouds of a rank (weak scaling: every rank owns different clouds)."""
    return int(rank)


def cloud_ids(rank, world, clouds_per_rank):
    """Global ids of the clouds a rank owns (contiguous blocks, disjoint across ranks)."""
    return list(range(rank * clouds_per_rank, (rank + 1) * clouds_per_rank))


def barrier(device):
    if _active():
        dist.barrier()
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value, device):
    """Step time of the job = the slowest rank's device time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if _active():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if _active():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(round(float(t.item())))


def is_reporter(rank):
    return rank == 0
