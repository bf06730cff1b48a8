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


class GradAllReducer(object):
    """The one collective of the path: the per-step all-reduce (mean) of the parameter gradients of data-parallel
    replicas (training loop tasks/SemSeg/train_dfaust_rot.py:262-275).

    All gradients live in ONE flat fp32 buffer (every `p.grad` is a view into it, so backward accumulates in place and
    nothing is copied); the buffer is cut into buckets in reverse parameter order (the order backward produces them),
    and a post-accumulate hook launches the asynchronous all-reduce of a bucket as soon as its last gradient has landed
    -- the transfer of the early buckets overlaps the rest of backward (NCCL runs on its own stream over NVLink /
    NVSwitch).  `finish()` waits for the handles before the optimiser step.  With one process it is a no-op apart from
    the flat buffer (which also makes clip_grad_norm_ / zero_grad single kernels)."""

    def __init__(self, params, bucket_bytes=8 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if _active() else 1
        if self.world == 1:
            # one replica: nothing to reduce.  No flat buffer either -- with `p.grad = None` autograd hands every
            # gradient over without a kernel, while a pre-set view costs one in-place add per parameter per step
            self.flat, self.buckets, self._handles, self._hooks = None, [], [], []
            self.bytes = 4 * sum(p.numel() for p in self.params)
            return
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.bytes = total * 4
        # reverse registration order ~ the order in which backward finishes the gradients
        order = list(reversed(self.params))
        self.buckets, self._bucket_of, self._pending0 = [], {}, []
        off, start, count = 0, 0, 0
        for p in order:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._bucket_of[id(p)] = len(self.buckets)
            off += n
            count += 1
            if (off - start) * 4 >= bucket_bytes:
                self.buckets.append((start, off))
                self._pending0.append(count)
                start, count = off, 0
        if off > start:
            self.buckets.append((start, off))
            self._pending0.append(count)
        self._pending = list(self._pending0)
        self._handles = []
        self._hooks = []
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p):
        b = self._bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            s, e = self.buckets[b]
            self._handles.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.AVG if self.flat.is_cuda else
                                                 dist.ReduceOp.SUM, async_op=True))

    def finish(self):
        """Waits for every bucket (call after backward, before clipping / the optimiser step)."""
        if self.world == 1:
            return
        if self.world > 1:
            # parameters that received no gradient this step never fired their hook: reduce their buckets now
            for b, left in enumerate(self._pending):
                if left > 0:
                    s, e = self.buckets[b]
                    self._handles.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.AVG if self.flat.is_cuda else
                                                         dist.ReduceOp.SUM, async_op=True))
            for h in self._handles:
                h.wait()
            if not self.flat.is_cuda:            # gloo has no AVG
                self.flat.div_(self.world)
        self._handles = []
        self._pending = list(self._pending0)

    def zero_grad(self):
        if self.flat is None:
            for p in self.params:
                p.grad = None
        else:
            self.flat.zero_()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
