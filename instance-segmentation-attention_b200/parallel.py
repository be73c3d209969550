"""Single-node data parallelism for the training step: one process per GPU (torchrun), parameters
replicated, the image batch split across ranks, and ONE exchange per step -- a flat-bucket gradient
all-reduce (NCCL over NVLink on the GPU box; gloo in the CPU tests).  The reference has no
distributed code at all; this is the part the build adds (section 8e).

The only cross-image coupling in the loss is the q-regulariser's denominator int(sum(target)) over the
whole batch (/root/reference/code/lib/losses/discriminative.py:153-159): `global_q_denominator`
all-reduces the foreground count so the rank-averaged loss equals the single-process loss.
BatchNorm: the designed hot path has none; the VGG-style backbone has none either.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=world)
    return rank, world, local_rank


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(global_batch, rank, world):
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradBucket(object):
    """All gradients of `params` in ONE contiguous buffer; after `allreduce()` every p.grad is a view
    into it, so later backward passes accumulate straight into the bucket (no copy in steady state)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, device=ref.device, dtype=torch.float32)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def attach(self):
        """Make p.grad alias the bucket (keeps the current gradient values)."""
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            elif p.grad is None:
                v.zero_()
            p.grad = v

    def zero(self):
        self.flat.zero_()

    def allreduce(self, average=True):
        self.attach()
        if is_distributed():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())
        return self.flat


def global_q_denominator(local_foreground_count):
    """local count (0-dim / 1-element tensor) -> 1-element float tensor global_count / world_size."""
    t = local_foreground_count.detach().reshape(1).to(torch.float32).clone()
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(dist.get_world_size())
    return t


def allreduce_epoch_metrics(sums, n, device):
    """Per-rank metric sums + batch count of one epoch -> the SAME mean metrics on every rank.  Every rank validates on its
    own shard, ReduceLROnPlateau's learning rate is a per-rank kernel argument of the fused optimizer and rank 0 decides
    which checkpoint is best: fed rank-local costs, the replicas would drop the learning rate on different epochs and
    drift apart silently.  Works on any backend (gloo in the CPU tests)."""
    keys = sorted(sums)
    packed = torch.stack([torch.as_tensor(sums[k], dtype=torch.float64, device=device).reshape(()) for k in keys]
                         + [torch.tensor(float(n), dtype=torch.float64, device=device)])
    if is_distributed():
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    packed = packed.cpu()
    total = float(packed[-1].item())
    return {k: float(packed[i]) / max(total, 1.0) for i, k in enumerate(keys)}


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
