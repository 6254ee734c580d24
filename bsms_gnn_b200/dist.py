"""Multi-GPU plumbing for the batch-parallel (replica) mode: one process per GPU, every rank owns its
own samples of the batch (the axis `nn.DataParallel` would have split, reference
src/trainer/trainer.py:15-18) and the parameter gradients are summed over ranks once per step.

The exchange is ONE all-reduce over a flat, persistent fp32 bucket holding every parameter gradient
of the processor (2.15 M values at depth 6, 8.6 MB) — NVSwitch makes its cost latency- not
link-bound, so one bucket beats per-tensor calls.  `torch.distributed` (NCCL on GPUs, gloo in the
CPU tests) is the plumbing; the bucket logic is device-agnostic so the gloo tests exercise it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradBucket:
    """Flat gradient bucket: `.grad` of every parameter becomes a view into one buffer."""

    def __init__(self, params, average: bool = False):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(self.sizes), dtype=dt, device=dev)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(self.sizes), self.params)]
        self.average = average

    def collect(self):
        """Copy the parameter gradients produced by backward into the bucket (missing grads = 0)."""
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        if len(have) != len(self.views):
            self.flat.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])

    def allreduce(self):
        """Sum (or average) the bucket over all ranks and point `.grad` at the reduced views."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
            if self.average:
                self.flat.div_(dist.get_world_size())
        for v, p in zip(self.views, self.params):
            p.grad = v
        return self.flat

    def step_sync(self):
        self.collect()
        return self.allreduce()


def shard_batch(n_samples: int, rank: int, world: int):
    """Contiguous split of a batch of samples over ranks (remainder to the lowest ranks)."""
    base, rem = divmod(n_samples, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
