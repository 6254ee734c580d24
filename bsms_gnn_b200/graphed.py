"""CUDA-graph capture of the processor forward for rollouts.

The reference's rollout (src/utils/rollout_utils.py:15-64) calls the model T-1 ≈ 599 times in
sequence on a 2 k–5 k node mesh: ~140 tiny kernels per processor forward, i.e. launch- and
Python-bound.  For a fixed mesh the whole `BSGMP.forward` is static — plans,
packed weights, workspaces and launch geometry do not change between steps — so it is captured
once into a CUDA graph and replayed with new inputs copied into static buffers.
"""
from __future__ import annotations

import torch


class GraphedBSGMP:
    """`y = graphed(h[, pos])` replays the captured forward; inference only (no autograd)."""

    def __init__(self, model, m_ids, m_gs, h_example, pos_example, warmup: int = 2):
        self.model, self.m_ids, self.m_gs = model, m_ids, m_gs
        self.h = h_example.detach().clone()
        self.pos = pos_example.detach().clone()
        side = torch.cuda.Stream(device=self.h.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):  # builds and caches the plans; sizes the workspaces
                model(self.h, m_ids, m_gs, self.pos)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = model(self.h, m_ids, m_gs, self.pos)

    def __call__(self, h, pos=None):
        self.h.copy_(h, non_blocking=True)
        if pos is not None:
            self.pos.copy_(pos, non_blocking=True)
        self.graph.replay()
        return self.out


class GraphedStep:
    """Captures a whole training step `fn()` over STATIC tensors — forward, backward and the collectives in
    between (halo exchanges, gradient all-reduce: NCCL operations are capturable) — into one CUDA graph.
    The partitioned large-mesh step is ~1.3 k small launches and ~60 point-to-point batches per rank,
    i.e. bound by CPU launch cost once the kernels are fast; replaying removes that cost.
    `fn` must not synchronise or read results on the host; inputs are updated in place before a replay."""

    def __init__(self, fn, warmup: int = 3):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread may query events while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out
