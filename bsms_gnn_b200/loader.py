"""Device-side feeding of the training step: what the reference does on the host every step
— `Trainer.move_to_device` re-uploads features AND every index tensor of the hierarchy
(src/trainer/trainer.py:100-117), the datapipe draws the training noise on the CPU
(src/datasets/base.py:274-289) — becomes:

  * the hierarchy `(m_gs, m_ids)` is uploaded ONCE (from tensors or from the reference's
    `mmesh_layer_{d}.dat` cache through `mmesh_io`) and stays resident; the per-step tuple carries zero-copy
    batch-dimension views of it, so the processor's identity cache hits without a launch;
  * per-step tensors (node_in, node_tar, node_mask) travel from pinned host staging on a copy stream into one of
    two device slots while the previous step computes;
  * the training noise is injected on the device (`bsms_inject_noise`, Philox stream keyed by seed / step).

The tuple layout is exactly what `BSMS_Simulator.forward(data, consistent_mesh=True, ...)` unpacks
(src/models/model.py:189-192): (node_in [B,N,C+P+1], node_tar [B,N,C], node_mask [B,N,1], m_gs, m_ids) with a
leading batch dimension on every index tensor.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr


class DeviceFeeder:
    def __init__(self, m_gs, m_ids, device, noise_level=None, noise_gamma=1.0, seed=0, slots=2):
        self.device = torch.device(device)
        self.m_gs = [torch.as_tensor(g).to(self.device, torch.int64).contiguous() for g in m_gs]
        self.m_ids = [torch.as_tensor(i).to(self.device, torch.int64).contiguous() for i in m_ids]
        # batch-dimension VIEWS: `g[0]` in model.forward is then the same storage / shape / version every step
        self._gs_b = [g.unsqueeze(0) for g in self.m_gs]
        self._ids_b = [i.unsqueeze(0) for i in self.m_ids]
        self.noise_level = None if noise_level is None else [float(v) for v in noise_level]
        self.noise_gamma, self.seed = float(noise_gamma), int(seed)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * slots
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.free = [torch.cuda.Event() for _ in range(slots)]
        for ev in self.free:
            ev.record(torch.cuda.current_stream(self.device))
        self.put_count = self.get_count = 0

    @classmethod
    def from_mmesh(cls, path, device, **kw):
        """Hierarchy from the reference's multi-level mesh cache file (src/datasets/base.py:98-122)."""
        from . import mmesh_io
        m_gs, m_ids = mmesh_io.load_mmesh(path)
        return cls(m_gs, m_ids, device, **kw)

    def put(self, node_in, node_tar, node_mask):
        """Stage one host batch (any float tensors; pinned memory makes the copy asynchronous) into the next slot."""
        k = self.put_count % len(self.slots)
        host = [t if t.is_pinned() else t.contiguous().pin_memory() for t in (node_in.float(), node_tar.float(), node_mask.float())]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[k])  # the step that last consumed this slot has finished
            if self.slots[k] is None or any(d.shape != h.shape for d, h in zip(self.slots[k], host)):
                self.slots[k] = [torch.empty(h.shape, dtype=torch.float32, device=self.device) for h in host]
            for d, h in zip(self.slots[k], host):
                d.copy_(h, non_blocking=True)
            if self.noise_level is not None:
                d_in, d_tar, d_mask = self.slots[k]
                Cc = d_tar.shape[-1]
                lv = (C.c_float * Cc)(*self.noise_level[:Cc])
                with torch.cuda.device(self.device):
                    check(lib.bsms_inject_noise(ptr(d_in), d_in.shape[-1], ptr(d_tar), Cc, ptr(d_mask), d_tar.numel() // Cc, lv,
                                                self.noise_gamma, self.seed, self.put_count, C.c_void_p(self.copy_stream.cuda_stream)))
            self.ready[k].record(self.copy_stream)
        self._host_keepalive = host
        self.put_count += 1

    def get(self):
        """-> the data tuple of the oldest staged batch, valid on the current stream; call `done()` after the step."""
        if self.get_count >= self.put_count:
            raise RuntimeError("DeviceFeeder.get() without a staged batch")
        k = self.get_count % len(self.slots)
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        d_in, d_tar, d_mask = self.slots[k]
        self._in_use = k
        self.get_count += 1
        return d_in, d_tar, d_mask, self._gs_b, self._ids_b

    def done(self):
        """The step that consumed the last `get()` has been enqueued: its slot may be refilled once it completes."""
        self.free[self._in_use].record(torch.cuda.current_stream(self.device))
