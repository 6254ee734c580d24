"""The rest of the training step around the processor, on the device (reference:
src/trainer/trainer.py:79-98,134-156): masked RMSE loss, global-norm gradient clipping, AdamW with the
warmup-cosine schedule (src/utils/basic.py:168-184, configs/opt/default.yaml).

`masked_rmse(pred, tar, mask)` is an autograd function over `bsms_masked_rmse`.
`FlatAdamW(params, ...)` re-homes every parameter into ONE flat fp32 buffer (the parameters become views,
names / shapes / state_dict are unchanged), keeps gradients and both moments flat as well, and does
clip + update in three launches (`bsms_clip_adamw_step`) with no host synchronisation: step counter,
learning rate, bias corrections and the clip coefficient are device scalars.  With torch.distributed
initialised the flat gradient is all-reduced first (one NCCL call, dist.GradBucket semantics).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


class _MaskedRMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, tar, mask):
        _lib.require_cuda(pred, tar, mask)
        pred_c, tar_c = pred.contiguous().float(), tar.contiguous().float()
        Cc = pred_c.shape[-1]
        rows = pred_c.numel() // Cc
        mask_c = mask.to(torch.float32).expand(*pred_c.shape[:-1], 1).contiguous()
        if mask_c.numel() != rows:
            raise _lib.BsmsError(f"mask must broadcast to [..., 1] over pred {tuple(pred.shape)}, got {tuple(mask.shape)}")
        acc = torch.empty(2, dtype=torch.float64, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            check(lib.bsms_masked_rmse(ptr(pred_c), ptr(tar_c), ptr(mask_c), rows, Cc, ptr(acc), None, ptr(loss), None, stream_ptr()))
        ctx.save_for_backward(pred_c, tar_c, mask_c, acc)
        return loss

    @staticmethod
    def backward(ctx, g):
        pred_c, tar_c, mask_c, acc = ctx.saved_tensors
        Cc = pred_c.shape[-1]
        rows = pred_c.numel() // Cc
        grad = torch.empty_like(pred_c)
        g32 = g.to(torch.float32).contiguous()
        acc2 = torch.empty_like(acc)
        with torch.cuda.device(pred_c.device):
            check(lib.bsms_masked_rmse(ptr(pred_c), ptr(tar_c), ptr(mask_c), rows, Cc, ptr(acc2), ptr(g32), None, ptr(grad), stream_ptr()))
        return grad, None, None


def masked_rmse(pred, tar, mask):
    """sqrt(((pred - tar)^2 * mask).sum() / mask.sum() / C)  — Trainer._loss_fn (trainer.py:96-98)."""
    return _MaskedRMSE.apply(pred, tar, mask)


class FlatAdamW:
    """clip_grad_norm_(max_norm) + torch.optim.AdamW(lr, betas, eps, weight_decay) + WarmupCosineDecayScheduler
    fused over flat buffers.  `warmup_steps = decay_steps = 0` gives a constant learning rate."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_norm=1.0, warmup_steps=0,
                 decay_steps=0, all_reduce=True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        _lib.require_cuda(*self.params)
        # every tensor starts on a 16-byte boundary inside the flat buffers (vectorised kernels)
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.n = off
        f = dict(dtype=torch.float32, device=dev)
        self.flat_p, self.flat_g = torch.zeros(off, **f), torch.zeros(off, **f)
        self.exp_avg, self.exp_avg_sq = torch.zeros(off, **f), torch.zeros(off, **f)
        self.state = torch.zeros(2, dtype=torch.float64, device=dev)
        self.hyper = torch.zeros(4, **f)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view  # same Parameter object (names, state_dict, autograd identity), new home
        self.grad_views = [self.flat_g[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]
        self.cfg = (float(lr), float(warmup_steps), float(decay_steps), float(betas[0]), float(betas[1]), float(eps),
                    float(weight_decay), float(max_norm))
        self.all_reduce = all_reduce

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def step(self):
        """Collect .grad of every parameter into the flat buffer (missing grads = 0), all-reduce it when
        distributed, clip, update; gradients are cleared (p.grad = None)."""
        have = [(v, p.grad) for v, p in zip(self.grad_views, self.params) if p.grad is not None]
        if len(have) != len(self.params):
            self.flat_g.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if self.all_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g)
        lr, wu, dc, b1, b2, eps, wd, mn = self.cfg
        with torch.cuda.device(self.flat_p.device):
            check(lib.bsms_clip_adamw_step(ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.n,
                                           ptr(self.state), ptr(self.hyper), lr, wu, dc, b1, b2, eps, wd, mn, 0, stream_ptr()))
        _lib.WEIGHTS_EPOCH[0] += 1  # packed weight images cached for inference are stale now
        self.zero_grad()

    # read-backs (synchronise): for logging / tests only
    def last_lr(self):
        return float(self.hyper[0])

    def last_clip_coef(self):
        return float(self.hyper[3])

    def steps_done(self):
        return int(self.state[0])
