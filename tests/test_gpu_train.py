"""The rest of the training step on the device (bsms_gnn_b200/train.py, csrc/train_step.cu) against the
torch CPU statement of what the reference's Trainer does (src/trainer/trainer.py:79-98,134-156):
masked RMSE and its gradient, clip_grad_norm_ + torch.optim.AdamW + the warmup-cosine schedule
(src/utils/basic.py:168-184).  fp32 elementwise arithmetic: 2e-6 relative."""
import math

import pytest
import torch

from tests.util import max_rel

pytestmark = pytest.mark.gpu


def ref_rmse(pred, tar, mask):
    se = (pred - tar) ** 2
    return torch.sqrt((se * mask).sum() / mask.sum() / se.shape[-1])  # trainer.py:96-98


@pytest.mark.parametrize("shape,C", [((3, 144), 2), ((1, 5184), 3), ((7,), 1)])
def test_masked_rmse_forward_backward(shape, C):
    from bsms_gnn_b200.train import masked_rmse
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(1)
    pred = torch.randn(*shape, C, generator=gen)
    tar = torch.randn(*shape, C, generator=gen)
    mask = (torch.rand(*shape, 1, generator=gen) > 0.3).float()
    pr = pred.double().requires_grad_(True)
    want = ref_rmse(pr, tar.double(), mask.double())
    (want * 1.7).backward()
    pg = pred.to(dev).requires_grad_(True)
    got = masked_rmse(pg, tar.to(dev), mask.to(dev))
    (got * 1.7).backward()
    assert abs(float(got) - float(want)) < 2e-6 * abs(float(want))
    assert max_rel(pg.grad.cpu(), pr.grad) < 2e-6


def factor(step, warmup, max_iters):
    if step <= warmup:
        return step / warmup
    return 0.5 * (1 + math.cos(math.pi * (step - warmup) / (max_iters - warmup)))  # basic.py:178-184


@pytest.mark.parametrize("warmup,max_iters,grad_scale", [(3, 10, 1.0), (2, 8, 1e-3), (0, 0, 30.0)])
def test_flat_adamw_matches_torch(warmup, max_iters, grad_scale):
    from bsms_gnn_b200.train import FlatAdamW
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(2)
    shapes = [(128, 259), (128,), (128, 128), (3, 5), (7,), (128, 256)]
    cpu = [torch.nn.Parameter(torch.randn(*s, generator=gen)) for s in shapes]
    gpu = [torch.nn.Parameter(p.detach().clone().to(dev)) for p in cpu]
    lr, wd, clip = 1e-2, 1e-2, 1.0
    opt = torch.optim.AdamW(cpu, lr=lr, weight_decay=wd)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, (lambda s: factor(s, warmup, max_iters)) if max_iters else (lambda s: 1.0))
    flat = FlatAdamW(gpu, lr=lr, weight_decay=wd, max_norm=clip, warmup_steps=warmup, decay_steps=max_iters)
    for g, c in zip(gpu, cpu):
        assert torch.equal(g.detach().cpu(), c.detach())  # re-homing kept the values
    for step in range(6):
        grads = [grad_scale * torch.randn(*s, generator=gen) for s in shapes]
        for p, g in zip(cpu, grads):
            p.grad = g.clone()
        for p, g in zip(gpu, grads):
            p.grad = g.to(dev)
        norm = torch.nn.utils.clip_grad_norm_(cpu, clip)
        opt.step()
        lr_used = opt.param_groups[0]["lr"]
        sched.step()
        opt.zero_grad()
        flat.step()
        assert abs(flat.last_lr() - lr_used) <= 1e-6 * max(lr_used, 1e-12), (step, flat.last_lr(), lr_used)
        want_coef = min(1.0, clip / (float(norm) + 1e-6))
        assert abs(flat.last_clip_coef() - want_coef) < 2e-6
        for i, (g, c) in enumerate(zip(gpu, cpu)):
            assert max_rel(g.detach().cpu(), c.detach()) < 2e-6, (step, i)
    assert flat.steps_done() == 6
