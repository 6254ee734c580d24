"""Shared helpers for the tests: golden loading and error metrics (SURVEY.md §8c metric)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_hier(name, device="cpu"):
    z = np.load(os.path.join(GOLDEN, f"hier_{name}.npz"))
    d = int(z["depth"])
    m_gs = [torch.from_numpy(z[f"g{l}"].astype(np.int64)).to(device) for l in range(d + 1)]
    m_ids = [torch.from_numpy(z[f"ids{l}"].astype(np.int64)).to(device) for l in range(d)]
    return m_gs, m_ids, torch.from_numpy(z["pos"]).to(device), d


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def bsgmp_inputs(rec, pos, n):
    """Re-create the seeded inputs make_golden.py fed the reference for a bsgmp_* record."""
    seed, batch, pb, P = int(rec["seed"]), int(rec["batch"]), int(rec["pos_batched"]), int(rec["P"])
    gen = torch.Generator().manual_seed(seed)
    h = torch.randn(*([batch, n, 128] if batch else [n, 128]), generator=gen)
    ps = pos.clone()
    if pb:
        ps = ps.unsqueeze(0) + 0.05 * torch.randn(batch, n, P, generator=gen)
    return h, ps


def max_rel(a, b):
    """max|a-b| / max|b|  — the survey's parity metric."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def l2_rel(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------------------------
# Kink-aware gradient comparison.  A ReLU pre-activation that lies closer to zero than the forward rounding noise
# of an implementation has an undetermined sign there, and the two one-sided derivatives differ by a whole gradient
# entry: one such entry among the millions of a level moves two rows of g_x by ~5e-3 and, because parameter
# gradients are random-sign sums over the rows, the parameter gradients by ~1/sqrt(rows).  The reference's own fp32
# runs differ from the fp64 oracle the same way.  The check below proves that a discrepancy is exactly that: it
# finds the oracle's pre-activations within `delta`·max|z| of zero whose graph neighbourhood contains a failing
# g_x row, and accepts the other one-sided derivative at those entries (and nowhere else).
class _MaskedRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, mask):
        ctx.save_for_backward(mask)
        return z.clamp_min(0)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask, None


class KinkRelu:
    """Drop-in for torch.relu in oracle.bsms_oracle.gmp: records every pre-activation tensor in call order and takes the
    other one-sided derivative at `flips` = {(call, flat index)}."""

    def __init__(self, flips=()):
        self.flips = {}
        for c, i in flips:
            self.flips.setdefault(c, []).append(i)
        self.z = []

    def __call__(self, z):
        c = len(self.z)
        self.z.append(z.detach())
        mask = z.detach() > 0
        if c in self.flips:
            m = mask.reshape(-1).clone()
            idx = torch.tensor(self.flips[c])
            m[idx] = ~m[idx]
            mask = m.view_as(mask)
        return _MaskedRelu.apply(z, mask.to(z.dtype))


def gmp_reference_kink_aware(x, g, pos, params, w, got_gx, tol, delta=1e-5, max_flips=8):
    """fp64 oracle gradients of sum(gmp(x) * w) for one GMP level, x [B, N, D].  Returns (out, g_x, {name: grad},
    flips): the plain oracle when `got_gx` already agrees within `tol`; otherwise the oracle with the other one-sided
    ReLU derivative at those near-zero pre-activations (|z| < delta·max|z|, delta = the forward tolerance) that
    touch a failing row and whose flip reduces the error."""
    from oracle import bsms_oracle as O

    def run(flips):
        pr = {"g." + k: v.double().clone().requires_grad_(True) for k, v in params.items()}
        xr = x.double().clone().requires_grad_(True)
        relu = KinkRelu(flips)
        out = O.gmp(xr, g, pos.double(), pr, "g", relu=relu)
        (out * w).sum().backward()
        return out.detach(), xr.grad, {k[2:]: v.grad for k, v in pr.items()}, relu.z

    out, gx, grads, zs = run(())
    err = max_rel(got_gx, gx)
    flips = []
    if err < tol:
        return out, gx, grads, flips
    B, N = x.shape[0], x.shape[1]
    src, dst = g[0], g[1]
    bad = ((got_gx.double() - gx).abs().amax(-1) > tol * gx.abs().max())  # [B, N]
    cands = []
    for c, z in enumerate(zs):
        near = (z.abs() < delta * z.abs().max()).nonzero()
        for b, r, col in near.tolist():
            if c < 3:  # edge MLP: row r is an edge; its gradient reaches x[src], x[dst]
                touched = [int(src[r]), int(dst[r])]
            else:  # node MLP: row r is a node; its gradient reaches x[r] and, through aggr, the senders of its in-edges
                touched = [r] + src[dst == r].tolist()
            if bool(bad[b, touched].any()):
                cands.append((float(z[b, r, col].abs()), c, (b * z.shape[1] + r) * z.shape[2] + col))
    # a flip is accepted when it removes error from the failing rows (summed, so that independent flips in
    # different rows are each recognised whichever of them carries the maximum)
    scale = float(tol * gx.abs().max())

    def miss(ref_gx):
        return float((got_gx.double() - ref_gx).abs().amax(-1)[bad].sum())

    cur = miss(gx)
    for _, c, i in sorted(cands)[:4 * max_flips]:
        trial = run(flips + [(c, i)])
        m2 = miss(trial[1])
        if m2 < cur - 0.5 * scale:
            flips.append((c, i))
            out, gx, grads, cur = trial[0], trial[1], trial[2], m2
            if max_rel(got_gx, gx) < tol or len(flips) >= max_flips:
                break
    return out, gx, grads, flips
