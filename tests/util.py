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
