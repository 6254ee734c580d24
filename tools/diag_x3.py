#!/usr/bin/env python
"""Per-level diagnostic of one GMP block fwd+bwd in a given mode against the fp64 oracle (development aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bsms_gnn_b200.ops import GMP  # noqa: E402
from oracle import bsms_oracle as O  # noqa: E402
from tests.util import load_hier, max_rel  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
hname = sys.argv[2] if len(sys.argv) > 2 else "grid44"
gscale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
dev = torch.device("cuda", 0)
m_gs, m_ids, pos0, d = load_hier(hname)
n = [pos0.shape[0]] + [len(i) for i in m_ids]
P = pos0.shape[1]
only = int(os.environ.get("DIAG_LEVEL", "-1"))
for level in range(d + 1):
    if only >= 0 and level != only:
        continue
    N, g = n[level], m_gs[level]
    gen = torch.Generator().manual_seed(100 + level)
    x = torch.randn(int(os.environ.get("DIAG_B", "2")), N, 128, generator=gen)
    pos = torch.randn(N, P, generator=gen)
    params = {k[len("bottom_gmp."):]: v for k, v in O.init_params(0, pos_dim=P, seed=4).items()}
    pr = {"g." + k: v.double().requires_grad_(True) for k, v in params.items()}
    xr = x.double().requires_grad_(True)
    ref = O.gmp(xr, g, pos.double(), pr, "g")
    w = gscale * torch.randn(ref.shape, generator=gen).double()
    (ref * w).sum().backward()
    m = GMP(128, 3, P, mode=mode).to(dev)
    m.load_state_dict(params)
    xg = x.to(dev).requires_grad_(True)
    out = m(xg, g.to(dev), pos.to(dev))
    (out * w.float().to(dev)).sum().backward()
    errs = {k.replace("mlp_", "").replace(".seq", ""): max_rel(v.grad.cpu(), pr["g." + k].grad) for k, v in m.named_parameters()}
    worst = max(errs, key=errs.get)
    deg = g.shape[1] / max(N, 1)
    print(f"[{mode} {hname} L{level} N={N} E={g.shape[1]} deg={deg:.1f}] out {max_rel(out.detach().cpu(), ref.detach()):.1e} g_x {max_rel(xg.grad.cpu(), xr.grad):.1e} "
          f"worst param {errs[worst]:.1e} ({worst}) | " + " ".join(f"{k}:{v:.0e}" for k, v in errs.items()), flush=True)
    if os.environ.get("DIAG_LOC"):
        dg = (xg.grad.cpu().double() - xr.grad).abs().amax(-1)  # [B, N]
        thr = 1e-4 * xr.grad.abs().max()
        bad = (dg > thr).nonzero()
        print(f"  g_x rows over 1e-4: {bad.shape[0]} of {dg.numel()}; first {bad[:12].tolist()} last {bad[-6:].tolist()}")
        dst = g[1]
        src = g[0]
        indeg = torch.bincount(dst, minlength=N)
        outdeg = torch.bincount(src, minlength=N)
        bn = bad[:, 1].unique()
        print(f"  bad nodes: {bn.numel()}; indeg of bad {indeg[bn].float().mean():.2f} (all {indeg.float().mean():.2f}) min/max {indeg[bn].min()}/{indeg[bn].max()}; outdeg {outdeg[bn].float().mean():.2f}; max indeg overall {indeg.max()}")
        for k in ["mlp_edge.seq.4.weight", "mlp_edge.seq.4.bias", "mlp_edge.seq.6.bias"]:
            a, b = dict(m.named_parameters())[k].grad.cpu().double(), pr["g." + k].grad
            dd = (a - b).abs()
            print(f"  {k}: max|ref| {b.abs().max():.3e} mean|ref| {b.abs().mean():.3e} max|diff| {dd.max():.3e} mean|diff| {dd.mean():.3e} mean signed diff {(a - b).mean():.3e}")
