"""fp64 model of BSMS_MODE_BF16 arithmetic.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

The reference has no bf16 mode (it trains in fp32, src/ops/basic.py:12-18), so the benchmarked bf16
configuration (BASELINE.json config 3) is checked twice: its forward against the reference goldens at
a bf16-sized tolerance, and its forward AND gradients against this file — the oracle's formulas
(oracle/bsms_oracle.py, pinned to the reference) evaluated in fp64 with bf16 rounding inserted at
exactly the points where the sm_100a kernels round: both operands of every tensor-core GEMM (the
per-node projections Ps / Pd, edge layers 1..3, the four node-MLP layers) and the edge biases b2..b4
(they ride in the MMA).  Fiber term, b1, node biases, ReLU, LayerNorm, segment sums and the
restriction / prolongation stay un-rounded.  Rounding uses a straight-through gradient, which is what
the kernels' backward implements (gradients of the rounded forward, gradient tiles themselves
rounded to bf16 inside the backward GEMMs — that last part is NOT modelled and is the residual the
tests' tolerance covers).
"""
import torch

from . import bsms_oracle as O


def rb(t):
    """round to bf16, straight-through gradient"""
    return t + (t.detach().float().bfloat16().to(t.dtype) - t.detach())


def gmp_bf16(x, g, pos, p, prefix):
    lin = torch.nn.functional.linear
    P = pos.shape[-1]
    i, j = g[0], g[1]
    W1, b1 = p[f"{prefix}.mlp_edge.seq.0.weight"], p[f"{prefix}.mlp_edge.seq.0.bias"]
    ps = lin(rb(x), rb(W1[:, P + 1:P + 1 + 128]))
    pd = lin(rb(x), rb(W1[:, P + 1 + 128:]))
    pp = pos if (pos.dim() == x.dim()) else pos.unsqueeze(0).expand(x.shape[0], -1, -1)
    dd = pp[..., i, :] - pp[..., j, :]
    fiber = torch.cat([dd, dd.norm(dim=-1, keepdim=True)], -1)
    h = torch.relu(ps[..., i, :] + pd[..., j, :] + b1 + lin(fiber, W1[:, :P + 1]))
    for l in (2, 4):
        h = torch.relu(lin(rb(h), rb(p[f"{prefix}.mlp_edge.seq.{l}.weight"]), rb(p[f"{prefix}.mlp_edge.seq.{l}.bias"])))
    y = lin(rb(h), rb(p[f"{prefix}.mlp_edge.seq.6.weight"]), rb(p[f"{prefix}.mlp_edge.seq.6.bias"]))
    mu = y.mean(-1, keepdim=True)
    e = (y - mu) / torch.sqrt(((y - mu) ** 2).mean(-1, keepdim=True) + O.EPS_LN)
    aggr = O.scatter_sum(e, j, -2, x.shape[-2])
    n = torch.cat([x, aggr], -1)
    for l in (0, 2, 4):
        n = torch.relu(lin(rb(n), rb(p[f"{prefix}.mlp_node.seq.{l}.weight"]), p[f"{prefix}.mlp_node.seq.{l}.bias"]))
    yn = lin(rb(n), rb(p[f"{prefix}.mlp_node.seq.6.weight"]), p[f"{prefix}.mlp_node.seq.6.bias"])
    mu = yn.mean(-1, keepdim=True)
    return (yn - mu) / torch.sqrt(((yn - mu) ** 2).mean(-1, keepdim=True) + O.EPS_LN) + x


def bsgmp_bf16(h, m_ids, m_gs, pos, p, unet_depth):
    """The schedule of oracle.bsgmp (src/ops/BSMS.py:39-104) with bf16-modelled GMP blocks."""
    down_outs, down_ps, cts = [], [], []
    w = pos.new_ones((pos.shape[-2], 1))
    for l in range(unet_depth):
        h = gmp_bf16(h, m_gs[l], pos, p, f"down_gmps.{l}")
        down_outs.append(h)
        down_ps.append(pos)
        ew, w = O.cal_ew(w, m_gs[l])
        h = O.edge_conv(h, m_gs[l], ew)
        pos = O.edge_conv(pos, m_gs[l], ew)
        cts.append(ew)
        h, pos, w = h[..., m_ids[l], :], pos[..., m_ids[l], :], w[m_ids[l]]
    h = gmp_bf16(h, m_gs[unet_depth], pos, p, "bottom_gmp")
    for k in range(unet_depth):
        l = unet_depth - 1 - k
        h = O.unpool(h, down_outs[l].shape[-2], m_ids[l])
        h = O.edge_conv(h, m_gs[l], cts[l], aggragating=False)
        h = gmp_bf16(h, m_gs[l], down_ps[l], p, f"up_gmps.{k}")
        h = h + down_outs[l]
    return h
