"""The fused tcgen05 edge stage (edge_chain.cu), stage by stage, against an fp64 evaluation of the
same formulas on the GPU (torch fp64 matmuls — test infrastructure only).

Tolerances (max-abs error / max-abs reference, per stage):
  fp16x3 (the fp32-parity mode): 2e-6 — two-way fp16 split keeps 22 significant bits per operand.
  bf16: 2e-2 — single bf16 MMA, the precision config 3 of BASELINE.json names; not a parity mode.
"""
import ctypes as C

import pytest
import torch

from oracle import bsms_oracle as O
from tests.util import load_hier, max_rel

pytestmark = pytest.mark.gpu


def stage_reference(params, x, pos, src, dst, P):
    """fp64 per-stage values in dst-sorted edge order: a0, a1, a2, y (pre-LN), aggr."""
    W = {k: v.double() for k, v in params.items()}
    x = x.double()
    pos = pos.double()
    B, N, _ = x.shape
    W1 = W["mlp_edge.seq.0.weight"]
    xi, xj = x[:, src], x[:, dst]
    pp = pos if pos.dim() == 3 else pos.unsqueeze(0).expand(B, -1, -1)
    d = pp[:, src] - pp[:, dst]
    fiber = torch.cat([d, d.norm(dim=-1, keepdim=True)], -1)
    u0 = torch.cat([fiber, xi, xj], -1) @ W1.T + W["mlp_edge.seq.0.bias"]
    a0 = u0.relu()
    a1 = (a0 @ W["mlp_edge.seq.2.weight"].T + W["mlp_edge.seq.2.bias"]).relu()
    a2 = (a1 @ W["mlp_edge.seq.4.weight"].T + W["mlp_edge.seq.4.bias"]).relu()
    y = a2 @ W["mlp_edge.seq.6.weight"].T + W["mlp_edge.seq.6.bias"]
    m = (y - y.mean(-1, keepdim=True)) / torch.sqrt(y.var(-1, unbiased=False, keepdim=True) + 1e-5)
    aggr = torch.zeros(B, N, 128, dtype=torch.float64, device=x.device).index_add_(1, dst, m)
    return [a0, a1, a2, y], aggr


@pytest.mark.parametrize("mode,tol", [("bf16", 2e-2), ("fp16x3", 2e-6)])
@pytest.mark.parametrize("hname,level,B,P,pos_batched", [("grid12", 0, 1, 2, False), ("grid44", 0, 2, 2, True),
                                                         ("ico3", 1, 3, 3, False), ("grid72", 4, 2, 2, True)])
def test_edge_stage_by_stage(mode, tol, hname, level, B, P, pos_batched):
    from bsms_gnn_b200 import _lib
    from bsms_gnn_b200.ops import GMP, _weights_struct
    from bsms_gnn_b200.plan import LevelPlan
    dev = torch.device("cuda:0")
    m_gs, m_ids, pos0, d = load_hier(hname)
    n = [pos0.shape[0]] + [len(i) for i in m_ids]
    N, g = n[level], m_gs[level].to(dev)
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(B, N, 128, generator=gen).to(dev)
    pos = (torch.randn(B, N, P, generator=gen) if pos_batched else torch.randn(N, P, generator=gen)).to(dev)
    params = {k[len("bottom_gmp."):]: v.to(dev) for k, v in O.init_params(0, pos_dim=P, seed=11).items()}
    gmp = GMP(128, 3, P).to(dev)
    gmp.load_state_dict(params)
    plan = LevelPlan(g, N)
    E = plan.n_edges
    refs, aggr_ref = stage_reference(params, x, pos, plan.src_d.long(), plan.dst_d.long(), P)
    plist = [p.detach().contiguous() for p in gmp._params()]
    w = _weights_struct(plist)
    ws = torch.empty(B * N * 256 * 4 + (1 << 20), dtype=torch.uint8, device=dev)
    errs = {}
    for stage in range(4):
        dbg = torch.zeros(B * E, 128, device=dev)
        aggr = torch.empty(B, N, 128, device=dev)
        _lib.check(_lib.lib.bsms_debug_edge_stage(plan.byref(), C.byref(w), _lib.ptr(x), _lib.ptr(pos),
                                                  1 if pos_batched else 0, B, P, _lib.MODES[mode], stage,
                                                  _lib.ptr(dbg), _lib.ptr(aggr), _lib.ptr(ws), ws.numel(),
                                                  _lib.stream_ptr()))
        torch.cuda.synchronize()
        errs[f"stage{stage}"] = max_rel(dbg.view(B, E, 128), refs[stage])
        errs["aggr"] = max_rel(aggr, aggr_ref)
    print(f"\n[{mode} {hname} L{level} B{B}] " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    assert errs["stage0"] < 2e-6
    for k, v in errs.items():
        assert v < tol, (k, v)


@pytest.mark.parametrize("mode,tol", [("fp16x3", 1e-5), ("bf16", 3e-2)])
def test_bsgmp_forward_tensor_modes(mode, tol):
    """Whole processor forward with the fused edge stage against the reference golden."""
    from bsms_gnn_b200.ops import BSGMP
    from tests.util import bsgmp_inputs, load_npz
    dev = torch.device("cuda:0")
    for case, hname in [("grid12_b3", "grid12"), ("grid44", "grid44"), ("ico3", "ico3"), ("grid72d7", "grid72d7")]:
        rec = load_npz(f"bsgmp_{case}.npz")
        m_gs, m_ids, pos, d = load_hier(hname)
        h, ps = bsgmp_inputs(rec, pos, pos.shape[0])
        model = BSGMP(d, 128, 3, int(rec["P"]), mode=mode).to(dev)
        model.load_state_dict(O.init_params(d, pos_dim=int(rec["P"]), seed=int(rec["seed"])))
        with torch.no_grad():
            out = model(h.to(dev), [i.to(dev) for i in m_ids], [g.to(dev) for g in m_gs], ps.to(dev))
        rs = int(rec["row_stride"])
        err = max_rel(out.cpu()[..., ::rs, :], rec["out"])
        print(f"\n[{mode}] {case}: forward max-rel {err:.2e}")
        assert err < tol, (case, err)
