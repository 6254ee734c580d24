"""GPU parity of the individual operators (through the C-ABI) against the CPU oracle and the
reference goldens.  Tolerances: forward 1e-5 max-rel (BASELINE.json north_star: "within 1e-5
relative fp32"); index/topology outputs exact; gradients 5e-4 (fp32 noise floor of the reference's
own backward, see tests/test_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import bsms_oracle as O
from tests.util import load_hier, load_npz, max_rel

pytestmark = pytest.mark.gpu
FWD_TOL = 1e-5
GRAD_TOL = 5e-4


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def test_plan_build_matches_numpy(dev):
    from bsms_gnn_b200.plan import LevelPlan
    for name in ["chain11", "grid12", "ico3", "grid44"]:
        m_gs, m_ids, pos, d = load_hier(name)
        n = [pos.shape[0]] + [len(i) for i in m_ids]
        for l, g in enumerate(m_gs):
            if g.shape[1] == 0:
                continue
            p = LevelPlan(g.to(dev), n[l])
            gn = g.numpy()
            od = np.argsort(gn[1], kind="stable")
            os_ = np.argsort(gn[0], kind="stable")
            assert np.array_equal(p.perm_d.cpu().numpy(), od)
            assert np.array_equal(p.src_d.cpu().numpy(), gn[0][od])
            assert np.array_equal(p.dst_d.cpu().numpy(), gn[1][od])
            assert np.array_equal(p.src_s.cpu().numpy(), gn[0][os_])
            assert np.array_equal(p.dst_s.cpu().numpy(), gn[1][os_])
            rp = np.concatenate([[0], np.cumsum(np.bincount(gn[1], minlength=n[l]))])
            assert np.array_equal(p.rowptr_d.cpu().numpy(), rp)
            rs = np.concatenate([[0], np.cumsum(np.bincount(gn[0], minlength=n[l]))])
            assert np.array_equal(p.rowptr_s.cpu().numpy(), rs)
            inv = np.empty_like(od)
            inv[od] = np.arange(len(od))
            assert np.array_equal(p.s2d.cpu().numpy(), inv[os_])


def test_plan_rejects_bad_index(dev):
    from bsms_gnn_b200.plan import LevelPlan
    with pytest.raises(IndexError):
        LevelPlan(torch.tensor([[0, 5], [1, 0]], device=dev), 3)


def test_cal_ew_conv_unpool_against_reference_golden(dev):
    from bsms_gnn_b200.ops import Unpool, WeightedEdgeConv
    z = load_npz("ops_grid12.npz")
    m_gs, m_ids, pos, d = load_hier("grid12")
    g0, g1 = m_gs[0].to(dev), m_gs[1].to(dev)
    conv = WeightedEdgeConv()
    ew0, aw0 = conv.cal_ew(torch.ones(144, 1, device=dev), g0)
    assert max_rel(ew0.cpu(), z["ew0"]) < 1e-6 and max_rel(aw0.cpu(), z["aggr_w0"]) < 1e-6
    ew1, aw1 = conv.cal_ew(aw0[m_ids[0].to(dev)], g1)
    assert max_rel(ew1.cpu(), z["ew1"]) < 1e-6 and max_rel(aw1.cpu(), z["aggr_w1"]) < 1e-6
    x2, x3, pos3 = (torch.from_numpy(z[k]).to(dev) for k in ("x2", "x3", "pos3"))
    assert max_rel(conv(x2, g0, ew0).cpu(), z["conv_down_x2"]) < FWD_TOL
    assert max_rel(conv(x3, g0, ew0).cpu(), z["conv_down_x3"]) < FWD_TOL
    assert max_rel(conv(x2, g0, ew0, aggragating=False).cpu(), z["conv_up_x2"]) < FWD_TOL
    assert max_rel(conv(x3, g0, ew0, False).cpu(), z["conv_up_x3"]) < FWD_TOL
    assert max_rel(conv(pos3, g0, ew0).cpu(), z["conv_down_pos3"]) < FWD_TOL
    ids = m_ids[0].to(dev)
    up = Unpool()(x3[:, ids].contiguous(), 144, ids)
    assert np.array_equal(up.cpu().numpy(), z["unpool_x3"])


def test_chain11_known_answers(dev):
    from bsms_gnn_b200.ops import WeightedEdgeConv
    z = load_npz("chain11_known.npz")
    m_gs, m_ids, pos, d = load_hier("chain11")
    conv = WeightedEdgeConv()
    ew, aw = conv.cal_ew(torch.ones(11, 1, device=dev), m_gs[0].to(dev))
    assert np.array_equal(ew.cpu().numpy(), z["ew"])  # sequential sums in the reference's order: bit-exact
    assert np.array_equal(aw.cpu().numpy(), z["aggr_w"])
    px = conv(pos[:, :1].contiguous().to(dev), m_gs[0].to(dev), ew)
    assert max_rel(px.cpu(), z["conv_posx"]) < 1e-6


def test_cal_ew_error_parity(dev):
    from bsms_gnn_b200.ops import WeightedEdgeConv
    # highest-numbered node has no out-edge -> the reference's degree() is too short and it raises
    with pytest.raises(RuntimeError):
        WeightedEdgeConv().cal_ew(torch.ones(3, 1, device=dev), torch.tensor([[0, 1], [1, 2]], device=dev))
    with pytest.raises(RuntimeError):
        WeightedEdgeConv().cal_ew(torch.ones(3, 1, device=dev), torch.zeros(2, 0, dtype=torch.long, device=dev))
    with pytest.raises(NotImplementedError):
        WeightedEdgeConv()(torch.zeros(3, device=dev), torch.tensor([[0, 1], [1, 2]], device=dev),
                           torch.ones(2, device=dev))


def test_conv_autograd_is_the_adjoint(dev):
    from bsms_gnn_b200.ops import Unpool, WeightedEdgeConv
    m_gs, m_ids, pos, d = load_hier("ico3")
    g = m_gs[0]
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, pos.shape[0], 128, generator=gen)
    ew = torch.rand(g.shape[1], generator=gen)
    for aggr in (True, False):
        xr = x.clone().requires_grad_(True)
        O.edge_conv(xr, g, ew, aggr).square().sum().backward()
        xg = x.to(dev).requires_grad_(True)
        WeightedEdgeConv()(xg, g.to(dev), ew.to(dev), aggr).square().sum().backward()
        assert max_rel(xg.grad.cpu(), xr.grad) < FWD_TOL
    ids = m_ids[0]
    hc = x[:, ids].clone().requires_grad_(True)
    O.unpool(hc, pos.shape[0], ids).mul(x).sum().backward()
    hg = x[:, ids].to(dev).requires_grad_(True)
    Unpool()(hg, pos.shape[0], ids.to(dev)).mul(x.to(dev)).sum().backward()
    assert torch.equal(hg.grad.cpu(), hc.grad)


@pytest.mark.parametrize("key,xk,posk", [("gmp_x2_pos2", "x2", None), ("gmp_x3_pos2", "x3", None),
                                         ("gmp_x3_pos3", "x3", "pos3")])
def test_gmp_forward_against_reference_golden(dev, key, xk, posk):
    from bsms_gnn_b200.ops import BSGMP
    z = load_npz("ops_grid12.npz")
    m_gs, m_ids, pos, d = load_hier("grid12")
    model = BSGMP(2, 128, 3, 2).to(dev)
    model.load_state_dict(O.init_params(2, pos_dim=2, seed=3))
    x = torch.from_numpy(z[xk]).to(dev)
    p = torch.from_numpy(z[posk]).to(dev) if posk else pos.to(dev)
    with torch.no_grad():
        out = model.down_gmps[0](x, m_gs[0].to(dev), p)
    assert out.shape == x.shape
    assert max_rel(out.cpu(), z[key]) < FWD_TOL


def test_gmp_backward_against_oracle(dev):
    from bsms_gnn_b200.ops import GMP
    m_gs, m_ids, pos, d = load_hier("ico3")
    g = m_gs[1]
    n = len(m_ids[0])
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, n, 128, generator=gen)
    ps = torch.randn(2, n, 3, generator=gen)
    params = {k[len("bottom_gmp."):]: v for k, v in O.init_params(0, pos_dim=3, seed=4).items()}
    pr = {"g." + k: v.double().requires_grad_(True) for k, v in params.items()}
    xr = x.double().requires_grad_(True)
    ref = O.gmp(xr, g, ps.double(), pr, "g")
    w = torch.randn(ref.shape, generator=gen).double()
    (ref * w).sum().backward()
    m = GMP(128, 3, 3).to(dev)
    m.load_state_dict(params)
    xg = x.to(dev).requires_grad_(True)
    out = m(xg, g.to(dev), ps.to(dev))
    (out * w.float().to(dev)).sum().backward()
    assert max_rel(out.detach().cpu(), ref.detach()) < FWD_TOL
    assert max_rel(xg.grad.cpu(), xr.grad) < GRAD_TOL
    for k, v in m.named_parameters():
        assert max_rel(v.grad.cpu(), pr["g." + k].grad) < GRAD_TOL, k


def test_gmp_empty_graph_and_isolated_nodes(dev):
    from bsms_gnn_b200.ops import GMP
    params = {k[len("bottom_gmp."):]: v for k, v in O.init_params(0, pos_dim=2, seed=5).items()}
    m = GMP(128, 3, 2).to(dev)
    m.load_state_dict(params)
    x = torch.randn(5, 128, generator=torch.Generator().manual_seed(1))
    pos = torch.randn(5, 2, generator=torch.Generator().manual_seed(2))
    for g in (torch.zeros(2, 0, dtype=torch.long), torch.tensor([[0, 1, 1], [1, 0, 3]])):
        ref = O.gmp(x, g, pos, {"g." + k: v for k, v in params.items()}, "g")
        with torch.no_grad():
            out = m(x.to(dev), g.to(dev), pos.to(dev))
        assert max_rel(out.cpu(), ref) < FWD_TOL
