"""(f)1: the whole `_forward` of the reference model (src/models/model.py:127-164) and its rollout loop
(src/utils/rollout_utils.py:15-64) through bsms_gnn_b200.simulator — encoder / decoder / normalisers / mask /
residual fused around the processor, feedback and boundary re-imposition on the device, one CUDA graph per
step — against the UNMODIFIED reference model running its own ops on the CPU."""
import types

import pytest
import torch

from oracle import ref_import
from tests.util import load_hier, max_rel

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_import.available(), reason="reference modules not present")]


def _cfg(depth, C, P):
    return types.SimpleNamespace(consistent_mesh=True, latent_dim=128, hidden_layer=3, unet_depth=depth, out_dim=C, pos_dim=P,
                                 accumulation_steps=2)


def _models(depth, C, P, dev, seed=0):
    from bsms_gnn_b200 import ops as b200_ops
    torch.manual_seed(seed)
    ref_model = ref_import.load_simulator(None, device="cpu").BSMS_Simulator(_cfg(depth, C, P))
    new_model = ref_import.load_simulator(b200_ops, device=dev).BSMS_Simulator(_cfg(depth, C, P))
    return ref_model, new_model


def _inputs(B, pos, C, seed):
    gen = torch.Generator().manual_seed(seed)
    N, P = pos.shape
    state = torch.randn(B, N, C, generator=gen)
    ntype = (torch.rand(B, N, 1, generator=gen) > 0.8).float()
    node_in = torch.cat([state, pos.unsqueeze(0).expand(B, -1, -1), ntype], -1)
    mask = (ntype == 0).float()
    return node_in, mask


@pytest.mark.parametrize("hname,C,mode,tol", [("grid12", 2, "fp32", 1e-5), ("grid12", 2, "fp16x3", 1e-5), ("ico3", 3, "fp16x3", 1e-5),
                                              ("grid44", 2, "bf16", 3e-2)])
def test_fused_forward_and_rollout_match_reference(hname, C, mode, tol):
    from bsms_gnn_b200.simulator import FusedSimulator, GraphedRollout
    dev = torch.device("cuda", 0)
    m_gs, m_ids, pos, d = load_hier(hname)
    P = pos.shape[1]
    ref_model, new_model = _models(d, C, P, dev)
    gs_b = lambda B, device: [g.unsqueeze(0).expand(B, -1, -1).contiguous().to(device) for g in m_gs]
    ids_b = lambda B, device: [i.unsqueeze(0).expand(B, -1).contiguous().to(device) for i in m_ids]
    for k in range(2):  # normaliser statistics (model.py:108-125), then identical parameters on both sides
        node_in, mask = _inputs(3, pos, C, 50 + k)
        tar = node_in[..., :C] + 0.1 * torch.randn(3, pos.shape[0], C, generator=torch.Generator().manual_seed(60 + k))
        ref_model((node_in, tar, mask, gs_b(3, "cpu"), ids_b(3, "cpu")), True, True)
    new_model.load_state_dict(ref_model.state_dict())
    new_model = new_model.to(dev)
    sim = FusedSimulator(new_model, mode=mode)
    gs, ids = [g.to(dev) for g in m_gs], [i.to(dev) for i in m_ids]
    # ---- one forward, batched
    node_in, mask = _inputs(2, pos, C, 70)
    want = ref_model._forward(m_ids, m_gs, node_in, mask)
    got = sim(node_in.to(dev), mask.to(dev), gs, ids)
    e_fwd = max_rel(got.cpu(), want.detach())
    # ---- rollout with feedback and boundary re-imposition
    T = 5
    ic, mask1 = _inputs(1, pos, C, 71)
    cfg = None
    res_ref = ref_import.load().utils.rollout_one_traj(types.SimpleNamespace(model=ref_model), ic, torch.zeros(T, pos.shape[0], C), mask1,
                                                       gs_b(1, "cpu"), ids_b(1, "cpu"), cfg)
    res_eager = sim.rollout(ic.to(dev), mask1.to(dev), gs, ids, T)
    gr = GraphedRollout(sim, ic.to(dev), mask1.to(dev), gs, ids)
    res_graph = gr.run(T)
    torch.cuda.synchronize()
    e_roll = max(max_rel(res_eager[t].cpu(), res_ref[t]) for t in range(T))
    e_graph = max_rel(res_graph, res_eager)
    print(f"\n[{mode} {hname}] _forward max-rel {e_fwd:.2e}; {T}-step rollout max-rel {e_roll:.2e}; graph vs eager {e_graph:.2e}")
    assert e_fwd < tol
    assert e_roll < 4 * tol
    assert e_graph < (1e-6 if mode != "bf16" else 1e-2)
    gr.reset()
    again = gr.run(2)
    torch.cuda.synchronize()
    assert max_rel(again[1], res_graph[1]) < (1e-6 if mode != "bf16" else 1e-2)
