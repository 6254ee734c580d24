"""Drop-in test: the reference's OWN `BSMS_Simulator` (src/models/model.py, unmodified — imported from
/root/reference or from the byte-for-byte copy oracle/_ref made by oracle/build_ref.py) runs on top of
`bsms_gnn_b200.ops` exactly as INTEGRATION.md describes (the `from ops import MLP, BSGMP` swap), is fed
FRESH device copies of `m_gs` / `m_ids` every step the way `Trainer.move_to_device` does
(src/trainer/trainer.py:100-117), in both data modes of `model.forward` (src/models/model.py:189-200), and
must (a) match the all-reference model on the CPU and (b) never re-plan after the first step."""
import types

import pytest
import torch

from oracle import ref_import
from tests.util import load_hier, max_rel

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference modules not present (run oracle/build_ref.py)")

OUT_DIM, POS_DIM = 2, 2


def _cfg(depth):
    return types.SimpleNamespace(consistent_mesh=True, latent_dim=128, hidden_layer=3, unet_depth=depth, out_dim=OUT_DIM,
                                 pos_dim=POS_DIM, accumulation_steps=2)


def _build_pair(depth, dev):
    from bsms_gnn_b200 import ops as b200_ops
    torch.manual_seed(0)
    ref_mod = ref_import.load_simulator(None, device="cpu")
    ref_model = ref_mod.BSMS_Simulator(_cfg(depth))
    new_mod = ref_import.load_simulator(b200_ops, device=dev)
    assert new_mod.BSGMP is b200_ops.BSGMP and new_mod.MLP is b200_ops.MLP  # the swap took
    new_model = new_mod.BSMS_Simulator(_cfg(depth))
    assert sorted(new_model.state_dict()) == sorted(ref_model.state_dict())
    new_model.load_state_dict(ref_model.state_dict())
    return ref_model, new_model.to(dev)


def _batch(B, pos, m_gs, m_ids, seed):
    gen = torch.Generator().manual_seed(seed)
    N = pos.shape[0]
    state = torch.randn(B, N, OUT_DIM, generator=gen)
    ntype = (torch.rand(B, N, 1, generator=gen) > 0.8).float()
    node_in = torch.cat([state, pos.unsqueeze(0).expand(B, -1, -1), ntype], -1)
    node_tar = state + 0.1 * torch.randn(B, N, OUT_DIM, generator=gen)
    mask = (ntype == 0).float()
    # the collated consistent-mesh batch: every index tensor carries a leading batch dimension (model.py:190-192)
    gs = [g.unsqueeze(0).expand(B, -1, -1).contiguous() for g in m_gs]
    ids = [i.unsqueeze(0).expand(B, -1).contiguous() for i in m_ids]
    return node_in, node_tar, mask, gs, ids


def _to(data, dev):
    """Trainer.move_to_device (trainer.py:100-117): a fresh device copy of everything, every step."""
    if isinstance(data, (list, tuple)):
        return [_to(d, dev) for d in data]
    return data.to(dev)


@needs_ref
def test_reference_simulator_tuple_mode_fresh_copies_no_replan():
    from bsms_gnn_b200 import plan as P
    dev = torch.device("cuda", 0)
    m_gs, m_ids, pos, d = load_hier("grid12")
    ref_model, new_model = _build_pair(d, dev)
    P.clear_caches()
    for k in range(2):  # normaliser warm-up on identical data (model.py:108-125): no processor call
        data = _batch(3, pos, m_gs, m_ids, 10 + k)
        ref_model(data, True, True)
        new_model(_to(data, dev), True, True)
    builds = None
    for step in range(4):
        data = _batch(3, pos, m_gs, m_ids, 20 + step)
        want = ref_model(data, True, False)
        got = new_model(_to(data, dev), True, False)
        assert max_rel(got.detach().cpu(), want.detach()) < 1e-5, f"step {step}"
        if step == 0:
            builds = (P.STATS["level_builds"], P.STATS["hierarchy_builds"])
            assert builds[1] >= 1
    assert (P.STATS["level_builds"], P.STATS["hierarchy_builds"]) == builds, "fresh copies of a known mesh must not re-plan"
    assert P.STATS["content_hits"] >= 3
    # gradients through the whole reference model (encoder -> processor -> decoder)
    data = _batch(2, pos, m_gs, m_ids, 99)
    ref_model.zero_grad()
    new_model.zero_grad()
    ref_model(data, True, False).square().mean().backward()
    new_model(_to(data, dev), True, False).square().mean().backward()
    refp, newp = dict(ref_model.named_parameters()), dict(new_model.named_parameters())
    for name in ["encode.seq.0.weight", "process.down_gmps.0.mlp_edge.seq.0.weight", "process.bottom_gmp.mlp_node.seq.6.bias",
                 "process.up_gmps.1.mlp_edge.seq.4.weight", "decode.seq.6.weight"]:
        assert max_rel(newp[name].grad.cpu(), refp[name].grad) < 5e-4, name


@needs_ref
def test_reference_simulator_data_list_mode():
    """consistent_mesh=False: a list of per-level PyG-`Data`-like objects (x, y, mask, edge_index, face)."""
    from bsms_gnn_b200 import plan as P
    dev = torch.device("cuda", 0)
    m_gs, m_ids, pos, d = load_hier("grid12")
    ref_model, new_model = _build_pair(d, dev)

    def data_list(seed, device):
        node_in, node_tar, mask, _, _ = _batch(1, pos, m_gs, m_ids, seed)
        lst = []
        for l in range(d + 1):
            ns = types.SimpleNamespace(edge_index=m_gs[l].clone().to(device), face=(m_ids[l].clone().to(device) if l < d else None))
            if l == 0:
                ns.x, ns.y, ns.mask = node_in[0].to(device), node_tar[0].to(device), mask[0].to(device)
            lst.append(ns)
        return lst

    for k in range(2):
        ref_model(data_list(5 + k, "cpu"), False, True)
        new_model(data_list(5 + k, dev), False, True)
    P.clear_caches()
    for step in range(3):
        want = ref_model(data_list(30 + step, "cpu"), False, False)
        got = new_model(data_list(30 + step, dev), False, False)
        assert got.shape == want.shape
        assert max_rel(got.detach().cpu(), want.detach()) < 1e-5, f"step {step}"
    assert P.STATS["hierarchy_builds"] >= 1


def test_bind_mesh_and_identity_cache():
    from bsms_gnn_b200 import plan as P
    from bsms_gnn_b200.ops import BSGMP
    from oracle import bsms_oracle as O
    dev = torch.device("cuda", 0)
    m_gs, m_ids, pos, d = load_hier("grid12")
    model = BSGMP(d, 128, 3, 2).to(dev)
    model.load_state_dict(O.init_params(d, pos_dim=2, seed=3))
    h = torch.randn(pos.shape[0], 128, generator=torch.Generator().manual_seed(4)).to(dev)
    gs, ids, p = [g.to(dev) for g in m_gs], [i.to(dev) for i in m_ids], pos.to(dev)
    P.clear_caches()
    with torch.no_grad():
        want = model(h, ids, gs, p)
        f0 = P.STATS["fingerprints"]
        model(h, ids, gs, p)  # same tensors: identity hit, no fingerprint launch
        assert P.STATS["fingerprints"] == f0
        model.bind_mesh(gs, ids, pos.shape[0])
        got = model(h, [i.clone() for i in ids], [g.clone() for g in gs], p)  # bound: fresh copies cost nothing
        assert P.STATS["fingerprints"] == f0
        assert max_rel(got, want) < 2e-6  # the default mode's edge stage reduces with red.add: order varies run to run
        with pytest.raises(RuntimeError):
            model(h[:-1], ids, gs, p[:-1])
        model.unbind_mesh()
        # a DIFFERENT mesh with the same shapes must not alias the cached plan
        g_alt = [g.clone() for g in gs]
        g_alt[0] = g_alt[0].flip(0).contiguous()  # reversed edge directions at level 0: other content, same shape
        b0 = P.STATS["level_builds"]
        model(h, ids, g_alt, p)
        assert P.STATS["level_builds"] > b0
