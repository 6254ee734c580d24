"""(f)3: device-side feeding of the training step (bsms_gnn_b200/loader.py): resident hierarchy (loaded from the
reference's mmesh cache format), pinned double-buffered H2D of the per-step tensors, training noise on the device
with the semantics of src/datasets/base.py:274-289."""
import os

import pytest
import torch

from tests.util import load_hier, max_rel

pytestmark = pytest.mark.gpu


def _batch(B, N, C, P, seed):
    gen = torch.Generator().manual_seed(seed)
    node_in = torch.randn(B, N, C + P + 1, generator=gen)
    node_in[..., -1] = (torch.rand(B, N, generator=gen) > 0.7).float()
    node_tar = torch.randn(B, N, C, generator=gen)
    mask = (node_in[..., -1:] == 0).float()
    return node_in, node_tar, mask


def test_feeder_roundtrip_noise_semantics_and_resident_hierarchy(tmp_path):
    from bsms_gnn_b200 import mmesh_io, plan as P_
    from bsms_gnn_b200.loader import DeviceFeeder
    from bsms_gnn_b200.ops import BSGMP
    from oracle import bsms_oracle as O
    dev = torch.device("cuda", 0)
    m_gs, m_ids, pos, d = load_hier("grid12")
    path = os.path.join(tmp_path, "mmesh_layer_%d.dat" % d)
    mmesh_io.save_mmesh(path, m_gs, m_ids)
    N, C, P = pos.shape[0], 2, 2
    # ---- no noise: what comes out is what went in; the index tensors carry a batch dimension over resident storage
    fd = DeviceFeeder.from_mmesh(path, dev)
    batches = [_batch(3, N, C, P, 10 + k) for k in range(3)]
    fd.put(*batches[0])
    for k in range(3):
        if k + 1 < 3:
            fd.put(*batches[k + 1])  # staged on the copy stream while "step k" runs
        node_in, node_tar, mask, gs, ids = fd.get()
        assert torch.equal(node_in.cpu(), batches[k][0]) and torch.equal(node_tar.cpu(), batches[k][1]) and torch.equal(mask.cpu(), batches[k][2])
        assert all(g.shape[0] == 1 and torch.equal(g[0].cpu(), m) for g, m in zip(gs, m_gs))
        assert all(i.shape[0] == 1 and torch.equal(i[0].cpu(), m) for i, m in zip(ids, m_ids))
        fd.done()
    # the processor sees the SAME views every step: one plan, identity hits, no fingerprint launches after the first
    model = BSGMP(d, 128, 3, 2).to(dev)
    model.load_state_dict(O.init_params(d, pos_dim=2, seed=1))
    P_.clear_caches()
    h = torch.randn(3, N, 128, generator=torch.Generator().manual_seed(2)).to(dev)
    with torch.no_grad():
        for k in range(3):
            fd.put(*batches[k])
            node_in, _, _, gs, ids = fd.get()
            model(h, [i[0] for i in ids], [g[0] for g in gs], node_in[..., C:C + P].contiguous())  # model.py:190-192
            fd.done()
            if k == 0:
                f0, b0 = P_.STATS["fingerprints"], P_.STATS["hierarchy_builds"]
    assert P_.STATS["fingerprints"] == f0 and P_.STATS["hierarchy_builds"] == b0
    # ---- noise: zero on masked-out nodes, tar noise = (1 - gamma) * input noise, std = level, reproducible
    gamma, level = 0.25, [0.5, 2.0]
    big = _batch(8, 4096, C, P, 99)
    outs = []
    for rep in range(2):
        fn = DeviceFeeder(m_gs, m_ids, dev, noise_level=level, noise_gamma=gamma, seed=7)
        fn.put(*big)
        node_in, node_tar, mask, _, _ = fn.get()
        torch.cuda.synchronize()
        outs.append((node_in.cpu(), node_tar.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    d_in = outs[0][0][..., :C] - big[0][..., :C]
    d_tar = outs[0][1] - big[1]
    keep = big[2].expand(-1, -1, C) > 0
    assert float(d_in[~keep].abs().max()) == 0.0 and float(d_tar[~keep].abs().max()) == 0.0
    assert torch.equal(outs[0][0][..., C:], big[0][..., C:])  # mesh_pos / node_type untouched
    assert max_rel(d_tar, (1 - gamma) * d_in) < 1e-5
    for c in range(C):
        x = d_in[..., c][keep[..., c]]
        assert abs(float(x.mean())) < 0.05 * level[c]
        assert abs(float(x.std()) / level[c] - 1) < 0.03
    fn2 = DeviceFeeder(m_gs, m_ids, dev, noise_level=level, noise_gamma=gamma, seed=8)
    fn2.put(*big)
    other = fn2.get()[0].cpu()
    assert not torch.equal(other, outs[0][0])
