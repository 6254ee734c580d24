"""Generate golden vectors by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src through oracle/ref_import.py, feeds it seeded inputs and the
deterministic parameters of oracle.bsms_oracle.init_params (loaded with load_state_dict, so the
reference modules compute with known weights), and writes small .npz fixtures next to this file:

  hier_<case>.npz   reference BistrideMultiLayerGraph output (m_gs, m_ids as int32) + positions
  ops_grid12.npz    per-op outputs (GMP, cal_ew, WeightedEdgeConv both directions, Unpool)
  bsgmp_<case>.npz  BSGMP forward outputs (+ gradients of out.square().mean() for small cases)

The GPU box has no reference tree; tests there compare against these files.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bsms_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def load_pkg_module(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "bsms_gnn_b200", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    R = ref_import.load()
    meshgen = load_pkg_module("meshgen")
    torch.set_num_threads(8)

    def hierarchy(fe, depth, n, pos):
        _, m_es, m_ids = R.graph_wrappers.BistrideMultiLayerGraph(fe, depth, n, pos).get_multi_layer_graphs()
        return [np.asarray(e, dtype=np.int64).reshape(2, -1) for e in m_es], [np.asarray(i, dtype=np.int64) for i in m_ids]

    cases = {}
    fe = np.array([[0, 1, 2, 3, 4, 5, 6, 7, 8, 9], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]])
    fe = np.concatenate((fe, fe[::-1]), axis=1)
    pos = np.stack([np.arange(11.0), np.zeros(11), np.zeros(11)], 1).astype(np.float32)
    cases["chain11"] = (fe, 2, pos)
    for nx, d in [(12, 2), (44, 5), (72, 6)]:
        p, cells = meshgen.tri_grid(nx, nx)
        cases[f"grid{nx}"] = (R.mesh_convertions.to_flat_edge(cells, "tri"), d, p)
    p, cells = meshgen.icosphere(3)
    cases["ico3"] = (R.mesh_convertions.to_flat_edge(cells, "tri"), 3, p)
    p1, c1 = meshgen.tri_grid(9, 7)
    p2, c2 = meshgen.tri_grid(5, 6, seed=1)
    cells = np.concatenate([c1, c2 + p1.shape[0]])
    cases["twoclusters"] = (R.mesh_convertions.to_flat_edge(cells, "tri"), 3, np.concatenate([p1, p2 + 20]))
    # depth-7 airfoil-like grid: deepest level has 1 node / 0 edges
    p, cells = meshgen.tri_grid(72, 72)
    cases["grid72d7"] = (R.mesh_convertions.to_flat_edge(cells, "tri"), 7, p)

    hier = {}
    for name, (fe, d, pos) in cases.items():
        m_gs, m_ids = hierarchy(fe, d, pos.shape[0], pos)
        hier[name] = (m_gs, m_ids, pos, d)
        out = {"depth": np.int64(d), "pos": pos.astype(np.float32)}
        for l, g in enumerate(m_gs):
            out[f"g{l}"] = g.astype(np.int32)
        for l, i in enumerate(m_ids):
            out[f"ids{l}"] = i.astype(np.int32)
        np.savez_compressed(os.path.join(HERE, f"hier_{name}.npz"), **out)
        print(name, [g.shape[1] for g in m_gs], [len(i) for i in m_ids])

    def ref_bsgmp(depth, P, params):
        m = R.ops.BSGMP(depth, 128, 3, P)
        m.load_state_dict(params)
        return m

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a))

    # ---------------- per-op goldens on grid12 ----------------
    m_gs, m_ids, pos, d = hier["grid12"]
    g0, g1 = T(m_gs[0]), T(m_gs[1])
    gen = torch.Generator().manual_seed(1)
    x2 = torch.randn(144, 128, generator=gen)
    x3 = torch.randn(3, 144, 128, generator=gen)
    pos2 = T(pos)
    pos3 = pos2.unsqueeze(0) + 0.05 * torch.randn(3, 144, 2, generator=gen)
    params = O.init_params(2, pos_dim=2, seed=3)
    bs = ref_bsgmp(2, 2, params)
    conv = R.ops.WeightedEdgeConv()
    ops = {"x2": x2, "x3": x3, "pos3": pos3}
    with torch.no_grad():
        ops["gmp_x2_pos2"] = bs.down_gmps[0](x2, g0, pos2)
        ops["gmp_x3_pos2"] = bs.down_gmps[0](x3, g0, pos2)
        ops["gmp_x3_pos3"] = bs.down_gmps[0](x3, g0, pos3)
        ew0, aw0 = conv.cal_ew(torch.ones(144, 1), g0)
        ops["ew0"], ops["aggr_w0"] = ew0, aw0
        w1 = aw0[T(m_ids[0])]
        ew1, aw1 = conv.cal_ew(w1, g1)
        ops["ew1"], ops["aggr_w1"] = ew1, aw1
        ops["conv_down_x2"] = conv(x2, g0, ew0)
        ops["conv_down_x3"] = conv(x3, g0, ew0)
        ops["conv_up_x2"] = conv(x2, g0, ew0, aggragating=False)
        ops["conv_up_x3"] = conv(x3, g0, ew0, aggragating=False)
        ops["conv_down_pos3"] = conv(pos3.clone(), g0, ew0)
        hc = x3[:, T(m_ids[0])]
        ops["unpool_x3"] = R.ops.Unpool()(hc, 144, T(m_ids[0]))
    np.savez_compressed(os.path.join(HERE, "ops_grid12.npz"), **{k: v.numpy() for k, v in ops.items()})

    # ---------------- BSGMP goldens ----------------
    def bsgmp_case(name, hname, P, seed, batch, pos_batched, grads, row_stride=1):
        m_gs, m_ids, pos, d = hier[hname]
        n = pos.shape[0]
        gen = torch.Generator().manual_seed(seed)
        h = torch.randn(*( [batch, n, 128] if batch else [n, 128]), generator=gen)
        ps = T(pos)
        if pos_batched:
            ps = ps.unsqueeze(0) + 0.05 * torch.randn(batch, n, P, generator=gen)
        params = O.init_params(d, pos_dim=P, seed=seed)
        model = ref_bsgmp(d, P, params)
        gs = [T(g) for g in m_gs]
        ids = [T(i) for i in m_ids]
        h.requires_grad_(grads)
        out = model(h, ids, gs, ps)
        rec = {"seed": np.int64(seed), "batch": np.int64(batch), "pos_batched": np.int64(pos_batched),
               "P": np.int64(P), "row_stride": np.int64(row_stride),
               "out": out.detach()[..., ::row_stride, :].numpy(),
               "out_abs_sum": np.float64(out.detach().double().abs().sum().item()),
               "out_sq_sum": np.float64(out.detach().double().square().sum().item())}
        if grads:
            loss = out.square().mean()
            loss.backward()
            rec["loss"] = np.float64(loss.item())
            rec["grad_h"] = h.grad[..., ::row_stride, :].numpy()
            sd = dict(model.named_parameters())
            for k in ["bottom_gmp.mlp_edge.seq.0.weight", "bottom_gmp.mlp_edge.seq.0.bias",
                      "bottom_gmp.mlp_node.seq.0.weight", "down_gmps.0.mlp_edge.seq.0.weight",
                      "down_gmps.0.mlp_edge.seq.2.weight", "down_gmps.0.mlp_edge.seq.6.weight",
                      "down_gmps.0.mlp_edge.seq.6.bias", "down_gmps.0.mlp_node.seq.0.weight",
                      "down_gmps.0.mlp_node.seq.4.weight", "down_gmps.0.mlp_node.seq.6.bias",
                      f"up_gmps.{d - 1}.mlp_edge.seq.4.weight", f"up_gmps.{d - 1}.mlp_node.seq.2.bias"]:
                if k in sd:
                    rec["grad:" + k] = sd[k].grad.numpy()
            rec["grad_norms"] = np.array([float(v.grad.double().norm()) for _, v in sorted(sd.items())])
        np.savez_compressed(os.path.join(HERE, f"bsgmp_{name}.npz"), **rec)
        print("bsgmp", name, tuple(out.shape), rec["out_abs_sum"])

    bsgmp_case("chain11", "chain11", 3, 0, 0, False, True)
    bsgmp_case("grid12", "grid12", 2, 5, 0, False, True)
    bsgmp_case("grid12_b3", "grid12", 2, 6, 3, True, True)
    bsgmp_case("grid12_b2_sharedpos", "grid12", 2, 7, 2, False, True)
    bsgmp_case("ico3", "ico3", 3, 8, 0, False, True)
    bsgmp_case("twoclusters", "twoclusters", 2, 9, 0, False, False)
    bsgmp_case("grid44", "grid44", 2, 10, 0, False, True, row_stride=4)
    bsgmp_case("grid72", "grid72", 2, 11, 0, False, True, row_stride=8)
    bsgmp_case("grid72d7", "grid72d7", 2, 12, 0, False, False, row_stride=8)

    # chain11 known answers quoted in SURVEY.md §4
    m_gs, m_ids, pos, d = hier["chain11"]
    with torch.no_grad():
        ew, aw = conv.cal_ew(torch.ones(11, 1), T(m_gs[0]))
        px = conv(T(pos)[:, :1].clone(), T(m_gs[0]), ew)
    np.savez_compressed(os.path.join(HERE, "chain11_known.npz"), ew=ew.numpy(), aggr_w=aw.numpy(), conv_posx=px.numpy())


if __name__ == "__main__":
    main()
