"""Node-partition planner (bsms_gnn_b200/partition.py) checked on CPU: R virtual ranks in one process
run the partitioned schedule with the ORACLE ops on their local graphs and an in-memory halo
exchange; the stitched result must equal the unpartitioned oracle ("multi-GPU without
a cluster").  Tolerance 2e-6: the same fp32 formulas, different summation order."""
import numpy as np
import pytest
import torch

from bsms_gnn_b200 import partition
from oracle import bsms_oracle as O
from tests.util import load_hier, max_rel


def emulate(plans, params, h, pos, d):
    R = len(plans)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    def exchange(level, owned):
        out = []
        for r in range(R):
            lp = plans[r].levels[level]
            parts = [owned[r]]
            for q in range(R):
                idx = plans[q].levels[level].send_idx[r]
                assert len(idx) == lp.recv_counts[q]
                parts.append(owned[q][T(idx)])
            loc = torch.cat(parts, 0)
            assert loc.shape[0] == lp.n_local
            out.append(loc)
        return out

    x_own = [h[T(p.levels[0].nodes[:p.levels[0].n_own])] for p in plans]
    p_own = [pos[T(p.levels[0].nodes[:p.levels[0].n_own])] for p in plans]
    skips, pos_loc = [], []
    for l in range(d):
        x_loc, p_loc = exchange(l, x_own), exchange(l, p_own)
        pos_loc.append(p_loc)
        y_own = []
        for r in range(R):
            lp = plans[r].levels[l]
            y_own.append(O.gmp(x_loc[r], T(lp.edges[:, lp.gmp_sel]), p_loc[r], params, f"down_gmps.{l}")[:lp.n_own])
        skips.append(y_own)
        y_loc = exchange(l, y_own)
        x_own, p_own = [], []
        for r in range(R):
            lp = plans[r].levels[l]
            e, ew, ids = T(lp.edges), T(lp.ew), T(lp.ids)
            x_own.append(O.edge_conv(y_loc[r], e, ew)[ids])
            p_own.append(O.edge_conv(p_loc[r], e, ew)[ids])
    x_loc, p_loc = exchange(d, x_own), exchange(d, p_own)
    x_own = [O.gmp(x_loc[r], T(plans[r].levels[d].edges[:, plans[r].levels[d].gmp_sel]), p_loc[r], params,
                   "bottom_gmp")[:plans[r].levels[d].n_own] for r in range(R)]
    for k in range(d):
        l = d - 1 - k
        hc_loc = exchange(l + 1, x_own)
        u_own = []
        for r in range(R):
            lp = plans[r].levels[l]
            inv = T(lp.inv)
            U = torch.zeros(lp.n_local, h.shape[-1])
            kept = inv >= 0
            U[kept] = hc_loc[r][inv[kept]]
            u_own.append(O.edge_conv(U, T(lp.edges), T(lp.ew), aggragating=False)[:lp.n_own])
        u_loc = exchange(l, u_own)
        x_own = [O.gmp(u_loc[r], T(plans[r].levels[l].edges[:, plans[r].levels[l].gmp_sel]), pos_loc[l][r], params,
                       f"up_gmps.{k}")[:plans[r].levels[l].n_own] + skips[l][r] for r in range(R)]
    out = torch.zeros_like(h)
    for r in range(R):
        lp = plans[r].levels[0]
        out[T(lp.nodes[:lp.n_own])] = x_own[r]
    return out


@pytest.mark.parametrize("hname,world", [("grid12", 2), ("grid12", 3), ("grid44", 4), ("ico3", 2), ("twoclusters", 2),
                                         ("grid72d7", 8)])
def test_partitioned_schedule_equals_global(hname, world):
    m_gs, m_ids, pos, d = load_hier(hname)
    n0 = pos.shape[0]
    gs = [g.numpy() for g in m_gs]
    ids = [i.numpy() for i in m_ids]
    plans = partition.build_all_plans(gs, ids, n0, world)
    # structural checks
    for l in range(d + 1):
        owned = np.concatenate([p.levels[l].nodes[:p.levels[l].n_own] for p in plans])
        n_l = n0 if l == 0 else len(ids[l - 1])
        assert np.array_equal(np.sort(owned), np.arange(n_l))  # every node owned exactly once
        for p in plans:
            lp = p.levels[l]
            assert lp.recv_counts.sum() == lp.n_local - lp.n_own
        # the receiver-owned edge subsets of all ranks partition the level's edges: no edge-MLP row is computed twice
        assert sum(int(p.levels[l].gmp_sel.sum()) for p in plans) == gs[l].shape[1]
    ew = partition.cal_ew_global(gs, ids, n0)
    w = torch.ones(n0, 1)
    for l in range(d):
        ew_o, aw = O.cal_ew(w, m_gs[l])
        assert max_rel(torch.from_numpy(ew[l]), ew_o) < 1e-6
        w = aw[m_ids[l]]
    P = pos.shape[1]
    params = O.init_params(d, pos_dim=P, seed=3)
    h = torch.randn(n0, 128, generator=torch.Generator().manual_seed(5))
    ref = O.bsgmp(h, m_ids, m_gs, pos, params, d)
    out = emulate(plans, params, h, pos, d)
    assert max_rel(out, ref) < 2e-6
