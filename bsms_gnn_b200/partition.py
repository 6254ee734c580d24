"""Node partition of a BSMS hierarchy across ranks, with per-level halo (ghost) plans.

The reference has no partitioning; this is the multi-GPU decomposition of the
processor for meshes that do not fit — or are too slow on — one GPU (BASELINE.json config 5).

Ownership: level-0 nodes are split into `world` contiguous blocks of the caller's node order (a
locality-preserving order: row-major for the synthetic grids), and a coarse node inherits the owner
of the fine node it was pooled from (`m_ids[l]` indexes level l, src/graph_wrappers/
bsms_graph_wrapper.py:39-44), so pooling and unpooling are rank-local.

Per level every rank holds LOCAL rows `[owned (ascending global id) | ghosts (sorted by owner, then
global id)]`.  Ghosts are (a) every endpoint of an edge that touches an owned node — the GMP
aggregation and the restriction read in-neighbours, the prolongation reads out-neighbours — and
(b) the coarse images of the kept local nodes of the finer level (the prolongation reads the coarse
value of a kept ghost).  Local edges are the edges with an owned endpoint, re-indexed locally, with
the GLOBAL transfer weights (cal_ew depends on whole neighbourhoods, src/ops/basic.py:142-167).

All exchanges then have one shape: "refresh the ghost rows of a level-l tensor from their owners"
(`requests[l][q]` = what this rank needs from rank q, `send_idx[l][q]` = which owned rows it sends
to q).  Pure numpy; runs once per mesh.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def cal_ew_global(m_gs, m_ids, n0):
    """Transfer weights of every pooling level on the GLOBAL graphs, fp32 like the reference
    (src/ops/basic.py:142-167, src/ops/BSMS.py:64-89).  -> list of ew [E_l] float32."""
    w = np.ones(n0, dtype=np.float32)
    out = []
    n = n0
    for l, ids in enumerate(m_ids):
        g = m_gs[l]
        deg = np.bincount(g[0], minlength=n).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            normed = (w / deg).astype(np.float32)
        s = normed[g[0]]
        aggr = np.zeros(n, dtype=np.float32)
        np.add.at(aggr, g[1], s)
        aggr = aggr + np.float32(1e-12)
        out.append((s / aggr[g[1]]).astype(np.float32))
        w = aggr[ids]
        n = len(ids)
    return out


def node_work(m_gs, m_ids, n0):
    """Edge-MLP rows a level-0 node is responsible for over one forward: the in-degree of the node and of every
    coarse image of it, counted twice on the levels that have a down AND an up GMP (src/ops/BSMS.py:67-102)."""
    d = len(m_ids)
    sizes = [n0] + [len(i) for i in m_ids]
    w = [np.bincount(np.asarray(m_gs[l]).reshape(2, -1)[1], minlength=sizes[l]).astype(np.float64) * (2.0 if l < d else 1.0)
         for l in range(d + 1)]
    for l in range(d, 0, -1):  # a coarse node is the fine node m_ids[l-1][c]: push its work down to level 0
        np.add.at(w[l - 1], np.asarray(m_ids[l - 1]), w[l])
    return w[0]


def level_owners(m_ids, n0, world, m_gs=None):
    """owner[l][node] for every level: contiguous blocks of the level-0 order, inherited below.  With the level graphs
    the block boundaries equalise the edge-MLP rows per rank (the ranks at the ends of a band partition have one
    cut instead of two and would otherwise finish ~5 % early); without them the blocks have equal node counts."""
    if m_gs is not None and world > 1:
        cw = np.cumsum(node_work(m_gs, m_ids, n0))
        cuts = np.searchsorted(cw, cw[-1] * np.arange(1, world) / world, side="left") + 1
        bounds = np.concatenate([[0], np.minimum(cuts, n0), [n0]])
        bounds = np.maximum.accumulate(bounds)
        sizes = np.diff(bounds).astype(np.int64)
    else:
        base, rem = divmod(n0, world)
        sizes = np.array([base + (1 if r < rem else 0) for r in range(world)], dtype=np.int64)
    own0 = np.repeat(np.arange(world, dtype=np.int32), sizes)
    owners = [own0]
    for ids in m_ids:
        owners.append(owners[-1][ids])
    return owners


@dataclass
class LevelPart:
    n_own: int
    n_local: int
    nodes: np.ndarray          # [n_local] global ids: owned ascending, then ghosts by (owner, id)
    edges: np.ndarray          # [2, E_loc] local indices
    edge_ids: np.ndarray       # [E_loc] global edge ids (order preserved)
    gmp_sel: np.ndarray        # [E_loc] bool: the edge's RECEIVER is owned.  The GMP aggregates onto owned nodes only, so
                               # it runs on this subset (exactly E_l / world edges over all ranks: no redundant edge-MLP
                               # rows); the transfers also need the sender-owned edges (prolongation reads out-neighbours)
    ew: np.ndarray | None      # [E_loc] global transfer weights (None at the deepest level)
    ids: np.ndarray | None     # local fine indices of the OWNED kept nodes (-> coarse local 0..n_own'-1)
    inv: np.ndarray | None     # [n_local] coarse local index of a kept local node, -1 otherwise
    recv_counts: np.ndarray    # [world] ghosts owned by each rank (ghost block is ordered by owner)
    requests: list             # [world] global ids requested from each owner (= ghost ids, by owner)
    send_idx: list = field(default_factory=list)  # [world] owned local rows this rank sends to each peer


@dataclass
class RankPlan:
    rank: int
    world: int
    levels: list  # LevelPart per level 0..d

    def finalize(self, incoming):
        """incoming[l][q] = global ids rank q requested from this rank at level l."""
        for l, lp in enumerate(self.levels):
            g2l = {int(g): k for k, g in enumerate(lp.nodes[:lp.n_own])}
            lp.send_idx = [np.array([g2l[int(g)] for g in incoming[l][q]], dtype=np.int64)
                           if len(incoming[l][q]) else np.zeros(0, dtype=np.int64) for q in range(self.world)]


def build_rank_plan(m_gs, m_ids, n0, world, rank, ew_global=None, owners=None) -> RankPlan:
    d = len(m_ids)
    owners = owners if owners is not None else level_owners(m_ids, n0, world, m_gs)
    ew_global = ew_global if ew_global is not None else cal_ew_global(m_gs, m_ids, n0)
    levels = []
    prev_local_kept_coarse = None  # coarse (level l) ids that the finer level needs locally
    for l in range(d + 1):
        own_l = owners[l]
        n_l = own_l.shape[0]
        g = np.asarray(m_gs[l]).reshape(2, -1)
        src, dst = g[0], g[1]
        mine = own_l == rank
        e_ids = np.nonzero(mine[dst] | mine[src])[0]
        touched = np.unique(np.concatenate([src[e_ids], dst[e_ids]])) if e_ids.size else np.zeros(0, dtype=np.int64)
        ghosts = touched[~mine[touched]]
        if prev_local_kept_coarse is not None and prev_local_kept_coarse.size:
            extra = prev_local_kept_coarse[~mine[prev_local_kept_coarse]]
            ghosts = np.union1d(ghosts, extra)
        order = np.lexsort((ghosts, own_l[ghosts]))
        ghosts = ghosts[order].astype(np.int64)
        owned = np.nonzero(mine)[0].astype(np.int64)
        nodes = np.concatenate([owned, ghosts])
        g2l = np.full(n_l, -1, dtype=np.int64)
        g2l[nodes] = np.arange(nodes.shape[0])
        edges = np.stack([g2l[src[e_ids]], g2l[dst[e_ids]]]) if e_ids.size else np.zeros((2, 0), dtype=np.int64)
        recv_counts = np.bincount(own_l[ghosts], minlength=world).astype(np.int64)
        bounds = np.concatenate([[0], np.cumsum(recv_counts)])
        requests = [ghosts[bounds[q]:bounds[q + 1]] for q in range(world)]
        ids = inv = ew = None
        if l < d:
            keep = np.asarray(m_ids[l]).astype(np.int64)          # global fine ids of coarse nodes 0..n'-1
            inv_g = np.full(n_l, -1, dtype=np.int64)
            inv_g[keep] = np.arange(keep.shape[0])
            local_coarse = inv_g[nodes]                             # coarse global id of each local node / -1
            prev_local_kept_coarse = np.unique(local_coarse[local_coarse >= 0])
            ids = np.nonzero(local_coarse[:owned.shape[0]] >= 0)[0].astype(np.int64)
            ew = ew_global[l][e_ids]
            levels.append(LevelPart(owned.shape[0], nodes.shape[0], nodes, edges, e_ids, mine[dst[e_ids]], ew, ids, local_coarse,
                                    recv_counts, requests))
        else:
            prev_local_kept_coarse = None
            levels.append(LevelPart(owned.shape[0], nodes.shape[0], nodes, edges, e_ids, mine[dst[e_ids]], None, None, None,
                                    recv_counts, requests))
    # inv: translate "coarse GLOBAL id" to "coarse LOCAL index" now that the coarser level's nodes are known
    for l in range(d):
        nxt = levels[l + 1]
        n_next = owners[l + 1].shape[0]
        g2l_next = np.full(n_next, -1, dtype=np.int64)
        g2l_next[nxt.nodes] = np.arange(nxt.n_local)
        cg = levels[l].inv
        levels[l].inv = np.where(cg >= 0, g2l_next[np.maximum(cg, 0)], -1).astype(np.int64)
        # the k-th owned kept fine node must be coarse local row k
        assert np.array_equal(levels[l].inv[levels[l].ids], np.arange(nxt.n_own)), "ownership inheritance broken"
    return RankPlan(rank, world, levels)


def build_all_plans(m_gs, m_ids, n0, world):
    """Single-process construction of every rank's plan (tests, small meshes)."""
    owners = level_owners(m_ids, n0, world, m_gs)
    ew = cal_ew_global(m_gs, m_ids, n0)
    plans = [build_rank_plan(m_gs, m_ids, n0, world, r, ew, owners) for r in range(world)]
    d = len(m_ids)
    for r in range(world):
        incoming = [[plans[q].levels[l].requests[r] for q in range(world)] for l in range(d + 1)]
        plans[r].finalize(incoming)
    return plans
