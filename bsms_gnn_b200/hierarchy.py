"""Bi-stride multi-level graph builder (once per mesh, off the hot path).

Two implementations of the same function, both exact against the reference's goldens (tests/test_hierarchy.py):
`native` (default) — the integer work (clusters, BFS parity, kept-row (A+I)^2, re-indexing) runs in the library's
host code (csrc/hierarchy_host.cpp, OpenMP; `bsms_components_host`, `bsms_bistride_level_host`), only the
floating-point seed choice stays in numpy; `numpy` — the vectorised numpy/scipy version below, kept as the
cross-check.  `BSMS_HIERARCHY=numpy` selects the latter.

Produces the same `(m_gs, m_ids)` the reference builds in pure Python with
`BistrideMultiLayerGraph(flat_edge, num_layers, num_nodes, pos).get_multi_layer_graphs()`
(reference: src/graph_wrappers/bsms_graph_wrapper.py:30-154, src/graph_wrappers/graph_wrapper.py:67-134):

  per level: connected clusters -> per cluster the seed is the node nearest the cluster centroid
  (:107-126) -> BFS distance parity from the seed -> keep the smaller of the even/odd sets, even on
  ties or when there is no odd node (:80-95) -> new adjacency = pattern of (A+I)^2 without the
  diagonal, restricted to kept nodes and re-indexed (:99-102, :129-154).

Differences from the reference, none of which change the graph: (i) edges of levels >= 1 are emitted
row-major with SORTED columns (the reference inherits whatever column order the MKL/scipy SpGEMM
leaves inside a row); (ii) clusters are weakly-connected components, which equals the reference's
"reachable from the first unvisited node" for the symmetric graphs every mesh produces;
(iii) only the kept rows/columns of (A+I)^2 are formed.  tests/test_hierarchy.py checks node ids
exactly and edge sets exactly against golden outputs of the reference builder.

The reference needs 96 s / 5.5 GB for a 2 M-node mesh; this one is what
bench.py uses to build the large synthetic hierarchies on the GPU box, where /root/reference does
not exist.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components


def _csr_pattern(flat_edge: np.ndarray, n: int) -> sp.csr_matrix:
    a = sp.csr_matrix((np.ones(flat_edge.shape[1], dtype=np.float32),
                       (flat_edge[0], flat_edge[1])), shape=(n, n))
    a.sum_duplicates()
    a.data[:] = 1.0
    return a


def _bfs_levels(a: sp.csr_matrix, seeds: np.ndarray) -> np.ndarray:
    """Multi-source BFS depth over directed edges row->col; -1 = unreachable."""
    n = a.shape[0]
    indptr, indices = a.indptr, a.indices
    dist = np.full(n, -1, dtype=np.int64)
    dist[seeds] = 0
    frontier = np.asarray(seeds, dtype=np.int64)
    depth = 0
    while frontier.size:
        depth += 1
        starts = indptr[frontier]
        counts = indptr[frontier + 1] - starts
        total = int(counts.sum())
        if total == 0:
            break
        # concatenated neighbour lists of the frontier without a Python loop
        offs = np.repeat(starts - np.concatenate([[0], np.cumsum(counts)[:-1]]), counts)
        nbr = indices[np.arange(total, dtype=np.int64) + offs]
        nbr = nbr[dist[nbr] < 0]
        if nbr.size == 0:
            break
        nbr = np.unique(nbr)
        dist[nbr] = depth
        frontier = nbr
    return dist


def _cluster_seeds(labels: np.ndarray, ncomp: int, pos: np.ndarray) -> np.ndarray:
    """Per cluster the node nearest the cluster centroid (bsms_graph_wrapper.py:107-126), numpy arithmetic."""
    # clusters ordered by their smallest node id, members ascending (graph_wrapper.py:107-134)
    order = np.argsort(labels, kind="stable")
    bounds = np.concatenate([[0], np.cumsum(np.bincount(labels, minlength=ncomp))])
    seeds = np.empty(ncomp, dtype=np.int64)
    for c in range(ncomp):
        members = order[bounds[c]:bounds[c + 1]]
        if members.size == 1:
            seeds[c] = members[0]
            continue
        pc = pos[members]
        center = np.mean(pc, axis=0)
        d = np.linalg.norm(pc - center[None, :], 2, axis=-1)
        seeds[c] = members[np.argmin(d)]
    return seeds


def bistride_level_native(flat_edge: np.ndarray, pos: np.ndarray, n: int):
    """One pooling level through the library's host code; same result as `bistride_level_numpy`."""
    import ctypes as C

    from ._lib import check, lib
    fe = np.ascontiguousarray(np.asarray(flat_edge, dtype=np.int64).reshape(2, -1))
    E = int(fe.shape[1])
    labels = np.empty(n, dtype=np.int64)
    ncomp = C.c_int64()
    check(lib.bsms_components_host(fe.ctypes.data, E, n, labels.ctypes.data, C.byref(ncomp)))
    seeds = np.ascontiguousarray(_cluster_seeds(labels, int(ncomp.value), pos))
    keep = np.empty(n, dtype=np.int64)
    nk, ne, eptr = C.c_int64(), C.c_int64(), C.c_void_p()
    check(lib.bsms_bistride_level_host(fe.ctypes.data, E, n, labels.ctypes.data, int(ncomp.value), seeds.ctypes.data,
                                       keep.ctypes.data, C.byref(nk), C.byref(eptr), C.byref(ne)))
    try:
        buf = (C.c_int64 * max(2 * ne.value, 1)).from_address(eptr.value)
        new_e = np.frombuffer(buf, dtype=np.int64, count=2 * ne.value).reshape(2, ne.value).copy()
    finally:
        lib.bsms_host_free(eptr)
    return keep[:nk.value].copy(), new_e


def bistride_level(flat_edge: np.ndarray, pos: np.ndarray, n: int):
    import os
    if os.environ.get("BSMS_HIERARCHY", "native") == "numpy":
        return bistride_level_numpy(flat_edge, pos, n)
    return bistride_level_native(flat_edge, pos, n)  # "levels" (and the per-level entry point of "native")


def bistride_level_numpy(flat_edge: np.ndarray, pos: np.ndarray, n: int):
    """One pooling level: returns (kept node ids sorted [n'], new flat edges [2,E'] int64)."""
    flat_edge = np.asarray(flat_edge, dtype=np.int64).reshape(2, -1)
    a = _csr_pattern(flat_edge, n)
    ncomp, labels = connected_components(a, directed=True, connection="weak")
    seeds = _cluster_seeds(labels, ncomp, pos)
    dist = _bfs_levels(a, seeds)
    reach = dist >= 0
    even = reach & (dist % 2 == 0)
    odd = reach & (dist % 2 == 1)
    n_even = np.bincount(labels[even], minlength=ncomp)
    n_odd = np.bincount(labels[odd], minlength=ncomp)
    keep_even = (n_even <= n_odd) | (n_odd == 0)
    kept_mask = np.where(keep_even[labels], even, odd)
    keep = np.nonzero(kept_mask)[0].astype(np.int64)
    a1 = (a + sp.identity(n, dtype=np.float32, format="csr")).tocsr()
    a2 = (a1[keep, :] @ a1[:, keep]).tocsr()
    a2.setdiag(0)
    a2.eliminate_zeros()
    a2.sort_indices()
    coo = a2.tocoo()
    new_e = np.stack([coo.row.astype(np.int64), coo.col.astype(np.int64)])
    return keep, new_e


class _HierarchyHandle:
    """Owns one native hierarchy (bsms_hierarchy_build_host); the numpy views handed out keep it alive."""

    def __init__(self, ptr, free):
        self.ptr, self._free = ptr, free

    def __del__(self):
        if self.ptr:
            self._free(self.ptr)
            self.ptr = None


def build_hierarchy_native(flat_edge: np.ndarray, num_layers: int, num_nodes: int, pos: np.ndarray):
    """All levels in ONE call of the library's host code (csrc/hierarchy_host.cpp bsms_hierarchy_build_host): int32 CSR
    between levels, parallel union-find, the seed choice in the positions' own floating-point type with numpy's
    operation order, one-pass kept-row (A+I)^2.  The returned arrays are zero-copy views of the native buffers."""
    import ctypes as C

    from ._lib import check, lib
    g = np.ascontiguousarray(np.asarray(flat_edge, dtype=np.int64).reshape(2, -1))
    p = np.asarray(pos)
    if p.dtype not in (np.float32, np.float64):
        p = p.astype(np.float64)
    p = np.ascontiguousarray(p.reshape(int(num_nodes), -1))
    h = C.c_void_p()
    check(lib.bsms_hierarchy_build_host(g.ctypes.data, int(g.shape[1]), int(num_nodes), p.ctypes.data, int(p.shape[1]),
                                        1 if p.dtype == np.float64 else 0, int(num_layers), C.byref(h)))
    owner = _HierarchyHandle(h.value, lib.bsms_hierarchy_free_host)
    m_gs, m_ids = [g], []
    for lvl in range(1, int(num_layers) + 1):
        nn, ne, ep, ip = C.c_int64(), C.c_int64(), C.c_void_p(), C.c_void_p()
        check(lib.bsms_hierarchy_level_host(owner.ptr, lvl, C.byref(nn), C.byref(ne), C.byref(ep), C.byref(ip)))

        def view(addr, count, shape):
            if count == 0:
                return np.zeros(shape, dtype=np.int64)
            buf = (C.c_int64 * count).from_address(addr)
            buf._bsms_owner = owner  # the numpy view references buf, buf references the handle
            return np.frombuffer(buf, dtype=np.int64, count=count).reshape(shape)
        m_gs.append(view(ep.value, 2 * ne.value, (2, ne.value)))
        m_ids.append(view(ip.value, nn.value, (nn.value,)))
    return m_gs, m_ids


def build_hierarchy(flat_edge: np.ndarray, num_layers: int, num_nodes: int, pos: np.ndarray):
    """-> (m_gs: list[num_layers+1] of int64 [2,E_l], m_ids: list[num_layers] of int64 [n_{l+1}]).
    BSMS_HIERARCHY = native (default: one native call) | levels (native, level by level with numpy seeds) | numpy."""
    import os
    if os.environ.get("BSMS_HIERARCHY", "native") == "native":
        return build_hierarchy_native(flat_edge, num_layers, num_nodes, pos)
    g = np.asarray(flat_edge, dtype=np.int64).reshape(2, -1)
    pos_l = np.asarray(pos)
    n = int(num_nodes)
    m_gs, m_ids = [g], []
    for _ in range(num_layers):
        keep, g = bistride_level(g, pos_l, n)
        pos_l = pos_l[keep]
        n = int(keep.shape[0])
        m_gs.append(g)
        m_ids.append(keep)
    return m_gs, m_ids
