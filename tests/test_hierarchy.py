"""bsms_gnn_b200.hierarchy / meshgen against golden outputs of the reference builder
(src/graph_wrappers/bsms_graph_wrapper.py) — integer work, so the bar is exact equality."""
import numpy as np
import pytest

from bsms_gnn_b200 import hierarchy, meshgen
from tests.util import load_hier


def canon(e):
    e = np.asarray(e).reshape(2, -1)
    return e[:, np.lexsort((e[1], e[0]))]


def mesh(name):
    if name.startswith("grid"):
        nx = int(name[4:6])
        pos, cells = meshgen.tri_grid(nx, nx)
    elif name == "ico3":
        pos, cells = meshgen.icosphere(3)
    elif name == "twoclusters":
        p1, c1 = meshgen.tri_grid(9, 7)
        p2, c2 = meshgen.tri_grid(5, 6, seed=1)
        cells = np.concatenate([c1, c2 + p1.shape[0]])
        pos = np.concatenate([p1, p2 + 20])
    return pos, meshgen.cells_to_flat_edge(cells)


@pytest.mark.parametrize("name", ["grid12", "grid44", "grid72", "grid72d7", "ico3", "twoclusters"])
def test_hierarchy_matches_reference(name):
    m_gs, m_ids, pos_g, d = load_hier(name)
    pos, fe = mesh(name)
    assert np.array_equal(pos, pos_g.numpy())
    assert np.array_equal(fe, m_gs[0].numpy())  # level-0 edge ORDER is also the reference's
    gs, ids = hierarchy.build_hierarchy(fe, d, pos.shape[0], pos)
    assert len(gs) == d + 1 and len(ids) == d
    for a, b in zip(ids, m_ids):
        assert np.array_equal(a, b.numpy())
    for a, b in zip(gs, m_gs):
        assert np.array_equal(canon(a), canon(b.numpy()))


def test_chain_demo():
    # the reference's own demo (bsms_graph_wrapper.py:157-175)
    fe = np.array([[0, 1, 2, 3, 4, 5, 6, 7, 8, 9], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]])
    fe = np.concatenate((fe, fe[::-1]), axis=1)
    pos = np.stack([np.arange(11.0), np.zeros(11), np.zeros(11)], 1)
    gs, ids = hierarchy.build_hierarchy(fe, 2, 11, pos)
    assert [i.tolist() for i in ids] == [[1, 3, 5, 7, 9], [1, 3]]
    assert canon(gs[1]).tolist() == [[0, 1, 1, 2, 2, 3, 3, 4], [1, 0, 2, 1, 3, 2, 4, 3]]
    assert canon(gs[2]).tolist() == [[0, 1], [1, 0]]


def test_mesh_sizes():
    for nx, n, e in [(44, 1936, 11266), (72, 5184, 30530)]:
        pos, cells = meshgen.tri_grid(nx, nx)
        assert pos.shape == (n, 2) and meshgen.cells_to_flat_edge(cells).shape == (2, e)
    pos, cells = meshgen.icosphere(2)
    assert pos.shape == (162, 3) and meshgen.cells_to_flat_edge(cells).shape == (2, 960)


def _many_clusters(dtype, P):
    """40 small grids of different shapes scattered in space (+ isolated nodes): many clusters, ties, single-node clusters."""
    rng = np.random.default_rng(5)
    pos_l, cells_l, off = [], [], 0
    for k in range(40):
        nx, ny = int(rng.integers(2, 9)), int(rng.integers(2, 9))
        p, c = meshgen.tri_grid(nx, ny, seed=k)
        p = np.concatenate([p, np.zeros((p.shape[0], P - 2), dtype=p.dtype)], 1) if P > 2 else p
        pos_l.append(p.astype(dtype) + rng.uniform(-50, 50, size=(1, P)).astype(dtype))
        cells_l.append(c + off)
        off += p.shape[0]
    pos_l.append(rng.uniform(-50, 50, size=(3, P)).astype(dtype))  # three isolated nodes
    return np.concatenate(pos_l), meshgen.cells_to_flat_edge(np.concatenate(cells_l))


def test_native_builder_equals_numpy_builder(monkeypatch):
    """csrc/hierarchy_host.cpp against the numpy/scipy implementation: identical ids and identical edge ARRAYS (both
    emit row-major edges with sorted columns), on a mesh larger than the goldens and on disconnected / directed inputs.
    `native` = the whole hierarchy in one native call including the seed choice (float32 and float64 positions, 2-D and
    3-D); `levels` = native integer work level by level with the numpy seed choice."""
    cases = []
    pos, cells = meshgen.tri_grid(120, 90)
    cases.append((pos, meshgen.cells_to_flat_edge(cells), 5))
    cases.append((pos.astype(np.float64) * 1.37, meshgen.cells_to_flat_edge(cells), 5))
    pos, cells = meshgen.icosphere(4)
    cases.append((pos, meshgen.cells_to_flat_edge(cells), 4))
    p1, c1 = meshgen.tri_grid(9, 7)
    p2, c2 = meshgen.tri_grid(5, 6, seed=1)
    cases.append((np.concatenate([p1, p2 + 20, np.array([[99.0, 99.0]], dtype=np.float32)]),
                  meshgen.cells_to_flat_edge(np.concatenate([c1, c2 + p1.shape[0]])), 3))  # + one isolated node
    for dtype, P in [(np.float32, 2), (np.float64, 3), (np.float32, 3)]:
        pos, fe = _many_clusters(dtype, P)
        cases.append((pos, fe, 3))
    # an edge list in random order with duplicates (the level-0 conversion's general path)
    pos, cells = meshgen.tri_grid(40, 31)
    fe = meshgen.cells_to_flat_edge(cells)
    rng = np.random.default_rng(1)
    fe = np.concatenate([fe, fe[:, :100]], 1)[:, rng.permutation(fe.shape[1] + 100)]
    cases.append((pos, fe, 4))
    for pos, fe, d in cases:
        monkeypatch.setenv("BSMS_HIERARCHY", "numpy")
        gs_n, ids_n = hierarchy.build_hierarchy(fe, d, pos.shape[0], pos)
        for impl in ("native", "levels"):
            monkeypatch.setenv("BSMS_HIERARCHY", impl)
            gs_c, ids_c = hierarchy.build_hierarchy(fe, d, pos.shape[0], pos)
            assert len(gs_c) == d + 1 and len(ids_c) == d
            for a, b in zip(ids_c, ids_n):
                assert a.dtype == np.int64 and np.array_equal(a, b), impl
            for a, b in zip(gs_c[1:], gs_n[1:]):
                assert a.dtype == np.int64 and np.array_equal(a, b), impl
    # a one-directional chain: nodes the seed cannot reach are in neither parity class
    fe = np.array([[0, 1, 2, 3], [1, 2, 3, 4]])
    pos = np.stack([np.arange(5.0), np.zeros(5)], 1).astype(np.float32)
    k_n, e_n = hierarchy.bistride_level_numpy(fe, pos, 5)
    k_c, e_c = hierarchy.bistride_level_native(fe, pos, 5)
    assert np.array_equal(k_n, k_c) and np.array_equal(e_n, e_c)
    monkeypatch.setenv("BSMS_HIERARCHY", "numpy")
    gs_n, ids_n = hierarchy.build_hierarchy(fe, 3, 5, pos)
    gs_c, ids_c = hierarchy.build_hierarchy_native(fe, 3, 5, pos)
    for a, b in zip(ids_c + gs_c[1:], ids_n + gs_n[1:]):
        assert np.array_equal(a, b)
    # errors: an out-of-range index is reported, not read
    with pytest.raises(IndexError):
        hierarchy.build_hierarchy_native(np.array([[0, 1], [1, 7]]), 1, 5, pos)


def test_native_views_outlive_the_call():
    """The arrays handed out are zero-copy views of native buffers; they must stay valid after everything else is gone."""
    import gc
    pos, cells = meshgen.tri_grid(30, 30)
    gs, ids = hierarchy.build_hierarchy_native(meshgen.cells_to_flat_edge(cells), 3, pos.shape[0], pos)
    want = [g.copy() for g in gs], [i.copy() for i in ids]
    last_g, last_i = gs[-1], ids[0]
    del gs, ids
    gc.collect()
    junk = [np.zeros(1 << 20, dtype=np.int64) for _ in range(8)]  # churn the allocator
    assert np.array_equal(last_g, want[0][-1]) and np.array_equal(last_i, want[1][0])
    del junk


def test_reference_graph_selfcheck_clusters():
    """The reference's own `Graph` self-check (src/graph_wrappers/graph_wrapper.py:216-241): two directed 3-cycles on six
    nodes form the clusters [[0, 1, 2], [3, 4, 5]] — through the library's host code; then one bi-stride level of all
    three builders on it: from the seed a directed 3-cycle has one node at depth 1 and one at depth 2, so the parity
    classes are {seed, depth 2} and {depth 1}, and the smaller (odd) one is kept."""
    import ctypes as C

    from bsms_gnn_b200._lib import check, lib
    fe = np.array([[0, 1, 2, 3, 4, 5], [1, 2, 0, 4, 5, 3]], dtype=np.int64)
    labels = np.empty(6, dtype=np.int64)
    nc = C.c_int64()
    check(lib.bsms_components_host(fe.ctypes.data, 6, 6, labels.ctypes.data, C.byref(nc)))
    assert nc.value == 2 and labels.tolist() == [0, 0, 0, 1, 1, 1]
    pos = np.array([[0.0, 0.0], [1.0, 0.0], [0.4, 0.1], [5.0, 5.0], [6.0, 5.0], [5.4, 5.1]], dtype=np.float32)
    k_n, e_n = hierarchy.bistride_level_numpy(fe, pos, 6)
    k_c, e_c = hierarchy.bistride_level_native(fe, pos, 6)
    gs, ids = hierarchy.build_hierarchy_native(fe, 1, 6, pos)
    assert np.array_equal(k_n, k_c) and np.array_equal(e_n, e_c)
    assert np.array_equal(ids[0], k_n) and np.array_equal(gs[1], e_n)
    # seeds are the nodes nearest the cluster centroids (2 and 5); depth-1 nodes (their successors 0 and 3) are the
    # smaller parity class of each cluster
    assert k_n.tolist() == [0, 3]
