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


def test_native_builder_equals_numpy_builder(monkeypatch):
    """csrc/hierarchy_host.cpp against the numpy/scipy implementation: identical ids and identical edge ARRAYS (both
    emit row-major edges with sorted columns), on a mesh larger than the goldens and on disconnected / directed inputs."""
    cases = []
    pos, cells = meshgen.tri_grid(120, 90)
    cases.append((pos, meshgen.cells_to_flat_edge(cells), 5))
    pos, cells = meshgen.icosphere(4)
    cases.append((pos, meshgen.cells_to_flat_edge(cells), 4))
    p1, c1 = meshgen.tri_grid(9, 7)
    p2, c2 = meshgen.tri_grid(5, 6, seed=1)
    cases.append((np.concatenate([p1, p2 + 20, np.array([[99.0, 99.0]], dtype=np.float32)]),
                  meshgen.cells_to_flat_edge(np.concatenate([c1, c2 + p1.shape[0]])), 3))  # + one isolated node
    for pos, fe, d in cases:
        monkeypatch.setenv("BSMS_HIERARCHY", "numpy")
        gs_n, ids_n = hierarchy.build_hierarchy(fe, d, pos.shape[0], pos)
        monkeypatch.setenv("BSMS_HIERARCHY", "native")
        gs_c, ids_c = hierarchy.build_hierarchy(fe, d, pos.shape[0], pos)
        for a, b in zip(ids_c, ids_n):
            assert np.array_equal(a, b)
        for a, b in zip(gs_c[1:], gs_n[1:]):
            assert np.array_equal(a, b)
    # a one-directional chain: nodes the seed cannot reach are in neither parity class
    fe = np.array([[0, 1, 2, 3], [1, 2, 3, 4]])
    pos = np.stack([np.arange(5.0), np.zeros(5)], 1).astype(np.float32)
    k_n, e_n = hierarchy.bistride_level_numpy(fe, pos, 5)
    k_c, e_c = hierarchy.bistride_level_native(fe, pos, 5)
    assert np.array_equal(k_n, k_c) and np.array_equal(e_n, e_c)
