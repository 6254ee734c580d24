"""The multi-level mesh cache file of the reference (src/datasets/base.py:98-122): written here, read the
way the reference reads it, and back."""
import pickle

import numpy as np
import pytest
import torch

from bsms_gnn_b200 import hierarchy, meshgen, mmesh_io


def test_roundtrip_and_reference_layout(tmp_path):
    pos, cells = meshgen.tri_grid(12, 12)
    fe = meshgen.cells_to_flat_edge(cells)
    m_gs, m_ids = hierarchy.build_hierarchy(fe, 3, pos.shape[0], pos)
    path = mmesh_io.mmesh_path(str(tmp_path), 3)
    assert path.endswith("mmesh_layer_3.dat")
    mmesh_io.save_mmesh(path, m_gs, m_ids)
    # exactly what the reference does with the file (base.py:117-120)
    with open(path, "rb") as f:
        m = pickle.load(f)
    assert set(m) == {"m_gs", "m_ids"}
    assert all(isinstance(g, torch.Tensor) and g.dtype == torch.long and g.shape[0] == 2 for g in m["m_gs"])
    assert all(isinstance(i, torch.Tensor) and i.dtype == torch.long and i.dim() == 1 for i in m["m_ids"])
    gs, ids = mmesh_io.load_mmesh(path)
    for a, b in zip(gs, m_gs):
        assert np.array_equal(a.numpy(), b)
    for a, b in zip(ids, m_ids):
        assert np.array_equal(a.numpy(), b)


def test_reads_a_file_written_the_reference_way(tmp_path):
    # base.py:107-115: python lists of numpy arrays -> torch.tensor(..., dtype=torch.long) -> pickle
    m_gs = [np.array([[0, 1, 1, 2], [1, 0, 2, 1]]), np.array([[0, 1], [1, 0]])]
    m_ids = [np.array([0, 2])]
    path = mmesh_io.mmesh_path(str(tmp_path), 1, consist_mesh=False, traj_file="/data/traj_7.h5")
    assert path.endswith("traj_7.h5_mmesh_layer_1.dat")
    with open(path, "wb") as f:
        pickle.dump({"m_gs": [torch.tensor(g, dtype=torch.long) for g in m_gs],
                     "m_ids": [torch.tensor(i, dtype=torch.long) for i in m_ids]}, f)
    gs, ids = mmesh_io.load_mmesh(path)
    assert [tuple(g.shape) for g in gs] == [(2, 4), (2, 2)] and ids[0].tolist() == [0, 2]


def test_rejects_malformed_hierarchies(tmp_path):
    g = np.array([[0, 1], [1, 0]])
    with pytest.raises(ValueError):
        mmesh_io.save_mmesh(str(tmp_path / "a.dat"), [g], [np.array([0])])          # graph count
    with pytest.raises(ValueError):
        mmesh_io.save_mmesh(str(tmp_path / "b.dat"), [g, g], [np.array([1, 0])])    # ids not increasing
    with pytest.raises(ValueError):
        mmesh_io.save_mmesh(str(tmp_path / "c.dat"), [g, g, g], [np.array([0, 1]), np.array([5])])  # out of range
