"""The reference's on-disk cache of a multi-level mesh, `mmesh_layer_{depth}.dat`
(reference: src/datasets/base.py:98-122): a pickle of `{"m_gs": [LongTensor[2,E_l]] * (depth+1),
"m_ids": [LongTensor[n_{l+1}]] * depth}`.  `load_mmesh` / `save_mmesh` read and write exactly that file, so
a hierarchy cached by the reference feeds `bsms_gnn_b200.ops.BSGMP` directly and a hierarchy built here
(`hierarchy.build_hierarchy`, seconds instead of minutes on large meshes) is picked up by the reference's
datasets unchanged.  `mmesh_path` reproduces the reference's file naming.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch


def mmesh_path(data_dir: str, unet_depth: int, consist_mesh: bool = True, traj_file: str | None = None) -> str:
    """`<data_dir>/mmesh_layer_<depth>.dat`, prefixed by the trajectory file name when every trajectory has its own
    mesh (base.py:98-101)."""
    prefix = "" if consist_mesh else os.path.basename(traj_file) + "_"
    return os.path.join(data_dir, f"{prefix}mmesh_layer_{unet_depth}.dat")


def _check(m_gs, m_ids):
    if len(m_gs) != len(m_ids) + 1:
        raise ValueError(f"a depth-{len(m_ids)} hierarchy has {len(m_ids) + 1} graphs, got {len(m_gs)}")
    n_prev = None
    for l, g in enumerate(m_gs):
        if g.dim() != 2 or g.shape[0] != 2:
            raise ValueError(f"m_gs[{l}] must be [2, E], got {tuple(g.shape)}")
        if l < len(m_ids):
            ids = m_ids[l]
            if ids.dim() != 1:
                raise ValueError(f"m_ids[{l}] must be 1-D, got {tuple(ids.shape)}")
            if ids.numel() > 1 and not bool((ids[1:] > ids[:-1]).all()):
                raise ValueError(f"m_ids[{l}] must be strictly increasing (bsms_graph_wrapper.py:97-98)")
            if n_prev is not None and ids.numel() and int(ids.max()) >= n_prev:
                raise ValueError(f"m_ids[{l}] indexes past the {n_prev} nodes of level {l}")
            n_prev = int(ids.numel())


def save_mmesh(path: str, m_gs, m_ids) -> None:
    """Write the reference's cache file (LongTensors, pickle protocol default — what base.py:113-115 writes)."""
    gs = [torch.as_tensor(np.asarray(g), dtype=torch.long).reshape(2, -1) for g in m_gs]
    ids = [torch.as_tensor(np.asarray(i), dtype=torch.long).reshape(-1) for i in m_ids]
    _check(gs, ids)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        pickle.dump({"m_gs": gs, "m_ids": ids}, f)
    os.replace(tmp, path)


def load_mmesh(path: str, device=None):
    """-> (m_gs, m_ids) as int64 tensors (on `device` if given), validated; the argument order of
    `BSGMP.forward(h, m_ids, m_gs, pos)` is the reference's (src/ops/BSMS.py:39)."""
    with open(path, "rb") as f:
        m = pickle.load(f)
    gs = [torch.as_tensor(g, dtype=torch.long) for g in m["m_gs"]]
    ids = [torch.as_tensor(i, dtype=torch.long) for i in m["m_ids"]]
    _check(gs, ids)
    if device is not None:
        gs, ids = [g.to(device) for g in gs], [i.to(device) for i in ids]
    return gs, ids
