"""Synthetic mesh generators in the reference's data layout.

The reference reads triangle cells from .h5 files and turns them into a two-way edge list with
`to_flat_edge(cells, "tri")` (reference: src/utils/mesh_convertions.py:4-21, :88-100).  There are no
datasets on the build or GPU boxes, so benchmarks and tests use the meshes below:
a jittered structured triangle grid (CylinderFlow/Airfoil stand-ins, pos_dim=2) and an icosphere
(InflatingSphere stand-in, pos_dim=3).  Everything is numpy, deterministic, and runs once per mesh.
"""
from __future__ import annotations

import numpy as np


def cells_to_flat_edge(cells: np.ndarray) -> np.ndarray:
    """Unique undirected edges of simplicial cells -> two-way [2, E] int64 edge list.

    Same result (including edge ORDER) as the reference `triangles_to_edges`
    (src/utils/mesh_convertions.py:4-21): pack each edge as (max, min), lexicographically unique
    them, then emit [senders ‖ receivers ; receivers ‖ senders].
    """
    cells = np.asarray(cells, dtype=np.int64)
    k = cells.shape[1]
    if k == 3:
        pairs = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [2, 0]]], 0)
    elif k == 4:  # tetrahedra (src/utils/mesh_convertions.py:24-50)
        pairs = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [2, 3]],
                                cells[:, [3, 0]], cells[:, [0, 2]], cells[:, [1, 3]]], 0)
    else:
        raise ValueError(f"unsupported cell width {k}")
    s = pairs.max(1)
    r = pairs.min(1)
    base = int(cells.max()) + 1
    key = np.unique(s * base + r)  # lexicographic in (s, r), like the reference's row-wise unique
    s, r = key // base, key % base
    return np.stack([np.concatenate([s, r]), np.concatenate([r, s])]).astype(np.int64)


def tri_grid(nx: int, ny: int, jitter: float = 0.1, seed: int = 0):
    """nx×ny structured grid, every quad split into two triangles, positions jittered.

    Returns (pos [N,2] float32, cells [2(nx-1)(ny-1),3] int64).  44×44 gives 1 936 nodes / 11 266
    directed edges, 72×72 gives 5 184 / 30 530, 1414×1414 gives 1 999 396 / 11 985 066.
    """
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    pos = np.stack([ix.ravel(), iy.ravel()], 1).astype(np.float64)
    rng = np.random.default_rng(seed)
    pos = pos + jitter * rng.standard_normal(pos.shape)
    v00 = (iy[:-1, :-1] * nx + ix[:-1, :-1]).ravel()
    v10 = v00 + 1
    v01 = v00 + nx
    v11 = v01 + 1
    cells = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], 0)
    return pos.astype(np.float32), cells.astype(np.int64)


def icosphere(subdiv: int):
    """Unit icosphere: (pos [N,3] float32, cells [F,3] int64); subdiv=5 -> 10 242 nodes."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]],
                 dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdiv):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        e.sort(1)
        ue, inv = np.unique(e, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        mid = v[ue[:, 0]] + v[ue[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        nf = f.shape[0]
        m01 = inv[:nf] + len(v)
        m12 = inv[nf:2 * nf] + len(v)
        m20 = inv[2 * nf:] + len(v)
        v = np.concatenate([v, mid], 0)
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], 0)
    return v.astype(np.float32), f
