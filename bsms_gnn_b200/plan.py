"""Graph plans: device-resident int32 index structures built once per mesh and cached.

The reference re-derives everything from the raw int64 `[2,E]` edge lists on every call
(src/ops/basic.py:66,130-137) and recomputes `cal_ew` every forward although it only depends on the
topology (src/ops/BSMS.py:73).  Here a `LevelPlan` holds, per level, the dst-sorted and src-sorted
CSR views the kernels consume (include/bsms_b200.h), and a `HierarchyPlan` adds the pooled-id maps
and the cached transfer weights of a whole `(m_gs, m_ids)` hierarchy.  Plans are keyed on the
identity of the caller's tensors, so the unchanged `BSMS_Simulator.forward` can hand the same
`m_gs`/`m_ids` in every step without re-planning.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr


class LevelPlan:
    """dst-/src-sorted CSR views of one level graph g = m_gs[l] (int64 [2,E] on the GPU)."""

    def __init__(self, g: torch.Tensor, n_nodes: int):
        _lib.require_cuda(g)
        if g.dim() != 2 or g.shape[0] != 2:
            raise _lib.BsmsError(f"edge list must be [2,E], got {tuple(g.shape)}")
        if g.dtype != torch.int64:
            g = g.long()
        g = g.contiguous()
        dev = g.device
        E, N = int(g.shape[1]), int(n_nodes)
        self.n_nodes, self.n_edges, self.device = N, E, dev
        i32 = dict(dtype=torch.int32, device=dev)
        e = max(E, 1)
        self.src_d, self.dst_d, self.perm_d = (torch.empty(e, **i32) for _ in range(3))
        self.src_s, self.dst_s, self.s2d = (torch.empty(e, **i32) for _ in range(3))
        self.rowptr_d = torch.empty(N + 1, **i32)
        self.rowptr_s = torch.empty(N + 1, **i32)
        self.c = _lib.LevelPlanC(N, E, *(C.c_void_p(t.data_ptr()) for t in (
            self.src_d, self.dst_d, self.rowptr_d, self.perm_d, self.src_s, self.dst_s, self.rowptr_s, self.s2d)))
        ws_bytes = int(lib.bsms_plan_workspace_bytes(E, N))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        status = torch.zeros(4, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.bsms_plan_build(ptr(g) if E else None, E, N, C.byref(self.c), ptr(status), ptr(ws), ws_bytes,
                                      stream_ptr()))
        # the reference's degree() sizes itself by max(index)+1 and ignores num_nodes
        # (src/utils/basic.py:305-307): remember it so cal_ew can raise where the reference does.
        self.max_src = int(g[0].max()) if E else -1

    def byref(self):
        return C.byref(self.c)


def _key_of(t: torch.Tensor):
    return (t.data_ptr(), tuple(t.shape), t._version, t.device.index)


class _Cache:
    def __init__(self, cap):
        self.cap, self.d = cap, OrderedDict()

    def get(self, key):
        v = self.d.get(key)
        if v is not None:
            self.d.move_to_end(key)
        return v

    def put(self, key, val):
        self.d[key] = val
        while len(self.d) > self.cap:
            self.d.popitem(last=False)


_LEVELS = _Cache(64)
_HIERS = _Cache(8)


def level_plan(g: torch.Tensor, n_nodes: int) -> LevelPlan:
    key = (_key_of(g), int(n_nodes))
    hit = _LEVELS.get(key)
    if hit is None:
        hit = (LevelPlan(g, n_nodes), g)  # keep g alive so its data_ptr cannot be recycled
        _LEVELS.put(key, hit)
    return hit[0]


def cal_ew_raw(level: LevelPlan, w: torch.Tensor, want_orig: bool):
    """-> (ew_orig|None, ew_d, ew_s, aggr_w) on the level's device; w: fp32 [N]."""
    E, N = level.n_edges, level.n_nodes
    if E == 0:
        raise RuntimeError("cal_ew on a graph without edges (the reference fails in torch.max, "
                           "src/utils/basic.py:305)")
    if level.max_src + 1 != N:
        raise RuntimeError(
            f"cal_ew: the size of w ({N}) must match the out-degree vector ({level.max_src + 1}); the reference's "
            "degree() ignores num_nodes (src/utils/basic.py:305-307) and fails the same way")
    f32 = dict(dtype=torch.float32, device=level.device)
    ew_d, ew_s = torch.empty(E, **f32), torch.empty(E, **f32)
    ew_o = torch.empty(E, **f32) if want_orig else None
    aggr_w = torch.empty(N, **f32)
    check(lib.bsms_cal_ew(level.byref(), ptr(w), ptr(ew_o), ptr(ew_d), ptr(ew_s), ptr(aggr_w), stream_ptr()))
    return ew_o, ew_d, ew_s, aggr_w


class HierarchyPlan:
    """Everything topology-only for one `(m_gs, m_ids)`: level plans, pooled ids (+ inverse), cached ew."""

    def __init__(self, m_gs, m_ids, n0: int):
        depth = len(m_ids)
        if len(m_gs) < depth + 1:
            raise _lib.BsmsError(f"need {depth + 1} level graphs for {depth} pooling levels, got {len(m_gs)}")
        self.depth = depth
        self.n = [int(n0)] + [int(i.shape[0]) for i in m_ids]
        self.levels = [level_plan(m_gs[l], self.n[l]) for l in range(depth + 1)]
        dev = self.levels[0].device
        self.ids, self.inv, self.ew_d, self.ew_s = [], [], [], []
        w = torch.ones(self.n[0], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), torch.no_grad():
            for l in range(depth):
                ids64 = m_ids[l].to(device=dev, dtype=torch.int64).contiguous()
                if ids64.numel() and (int(ids64.min()) < 0 or int(ids64.max()) >= self.n[l]):
                    raise IndexError(f"m_ids[{l}] out of range for a level with {self.n[l]} nodes")
                ids32 = ids64.to(torch.int32)
                inv = torch.full((self.n[l],), -1, dtype=torch.int32, device=dev)
                inv[ids64] = torch.arange(ids64.numel(), dtype=torch.int32, device=dev)
                _, ew_d, ew_s, aggr_w = cal_ew_raw(self.levels[l], w, False)
                self.ids.append(ids32)
                self.inv.append(inv)
                self.ew_d.append(ew_d)
                self.ew_s.append(ew_s)
                w = aggr_w[ids64].contiguous()  # src/ops/BSMS.py:89
        self._keep = (list(m_gs), list(m_ids))

    def edge_rows_per_forward(self) -> int:
        d = self.depth
        return 2 * sum(p.n_edges for p in self.levels[:d]) + self.levels[d].n_edges

    def node_rows_per_forward(self) -> int:
        d = self.depth
        return 2 * sum(self.n[:d]) + self.n[d]


def hierarchy_plan(m_gs, m_ids, n0: int) -> HierarchyPlan:
    depth = len(m_ids)
    key = (tuple(_key_of(g) for g in m_gs[:depth + 1]), tuple(_key_of(i) for i in m_ids), int(n0))
    hit = _HIERS.get(key)
    if hit is None:
        hit = HierarchyPlan(m_gs, m_ids, n0)
        _HIERS.put(key, hit)
    return hit
