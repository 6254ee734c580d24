"""Graph plans: device-resident int32 index structures built once per mesh and cached BY CONTENT.

The reference re-derives everything from the raw int64 `[2,E]` edge lists on every call
(src/ops/basic.py:66,130-137) and recomputes `cal_ew` every forward although it only depends on the
topology (src/ops/BSMS.py:73).  Here a `LevelPlan` holds, per level, the dst-sorted and src-sorted
CSR views the kernels consume (include/bsms_b200.h), and a `HierarchyPlan` adds the pooled-id maps
and the cached transfer weights of a whole `(m_gs, m_ids)` hierarchy.

Cache structure (what makes the unchanged `BSMS_Simulator.forward` a drop-in caller):

* The real caller re-creates the index tensors on the device EVERY step — `Trainer.move_to_device`
  copies the whole batch (src/trainer/trainer.py:100-117) and `model.forward` slices `g[0]` views of
  it (src/models/model.py:189-200) — so tensor identity is useless as a key there.  Plans are keyed
  on a 64-bit CONTENT fingerprint of every index tensor (`bsms_fingerprint`: one launch for the whole
  hierarchy + one 8·n-byte read-back).  A step with fresh copies of a known mesh costs that one
  launch and one small synchronising copy; nothing is re-sorted, `cal_ew` is not recomputed.
* In front of it sits a small identity cache (data_ptr / shape / version, a handful of entries) for
  callers that keep their tensors resident (bench.py, rollouts, CUDA-graph capture): zero launches,
  zero synchronisation.  It holds references to the caller's tensors only for those few entries;
  the content cache holds the plans alone.
* `BSGMP.bind_mesh(m_gs, m_ids)` (ops.py) pins one hierarchy explicitly and skips both look-ups.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr

# counters the drop-in tests read: how often index structures were (re)built / fingerprinted
STATS = {"level_builds": 0, "hierarchy_builds": 0, "fingerprints": 0, "identity_hits": 0, "content_hits": 0}


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
        # (status[1] was produced by the build, which has already synchronised the stream.)
        self.max_src = int(status[1]) if E else -1
        STATS["level_builds"] += 1

    def byref(self):
        return C.byref(self.c)


def _ident(t: torch.Tensor):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t._version, t.device.index, t.dtype)


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
        self.d.move_to_end(key)
        while len(self.d) > self.cap:
            self.d.popitem(last=False)

    def clear(self):
        self.d.clear()


# identity caches (values keep the caller's tensors alive so a data_ptr cannot be recycled): small
_ID_LEVELS = _Cache(16)
_ID_HIERS = _Cache(2)
# content caches: plans only
_LEVELS = _Cache(64)
_HIERS = _Cache(8)


def clear_caches():
    for c in (_ID_LEVELS, _ID_HIERS, _LEVELS, _HIERS):
        c.clear()


def fingerprints(tensors):
    """64-bit content fingerprints of int64 CUDA tensors: one launch, one synchronising read-back."""
    dev = tensors[0].device
    ts = [t if (t.dtype == torch.int64 and t.is_contiguous()) else t.long().contiguous() for t in tensors]
    out = torch.empty(len(ts), dtype=torch.int64, device=dev)
    fps = []
    with torch.cuda.device(dev):
        for lo in range(0, len(ts), 32):
            chunk = ts[lo:lo + 32]
            ptrs = (C.c_void_p * len(chunk))(*[t.data_ptr() for t in chunk])
            sizes = (C.c_int64 * len(chunk))(*[t.numel() * 8 for t in chunk])
            check(lib.bsms_fingerprint(ptrs, sizes, len(chunk), C.c_void_p(out[lo:].data_ptr()), stream_ptr()))
    STATS["fingerprints"] += 1
    fps = out.tolist()
    return [(fp, tuple(t.shape)) for fp, t in zip(fps, ts)]


def level_plan(g: torch.Tensor, n_nodes: int) -> LevelPlan:
    """Plan of one level graph (the standalone GMP / WeightedEdgeConv modules come through here)."""
    _lib.require_cuda(g)
    ikey = (_ident(g), int(n_nodes))
    hit = _ID_LEVELS.get(ikey)
    if hit is not None:
        STATS["identity_hits"] += 1
        return hit[0]
    ckey = (fingerprints([g])[0], int(n_nodes), g.device.index)
    plan = _LEVELS.get(ckey)
    if plan is None:
        plan = LevelPlan(g, n_nodes)
        _LEVELS.put(ckey, plan)
    else:
        STATS["content_hits"] += 1
    _ID_LEVELS.put(ikey, (plan, g))  # g kept alive: its data_ptr cannot be recycled while the entry lives
    return plan


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

    def __init__(self, m_gs, m_ids, n0: int, level_keys=None):
        depth = len(m_ids)
        if len(m_gs) < depth + 1:
            raise _lib.BsmsError(f"need {depth + 1} level graphs for {depth} pooling levels, got {len(m_gs)}")
        self.depth = depth
        self.n = [int(n0)] + [int(i.shape[0]) for i in m_ids]
        self.levels = []
        for l in range(depth + 1):
            plan = None
            if level_keys is not None:
                ckey = (level_keys[l], self.n[l], m_gs[l].device.index)
                plan = _LEVELS.get(ckey)
            if plan is None:
                plan = LevelPlan(m_gs[l], self.n[l])
                if level_keys is not None:
                    _LEVELS.put(ckey, plan)
            self.levels.append(plan)
        dev = self.levels[0].device
        self.ids, self.inv, self.ew_d, self.ew_s = [], [], [], []
        w = torch.ones(self.n[0], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), torch.no_grad():
            # one synchronising range check for all pooled-id tensors (one-time per mesh)
            if depth:
                lim = torch.stack([torch.stack([i.min(), i.max()]) if i.numel() else i.new_zeros(2) for i in m_ids]).tolist()
                for l, (lo, hi) in enumerate(lim):
                    if m_ids[l].numel() and (lo < 0 or hi >= self.n[l]):
                        raise IndexError(f"m_ids[{l}] out of range for a level with {self.n[l]} nodes")
            for l in range(depth):
                ids64 = m_ids[l].to(device=dev, dtype=torch.int64).contiguous()
                ids32 = ids64.to(torch.int32)
                inv = torch.full((self.n[l],), -1, dtype=torch.int32, device=dev)
                inv[ids64] = torch.arange(ids64.numel(), dtype=torch.int32, device=dev)
                _, ew_d, ew_s, aggr_w = cal_ew_raw(self.levels[l], w, False)
                self.ids.append(ids32)
                self.inv.append(inv)
                self.ew_d.append(ew_d)
                self.ew_s.append(ew_s)
                w = aggr_w[ids64].contiguous()  # src/ops/BSMS.py:89
        STATS["hierarchy_builds"] += 1

    def edge_rows_per_forward(self) -> int:
        d = self.depth
        return 2 * sum(p.n_edges for p in self.levels[:d]) + self.levels[d].n_edges

    def node_rows_per_forward(self) -> int:
        d = self.depth
        return 2 * sum(self.n[:d]) + self.n[d]


def hierarchy_plan(m_gs, m_ids, n0: int) -> HierarchyPlan:
    depth = len(m_ids)
    gs = list(m_gs[:depth + 1])
    ikey = (tuple(_ident(g) for g in gs), tuple(_ident(i) for i in m_ids), int(n0))
    hit = _ID_HIERS.get(ikey)
    if hit is not None:
        STATS["identity_hits"] += 1
        return hit[0]
    if torch.cuda.is_current_stream_capturing():
        raise _lib.BsmsError("a new (m_gs, m_ids) reached BSGMP.forward during CUDA-graph capture: run one eager "
                             "forward with these tensors (or bind_mesh) before capturing")
    fps = fingerprints(gs + list(m_ids))
    ckey = (tuple(fps), int(n0), gs[0].device.index)
    plan = _HIERS.get(ckey)
    if plan is None:
        plan = HierarchyPlan(gs, list(m_ids), n0, level_keys=fps[:depth + 1])
        _HIERS.put(ckey, plan)
    else:
        STATS["content_hits"] += 1
    _ID_HIERS.put(ikey, (plan, gs, list(m_ids)))
    return plan
