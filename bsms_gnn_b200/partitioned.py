"""Node-partitioned BSMS processor: every rank runs the reference schedule (src/ops/BSMS.py:39-104)
on its LOCAL rows `[owned | ghosts]` with the same sm_100a kernels, and ghost rows are refreshed from
their owners before every operator that reads neighbours (GMP, restriction, prolongation) —
`4·depth + 1` feature exchanges per forward, mirrored in backward.

    plans = partition.build_rank_plan(...) (+ exchange_requests)      # once per mesh
    pmodel = PartitionedBSGMP(bsgmp_module, [plan], DistExchanger(...))
    out_own, = pmodel([h_own], [pos_own])                              # [n_own, 128] per rank

The schedule is written over a LIST of rank states so the same code runs (a) distributed, one state
per process, ghost traffic as NCCL point-to-point sends between GPUs over NVLink
(`DistExchanger`), and (b) single-process with R virtual ranks and an in-memory, autograd-visible
exchange (`LocalExchanger`) — the way the multi-rank path is tested on one GPU.
Un-batched tensors only ([n, C]; the large partitioned meshes run at B = 1).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .ops import BSGMP
from .plan import LevelPlan


class LocalLevel:
    """Device-side index structures of one level of one rank."""

    def __init__(self, lp, device):
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
        self.n_own, self.n_local = lp.n_own, lp.n_local
        # a rank can own nothing at a deep level (e.g. 52 nodes over 8 ranks): no plan, operators return empties
        self.plan = LevelPlan(t(lp.edges, torch.int64), lp.n_local) if lp.n_local > 0 else None
        # the GMP aggregates onto OWNED receivers only: its plan holds the receiver-owned edges (no redundant edge rows)
        self.plan_gmp = (LevelPlan(t(lp.edges[:, np.asarray(lp.gmp_sel, dtype=bool)], torch.int64), lp.n_local)
                         if lp.n_local > 0 else None)
        self.recv_counts = [int(c) for c in lp.recv_counts]
        self.send_idx = [t(s, torch.int64) for s in lp.send_idx]
        self.send_counts = [int(s.numel()) for s in self.send_idx]
        self.send_cat = torch.cat(self.send_idx) if self.send_idx else torch.zeros(0, dtype=torch.int64, device=device)
        self.has_transfer = lp.ew is not None
        if self.has_transfer and self.plan is None:
            self.has_transfer = False
            self.n_own_next = 0
        if self.has_transfer:
            E = max(self.plan.n_edges, 1)
            ew = t(lp.ew, torch.float32)
            self.ew_d = torch.empty(E, dtype=torch.float32, device=device)
            self.ew_s = torch.empty(E, dtype=torch.float32, device=device)
            with torch.cuda.device(device):
                check(lib.bsms_permute_ew(self.plan.byref(), ptr(ew), ptr(self.ew_d), ptr(self.ew_s), stream_ptr()))
            self.ids = t(lp.ids, torch.int32)                      # owned kept fine rows -> coarse rows 0..n_own'-1
            inv = np.asarray(lp.inv)
            self.inv = t(inv, torch.int32)                         # local fine row -> local coarse row / -1
            n_own_next = int(lp.ids.shape[0])
            self.inv_own = t(np.where(inv < n_own_next, inv, -1), torch.int32)  # ... owned coarse rows only
            self.n_own_next = n_own_next

    def set_next(self, n_local_next):
        """ids_full[c] = local fine row of local coarse row c, or -1 when that fine node is remote."""
        self.n_local_next = n_local_next
        if self.plan is None or not self.has_transfer:
            return
        inv = self.inv.cpu().numpy()
        full = np.full(n_local_next, -1, dtype=np.int32)
        rows = np.nonzero(inv >= 0)[0]
        full[inv[rows]] = rows
        self.ids_full = torch.from_numpy(full).to(self.inv.device)
        self.n_local_next = n_local_next


class RankState:
    def __init__(self, plan, device):
        self.rank, self.world = plan.rank, plan.world
        self.levels = [LocalLevel(lp, device) for lp in plan.levels]
        for l in range(len(self.levels) - 1):
            self.levels[l].set_next(self.levels[l + 1].n_local)


# ------------------------------------------------------------------------------------------ transfers
class _PRestrict(torch.autograd.Function):
    """coarse_owned = conv_down(x_local)[ids]; backward = prolongation kernel over the owned coarse rows."""

    @staticmethod
    def forward(ctx, x, lv):
        Cc = x.shape[-1]
        out = torch.empty(lv.n_own_next, Cc, dtype=x.dtype, device=x.device)
        ctx.lv = lv
        if lv.n_own_next == 0:
            return out
        with torch.cuda.device(x.device):
            check(lib.bsms_conv_down_pool(lv.plan.byref(), ptr(lv.ew_d), ptr(lv.ids), lv.n_own_next, ptr(x), ptr(out), 1,
                                          Cc, stream_ptr()))
        ctx.lv = lv
        return out

    @staticmethod
    def backward(ctx, g):
        lv = ctx.lv
        g = g.contiguous()
        if lv.n_own_next == 0:
            return g.new_zeros(lv.n_local, g.shape[-1]), None
        gx = torch.empty(lv.n_local, g.shape[-1], dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            check(lib.bsms_unpool_conv_up(lv.plan.byref(), ptr(lv.ew_s), ptr(lv.inv_own), lv.n_own_next, ptr(g), ptr(gx), 1,
                                          g.shape[-1], stream_ptr()))
        return gx, None


class _PProlong(torch.autograd.Function):
    """fine_local = conv_up(unpool(coarse_local)); backward = restriction kernel onto every local coarse row."""

    @staticmethod
    def forward(ctx, hc, lv):
        Cc = hc.shape[-1]
        ctx.lv = lv
        if lv.n_local_next == 0:
            return hc.new_zeros(lv.n_local, Cc)
        out = torch.empty(lv.n_local, Cc, dtype=hc.dtype, device=hc.device)
        with torch.cuda.device(hc.device):
            check(lib.bsms_unpool_conv_up(lv.plan.byref(), ptr(lv.ew_s), ptr(lv.inv), lv.n_local_next, ptr(hc), ptr(out), 1,
                                          Cc, stream_ptr()))
        ctx.lv = lv
        return out

    @staticmethod
    def backward(ctx, g):
        lv = ctx.lv
        g = g.contiguous()
        gh = torch.empty(lv.n_local_next, g.shape[-1], dtype=g.dtype, device=g.device)
        if lv.n_local_next == 0:
            return gh, None
        with torch.cuda.device(g.device):
            check(lib.bsms_conv_down_pool(lv.plan.byref(), ptr(lv.ew_d), ptr(lv.ids_full), lv.n_local_next, ptr(g), ptr(gh),
                                          1, g.shape[-1], stream_ptr()))
        return gh, None


# ------------------------------------------------------------------------------------------ exchanges
class LocalExchanger:
    """R virtual ranks in one process: ghosts are copied in memory with autograd-visible torch ops."""

    def exchange(self, states, level, owned):
        out = []
        for r, st in enumerate(states):
            lv = st.levels[level]
            parts = [owned[r]]
            for q, sq in enumerate(states):
                idx = sq.levels[level].send_idx[r]
                if idx.numel():
                    parts.append(owned[q].index_select(0, idx))
            out.append(torch.cat(parts, 0) if len(parts) > 1 else owned[r])
            assert out[-1].shape[0] == lv.n_local
        return out


class _HaloP2P(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_own, lv, rank, group):
        Cc = x_own.shape[-1]
        out = x_own.new_empty(lv.n_local, Cc)
        out[:lv.n_own] = x_own
        send = x_own.index_select(0, lv.send_cat) if lv.send_cat.numel() else x_own.new_empty(0, Cc)
        ops, so, ro = [], 0, lv.n_own
        for q in range(len(lv.recv_counts)):
            sc, rc = lv.send_counts[q], lv.recv_counts[q]
            if q != rank and sc:
                ops.append(dist.P2POp(dist.isend, send[so:so + sc], q, group))
            if q != rank and rc:
                ops.append(dist.P2POp(dist.irecv, out[ro:ro + rc], q, group))
            so += sc
            ro += rc
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        ctx.lv, ctx.rank, ctx.group = lv, rank, group
        return out

    @staticmethod
    def backward(ctx, g):
        lv, rank, group = ctx.lv, ctx.rank, ctx.group
        g = g.contiguous()
        Cc = g.shape[-1]
        g_own = g[:lv.n_own].clone()
        back = g.new_empty(int(lv.send_cat.numel()), Cc)  # gradients of the rows this rank sent out
        ops, so, ro = [], 0, lv.n_own
        for q in range(len(lv.recv_counts)):
            sc, rc = lv.send_counts[q], lv.recv_counts[q]
            if q != rank and rc:
                ops.append(dist.P2POp(dist.isend, g[ro:ro + rc], q, group))
            if q != rank and sc:
                ops.append(dist.P2POp(dist.irecv, back[so:so + sc], q, group))
            so += sc
            ro += rc
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if back.numel():
            g_own.index_add_(0, lv.send_cat, back)
        return g_own, None, None, None


class DistExchanger:
    """One rank per process: ghost rows travel as point-to-point sends (NCCL over NVLink on GPUs)."""

    def __init__(self, group=None):
        self.group = group

    def exchange(self, states, level, owned):
        (st,), (x,) = states, owned
        lv = st.levels[level]
        if lv.n_local == lv.n_own and not lv.send_cat.numel():
            return [x]
        return [_HaloP2P.apply(x.contiguous(), lv, st.rank, self.group)]


def exchange_requests(plan, group=None):
    """Distributed completion of a RankPlan: tell every owner which rows this rank needs."""
    world = plan.world
    mine = [[np.asarray(lp.requests[q]) for q in range(world)] for lp in plan.levels]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    incoming = [[gathered[q][l][plan.rank] for q in range(world)] for l in range(len(plan.levels))]
    plan.finalize(incoming)
    return plan


# ------------------------------------------------------------------------------------------ the schedule
class PartitionedBSGMP(torch.nn.Module):
    def __init__(self, model: BSGMP, plans, exchanger, device=None, pos_exchanger=None):
        super().__init__()
        self.model = model
        device = device or next(model.parameters()).device
        self.states = [RankState(p, device) for p in plans]
        self.ex = exchanger
        self.depth = model.unet_depth
        # positions of a static mesh are exchanged once (cached): they may travel over a different exchanger than
        # the features (the push exchanger lays out one buffer per feature call site)
        self.ex_pos = pos_exchanger or exchanger
        if hasattr(exchanger, "setup") and not exchanger.ready:
            exchanger.setup(self.states[0], self.site_levels(), 128)
        self._pos_key, self._pos_cache, self._pos_ref = None, None, None

    def site_levels(self):
        """Level of every feature exchange of one forward, in call order (4·depth + 1 sites)."""
        d = self.depth
        seq = []
        for l in range(d):
            seq += [l, l]
        seq.append(d)
        for k in range(d):
            l = d - 1 - k
            seq += [l + 1, l]
        return seq

    @staticmethod
    def _gmp(gmp, x_loc, lv, p_loc):
        if lv.plan is None:
            return x_loc[:0]
        return gmp._run(x_loc, lv.plan_gmp, p_loc)[:lv.n_own]

    @staticmethod
    def _restrict(x_loc, lv):
        if lv.plan is None:
            return x_loc.new_zeros(0, x_loc.shape[-1])
        return _PRestrict.apply(x_loc.contiguous(), lv)

    @staticmethod
    def _prolong(hc_loc, lv):
        if lv.plan is None:
            return hc_loc.new_zeros(0, hc_loc.shape[-1])
        return _PProlong.apply(hc_loc.contiguous(), lv)[:lv.n_own]

    def _positions(self, pos_own):
        """Local positions [owned | ghosts] of every level.  They depend on the positions only (the transfer
        weights are topology-only), so for a static mesh they are exchanged and restricted ONCE and reused
        while the caller passes the same, unmodified position tensors (identity + version check)."""
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in pos_own)
        if key == self._pos_key:
            return self._pos_cache
        # the key is only meaningful while the tensors it was taken from are alive: keep them, otherwise the
        # caching allocator can hand the same address (and version 0) to a DIFFERENT position tensor
        self._pos_ref = list(pos_own)
        d, S, ex = self.depth, self.states, self.ex_pos
        R = range(len(S))
        with torch.no_grad():
            p = [t.detach().to(torch.float32).contiguous() for t in pos_own]
            pos_loc = []
            for l in range(d):
                p_loc = ex.exchange(S, l, p)
                pos_loc.append(p_loc)
                p = [self._restrict(p_loc[r], S[r].levels[l]) for r in R]
            pos_loc.append(ex.exchange(S, d, p))
        self._pos_key, self._pos_cache = key, pos_loc
        return pos_loc

    def forward(self, h_own, pos_own):
        d, S, ex, m = self.depth, self.states, self.ex, self.model
        R = range(len(S))
        x = [t.contiguous() for t in h_own]
        pos_loc = self._positions(pos_own)
        if hasattr(ex, "begin"):
            ex.begin()
        skips = []
        for l in range(d):
            x_loc, p_loc = ex.exchange(S, l, x), pos_loc[l]
            y = [self._gmp(m.down_gmps[l], x_loc[r], S[r].levels[l], p_loc[r]) for r in R]
            skips.append(y)
            y_loc = ex.exchange(S, l, y)
            x = [self._restrict(y_loc[r], S[r].levels[l]) for r in R]
        x_loc, p_loc = ex.exchange(S, d, x), pos_loc[d]
        x = [self._gmp(m.bottom_gmp, x_loc[r], S[r].levels[d], p_loc[r]) for r in R]
        for k in range(d):
            l = d - 1 - k
            hc_loc = ex.exchange(S, l + 1, x)
            u = [self._prolong(hc_loc[r], S[r].levels[l]) for r in R]
            u_loc = ex.exchange(S, l, u)
            x = [self._gmp(m.up_gmps[k], u_loc[r], S[r].levels[l], pos_loc[l][r]) + skips[l][r] for r in R]
        return x
