"""K6 on the host side: the halo exchange of the node-partitioned processor as ONE push kernel per exchange
over peer-mapped memory (`bsms_halo_exchange`, csrc/halo.cu) instead of index_select + an NCCL point-to-point
group + index_add_ (partitioned._HaloP2P).

One process per GPU of one NVLink/NVSwitch box.  Every rank allocates an arena (cudaMalloc through the C-ABI),
exports it with CUDA IPC, and maps the arenas of its peers; `torch.distributed` only carries the 64-byte handles
and the layout tables once at set-up.  Per call site of the schedule (4·depth + 1 per forward, fixed order) the
arena holds this rank's `[owned | ghost]` output buffer (peers store the ghost rows straight into it), the
`back` region its peers return ghost gradients into, two flag vectors and two control words.

Buffer reuse across steps is safe because every training step ends in a collective over all ranks (the
gradient all-reduce): no rank can start the next step's call site k before every rank has finished the
previous step's backward, which is the last reader of the buffers of call site k.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib, stream_ptr


class _DevMem:
    """torch view of raw device memory (no ownership) through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def _view(ptr, nbytes, dtype, shape, device):
    t = torch.as_tensor(_DevMem(ptr, nbytes), device=device)
    return t.view(dtype).view(*shape)


def _align(x, a=256):
    return (x + a - 1) // a * a


class PushExchanger:
    """`exchange(states, level, owned)` of partitioned.PartitionedBSGMP for one rank per process."""

    def __init__(self, group=None):
        self.group = group
        self.ready = False

    # ------------------------------------------------------------------ set-up (collective)
    def setup(self, state, site_levels, channels=128):
        """state: partitioned.RankState of this rank; site_levels[s] = level of the s-th exchange of a forward."""
        self.state, self.site_levels, self.Cc = state, list(site_levels), int(channels)
        self.rank, self.world = state.rank, state.world
        dev = state.levels[0].send_cat.device
        self.device = dev
        row = self.Cc * 4
        lv = state.levels
        # arena layout of this rank
        off, self.sites = 0, []
        for l in self.site_levels:
            n_local, n_send = lv[l].n_local, int(lv[l].send_cat.numel())
            s = {"level": l, "out": off}
            off += _align(max(n_local, 1) * row)
            s["back"] = off
            off += _align(max(n_send, 1) * row)
            s["flags_f"], s["flags_b"] = off, off + 64
            s["ctrl_f"], s["ctrl_b"] = off + 128, off + 160
            off += 256
            self.sites.append(s)
        self.arena_bytes = off
        base = C.c_void_p()
        with torch.cuda.device(dev):
            check(lib.bsms_ipc_alloc(self.arena_bytes, C.byref(base)))
            handle = C.create_string_buffer(64)
            check(lib.bsms_ipc_export(base, handle))
        self.base = int(base.value)
        mine = {"handle": bytes(handle.raw), "sites": [(s["out"], s["back"], s["flags_f"], s["flags_b"]) for s in self.sites],
                "n_own": [l_.n_own for l_ in lv], "recv": [list(l_.recv_counts) for l_ in lv],
                "send": [list(l_.send_counts) for l_ in lv]}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        self.peer_base = [0] * self.world
        with torch.cuda.device(dev):
            for q in range(self.world):
                if q == self.rank:
                    self.peer_base[q] = self.base
                    continue
                p = C.c_void_p()
                check(lib.bsms_ipc_open(everyone[q]["handle"], C.byref(p)))
                self.peer_base[q] = int(p.value)
        # per site: the argument blocks of the forward and the backward launch
        self.args = []
        for si, s in enumerate(self.sites):
            l = s["level"]
            L = lv[l]
            n_ghost, n_send = L.n_local - L.n_own, int(L.send_cat.numel())
            send_off = np.concatenate([[0], np.cumsum(L.send_counts)]).astype(np.int64)
            recv_off = np.concatenate([[0], np.cumsum(L.recv_counts)]).astype(np.int64)
            out_t = _view(self.base + s["out"], max(L.n_local, 1) * row, torch.float32, (max(L.n_local, 1), self.Cc), dev)[:L.n_local]
            pair = []
            for backward in (0, 1):
                a = _lib.HaloArgsC()
                a.world, a.rank, a.channels, a.backward = self.world, self.rank, self.Cc, backward
                a.n_own, a.n_ghost, a.n_send = L.n_own, n_ghost, n_send
                a.send_idx = L.send_cat.data_ptr() if n_send else None
                for q in range(9):
                    a.send_off[q] = int(send_off[min(q, self.world)])
                    a.recv_off[q] = int(recv_off[min(q, self.world)])
                for q in range(self.world):
                    pq = everyone[q]
                    so, sb, sff, sfb = pq["sites"][si]
                    if backward == 0:
                        # my rows land in q's out buffer behind q's owned rows and the ghosts q gets from ranks < me
                        first = pq["n_own"][l] + sum(pq["recv"][l][:self.rank])
                        a.peer_dst[q] = self.peer_base[q] + so + first * row
                        a.peer_flag[q] = self.peer_base[q] + sff + 4 * self.rank
                    else:
                        # my ghost gradients land in q's back region behind what q sent to ranks < me
                        first = sum(pq["send"][l][:self.rank])
                        a.peer_dst[q] = self.peer_base[q] + sb + first * row
                        a.peer_flag[q] = self.peer_base[q] + sfb + 4 * self.rank
                a.back = self.base + s["back"]
                a.my_flags = self.base + (s["flags_b"] if backward else s["flags_f"])
                a.ctrl = self.base + (s["ctrl_b"] if backward else s["ctrl_f"])
                pair.append(a)
            self.args.append((pair[0], pair[1], out_t))
        dist.barrier(group=self.group)
        self.cursor = 0
        self.ready = True

    def begin(self):
        self.cursor = 0

    # ------------------------------------------------------------------ the exchange
    def exchange(self, states, level, owned):
        (st,), (x,) = states, owned
        s = self.cursor
        self.cursor += 1
        if self.sites[s]["level"] != level:
            raise _lib.BsmsError(f"halo call site {s} was laid out for level {self.sites[s]['level']}, got level {level}")
        return [_HaloPush.apply(x.contiguous(), self, s)]


class _HaloPush(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_own, ex, site):
        a_f, _, out = ex.args[site]
        if x_own.shape[0] != a_f.n_own or x_own.shape[1] != ex.Cc:
            raise _lib.BsmsError(f"halo site {site}: expected [{a_f.n_own}, {ex.Cc}] owned rows, got {tuple(x_own.shape)}")
        a_f.src = x_own.data_ptr() if x_own.numel() else ex.base  # never dereferenced when n_own == 0
        a_f.dst = out.data_ptr() if out.numel() else ex.base
        with torch.cuda.device(x_own.device):
            check(lib.bsms_halo_exchange(C.byref(a_f), stream_ptr()))
        ctx.ex, ctx.site = ex, site
        return out.view(out.shape)  # a fresh tensor object over the persistent arena region

    @staticmethod
    def backward(ctx, g):
        ex, site = ctx.ex, ctx.site
        _, a_b, _ = ex.args[site]
        g = g.contiguous()
        g_own = torch.empty(a_b.n_own, ex.Cc, dtype=g.dtype, device=g.device)
        a_b.src = g.data_ptr() if g.numel() else ex.base
        a_b.dst = g_own.data_ptr() if g_own.numel() else ex.base
        with torch.cuda.device(g.device):
            check(lib.bsms_halo_exchange(C.byref(a_b), stream_ptr()))
        return g_own, None, None
