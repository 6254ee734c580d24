"""Host-side mirror of the reference operator surface for the BSMS processor hot path.

Same class names, constructor arguments, forward signatures, parameter names and error behaviour
as the reference modules (src/ops/basic.py:6-201, src/ops/BSMS.py:8-104), so
`from bsms_gnn_b200.ops import MLP, BSGMP` can replace `from ops import MLP, BSGMP`
(src/models/model.py:2) and reference checkpoints load unchanged.  All arithmetic of the processor
runs in libbsms_b200.so through the C-ABI (include/bsms_b200.h); PyTorch supplies device memory,
streams and the autograd graph.  There is no CPU path: CPU tensors raise.

Limits compared with the reference modules (they raise `BsmsError` at construction / call time, they never
fall back): GMP / BSGMP are built for `latent_dim == 128`, `hidden_layer == 3`, `pos_dim` 1..3 (the values of
every configs/model/*.yaml); tensors must be fp32 CUDA tensors.  Arithmetic modes of the MLP contractions:
`fp32` (FFMA, exact fp32 products), `fp16x3` (tcgen05, 2-way fp16 split, 3 MMAs, fp32-grade: forward within
1e-5 of the reference) and `bf16` (tcgen05, one bf16 MMA; activations AND weights are rounded to bf16 at every
MMA input, the edge-layer biases b2..b4 too because they ride in the MMA; storage and accumulation stay fp32;
forward within ~5e-3 of the reference).  Standalone `MLP` accepts any shape (it runs the declared torch layers).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, plan as _plan
from ._lib import lib, check, ptr, stream_ptr, BsmsError

LATENT = 128
# Default arithmetic of the MLP contractions: the fp32-parity tensor-core mode (forward on tcgen05 with the 2-way fp16
# split, within 1e-5 of the reference — BASELINE.json north_star; its backward is the exact-fp32 path).  "bf16" is the
# training-throughput mode BASELINE.json config 3 names; "fp32" runs everything on FFMA.
_DEFAULT_MODE = "fp16x3"


def set_default_mode(mode: str):
    """Arithmetic of the MLP contractions for modules created afterwards: 'fp32' | 'fp16x3' | 'bf16'."""
    global _DEFAULT_MODE
    if mode not in _lib.MODES:
        raise ValueError(f"unknown mode {mode!r}; choose from {sorted(_lib.MODES)}")
    _DEFAULT_MODE = mode


def set_deterministic(on: bool):
    """Bitwise run-to-run reproducible forward and backward (the role torch.use_deterministic_algorithms plays for the
    reference, whose scatter_add_ is order-dependent on CUDA).  While on: `bf16` blocks keep their tensor-core kernels,
    but the fused edge kernels hand their per-edge-row results to order-fixed CSR segment sums instead of reducing with
    red.add, and every per-CTA atomic flush becomes a partial-sum block + one ordered reduction (about 1.5x the bf16
    step); `fp32` blocks commit their weight-gradient partial sums in ticket order; `fp16x3` blocks run as `fp32`
    (same 1e-5 grade).  The transfer operators are order-fixed CSR sums in every mode."""
    lib.bsms_set_deterministic(1 if on else 0)


def is_deterministic() -> bool:
    return bool(lib.bsms_get_deterministic())


def _as_b3(x: torch.Tensor, what: str):
    """[N,C] | [B,N,C] fp32 CUDA -> contiguous [B,N,C] view + original rank."""
    if x.dim() not in (2, 3):
        raise NotImplementedError("Only implemented for dim 2 and 3")  # src/ops/basic.py:74,81,134
    _lib.require_cuda(x)
    if x.dtype != torch.float32:
        raise BsmsError(f"{what} must be float32 (features are fp32 in HBM in every mode), got {x.dtype}")
    x3 = x if x.dim() == 3 else x.unsqueeze(0)
    return x3.contiguous()


class MLP(nn.Module):
    """hidden_layers×(Linear+ReLU) + Linear [+ LayerNorm without affine] (src/ops/basic.py:6-23).

    Inside GMP the parameters are consumed by the fused kernels; `forward` is only reached when the
    module is used standalone (the caller-side encoder/decoder of src/models/model.py:20-22, which is
    outside the processor path — SURVEY.md §8f rank 1) and then runs the torch layers as declared.
    """

    def __init__(self, input_dim, latent_dim, output_dim, hidden_layers, layer_normalized=True):
        super().__init__()
        modules = []
        for l in range(hidden_layers):
            modules.append(nn.Linear(input_dim if l == 0 else latent_dim, latent_dim))
            modules.append(nn.ReLU())
        modules.append(nn.Linear(latent_dim, output_dim))
        if layer_normalized:
            modules.append(nn.LayerNorm(output_dim, elementwise_affine=False))
        self.seq = nn.Sequential(*modules)

    def linears(self):
        return [m for m in self.seq if isinstance(m, nn.Linear)]

    def forward(self, x):
        return self.seq(x)


def _weights_struct(params):
    w = _lib.GmpWeightsC()
    for l in range(4):
        w.w_edge[l] = params[l].data_ptr()
        w.b_edge[l] = params[4 + l].data_ptr()
        w.w_node[l] = params[8 + l].data_ptr()
        w.b_node[l] = params[12 + l].data_ptr()
    return w


class _GMPFunction(torch.autograd.Function):
    """One GMP block through bsms_gmp_forward / bsms_gmp_backward (backward recomputes; nothing but
    the inputs is kept between the two)."""

    @staticmethod
    def forward(ctx, x3, pos, skip3, level, mode, P, packed, *params):
        B, N, _ = x3.shape
        pos_batched = 1 if pos.dim() == 3 else 0
        params = tuple(p.detach().contiguous() for p in params)
        out = torch.empty_like(x3)
        nbytes = int(lib.bsms_gmp_workspace_bytes(B, N, level.n_edges, mode, 0))
        ws = _lib.workspace(nbytes, x3.device)
        w = _weights_struct(params)
        # node-level intermediates kept for backward (fused tcgen05 modes; nothing per-edge is kept)
        saved = None
        if mode in (_lib.MODE_BF16, _lib.MODE_FP16X3) and any(ctx.needs_input_grad):
            saved = torch.empty(int(lib.bsms_gmp_saved_bytes(B, N)), dtype=torch.uint8, device=x3.device)
        with torch.cuda.device(x3.device):
            check(lib.bsms_gmp_forward_packed(level.byref(), C.byref(w), ptr(packed), ptr(x3), ptr(pos), pos_batched, ptr(skip3),
                                              ptr(out), ptr(saved), B, P, mode, ptr(ws), ws.numel(), stream_ptr()))
        ctx.saved_nodes = saved
        ctx.save_for_backward(x3, pos, *params)
        ctx.level, ctx.mode, ctx.P, ctx.has_skip = level, mode, P, skip3 is not None
        return out

    @staticmethod
    def backward(ctx, g_out):
        x3, pos, *params = ctx.saved_tensors
        level, mode, P = ctx.level, ctx.mode, ctx.P
        B, N, _ = x3.shape
        g_out = g_out.contiguous()
        g_x = torch.empty_like(x3)
        # one zero-filled flat buffer for the 16 parameter gradients (the kernels accumulate into them)
        flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=x3.device)
        grads = [v.view_as(p) for v, p in zip(flat.split([p.numel() for p in params]), params)]
        nbytes = int(lib.bsms_gmp_workspace_bytes(B, N, level.n_edges, mode, 1))
        ws = _lib.workspace(nbytes, x3.device)
        w, gw = _weights_struct(params), _weights_struct(grads)
        with torch.cuda.device(x3.device):
            check(lib.bsms_gmp_backward(level.byref(), C.byref(w), ptr(x3), ptr(pos), 1 if pos.dim() == 3 else 0,
                                        ptr(ctx.saved_nodes), ptr(g_out), ptr(g_x), C.byref(gw), B, P, mode, ptr(ws),
                                        ws.numel(), stream_ptr()))
        ctx.saved_nodes = None
        return (g_x, None, g_out if ctx.has_skip else None, None, None, None, None, *grads)


class GMP(nn.Module):
    """Graph message passing block (src/ops/basic.py:26-98)."""

    def __init__(self, latent_dim, hidden_layer, pos_dim, mode=None):
        super().__init__()
        if latent_dim != LATENT or hidden_layer != 3:
            raise BsmsError(f"the sm_100a kernels are built for latent_dim=128, hidden_layer=3 "
                            f"(configs/model/*.yaml); got {latent_dim}, {hidden_layer}")
        if not 1 <= pos_dim <= 3:
            raise BsmsError(f"pos_dim must be 1..3, got {pos_dim}")
        self.mlp_node = MLP(2 * latent_dim, latent_dim, latent_dim, hidden_layer)
        self.mlp_edge = MLP(2 * latent_dim + pos_dim + 1, latent_dim, latent_dim, hidden_layer)
        self.pos_dim = pos_dim
        self.mode = _lib.MODES[mode or _DEFAULT_MODE]

    def _params(self):
        le, ln = self.mlp_edge.linears(), self.mlp_node.linears()
        return [m.weight for m in le] + [m.bias for m in le] + [m.weight for m in ln] + [m.bias for m in ln]

    def _run(self, x, level, pos, skip=None):
        x3 = _as_b3(x, "x")
        if x3.shape[-1] != LATENT:
            raise BsmsError(f"x must have {LATENT} channels, got {x3.shape[-1]}")
        if pos.dim() not in (2, 3):
            raise NotImplementedError("Only implemented for dim 2 and 3")
        _lib.require_cuda(pos)
        if pos.dim() == 3 and x.dim() == 2:
            raise RuntimeError("batched pos with un-batched x: the reference's torch.cat fails on this too "
                               "(src/ops/basic.py:90)")
        if pos.shape[-1] != self.pos_dim or pos.shape[-2] != x3.shape[1]:
            raise RuntimeError(f"pos must be [..., {x3.shape[1]}, {self.pos_dim}], got {tuple(pos.shape)}")
        if pos.dim() == 3 and pos.shape[0] != x3.shape[0]:
            raise RuntimeError("pos and x disagree on the batch size")
        pos = pos.detach().to(torch.float32).contiguous()
        skip3 = None if skip is None else _as_b3(skip, "skip")
        if lib.bsms_get_deterministic() and self.mode == _lib.MODE_FP16X3:
            # the fp32-parity mode's deterministic form is the exact-fp32 mode (same 1e-5 grade, ordered reductions)
            out = _GMPFunction.apply(x3, pos, skip3, level, _lib.MODE_FP32, self.pos_dim, None, *self._params())
        else:
            out = _GMPFunction.apply(x3, pos, skip3, level, self.mode, self.pos_dim, self._packed_weights(), *self._params())
        return out if x.dim() == 3 else out.squeeze(0)

    def _packed_weights(self):
        """Inference only (grad disabled, tensor-core modes): the 16-bit operand images of this block's weights,
        packed once and reused until a parameter changes (its version counter moves) — the reference's rollout
        calls the same blocks 599 times per trajectory (src/utils/rollout_utils.py:48-62)."""
        if torch.is_grad_enabled() or self.mode == _lib.MODE_FP32:
            return None
        params = self._params()
        key = (self.mode, _lib.WEIGHTS_EPOCH[0], tuple((p.data_ptr(), p._version) for p in params))
        hit = getattr(self, "_pack_cache", None)
        if hit is None or hit[0] != key:
            dev = params[0].device
            packed = torch.empty(int(lib.bsms_gmp_packed_bytes()), dtype=torch.uint8, device=dev)
            w = _weights_struct([p.detach().contiguous() for p in params])
            with torch.cuda.device(dev):
                check(lib.bsms_gmp_pack(C.byref(w), self.pos_dim, self.mode, ptr(packed), stream_ptr()))
            hit = (key, packed)
            self._pack_cache = hit
        return hit[1]

    def forward(self, x, g, pos):
        if x.dim() not in (2, 3):
            raise NotImplementedError("Only implemented for dim 2 and 3")
        return self._run(x, _plan.level_plan(g, x.shape[-2]), pos)


_EW_PERM = _plan._Cache(16)   # (ew identity, plan) -> ew in the plan's two edge orders
_IDS_OK = _plan._Cache(16)    # Unpool index tensors whose range has been checked (identity-keyed)


class _ConvFunction(torch.autograd.Function):
    """out = conv(x) in one direction; its adjoint (the other direction) is the backward."""

    @staticmethod
    def forward(ctx, x3, level, ew_d, ew_s, up):
        B, N, Cc = x3.shape
        out = torch.empty_like(x3)
        with torch.cuda.device(x3.device):
            check(lib.bsms_edge_conv(level.byref(), ptr(ew_s if up else ew_d), ptr(x3), ptr(out), B, Cc, int(up),
                                     stream_ptr()))
        ctx.level, ctx.ew_d, ctx.ew_s, ctx.up = level, ew_d, ew_s, up
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        B, N, Cc = g.shape
        gx = torch.empty_like(g)
        up = not ctx.up
        with torch.cuda.device(g.device):
            check(lib.bsms_edge_conv(ctx.level.byref(), ptr(ctx.ew_s if up else ctx.ew_d), ptr(g), ptr(gx), B, Cc,
                                     int(up), stream_ptr()))
        return gx, None, None, None, None


class _RestrictFunction(torch.autograd.Function):
    """conv_down(x)[:, ids] fused (src/ops/BSMS.py:74,79-82); backward = prolongation kernel."""

    @staticmethod
    def forward(ctx, x3, hp, l):
        B, N, Cc = x3.shape
        nk = hp.n[l + 1]
        out = torch.empty(B, nk, Cc, dtype=x3.dtype, device=x3.device)
        with torch.cuda.device(x3.device):
            check(lib.bsms_conv_down_pool(hp.levels[l].byref(), ptr(hp.ew_d[l]), ptr(hp.ids[l]), nk, ptr(x3), ptr(out),
                                          B, Cc, stream_ptr()))
        ctx.hp, ctx.l = hp, l
        return out

    @staticmethod
    def backward(ctx, g):
        hp, l = ctx.hp, ctx.l
        g = g.contiguous()
        B, nk, Cc = g.shape
        gx = torch.empty(B, hp.n[l], Cc, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            check(lib.bsms_unpool_conv_up(hp.levels[l].byref(), ptr(hp.ew_s[l]), ptr(hp.inv[l]), nk, ptr(g), ptr(gx), B,
                                          Cc, stream_ptr()))
        return gx, None, None


class _ProlongFunction(torch.autograd.Function):
    """conv_up(unpool(hc)) fused (src/ops/BSMS.py:98-100); backward = restriction kernel."""

    @staticmethod
    def forward(ctx, hc3, hp, l):
        B, nk, Cc = hc3.shape
        out = torch.empty(B, hp.n[l], Cc, dtype=hc3.dtype, device=hc3.device)
        with torch.cuda.device(hc3.device):
            check(lib.bsms_unpool_conv_up(hp.levels[l].byref(), ptr(hp.ew_s[l]), ptr(hp.inv[l]), nk, ptr(hc3), ptr(out),
                                          B, Cc, stream_ptr()))
        ctx.hp, ctx.l = hp, l
        return out

    @staticmethod
    def backward(ctx, g):
        hp, l = ctx.hp, ctx.l
        g = g.contiguous()
        B, N, Cc = g.shape
        nk = hp.n[l + 1]
        gh = torch.empty(B, nk, Cc, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            check(lib.bsms_conv_down_pool(hp.levels[l].byref(), ptr(hp.ew_d[l]), ptr(hp.ids[l]), nk, ptr(g), ptr(gh), B,
                                          Cc, stream_ptr()))
        return gh, None, None


class WeightedEdgeConv(nn.Module):
    """Weighted edge convolution for the transfer between levels (src/ops/basic.py:101-167)."""

    def __init__(self, *args):
        super().__init__()

    def forward(self, x, g, ew, aggragating=True):
        x3 = _as_b3(x, "x")
        level = _plan.level_plan(g, x3.shape[1])
        _lib.require_cuda(ew)
        if ew.numel() != level.n_edges:
            raise RuntimeError(f"ew has {ew.numel()} entries for {level.n_edges} edges")
        # the weights in the plan's two edge orders are cached per (ew tensor, plan): BSMS.py:74-75,100 call this
        # three times per level with the same `ew`
        key = (_plan._ident(ew), id(level))
        hit = _EW_PERM.get(key)
        if hit is None:
            ew32 = ew.detach().to(torch.float32).contiguous()
            E = max(level.n_edges, 1)
            ew_d = torch.empty(E, dtype=torch.float32, device=x3.device)
            ew_s = torch.empty(E, dtype=torch.float32, device=x3.device)
            with torch.cuda.device(x3.device):
                check(lib.bsms_permute_ew(level.byref(), ptr(ew32), ptr(ew_d), ptr(ew_s), stream_ptr()))
            hit = (ew_d, ew_s, ew, level)  # ew / level kept alive: neither identity can be recycled
            _EW_PERM.put(key, hit)
        out = _ConvFunction.apply(x3, level, hit[0], hit[1], not aggragating)
        return out if x.dim() == 3 else out.squeeze(0)

    @torch.no_grad()
    def cal_ew(self, w, g):
        _lib.require_cuda(w, g)
        w1 = w.squeeze(-1) if w.dim() > 1 else w
        level = _plan.level_plan(g, w1.shape[0])
        with torch.cuda.device(w.device):
            ew, _, _, aggr_w = _plan.cal_ew_raw(level, w1.to(torch.float32).contiguous(), True)
        return ew, aggr_w


class _UnpoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h3, ids32, n_rows):
        B, nk, Cc = h3.shape
        out = torch.empty(B, n_rows, Cc, dtype=h3.dtype, device=h3.device)
        with torch.cuda.device(h3.device):
            check(lib.bsms_unpool_rows(ptr(h3), ptr(ids32), nk, n_rows, ptr(out), B, Cc, stream_ptr()))
        ctx.ids32, ctx.n_rows = ids32, n_rows
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        B, n_rows, Cc = g.shape
        nk = ctx.ids32.numel()
        gh = torch.empty(B, nk, Cc, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            check(lib.bsms_gather_rows(ptr(g), ptr(ctx.ids32), nk, n_rows, ptr(gh), B, Cc, stream_ptr()))
        return gh, None, None


class Unpool(nn.Module):
    """Zero-fill + row injection (src/ops/basic.py:170-201)."""

    def __init__(self, *args):
        super().__init__()

    def forward(self, h, pre_node_num, idx):
        if h.dim() not in (2, 3):
            return None  # the reference falls through both branches and returns an unbound name
        h3 = _as_b3(h, "h")
        _lib.require_cuda(idx)
        # range check + int32 copy once per index tensor (one host sync the first time it is seen, none after)
        key = (_plan._ident(idx), int(pre_node_num))
        hit = _IDS_OK.get(key)
        if hit is None:
            if idx.numel():
                lo, hi = torch.stack([idx.min(), idx.max()]).tolist()
                if lo < 0 or hi >= pre_node_num:
                    raise IndexError("index out of range in Unpool")
            hit = (idx.to(torch.int32).contiguous(), idx)
            _IDS_OK.put(key, hit)
        out = _UnpoolFunction.apply(h3, hit[0], int(pre_node_num))
        return out if h.dim() == 3 else out.squeeze(0)


class BSGMP(nn.Module):
    """Bi-stride multi-scale message passing U-Net (src/ops/BSMS.py:8-104)."""

    def __init__(self, unet_depth, latent_dim, hidden_layer, pos_dim, mode=None):
        super().__init__()
        self.bottom_gmp = GMP(latent_dim, hidden_layer, pos_dim, mode)
        self.down_gmps = nn.ModuleList()
        self.up_gmps = nn.ModuleList()
        self.unpools = nn.ModuleList()
        self.unet_depth = unet_depth
        self.edge_conv = WeightedEdgeConv()
        for _ in range(self.unet_depth):
            self.down_gmps.append(GMP(latent_dim, hidden_layer, pos_dim, mode))
            self.up_gmps.append(GMP(latent_dim, hidden_layer, pos_dim, mode))
            self.unpools.append(Unpool())

    def set_mode(self, mode: str):
        for m in self.modules():
            if isinstance(m, GMP):
                m.mode = _lib.MODES[mode]
        return self

    def bind_mesh(self, m_gs, m_ids, n0=None):
        """Pin one hierarchy: later forwards reuse its plan without looking at the index tensors they are
        handed (only the level sizes are compared).  For callers whose mesh is fixed for the whole run — the
        reference's consistent-mesh training and its rollout loop (src/utils/rollout_utils.py:48-62) — this
        removes even the per-step content fingerprint.  `unbind_mesh()` returns to content-keyed look-ups."""
        d = self.unet_depth
        _lib.require_cuda(*m_gs[:d + 1], *m_ids[:d])
        if n0 is None:
            n0 = int(m_gs[0].max()) + 1 if m_gs[0].numel() else 1
        self._bound = _plan.hierarchy_plan(list(m_gs[:d + 1]), list(m_ids[:d]), int(n0))
        return self

    def unbind_mesh(self):
        self._bound = None
        return self

    def _hierarchy(self, m_gs, m_ids, n0):
        d = self.unet_depth
        hp = getattr(self, "_bound", None)
        if hp is not None:
            sizes_ok = (hp.n[0] == n0 and all(int(m_ids[l].shape[0]) == hp.n[l + 1] for l in range(d))
                        and all(int(m_gs[l].shape[-1]) == hp.levels[l].n_edges for l in range(d + 1)))
            if not sizes_ok:
                raise BsmsError("bind_mesh: the hierarchy handed to forward has different level sizes than the bound one; "
                                "call bind_mesh again or unbind_mesh()")
            return hp
        return _plan.hierarchy_plan(list(m_gs[:d + 1]), list(m_ids[:d]), n0)

    def forward(self, h, m_ids, m_gs, pos):
        d = self.unet_depth
        if h.dim() not in (2, 3) or pos.dim() not in (2, 3):
            raise NotImplementedError("Only implemented for dim 2 and 3")
        _lib.require_cuda(h, pos, *m_gs[:d + 1], *m_ids[:d])
        if len(m_ids) < d or len(m_gs) < d + 1:
            raise IndexError("list index out of range")  # what the reference's m_gs[i] / m_ids[i] raise
        hp = self._hierarchy(m_gs, m_ids, h.shape[-2])
        squeeze = h.dim() == 2
        x = _as_b3(h, "h")
        p = pos.detach().to(torch.float32)
        p3 = (p if p.dim() == 3 else p.unsqueeze(0)).contiguous()  # conv kernels take [B', N, P]
        pos_is_batched = pos.dim() == 3
        if pos_is_batched and squeeze:
            raise RuntimeError("batched pos with un-batched h: the reference's torch.cat fails on this too")
        down_outs, down_ps = [], []
        for l in range(d):
            x = self.down_gmps[l]._run(x, hp.levels[l], p3 if pos_is_batched else p3[0])
            down_outs.append(x)
            down_ps.append(p3)
            x = _RestrictFunction.apply(x, hp, l)
            with torch.no_grad():
                p3 = _RestrictFunction.apply(p3, hp, l)
        x = self.bottom_gmp._run(x, hp.levels[d], p3 if pos_is_batched else p3[0])
        for k in range(d):
            l = d - 1 - k
            x = _ProlongFunction.apply(x, hp, l)
            pl = down_ps[l]
            x = self.up_gmps[k]._run(x, hp.levels[l], pl if pos_is_batched else pl[0], skip=down_outs[l])
        return x.squeeze(0) if squeeze else x
