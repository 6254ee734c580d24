"""Whole `_forward` of the reference model and its rollout loop, fused around the processor (inference).

`FusedSimulator(model)` wraps a `BSMS_Simulator` (the reference's own class, src/models/model.py, built on
`bsms_gnn_b200.ops` — or anything with the same attributes: `encode`, `process`, `decode`,
`_inputNormalizer`, `_targetNormalizer`, `pos_dim`) and evaluates model.py:127-164 as

    bsms_encode_in      split + input normalisation + first encoder Linear + ReLU, positions     (1 launch)
    bsms_dense128_stack encoder layers 1..3 + LayerNorm                                           (3-4 launches)
    BSGMP.forward       the processor (packed weights cached, static-mesh positions restricted once)
    bsms_dense128_stack decoder layers 0..2                                                       (3 launches)
    bsms_decode_out     last decoder Linear + inverse target normalisation + mask + residual,
                        and the next rollout input with the boundary re-imposed (rollout_utils.py:57-62)  (1 launch)

with no host synchronisation, so `GraphedRollout` captures one step into a CUDA graph and replays it T times
with the feedback (prediction -> next input) staying on the device.  The parameters are the model's own
tensors (state_dict unchanged); the normaliser statistics are read once at construction (`refresh()` re-reads).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def _darr(vals):
    return (C.c_double * len(vals))(*[float(v) for v in vals])


class FusedSimulator:
    def __init__(self, model, mode=None, pos_feedback=False):
        self.model = model
        self.pos_feedback = bool(pos_feedback)  # deforming meshes: the next input's mesh positions are the new state
        self.process = model.process
        if mode is not None:
            self.process.set_mode(mode)
        self.mode = self.process.bottom_gmp.mode
        self.P = int(model.pos_dim)
        enc = [m for m in model.encode.seq if isinstance(m, torch.nn.Linear)]
        dec = [m for m in model.decode.seq if isinstance(m, torch.nn.Linear)]
        if len(enc) != 4 or len(dec) != 4 or enc[1].weight.shape != (128, 128):
            raise _lib.BsmsError("FusedSimulator is built for hidden_layer=3, latent_dim=128 (configs/model/*.yaml)")
        self.enc, self.dec = enc, dec
        self.C = int(dec[3].weight.shape[0])
        if enc[0].weight.shape[1] != self.C + 1:
            raise _lib.BsmsError("encoder input width must be out_dim + 1 (model.py:20)")
        self.Cin = self.C + self.P + 1
        self._packs = {}
        self._pos_cache = None
        self.refresh()

    def refresh(self):
        """Re-read the normaliser statistics (fp64 -> host doubles; one synchronisation)."""
        m = self.model
        with torch.no_grad():
            self.in_mean = _darr(m._inputNormalizer.mean().double().cpu().tolist())
            self.in_std = _darr(m._inputNormalizer.std_with_epsilon().double().cpu().tolist())
            self.out_mean = _darr(m._targetNormalizer.mean().double().cpu().tolist())
            self.out_std = _darr(m._targetNormalizer.std_with_epsilon().double().cpu().tolist())

    def _packed(self, name, layers):
        if self.mode == _lib.MODE_FP32:
            return None
        key = (self.mode, _lib.WEIGHTS_EPOCH[0], tuple((l.weight.data_ptr(), l.weight._version) for l in layers))
        hit = self._packs.get(name)
        if hit is None or hit[0] != key:
            dev = layers[0].weight.device
            buf = torch.empty(int(lib.bsms_dense128_packed_bytes(self.mode)), dtype=torch.uint8, device=dev)
            W = (C.c_void_p * len(layers))(*[l.weight.data_ptr() for l in layers])
            with torch.cuda.device(dev):
                check(lib.bsms_dense128_pack(W, len(layers), self.mode, ptr(buf), stream_ptr()))
            hit = (key, buf)
            self._packs[name] = hit
        return hit[1]

    def _stack(self, x, layers, relu_mask, layer_norm, name):
        rows = x.shape[0] * x.shape[1]
        out = torch.empty_like(x)
        scratch = _lib.workspace(2 * rows * 128 * 4 + 256, x.device)
        W = (C.c_void_p * len(layers))(*[l.weight.data_ptr() for l in layers])
        b = (C.c_void_p * len(layers))(*[l.bias.data_ptr() for l in layers])
        with torch.cuda.device(x.device):
            check(lib.bsms_dense128_stack(ptr(x), rows, W, b, len(layers), relu_mask, int(layer_norm), self.mode,
                                          ptr(self._packed(name, layers)), ptr(out), ptr(scratch), stream_ptr()))
        return out

    @torch.no_grad()
    def forward(self, node_in, node_mask, m_gs, m_ids, ic=None, want_next=False):
        """model.py:127-164 for node_in [B,N,C+P+1], node_mask [B,N,1] -> pred [B,N,C]
        (+ next_in [B,N,C+P+1] = where(mask == 0, ic, cat[pred, pos, type]) when want_next)."""
        _lib.require_cuda(node_in, node_mask)
        if node_in.dim() != 3 or node_in.shape[-1] != self.Cin:
            raise _lib.BsmsError(f"node_in must be [B, N, {self.Cin}], got {tuple(node_in.shape)}")
        x = node_in.contiguous().float()
        B, N, _ = x.shape
        rows = B * N
        dev = x.device
        mask = node_mask.to(torch.float32).expand(B, N, 1).contiguous()
        a1 = torch.empty(B, N, 128, dtype=torch.float32, device=dev)
        pos = torch.empty(B, N, self.P, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.bsms_encode_in(ptr(x), rows, self.Cin, self.C, self.P, self.in_mean, self.in_std, ptr(self.enc[0].weight),
                                     ptr(self.enc[0].bias), ptr(a1), ptr(pos), stream_ptr()))
        h = self._stack(a1, self.enc[1:], 0b011, True, "enc")
        h = self.process(h, m_ids, m_gs, pos)
        y = self._stack(h, self.dec[:3], 0b111, False, "dec")
        pred = torch.empty(B, N, self.C, dtype=torch.float32, device=dev)
        nxt = torch.empty_like(x) if want_next else None
        icc = None if ic is None else ic.contiguous().float()
        with torch.cuda.device(dev):
            check(lib.bsms_decode_out(ptr(y), rows, self.Cin, self.C, ptr(self.dec[3].weight), ptr(self.dec[3].bias),
                                      self.out_mean, self.out_std, ptr(x), ptr(mask), ptr(icc), ptr(pred), ptr(nxt), int(self.pos_feedback),
                                      stream_ptr()))
        return (pred, nxt) if want_next else pred

    __call__ = forward

    @torch.no_grad()
    def rollout(self, ic, node_mask, m_gs, m_ids, steps, results=None):
        """rollout_one_traj (src/utils/rollout_utils.py:15-64), eagerly: results [steps, N, C]."""
        cur = ic.contiguous().float()
        if results is None:
            results = torch.empty(steps, ic.shape[1], self.C, dtype=torch.float32, device=ic.device)
        for t in range(steps):
            pred, cur = self.forward(cur, node_mask, m_gs, m_ids, ic=ic, want_next=True)
            results[t].copy_(pred[0])
        return results


class GraphedRollout:
    """One rollout step captured into a CUDA graph; the feedback stays on the device.

    step(): replays the graph — reads the current input buffer, writes the prediction and the next input, and
    copies the next input back into the current-input buffer inside the graph.  `run(T)` replays T steps and
    gathers the predictions ([T, N, C], device) with one asynchronous copy per step outside the graph."""

    def __init__(self, sim: FusedSimulator, ic, node_mask, m_gs, m_ids, warmup: int = 2):
        self.sim, self.m_gs, self.m_ids = sim, m_gs, m_ids
        self.ic = ic.detach().clone().contiguous().float()
        self.mask = node_mask.detach().clone()
        self.cur = self.ic.clone()
        side = torch.cuda.Stream(device=self.ic.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # plans, packed weights, workspaces
                sim.forward(self.cur, self.mask, m_gs, m_ids, ic=self.ic, want_next=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.pred, nxt = sim.forward(self.cur, self.mask, m_gs, m_ids, ic=self.ic, want_next=True)
            self.cur.copy_(nxt)

    def reset(self, ic=None):
        if ic is not None:
            self.ic.copy_(ic)
        self.cur.copy_(self.ic)

    def step(self):
        self.graph.replay()
        return self.pred

    def run(self, steps, results=None, host=None):
        """-> results [steps, N, C] on the device; `host` (pinned [steps, N, C]) also receives every step."""
        if results is None:
            results = torch.empty(steps, self.ic.shape[1], self.sim.C, dtype=torch.float32, device=self.ic.device)
        for t in range(steps):
            self.graph.replay()
            results[t].copy_(self.pred[0], non_blocking=True)
            if host is not None:
                host[t].copy_(self.pred[0], non_blocking=True)
        return results
