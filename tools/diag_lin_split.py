#!/usr/bin/env python
"""Standalone check of the split-operand tensor-core layer (bsms_debug_lin_split) against fp64 for several row counts."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bsms_gnn_b200 import _lib  # noqa: E402
from bsms_gnn_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402

dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)
W = (torch.rand(128, 128, generator=gen) * 2 - 1).mul(0.088).to(dev)
scratch = torch.zeros(65536 + 64, dtype=torch.uint8, device=dev)
for rows in [16, 128, 940, 1936, 6884, 10156, 14400, 18064, 18063, 18176, 22532, 40000]:
    for b_mn in (0, 1):
        for grad in (0, 1):
            X = torch.randn(rows, 128, generator=gen) * (1e-4 if grad else 1.0)
            X[::7] *= 30.0
            M = torch.randn(rows, 128, generator=gen).relu()
            Xd, Md = X.to(dev), M.to(dev)
            Y = torch.full((rows, 128), float("nan"), device=dev)
            check(lib.bsms_debug_lin_split(ptr(Xd), rows, ptr(W), b_mn, ptr(Md), grad, ptr(Y), ptr(scratch), stream_ptr()))
            torch.cuda.synchronize()
            Wd = W.double().cpu()
            ref = (X.double() @ (Wd if b_mn else Wd.T)) * (M.double() > 0)
            err = float((Y.cpu().double() - ref).abs().max() / ref.abs().max())
            rowerr = (Y.cpu().double() - ref).abs().amax(1) / ref.abs().max()
            bad = torch.nonzero(rowerr > 1e-5).flatten()
            print(f"rows {rows:6d} b_mn {b_mn} grad {grad}: max-rel {err:.2e} bad rows {bad.numel()} {bad[:6].tolist()} .. {bad[-3:].tolist() if bad.numel() else ''}", flush=True)
