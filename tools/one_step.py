#!/usr/bin/env python
"""One fwd+bwd step of the metric workload between cudaProfilerStart / Stop, for
`ncu --profile-from-start off ... python tools/one_step.py` (launch lists and --set full captures of exactly one
warm step).  Same workload as bench.py: airfoil-like 72x72, depth 6, B = 48, bf16."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import build_workload  # noqa: E402
from bsms_gnn_b200.ops import BSGMP  # noqa: E402
from oracle import bsms_oracle as O  # noqa: E402

mode = os.environ.get("BSMS_MODE", "bf16")
B = int(os.environ.get("BSMS_BATCH", "48"))
dev = torch.device("cuda", 0)
pos, m_gs, m_ids = build_workload(72, 6)
model = BSGMP(6, 128, 3, 2, mode=mode).to(dev)
model.load_state_dict(O.init_params(6, pos_dim=2, seed=0))
gs = [torch.from_numpy(g).to(dev) for g in m_gs]
ids = [torch.from_numpy(i).to(dev) for i in m_ids]
gen = torch.Generator().manual_seed(1234)
h = torch.randn(B, pos.shape[0], 128, generator=gen).to(dev).requires_grad_(True)
p = (torch.from_numpy(pos).unsqueeze(0) + 0.01 * torch.randn(B, pos.shape[0], 2, generator=gen)).to(dev)


def step():
    model.zero_grad(set_to_none=True)
    h.grad = None
    model(h, ids, gs, p).square().mean().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
