"""Real multi-process check of the node-partitioned processor on 2 GPUs (skipped on a 1-GPU box): one process per
GPU under torchrun, ghosts over (a) the K6 push kernel on CUDA-IPC peer memory and (b) NCCL point-to-point;
forward 1e-5 and all-reduced parameter gradients 5e-4 against the un-partitioned module (tests/run_partitioned_dist.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange,mode", [("push", "fp32"), ("nccl", "fp32"), ("push", "bf16")])
def test_partitioned_two_processes(exchange, mode):
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "run_partitioned_dist.py"), "--exchange", exchange, "--mode", mode,
           "--steps", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("partitioned")]
    print("\n".join(lines))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dist2_parity.log"), "a") as f:
        f.write("\n".join(lines) + "\n")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert lines
