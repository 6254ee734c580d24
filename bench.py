#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: M-edges/s per BSMS fwd+bwd step, airfoil L=6, latent 128.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fp32|fp16x3|bf16] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input: BSGMP forward + backward
(loss = mean(out^2)) on the airfoil-like mesh (72x72 jittered triangle grid: 5 184 nodes, 30 530
directed edges, unet_depth 6 -> 226 188 edge-MLP rows per sample) at the reference's training batch
of 48 samples sharing one mesh (configs/default.yaml:16).  `value` counts MESH edges (B * E_0) per
second, whole job; `edge_evals_per_s` is the same step in edge-MLP rows (B * (2*sum_{l<d} E_l + E_d)).

N > 1 (torchrun, one rank per GPU): every rank owns its own 48 samples (weak scaling, the batch
axis is embarrassingly parallel — what nn.DataParallel would have split, src/trainer/trainer.py:15-18)
and the parameter gradients are all-reduced over NCCL once per step.

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 128
F_EDGE_ROW = 164608  # SURVEY.md §8d algorithmic FLOPs (fwd) per edge row, P=2
F_NODE_ROW = 163840


def build_workload(nx, depth):
    from bsms_gnn_b200 import hierarchy, meshgen
    pos, cells = meshgen.tri_grid(nx, nx)
    fe = meshgen.cells_to_flat_edge(cells)
    m_gs, m_ids = hierarchy.build_hierarchy(fe, depth, pos.shape[0], pos)
    return pos, m_gs, m_ids


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            self.result = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                           "samples": len(sm)}


def cpu_reference_model(depth):
    """-> (step_fn_factory, kind): the reference's OWN `ops.BSGMP` module when its sources are present
    (/root/reference here, the byte-for-byte copy oracle/_ref on the GPU box — oracle/build_ref.py), else the
    oracle port.  Same deterministic weights as the B200 arm."""
    from oracle import bsms_oracle as O, ref_import
    params = O.init_params(depth, pos_dim=2, seed=0)
    if ref_import.available():
        ref = ref_import.load()
        model = ref.ops.BSGMP(depth, D, 3, 2)
        model.load_state_dict(params)

        def make_step(h, ids, gs, p):
            def step():
                model.zero_grad(set_to_none=True)
                h.grad = None
                model(h, ids, gs, p).square().mean().backward()
            return step
        return make_step, "reference"
    pr = {k: v.requires_grad_(True) for k, v in params.items()}

    def make_step(h, ids, gs, p):
        def step():
            for v in pr.values():
                v.grad = None
            h.grad = None
            O.bsgmp(h, ids, gs, p, pr, depth).square().mean().backward()
        return step
    return make_step, "port"


def cpu_reference_run(pos, m_gs, m_ids, depth, steps, warmup, b_cpu, budget_s):
    """The reference's implementation of the path on the host cores: fwd+bwd steps at batch b_cpu (batched
    positions, like the B200 arm).  Stops early when `budget_s` is spent (at least one timed step)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    make_step, kind = cpu_reference_model(depth)
    gs = [torch.from_numpy(g) for g in m_gs]
    ids = [torch.from_numpy(i) for i in m_ids]
    gen = torch.Generator().manual_seed(1234)
    h = torch.randn(b_cpu, pos.shape[0], D, generator=gen).requires_grad_(True)
    p = torch.from_numpy(pos).unsqueeze(0) + 0.01 * torch.randn(b_cpu, pos.shape[0], 2, generator=gen)
    step = make_step(h, ids, gs, p)
    t_all = time.perf_counter()
    done_w = 0
    for _ in range(warmup):
        step()
        done_w += 1
        if time.perf_counter() - t_all > 0.4 * budget_s:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    return times, cores, kind, done_w


def host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2**30
    except Exception:
        return 0.0


def run_mesh(args):
    """BASELINE.json config 5: one large synthetic mesh, B=1, fwd+bwd, node-partitioned over the ranks
    with halo exchanges (strong scaling: the mesh is fixed, N GPUs split it)."""
    import torch.distributed as dist
    from bsms_gnn_b200 import _lib, hierarchy, meshgen, partition
    from bsms_gnn_b200.dist import GradBucket
    from bsms_gnn_b200.ops import BSGMP
    from bsms_gnn_b200.partitioned import DistExchanger, PartitionedBSGMP, exchange_requests
    from oracle import bsms_oracle as O

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nx = args.nx if args.nx != 72 else 1414
    depth = args.depth
    cache = f"/tmp/bsms_mesh_{nx}_{depth}.npz"
    t0 = time.perf_counter()
    if rank == 0 and not os.path.exists(cache):
        pos, cells = meshgen.tri_grid(nx, nx)
        m_gs, m_ids = hierarchy.build_hierarchy(meshgen.cells_to_flat_edge(cells), depth, pos.shape[0], pos)
        np.savez(cache + ".tmp.npz", pos=pos, **{f"g{l}": g for l, g in enumerate(m_gs)},
                 **{f"i{l}": i for l, i in enumerate(m_ids)})
        os.replace(cache + ".tmp.npz", cache)
    if world > 1:
        dist.barrier()
    z = np.load(cache)
    pos = z["pos"]
    m_gs = [z[f"g{l}"] for l in range(depth + 1)]
    m_ids = [z[f"i{l}"] for l in range(depth)]
    n0, E0 = pos.shape[0], int(m_gs[0].shape[1])
    edge_rows = 2 * sum(int(g.shape[1]) for g in m_gs[:depth]) + int(m_gs[depth].shape[1])
    model = BSGMP(depth, D, 3, 2, mode=args.mode).to(dev)
    model.load_state_dict(O.init_params(depth, pos_dim=2, seed=0))
    params = list(model.parameters())
    gen = torch.Generator().manual_seed(7)
    h_all = torch.randn(n0, D, generator=gen)
    if world > 1:
        plan = exchange_requests(partition.build_rank_plan(m_gs, m_ids, n0, world, rank))
        pm = PartitionedBSGMP(model, [plan], DistExchanger(), dev)
        own = torch.from_numpy(plan.levels[0].nodes[:plan.levels[0].n_own])
        h_host = h_all[own].contiguous().pin_memory()
        p_dev = torch.from_numpy(pos)[own].to(dev)
        bucket = GradBucket(params)
        ghosts = [lv.n_local - lv.n_own for lv in pm.states[0].levels]
    else:
        h_host = h_all.pin_memory()
        p_dev = torch.from_numpy(pos).to(dev)
        gs = [torch.from_numpy(g).to(dev) for g in m_gs]
        ids = [torch.from_numpy(i).to(dev) for i in m_ids]
        ghosts = None
    setup_s = time.perf_counter() - t0
    h_dev = h_host.to(dev).requires_grad_(True)

    def step(h):
        for q in params:
            q.grad = None
        h.grad = None
        if world > 1:
            (out,) = pm([h], [p_dev])
            loss = out.square().sum() / (n0 * D)
            loss.backward()
            bucket.step_sync()
        else:
            out = model(h, ids, gs, p_dev)
            loss = out.square().mean()
            loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(h_dev)
    barrier()
    n0l = _lib.launch_count()
    step(h_dev)
    launches_per_step = _lib.launch_count() - n0l
    run = lambda: step(h_dev)
    if not args.no_graph:
        # the whole step (forward, backward, halo exchanges, gradient all-reduce) replayed from ONE CUDA graph
        from bsms_gnn_b200.graphed import GraphedStep
        run = GraphedStep(lambda: step(h_dev), warmup=1)
        for _ in range(2):
            run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as cs:
        e0.record()
        for _ in range(args.steps):
            run()
        e1.record()
        barrier()
    launches = launches_per_step * args.steps
    ms = e0.elapsed_time(e1) / args.steps
    # end to end: this step's features come from pinned host memory, the loss goes back to the host
    t1 = time.perf_counter()
    for _ in range(args.steps):
        with torch.no_grad():
            h_dev.copy_(h_host, non_blocking=True)
        _ = run().item()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t1) / args.steps
    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1]) / 1e3
    if rank == 0:
        clk = cs.result
        print(json.dumps({
            "metric": "M-edges/s per BSMS fwd+bwd step", "value": E0 / (ms * 1e-3) / 1e6, "unit": "M-edges/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16x3": "f32 (fp16x3 split MMA)", "bf16": "bf16 MMA operands, fp32 storage/accumulate"}[args.mode],
            "data": "synthetic",
            "config": {"workload": f"synthetic {nx}x{nx} tri-grid ({n0} nodes / {E0} directed edges), unet_depth {depth}, "
                                   f"latent 128, B=1, fwd+bwd", "mode": args.mode,
                       "parallelism": (f"node partition x{world}, {4 * depth + 1} halo exchanges per forward (NCCL p2p), "
                                       f"grad all-reduce") if world > 1 else "single GPU",
                       "execution": "eager" if args.no_graph else "whole step replayed from one CUDA graph",
                       "rank0_ghost_rows_per_level": ghosts, "setup_s": setup_s,
                       "l2": "per-step working set far exceeds the 126 MB L2"},
            "edge_evals_per_s": edge_rows / (ms * 1e-3),
            "clocks": {"sm_mhz": clk.get("sm_mhz"), "sm_max_mhz": clk.get("sm_max_mhz"), "reasons": clk.get("reasons")},
            "e2e": {"value": E0 / e2e_s / 1e6, "unit": "M-edges/s", "h2d_bytes_per_step": int(h_host.numel() * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches)}), flush=True)
    if world > 1:
        if not args.no_graph:
            # a process group whose NCCL work was captured into a live CUDA graph does not tear down cleanly:
            # drop the graph, drain the device and leave without the collective destructor
            del run
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


def run_rollout(args):
    """BASELINE.json config 2 (CylinderFlow-like 44x44, unet_depth 5, B=1): T sequential processor
    forwards with feedback, the reference's rollout pattern (src/utils/rollout_utils.py:48-62).
    Reports ms per forward eagerly and replayed from a CUDA graph."""
    from bsms_gnn_b200 import _lib
    from bsms_gnn_b200.graphed import GraphedBSGMP
    from bsms_gnn_b200.ops import BSGMP
    from oracle import bsms_oracle as O

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    nx = args.nx if args.nx != 72 else 44
    depth = args.depth if args.nx != 72 else 5
    pos, m_gs, m_ids = build_workload(nx, depth)
    E0 = int(m_gs[0].shape[1])
    edge_rows = 2 * sum(int(g.shape[1]) for g in m_gs[:depth]) + int(m_gs[depth].shape[1])
    mode = args.mode
    model = BSGMP(depth, D, 3, 2, mode=mode).to(dev)
    model.load_state_dict(O.init_params(depth, pos_dim=2, seed=0))
    gs = [torch.from_numpy(g).to(dev) for g in m_gs]
    ids = [torch.from_numpy(i).to(dev) for i in m_ids]
    p = torch.from_numpy(pos).to(dev)
    h0 = torch.randn(pos.shape[0], D, generator=torch.Generator().manual_seed(3)).to(dev)
    T = max(args.steps, 10) if args.steps != 10 else 599
    graphed = GraphedBSGMP(model, ids, gs, h0, p)

    def timed(fn):
        x = h0
        for _ in range(args.warmup):
            x = fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        x = h0
        for _ in range(T):
            x = fn(x)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / T, _lib.launch_count() - n0

    with torch.no_grad():
        ms_eager, launches = timed(lambda x: model(x, ids, gs, p))
        ms_graph, _ = timed(lambda x: graphed(x))
    print(json.dumps({
        "metric": "M-edges/s per BSMS forward (rollout)", "value": E0 / (ms_graph * 1e-3) / 1e6, "unit": "M-edges/s",
        "n_gpus": 1, "steps": T, "warmup": args.warmup, "ms_per_step": ms_graph, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": mode, "data": "synthetic",
        "config": {"workload": f"cylinder-like {nx}x{nx} tri-grid ({pos.shape[0]} nodes / {E0} directed edges), "
                               f"unet_depth {depth}, B=1, forward rollout of {T} sequential steps with feedback",
                   "mode": mode, "execution": "CUDA graph replay"},
        "ms_per_forward_eager": ms_eager, "ms_per_forward_graph": ms_graph,
        "edge_evals_per_s": edge_rows / (ms_graph * 1e-3), "gpu_launches": int(launches),
        "kernels_per_forward": launches / T}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("BSMS_MODE", "bf16"), choices=["fp32", "fp16x3", "bf16"],
                    help="bf16 (default) is the precision BASELINE.json's metric config names; fp16x3 / fp32 are the "
                         "fp32-parity modes")
    ap.add_argument("--batch", type=int, default=48)
    ap.add_argument("--nx", type=int, default=72)
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-self-check", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="mesh workload: run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--workload", default="airfoil", choices=["airfoil", "mesh", "rollout"],
                    help="airfoil: BASELINE.json's metric config (batch-parallel over GPUs); mesh: one large "
                         "node-partitioned mesh, B=1 (config 5: --nx 1414 = 2.0 M nodes / 12.0 M edges)")
    args = ap.parse_args()
    if args.workload == "mesh":
        return run_mesh(args)
    if args.workload == "rollout":
        return run_rollout(args)
    assert args.warmup >= 3 or args.impl == "reference", "timing rules: at least 3 warm-up steps"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    pos, m_gs, m_ids = build_workload(args.nx, args.depth)
    E0 = int(m_gs[0].shape[1])
    edge_rows = 2 * sum(int(g.shape[1]) for g in m_gs[:args.depth]) + int(m_gs[args.depth].shape[1])
    node_rows = 2 * sum([pos.shape[0]] + [len(i) for i in m_ids[:-1]]) + len(m_ids[-1])
    workload = (f"airfoil-like {args.nx}x{args.nx} tri-grid ({pos.shape[0]} nodes / {E0} directed edges), "
                f"unet_depth {args.depth}, latent 128, fwd+bwd")

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's CPU implementation at the SAME batch as the B200 arm when the host has the memory for
        # it (the reference materialises ~3.8 KB per edge row for fwd+bwd: ~41 GB at B = 48), preceded by a scan
        # over smaller batches so that the per-sample cost is measured, not assumed
        scan = {}
        for b in (1, 8):
            if b >= args.batch:
                continue
            tm, cores, kind, _ = cpu_reference_run(pos, m_gs, m_ids, args.depth, 3, 2, b, 30.0)
            scan[b] = {"ms_per_step": 1e3 * sum(tm) / len(tm), "ms_per_sample": 1e3 * sum(tm) / len(tm) / b, "steps": len(tm)}
        need_gb = 3.8e3 * args.batch * edge_rows / 2**30 * 1.3
        b_cpu = args.batch if host_ram_gb() > need_gb else max(scan) if scan else 1
        times, cores, kind, w_done = cpu_reference_run(pos, m_gs, m_ids, args.depth, args.steps, args.warmup, b_cpu, 150.0)
        t = sum(times) / len(times)
        val = b_cpu * E0 / t / 1e6
        scan[b_cpu] = {"ms_per_step": t * 1e3, "ms_per_sample": t * 1e3 / b_cpu, "steps": len(times)}
        sample = (f"batch {b_cpu} of {args.batch} (same mesh, same weights, batched positions), {len(times)} timed fwd+bwd steps after "
                  f"{w_done} warm-up (150 s budget), {cores} threads")
        print(json.dumps({
            "impl": "reference", "metric": "M-edges/s per BSMS fwd+bwd step", "value": val, "unit": "M-edges/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": w_done, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": args.batch, "batch_timed": b_cpu},
            "edge_evals_per_s": b_cpu * edge_rows / t, "batch_scan": scan,
            "cpu_baseline": {"value": val, "unit": "M-edges/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "M-edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch.distributed as dist
    from bsms_gnn_b200 import _lib
    from bsms_gnn_b200.ops import BSGMP
    from oracle import bsms_oracle as O  # parameter init only (deterministic weights of the named architecture)

    assert torch.cuda.is_available(), "bench.py measures the CUDA path; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    model = BSGMP(args.depth, D, 3, 2, mode=args.mode).to(dev)
    model.load_state_dict(O.init_params(args.depth, pos_dim=2, seed=0))
    params = [p for p in model.parameters()]
    gs = [torch.from_numpy(g).to(dev) for g in m_gs]
    ids = [torch.from_numpy(i).to(dev) for i in m_ids]
    gen = torch.Generator().manual_seed(1234 + rank)
    h_host = torch.randn(B, pos.shape[0], D, generator=gen).pin_memory()
    pos_host = (torch.from_numpy(pos).unsqueeze(0) + 0.01 * torch.randn(B, pos.shape[0], 2, generator=gen)).pin_memory()
    h_dev = h_host.to(dev).requires_grad_(True)
    pos_dev = pos_host.to(dev)
    from bsms_gnn_b200.dist import GradBucket
    bucket = GradBucket(params) if world > 1 else None

    def step(h, p):
        for q in params:
            q.grad = None
        h.grad = None
        out = model(h, ids, gs, p)
        loss = out.square().mean()
        loss.backward()
        if world > 1:  # data-parallel exchange: one all-reduce of the 2.15 M parameter gradients
            bucket.step_sync()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(h_dev, pos_dev)
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as cs:
        e0.record()
        for _ in range(args.steps):
            step(h_dev, pos_dev)
        e1.record()
        barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- end-to-end: host buffers in, loss out, through the public module API
    barrier()
    #      every step copies ITS inputs from pinned host memory and reads its loss back; the copy of step
    #      k+1 runs on a copy stream into the other of two device buffers while step k computes (what a
    #      prefetching data loader does), so the PCIe transfer overlaps the kernels instead of preceding them
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(h_dev).requires_grad_(True), torch.empty_like(pos_dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def upload(k):
        hb, pb = bufs[k & 1]
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(done[k & 1])  # the step that last used this buffer has finished
            hb.copy_(h_host, non_blocking=True)
            pb.copy_(pos_host, non_blocking=True)
            ready[k & 1].record(copy_stream)

    for ev in done:
        ev.record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    upload(0)
    for k in range(args.steps):
        if k + 1 < args.steps:
            upload(k + 1)
        torch.cuda.current_stream().wait_event(ready[k & 1])
        hb, pb = bufs[k & 1]
        loss = step(hb, pb)
        done[k & 1].record()
        _ = loss.item()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # ---- roofline pass: per-kernel CUDA-event timing inside the library (rank 0)
    roofline, breakdown, roofline_edge = None, None, None
    prof_steps = 2
    if rank == 0:
        _lib.prof_enable(True)
    for _ in range(prof_steps):  # every rank runs the steps (the all-reduce inside is collective)
        step(h_dev, pos_dev)
    barrier()
    if rank == 0:
        prof = _lib.prof_collect()
        _lib.prof_enable(False)
        pk = peaks()
        Re, Rn = B * edge_rows, B * node_rows
        tensor_mode = args.mode != "fp32"
        fwd_passes = 1 if tensor_mode else 2  # fp32 FFMA edge layers also run in backward's recompute
        # algorithmic work of each kernel class in ONE fwd+bwd step (DESIGN.md "Kernels"):
        #   FLOPs for the dense kernels, gather-counted bytes for the bandwidth kernels
        #   (bf16 mode: "dgrad" = fused node chain (3 dgrad + 3 wgrad GEMMs) + 4 layer-0 dgrad blocks per node row,
        #    "wgrad" = the 4 layer-0 weight-gradient blocks per node row; the edge GEMMs live in edge_chain*)
        flops = {"edge_fwd_gemm": fwd_passes * Re * 3 * 2 * D * D,
                 "dgrad": (Rn * 10 if args.mode == "bf16" else Re * 3 + Rn * 7) * 2 * D * D,
                 "wgrad": (Rn * 4 if args.mode == "bf16" else Re * 3 + Rn * 8) * 2 * D * D,
                 "node_fwd_gemm": 2 * Rn * 7 * 2 * D * D,
                 "edge_chain": Re * 3 * 2 * D * D,
                 "edge_chain_bwd": Re * 9 * 2 * D * D}  # 3 recompute + 3 data-gradient + 3 weight-gradient GEMMs
        byts = {"edge_combine": 2 * (Re * (2 * D * 4 + 16 + 8 + D * 4)), "ln_segsum": 2 * (Re * D * 4 + Rn * D * 4),
                "ln_bwd": Re * 3 * D * 4 + Rn * 3 * D * 4, "edge_grad_segsum": 2 * Re * D * 4 + Rn * 2 * D * 4,
                "edge_chain": Re * (2 * D * 4 + 2 * 2 * 4 + 2 * 4) + Rn * D * 4,
                # backward: the two projected rows are gathered ONCE (a0 stays in shared memory for its weight
                # gradient), one upstream-gradient row is gathered, one gradient row is scattered per edge (sender
                # side); the receiver side is reduced per destination run first (one row per node)
                "edge_chain_bwd": Re * (2 * D * 4 + D * 4 + D * 4 + 2 * 2 * 4 + 2 * 4) + Rn * D * 4}
        breakdown = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] / prof_steps}
                     for k, v in prof.items() if v[1]}
        total_ms = sum(v["ms_per_step"] for v in breakdown.values())

        def roof(kind):
            tms = breakdown[kind]["ms_per_step"]
            r = {"kernel": kind, "share_of_step": tms / total_ms, "traffic": None,
                 "avg_launch_us": 1e3 * tms / breakdown[kind]["launches_per_step"]}
            if kind in ("edge_chain", "edge_chain_bwd") and args.mode == "bf16":
                # 1x bf16 MMA: the governing bound is the slower of the gather-counted HBM floor
                # and the tensor-pipe floor; for both fused kernels that is HBM
                ach = byts[kind] / (tms * 1e-3) / 1e9
                tfl = flops[kind] / (tms * 1e-3) / 1e12
                hbm_floor_ms = byts[kind] / (pk["hbm_gbs"] * 1e9) * 1e3
                tensor_floor_ms = flops[kind] / (pk["bf16_tflops_sustained"] * 1e12) * 1e3
                r.update(bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"],
                         peak_source=pk["source"], tensor_tflops=tfl, hbm_floor_ms=hbm_floor_ms,
                         tensor_floor_ms=tensor_floor_ms, algorithmic_bytes_per_step=byts[kind])
                if tensor_floor_ms > hbm_floor_ms:
                    r.update(bound="tensor", achieved=tfl, peak=pk["bf16_tflops_sustained"], unit="TFLOP/s",
                             frac=tfl / pk["bf16_tflops_sustained"])
            elif kind in flops:
                ach = flops[kind] / (tms * 1e-3) / 1e12
                peak = pk["bf16_tflops_sustained"]
                note = None
                if kind == "edge_chain":  # fp16x3: every logical MAC costs three fp16 MACs
                    peak = peak / 3.0
                    note = "fp16x3 split: peak = measured bf16/fp16 dense peak / 3 (3 MMAs per logical MMA)"
                elif kind != "edge_chain":
                    note = "FFMA fp32 kernel; fraction is against the tensor-pipe peak the tcgen05 kernels target"
                r.update(bound="tensor", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak,
                         peak_source=pk["source"] + " bf16 sustained", note=note)
                if kind == "edge_chain":
                    r["gather_counted_gbs"] = byts[kind] / (tms * 1e-3) / 1e9
            else:
                ach = byts.get(kind, 0) / (tms * 1e-3) / 1e9
                r.update(bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"],
                         peak_source=pk["source"])
            return r

        top = max(breakdown, key=lambda k: breakdown[k]["ms_per_step"])
        roofline = roof(top)
        # DRAM traffic of the same kernel from the committed ncu --set full capture (per launch, level 0)
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        traffic_db = json.load(open(tpath)) if os.path.exists(tpath) else {}

        def add_traffic(r):
            tj = traffic_db.get(r["kernel"]) if r else None
            if tj:
                r["traffic"] = tj["dram_bytes_per_launch"]
                r["traffic_note"] = tj["note"]

        add_traffic(roofline)
        if "edge_chain" in breakdown and top != "edge_chain":
            roofline_edge = roof("edge_chain")
            add_traffic(roofline_edge)
        else:
            roofline_edge = None

    # ---- untimed self-check (rank 0): the step's own output and input gradient for ONE sample of the batch
    #      against the CPU oracle (forward) and, in bf16 mode, the fp64 model of the bf16 arithmetic (gradient)
    self_check = None
    if rank == 0 and not args.no_self_check:
        from tests.util import l2_rel, max_rel
        step(h_dev, pos_dev)
        torch.cuda.synchronize()
        gs_c = [torch.from_numpy(g) for g in m_gs]
        ids_c = [torch.from_numpy(i) for i in m_ids]
        h1, p1 = h_host[:1].double(), pos_host[:1].double()
        params64 = {k: v.double() for k, v in O.init_params(args.depth, pos_dim=2, seed=0).items()}
        with torch.no_grad():
            out_dev = model(h_dev[:1], ids, gs, pos_dev[:1]).cpu()
            ref = O.bsgmp(h1, ids_c, gs_c, p1, params64, args.depth)
        fwd_err = max_rel(out_dev, ref)
        fwd_tol = 3e-2 if args.mode == "bf16" else 1e-5
        self_check = {"sample": f"sample 0 of {B}", "forward_max_rel_vs_oracle": fwd_err, "forward_tol": fwd_tol}
        ok = fwd_err < fwd_tol
        if args.mode == "bf16":
            from oracle import bf16_model as M
            hq = h1.clone().requires_grad_(True)
            (M.bsgmp_bf16(hq, ids_c, gs_c, p1, params64, args.depth).square().sum() / (B * pos.shape[0] * D)).backward()
            g_err = l2_rel(h_dev.grad[:1].cpu(), hq.grad)
            self_check.update(grad_h_l2_rel_vs_bf16_model=g_err, grad_tol=2e-2)
            ok = ok and g_err < 2e-2
        self_check["pass"] = bool(ok)
        assert ok, f"bench self-check failed: {self_check}"

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b_cpu = min(8, B)
        times, cores, kind, w_done = cpu_reference_run(pos, m_gs, m_ids, args.depth, 6, 2, b_cpu, 25.0)
        t = sum(times) / len(times)
        cpu = {"value": b_cpu * E0 / t / 1e6, "unit": "M-edges/s", "cores": cores, "kind": kind,
               "sample": f"batch {b_cpu} of {B} (same mesh and weights), {len(times)} fwd+bwd steps after {w_done} warm-up, "
                         f"{t * 1e3:.0f} ms each; the full batch is timed by --impl reference"}

    if rank == 0:
        value = world * B * E0 / (ms * 1e-3) / 1e6
        clk = cs.result
        out = {
            "metric": "M-edges/s per BSMS fwd+bwd step", "value": value, "unit": "M-edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16x3": "f32 (fp16x3 split MMA, fp32 accumulate)",
                      "bf16": "bf16 MMA operands, fp32 storage/accumulate"}[args.mode],
            "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": B, "mode": args.mode,
                       "parallelism": f"dp{world} (batch axis, grad all-reduce)" if world > 1 else "single GPU",
                       "l2": "working set per step (>5 GB of activations) exceeds the 126 MB L2; no explicit flush"},
            "edge_evals_per_s": world * B * edge_rows / (ms * 1e-3),
            "clocks": {"sm_mhz": clk.get("sm_mhz"), "sm_max_mhz": clk.get("sm_max_mhz"), "reasons": clk.get("reasons")},
            "e2e": {"value": world * B * E0 / e2e_s / 1e6, "unit": "M-edges/s",
                    "h2d_bytes_per_step": int(h_host.numel() * 4 + pos_host.numel() * 4), "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_fused_edge_kernel": roofline_edge, "kernel_breakdown": breakdown, "cpu_baseline": cpu,
            "self_check": self_check,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
