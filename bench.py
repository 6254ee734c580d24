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
import datetime
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


def measure_mesh(args, rank, world, local_rank, dev, nx, depth, steps, warmup, want_single=False):
    """One large synthetic mesh, B=1, fwd+bwd, node-partitioned over the ranks (strong scaling).  Returns a dict
    (rank 0: filled; other ranks: timing fields only).  The process group must already be initialised for world > 1.
    want_single: rank 0 additionally times the SAME mesh un-partitioned on its own GPU (the N = 1 reference of the
    strong-scaling ratio) while the other ranks wait."""
    import torch.distributed as dist
    from bsms_gnn_b200 import _lib, hierarchy, meshgen, partition
    from bsms_gnn_b200.dist import GradBucket
    from bsms_gnn_b200.ops import BSGMP
    from bsms_gnn_b200.partitioned import DistExchanger, PartitionedBSGMP, exchange_requests
    from oracle import bsms_oracle as O

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
    exchange = "none"
    if world > 1:
        plan = exchange_requests(partition.build_rank_plan(m_gs, m_ids, n0, world, rank))
        if args.exchange == "push":
            from bsms_gnn_b200.halo import PushExchanger
            pm = PartitionedBSGMP(model, [plan], PushExchanger(), dev, pos_exchanger=DistExchanger())
            exchange = "one push kernel per exchange over CUDA-IPC peer memory (bsms_halo_exchange)"
        else:
            pm = PartitionedBSGMP(model, [plan], DistExchanger(), dev)
            exchange = "index_select + NCCL point-to-point group + index_add_"
        own = torch.from_numpy(plan.levels[0].nodes[:plan.levels[0].n_own])
        h_host = h_all[own].contiguous().pin_memory()
        p_dev = torch.from_numpy(pos)[own].to(dev)
        bucket = GradBucket(params)
        ghosts = [lv.n_local - lv.n_own for lv in pm.states[0].levels]
    else:
        h_host = h_all.pin_memory()
        p_dev = torch.from_numpy(pos).to(dev)
        ghosts = None
    gs = ids = None
    if world == 1 or (want_single and rank == 0):
        gs = [torch.from_numpy(g).to(dev) for g in m_gs]
        ids = [torch.from_numpy(i).to(dev) for i in m_ids]
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

    def timed(run, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / n

    for _ in range(warmup):
        step(h_dev)
    barrier()
    n0l = _lib.launch_count()
    step(h_dev)
    launches_per_step = _lib.launch_count() - n0l
    run = lambda: step(h_dev)
    graph = None
    if not args.no_graph:
        # the whole step (forward, backward, halo exchanges, gradient all-reduce) replayed from ONE CUDA graph
        from bsms_gnn_b200.graphed import GraphedStep
        graph = GraphedStep(lambda: step(h_dev), warmup=1)
        run = graph
        for _ in range(2):
            run()
    with ClockSampler(local_rank) as cs:
        ms = timed(run, steps)
    # end to end: this step's features come from pinned host memory, the loss goes back to the host
    barrier()
    t1 = time.perf_counter()
    for _ in range(steps):
        with torch.no_grad():
            h_dev.copy_(h_host, non_blocking=True)
        _ = run().item()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t1) / steps
    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1]) / 1e3
    kinds = None
    if args.prof:
        # per-kernel-class device time of one EAGER step (library profiler, CUDA events): where the partitioned step
        # spends its time — "transfer" contains the halo kernels, whose duration includes waiting for the slowest peer
        barrier()
        if rank == 0:
            _lib.prof_enable(True)
        for _ in range(2):
            step(h_dev)
        barrier()
        if rank == 0:
            prof = _lib.prof_collect()
            _lib.prof_enable(False)
            kinds = {k: {"ms": round(v[0] / 2, 3), "launches": v[1] // 2} for k, v in prof.items() if v[1]}
    res = {"kinds": kinds, "ms_per_step": ms, "e2e_ms_per_step": e2e_s * 1e3, "launches_per_step": int(launches_per_step), "ghosts": ghosts,
           "setup_s": setup_s, "n0": n0, "E0": E0, "edge_rows": edge_rows, "clocks": cs.result, "exchange": exchange,
           "h2d_bytes_per_step": int(h_host.numel() * 4), "graph": graph, "single_ms_per_step": None}
    if want_single and world > 1:
        # the N = 1 reference of the strong-scaling ratio: the same mesh, un-partitioned, on rank 0's GPU
        if rank == 0:
            h1 = h_all.to(dev).requires_grad_(True)
            p1 = torch.from_numpy(pos).to(dev)

            def step1():
                for q in params:
                    q.grad = None
                h1.grad = None
                model(h1, ids, gs, p1).square().mean().backward()

            for _ in range(2):
                step1()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(max(3, steps // 2)):
                step1()
            e1.record()
            torch.cuda.synchronize()
            res["single_ms_per_step"] = e0.elapsed_time(e1) / max(3, steps // 2)
            del h1, p1
        dist.barrier()
    return res


def run_mesh(args):
    """BASELINE.json config 5: one large synthetic mesh, B=1, fwd+bwd, node-partitioned over the ranks
    with halo exchanges (strong scaling: the mesh is fixed, N GPUs split it)."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    nx = args.nx if args.nx != 72 else 1414
    depth = args.depth
    r = measure_mesh(args, rank, world, local_rank, dev, nx, depth, args.steps, args.warmup, want_single=args.single_ref)
    if rank == 0:
        clk, ms, n0, E0 = r["clocks"], r["ms_per_step"], r["n0"], r["E0"]
        print(json.dumps({
            "metric": "M-edges/s per BSMS fwd+bwd step", "value": E0 / (ms * 1e-3) / 1e6, "unit": "M-edges/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16x3": "f32 (fp16x3 split MMA)", "bf16": "bf16 MMA operands, fp32 storage/accumulate"}[args.mode],
            "data": "synthetic",
            "config": {"workload": f"synthetic {nx}x{nx} tri-grid ({n0} nodes / {E0} directed edges), unet_depth {depth}, "
                                   f"latent 128, B=1, fwd+bwd", "mode": args.mode,
                       "parallelism": (f"node partition x{world}, {4 * depth + 1} halo exchanges per forward ({r['exchange']}), "
                                       f"grad all-reduce") if world > 1 else "single GPU",
                       "execution": "eager" if args.no_graph else "whole step replayed from one CUDA graph",
                       "rank0_ghost_rows_per_level": r["ghosts"], "setup_s": r["setup_s"],
                       "l2": "per-step working set far exceeds the 126 MB L2"},
            "edge_evals_per_s": r["edge_rows"] / (ms * 1e-3),
            "single_gpu_ms_per_step_same_run": r["single_ms_per_step"], "kernel_breakdown_eager": r["kinds"],
            "speedup_vs_single_gpu_same_run": (r["single_ms_per_step"] / ms) if r["single_ms_per_step"] else None,
            "clocks": {"sm_mhz": clk.get("sm_mhz"), "sm_max_mhz": clk.get("sm_max_mhz"), "reasons": clk.get("reasons")},
            "e2e": {"value": E0 / (r["e2e_ms_per_step"] * 1e-3) / 1e6, "unit": "M-edges/s", "h2d_bytes_per_step": r["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": 4, "ms_per_step": r["e2e_ms_per_step"]},
            "gpu_launches": int(r["launches_per_step"] * args.steps)}), flush=True)
    if world > 1:
        leave_group(r["graph"] is not None)


def leave_group(graph_captured_nccl):
    import torch.distributed as dist
    if graph_captured_nccl:
        # a process group whose NCCL work was captured into a live CUDA graph does not tear down cleanly:
        # drain the device and leave without the collective destructor
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)
    dist.destroy_process_group()


class _SimModel(torch.nn.Module):
    """The attributes of the reference's BSMS_Simulator (src/models/model.py:20-27) that the fused rollout reads:
    encoder / processor / decoder built from this package's ops, and fixed normaliser statistics."""

    class _Norm:
        def __init__(self, mean, std):
            self._m, self._s = mean, std

        def mean(self):
            return self._m

        def std_with_epsilon(self):
            return self._s

    def __init__(self, depth, C, P, mode, dev, seed=0):
        super().__init__()
        from bsms_gnn_b200.ops import BSGMP, MLP
        from oracle import bsms_oracle as O
        torch.manual_seed(seed)
        self.encode = MLP(C + 1, D, D, 3, True)
        self.process = BSGMP(depth, D, 3, P, mode=mode)
        self.decode = MLP(D, D, C, 3, False)
        self.process.load_state_dict(O.init_params(depth, pos_dim=P, seed=seed))
        self.pos_dim = P
        self.to(dev)
        self._inputNormalizer = self._Norm(torch.zeros(C + 1, dtype=torch.float64), torch.ones(C + 1, dtype=torch.float64))
        # small target std: the random-init network then perturbs the state by ~1e-2 per step, like a trained one
        self._targetNormalizer = self._Norm(torch.zeros(C, dtype=torch.float64), torch.full((C,), 1e-2, dtype=torch.float64))


def run_rollout(args):
    print(json.dumps(measure_rollout(args, args.mesh, args.mode, args.steps if args.steps != 10 else 599, args.warmup)))


def measure_rollout(args, mesh, mode, T, warmup):
    """BASELINE.json configs 2 and 4: T sequential whole-model forwards with feedback, B = 1 — the reference's
    rollout_one_traj (src/utils/rollout_utils.py:15-64): encoder -> processor -> decoder, normalisers, mask,
    residual, next input = prediction with the boundary nodes re-imposed.
      --mesh cylinder (config 2): 44x44 tri-grid, unet_depth 5, out_dim 2, static mesh positions;
      --mesh sphere   (config 4): icosphere-5 (10 242 nodes / 61 440 directed edges), pos_dim = out_dim = 3, unet_depth 6,
                                  the mesh positions the processor sees change EVERY step (= the predicted state).
    Device time per step from CUDA events over T graph replays; e2e adds a pinned-host copy of every step's state."""
    from bsms_gnn_b200 import _lib, hierarchy, meshgen
    from bsms_gnn_b200.simulator import FusedSimulator, GraphedRollout

    dev = torch.device("cuda", torch.cuda.current_device())
    if mesh == "sphere":
        pos, cells = meshgen.icosphere(5)
        depth, C, P, name = 6, 3, 3, "icosphere-5"
        m_gs, m_ids = hierarchy.build_hierarchy(meshgen.cells_to_flat_edge(cells), depth, pos.shape[0], pos)
    else:
        nx = args.nx if args.nx != 72 else 44
        depth = args.depth if args.nx != 72 else 5
        pos, m_gs, m_ids = build_workload(nx, depth)
        C, P, name = 2, 2, f"cylinder-like {nx}x{nx} tri-grid"
    N, E0 = pos.shape[0], int(m_gs[0].shape[1])
    edge_rows = 2 * sum(int(g.shape[1]) for g in m_gs[:depth]) + int(m_gs[depth].shape[1])
    model = _SimModel(depth, C, P, mode, dev)
    sim = FusedSimulator(model, mode=mode, pos_feedback=(mesh == "sphere"))
    gs = [torch.from_numpy(g).to(dev) for g in m_gs]
    ids = [torch.from_numpy(i).to(dev) for i in m_ids]
    gen = torch.Generator().manual_seed(3)
    ntype = (torch.rand(1, N, 1, generator=gen) > 0.9).float()
    p0 = torch.from_numpy(pos).unsqueeze(0)
    state0 = p0.clone() if mesh == "sphere" else torch.randn(1, N, C, generator=gen)
    ic = torch.cat([state0, p0, ntype], -1).to(dev)
    mask = (ntype == 0).float().to(dev)
    with torch.no_grad():
        # eager (no graph): python + launch bound
        for _ in range(warmup):
            sim.rollout(ic, mask, gs, ids, 3)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res_eager = sim.rollout(ic, mask, gs, ids, min(T, 100))
        e1.record()
        torch.cuda.synchronize()
        ms_eager = e0.elapsed_time(e1) / min(T, 100)
        launches = (_lib.launch_count() - n0) / min(T, 100)
        gr = GraphedRollout(sim, ic, mask, gs, ids)
        for _ in range(warmup):
            gr.step()
        gr.reset()
        torch.cuda.synchronize()
        with ClockSampler(dev.index) as cs:
            e0.record()
            res = gr.run(T)
            e1.record()
            torch.cuda.synchronize()
        ms_graph = e0.elapsed_time(e1) / T
        assert torch.isfinite(res).all()
        err = float((res[:min(T, 100)] - res_eager).abs().max() / res_eager.abs().max())
        # e2e: every step's state goes back to pinned host memory
        host = torch.empty(T, N, C, dtype=torch.float32).pin_memory()
        gr.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gr.run(T, host=host)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / T
    clk = cs.result
    return ({
        "metric": "M-edges/s per BSMS forward (rollout step, whole model)", "value": E0 / (ms_graph * 1e-3) / 1e6, "unit": "M-edges/s",
        "n_gpus": 1, "steps": T, "warmup": warmup, "ms_per_step": ms_graph, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": mode, "data": "synthetic",
        "config": {"workload": f"{name} ({N} nodes / {E0} directed edges), unet_depth {depth}, out_dim {C}, pos_dim {P}, B=1, "
                               f"rollout of {T} sequential whole-model steps with on-device feedback and boundary re-imposition"
                               + (", mesh positions updated every step" if mesh == "sphere" else ""),
                   "mode": mode, "execution": "one CUDA graph per step (encoder + processor + decoder + feedback)"},
        "ms_per_step_eager": ms_eager, "ms_per_step_graph": ms_graph, "graph_vs_eager_max_rel": err,
        "edge_evals_per_s": edge_rows / (ms_graph * 1e-3), "gpu_launches": int(launches * T), "kernels_per_step": launches,
        "clocks": {"sm_mhz": clk.get("sm_mhz"), "sm_max_mhz": clk.get("sm_max_mhz"), "reasons": clk.get("reasons")},
        "e2e": {"value": E0 / (e2e_ms * 1e-3) / 1e6, "unit": "M-edges/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": int(N * C * 4), "note": "the input of step k+1 is produced on the device by step k; every state is copied to pinned host memory"}})


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
    ap.add_argument("--mesh-nx", type=int, default=1414, help="N > 1: side of the strong-scaling mesh reported in extra.mesh_strong")
    ap.add_argument("--exchange", default="push", choices=["push", "nccl"], help="mesh workload, N > 1: halo exchange implementation")
    ap.add_argument("--prof", action="store_true", help="mesh workload: per-kernel-class breakdown of one eager step on rank 0")
    ap.add_argument("--single-ref", action="store_true", help="mesh workload, N > 1: rank 0 also times the un-partitioned mesh")
    ap.add_argument("--mesh", default="cylinder", choices=["cylinder", "sphere"], help="rollout workload: config 2 / config 4")
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
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
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
                rows_l = tj["edge_rows_in_launch"]
                per_row = byts[r["kernel"]] / Re  # includes the node-row share
                r["traffic_launch"] = {"edge_rows": rows_l, "algorithmic_bytes": per_row * rows_l,
                                       "dram_over_algorithmic": tj["dram_bytes_per_launch"] / (per_row * rows_l),
                                       "captured_at_commit": traffic_db.get("_captured_at_commit")}
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
        # rank-local: NO collective may run here (the other ranks are already past this point)
        for q in params:
            q.grad = None
        h_dev.grad = None
        model(h_dev, ids, gs, pos_dev).square().mean().backward()
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

    # ---- N > 1: BASELINE.json config 5 in the same run — the 2.0 M-node / 12.0 M-edge mesh node-partitioned over the
    #      N GPUs (strong scaling), with the un-partitioned step timed on rank 0 as the N = 1 reference
    mesh_strong, mesh_graph = None, False
    if world > 1 and os.environ.get("BSMS_BENCH_MESH", "1") != "0":
        del bufs
        torch.cuda.empty_cache()
        mr = measure_mesh(args, rank, world, local_rank, dev, args.mesh_nx, 6, max(5, min(args.steps, 10)), 3, want_single=True)
        mesh_graph = mr["graph"] is not None
        if rank == 0:
            mesh_strong = {"workload": f"synthetic {args.mesh_nx}x{args.mesh_nx} tri-grid ({mr['n0']} nodes / {mr['E0']} directed edges), "
                                       f"unet_depth 6, B=1, fwd+bwd, {args.mode}", "n_gpus": world, "ms_per_step": mr["ms_per_step"],
                           "value": mr["E0"] / (mr["ms_per_step"] * 1e-3) / 1e6, "unit": "M-edges/s",
                           "single_gpu_ms_per_step": mr["single_ms_per_step"],
                           "speedup_vs_n1": (mr["single_ms_per_step"] / mr["ms_per_step"]) if mr["single_ms_per_step"] else None,
                           "e2e": {"ms_per_step": mr["e2e_ms_per_step"], "h2d_bytes_per_step": mr["h2d_bytes_per_step"],
                                   "d2h_bytes_per_step": 4,
                                   "speedup_vs_n1_device_time": (mr["single_ms_per_step"] / mr["e2e_ms_per_step"]) if mr["single_ms_per_step"] else None},
                           "halo_exchange": mr["exchange"], "exchanges_per_step": 2 * (4 * 6 + 1),
                           "execution": "eager" if args.no_graph else "whole step replayed from one CUDA graph",
                           "rank0_ghost_rows_per_level": mr["ghosts"], "setup_s": mr["setup_s"],
                           "gpu_launches_per_step": mr["launches_per_step"],
                           "n1_reference": "the same mesh un-partitioned on rank 0's GPU, eager, CUDA events, same run"}

    # ---- N = 1: the same step in the fp32-PARITY mode (fp16x3: forward within 1e-5 of the reference, gradients 5e-4;
    #      every GEMM of forward and backward on tcgen05 with split operands), next to the bf16 headline
    parity_mode = None
    if world == 1 and args.mode != "fp16x3" and os.environ.get("BSMS_BENCH_PARITY_MODE", "1") != "0":
        from tests.util import max_rel
        model_p = BSGMP(args.depth, D, 3, 2, mode="fp16x3").to(dev)
        model_p.load_state_dict(O.init_params(args.depth, pos_dim=2, seed=0))
        params_p = list(model_p.parameters())

        def step_p():
            for q in params_p:
                q.grad = None
            h_dev.grad = None
            model_p(h_dev, ids, gs, pos_dev).square().mean().backward()

        for _ in range(3):
            step_p()
        torch.cuda.synchronize()
        ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ep0.record()
        for _ in range(5):
            step_p()
        ep1.record()
        torch.cuda.synchronize()
        ms_p = ep0.elapsed_time(ep1) / 5
        with torch.no_grad():
            out_p = model_p(h_dev[:1], ids, gs, pos_dev[:1]).cpu()
            ref_p = O.bsgmp(h_host[:1].double(), [torch.from_numpy(i) for i in m_ids], [torch.from_numpy(g) for g in m_gs],
                            pos_host[:1].double(), {k: v.double() for k, v in O.init_params(args.depth, pos_dim=2, seed=0).items()},
                            args.depth)
        parity_mode = {"mode": "fp16x3", "ms_per_step": ms_p, "value": B * E0 / (ms_p * 1e-3) / 1e6, "unit": "M-edges/s",
                       "fwd_max_rel": max_rel(out_p, ref_p), "fwd_tol": 1e-5, "steps": 5, "warmup": 3,
                       "note": "same workload, forward and backward GEMMs on tcgen05 with two-way split operands (3 MMAs per K step)"}
        assert parity_mode["fwd_max_rel"] < 1e-5, parity_mode
        del model_p, params_p
        torch.cuda.empty_cache()

    # ---- N = 1: the same step under the deterministic option (ops.set_deterministic): tensor-core kernels with
    #      rows + CSR-ordered segment sums and ordered partial-sum reductions; two steps must agree bit for bit
    det_mode = None
    if world == 1 and args.mode == "bf16" and os.environ.get("BSMS_BENCH_DET_MODE", "1") != "0":
        from bsms_gnn_b200 import ops as _ops
        _ops.set_deterministic(True)
        try:
            for _ in range(3):
                step(h_dev, pos_dev)
            g1 = [q.grad.clone() for q in params]
            step(h_dev, pos_dev)
            same = all(torch.equal(a_, q.grad) for a_, q in zip(g1, params))
            torch.cuda.synchronize()
            ed0, ed1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ed0.record()
            for _ in range(5):
                step(h_dev, pos_dev)
            ed1.record()
            torch.cuda.synchronize()
            ms_d = ed0.elapsed_time(ed1) / 5
        finally:
            _ops.set_deterministic(False)
        det_mode = {"mode": args.mode, "ms_per_step": ms_d, "value": B * E0 / (ms_d * 1e-3) / 1e6, "unit": "M-edges/s",
                    "bitwise_equal_steps": bool(same), "steps": 5, "warmup": 4,
                    "note": "bsms_set_deterministic(1): same workload, every parameter gradient of two consecutive steps compared bit for bit"}
        assert same, "deterministic option: two steps differ"
        del g1
        torch.cuda.empty_cache()

    # ---- N = 1: BASELINE.json configs 2 and 4 (whole-model rollouts, B = 1) in the same run, short form
    rollouts = None
    if world == 1 and os.environ.get("BSMS_BENCH_ROLLOUT", "1") != "0":
        rollouts = {}
        for cfg_name, mesh, rmode in (("config2_cylinder_fp16x3", "cylinder", "fp16x3"), ("config2_cylinder_bf16", "cylinder", "bf16"),
                                      ("config4_sphere_fp16x3", "sphere", "fp16x3"), ("config4_sphere_bf16", "sphere", "bf16")):
            rr = measure_rollout(args, mesh, rmode, 200, 3)
            rollouts[cfg_name] = {"workload": rr["config"]["workload"], "mode": rmode, "ms_per_step": rr["ms_per_step_graph"],
                                  "ms_per_step_eager": rr["ms_per_step_eager"], "kernels_per_step": rr["kernels_per_step"],
                                  "value": rr["value"], "unit": rr["unit"], "e2e_ms_per_step": rr["e2e"]["ms_per_step"],
                                  "d2h_bytes_per_step": rr["e2e"]["d2h_bytes_per_step"]}

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
                    "ms_per_step": e2e_s * 1e3,
                    "note": "every step uploads its features and positions from pinned host memory (double-buffered on a copy "
                            "stream) and reads back its 4-byte loss; outputs and gradients of a training step stay on the device "
                            "(the rollout configs in extra.rollouts return every state to the host)"},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_fused_edge_kernel": roofline_edge, "kernel_breakdown": breakdown, "cpu_baseline": cpu,
            "self_check": self_check, "extra": {"mesh_strong": mesh_strong, "rollouts": rollouts, "parity_mode": parity_mode,
                                                      "deterministic_mode": det_mode},
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        leave_group(mesh_graph)


if __name__ == "__main__":
    main()
