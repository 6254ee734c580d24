#!/usr/bin/env python
"""A/B harness for kernel variants selected by environment variables (development aid, GPU box).

    python tools/variant_bench.py --env BSMS_BWD_V --values 0,1,2 [--phases] [--batch 48] [--mode bf16]

For every value: a fresh process (the library reads the variable once) runs the metric workload
(airfoil-like 72x72, depth 6, fwd+bwd), prints the CUDA-event step time and the per-kernel-class
times from the library's own profiler, and stores the gradients; the parent then compares every
variant's gradients with the first one (same arithmetic, different kernel structure: they must
agree to reduction-order noise).  --phases adds one step under BSMS_PHASE_PROF=1 (per-phase cycles
of the fused kernels on stderr).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import torch
    from bench import build_workload
    from bsms_gnn_b200 import _lib
    from bsms_gnn_b200.ops import BSGMP
    from oracle import bsms_oracle as O
    dev = torch.device("cuda", 0)
    pos, m_gs, m_ids = build_workload(args.nx, args.depth)
    model = BSGMP(args.depth, 128, 3, 2, mode=args.mode).to(dev)
    model.load_state_dict(O.init_params(args.depth, pos_dim=2, seed=0))
    gs = [torch.from_numpy(g).to(dev) for g in m_gs]
    ids = [torch.from_numpy(i).to(dev) for i in m_ids]
    gen = torch.Generator().manual_seed(1234)
    h = torch.randn(args.batch, pos.shape[0], 128, generator=gen).to(dev).requires_grad_(True)
    p = (torch.from_numpy(pos).unsqueeze(0) + 0.01 * torch.randn(args.batch, pos.shape[0], 2, generator=gen)).to(dev)
    params = list(model.parameters())

    def step():
        for q in params:
            q.grad = None
        h.grad = None
        out = model(h, ids, gs, p)
        out.square().mean().backward()
        return out

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    _lib.prof_enable(True)
    for _ in range(2):
        step()
    prof = _lib.prof_collect()
    _lib.prof_enable(False)
    kinds = {k: round(v[0] / 2, 4) for k, v in prof.items() if v[1]}
    names = dict(model.named_parameters())
    keep = {"out": out.detach().float().cpu(), "gh": h.grad.detach().cpu()}
    for n in ["down_gmps.0.mlp_edge.seq.0.weight", "down_gmps.0.mlp_edge.seq.2.weight", "down_gmps.0.mlp_edge.seq.6.weight",
              "down_gmps.0.mlp_edge.seq.4.bias", "down_gmps.0.mlp_edge.seq.0.bias", "up_gmps.5.mlp_edge.seq.4.weight",
              "bottom_gmp.mlp_edge.seq.2.weight", "down_gmps.0.mlp_node.seq.0.weight", "down_gmps.3.mlp_node.seq.4.weight"]:
        if n in names:
            keep[n] = names[n].grad.detach().cpu()
    torch.save(keep, args.dump)
    print(json.dumps({"ms_per_step": round(ms, 4), "kinds_ms": kinds}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="BSMS_BWD_V")
    ap.add_argument("--values", default="0,1")
    ap.add_argument("--phases", action="store_true")
    ap.add_argument("--batch", type=int, default=48)
    ap.add_argument("--nx", type=int, default=72)
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--dump", default="")
    args = ap.parse_args()
    if args.child:
        return child(args)
    import torch
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    dumps = []
    for v in args.values.split(","):
        env = dict(os.environ)
        env[args.env] = v
        dump = f"/tmp/variant_{args.env}_{v}.pt"
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--dump", dump, "--batch", str(args.batch), "--nx", str(args.nx),
               "--depth", str(args.depth), "--steps", str(args.steps), "--mode", args.mode]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        print(f"[{args.env}={v}] rc={r.returncode} {line}", flush=True)
        if r.returncode != 0:
            print(r.stderr[-2000:], flush=True)
            continue
        dumps.append((v, dump))
        if args.phases:
            env["BSMS_PHASE_PROF"] = "1"
            r2 = subprocess.run(cmd[:-4] + ["--steps", "1", "--mode", args.mode], env=env, capture_output=True, text=True, timeout=600)
            lines = [l for l in r2.stderr.splitlines() if "phases]" in l]
            # level-0 launches have the most tiles: print the last forward / backward line with the largest tile count
            for tag in ("[fwd phases]", "[bwd phases]"):
                sel = [l for l in lines if l.startswith(tag)]
                if sel:
                    best = max(sel, key=lambda l: int(l.split("tiles")[1].split(":")[0]))
                    print(f"[{args.env}={v}] {best}", flush=True)
    if len(dumps) > 1:
        ref = torch.load(dumps[0][1])
        for v, d in dumps[1:]:
            cur = torch.load(d)
            errs = {k: float((cur[k].double() - ref[k].double()).norm() / ref[k].double().norm().clamp_min(1e-30)) for k in ref}
            worst = max(errs.values())
            print(f"[{args.env}={v}] l2-rel vs {args.env}={dumps[0][0]}: worst {worst:.3e} " +
                  " ".join(f"{k.split('.')[-3] if '.' in k else k}:{e:.1e}" for k, e in errs.items()), flush=True)


if __name__ == "__main__":
    main()
