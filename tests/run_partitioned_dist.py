"""Multi-process check of the node-partitioned processor (run under torchrun on N GPUs):
every rank builds its plan, ghosts travel over NCCL point-to-point, the stitched forward and the
all-reduced parameter gradients are compared with the unpartitioned module on rank 0.

    torchrun --nproc-per-node 2 tests/run_partitioned_dist.py [--nx 200 --depth 4]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=120)
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--mode", default="fp32")
    ap.add_argument("--exchange", default="push", choices=["push", "nccl"])
    ap.add_argument("--steps", type=int, default=2, help="repeat the step: the push exchanger reuses its buffers / epochs")
    a = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from bsms_gnn_b200 import hierarchy, meshgen, partition
    from bsms_gnn_b200.dist import GradBucket
    from bsms_gnn_b200.ops import BSGMP
    from bsms_gnn_b200.partitioned import DistExchanger, PartitionedBSGMP, exchange_requests
    from oracle import bsms_oracle as O
    from tests.util import max_rel

    pos, cells = meshgen.tri_grid(a.nx, a.nx)
    m_gs, m_ids = hierarchy.build_hierarchy(meshgen.cells_to_flat_edge(cells), a.depth, pos.shape[0], pos)
    n0 = pos.shape[0]
    model = BSGMP(a.depth, 128, 3, 2, mode=a.mode).to(dev)
    model.load_state_dict(O.init_params(a.depth, pos_dim=2, seed=1))
    h = torch.randn(n0, 128, generator=torch.Generator().manual_seed(2)).to(dev)
    post = torch.from_numpy(pos).to(dev)
    plan = exchange_requests(partition.build_rank_plan(m_gs, m_ids, n0, world, rank))
    if a.exchange == "push":
        from bsms_gnn_b200.halo import PushExchanger
        pm = PartitionedBSGMP(model, [plan], PushExchanger(), dev, pos_exchanger=DistExchanger())
    else:
        pm = PartitionedBSGMP(model, [plan], DistExchanger(), dev)
    own = torch.from_numpy(plan.levels[0].nodes[:plan.levels[0].n_own]).to(dev)
    bucket = GradBucket(list(model.parameters()))
    p_own = post[own]
    for _ in range(a.steps):  # every step ends in the all-reduce (the barrier the buffer reuse relies on)
        model.zero_grad()
        (out_own,) = pm([h[own]], [p_own])
        (out_own.square().sum() / h.numel()).backward()
        bucket.step_sync()
    part_grads = {k: v.grad.clone() for k, v in model.named_parameters()}
    full = torch.zeros_like(h)
    full[own] = out_own.detach()
    dist.all_reduce(full)
    model.zero_grad()
    ok = True
    if rank == 0:
        ref = model(h, [torch.from_numpy(i).to(dev) for i in m_ids], [torch.from_numpy(g).to(dev) for g in m_gs], post)
        ref.square().mean().backward()
        e_f = max_rel(full, ref.detach())
        e_g = max(max_rel(part_grads[k], v.grad) for k, v in model.named_parameters())
        ghosts = [lv.n_local - lv.n_own for lv in pm.states[0].levels]
        print(f"partitioned x{world} [{a.exchange}] ({n0} nodes, depth {a.depth}, {a.mode}, {a.steps} steps): forward max-rel {e_f:.2e}, "
              f"param-grad max-rel {e_g:.2e}; rank-0 ghosts per level {ghosts}")
        ok = e_f < 1e-5 and e_g < 5e-4 if a.mode != "bf16" else e_f < 3e-2
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
