"""CUDA-graph replay of the processor forward equals the eager forward (bit-exact: same kernels,
same order) and follows new inputs."""
import pytest
import torch

from oracle import bsms_oracle as O
from tests.util import load_hier

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fp32", "fp16x3", "bf16"])
def test_graph_replay_matches_eager(mode):
    from bsms_gnn_b200.graphed import GraphedBSGMP
    from bsms_gnn_b200.ops import BSGMP
    dev = torch.device("cuda:0")
    m_gs, m_ids, pos, d = load_hier("grid44")
    gs, ids = [g.to(dev) for g in m_gs], [i.to(dev) for i in m_ids]
    model = BSGMP(d, 128, 3, 2, mode=mode).to(dev)
    model.load_state_dict(O.init_params(d, pos_dim=2, seed=4))
    gen = torch.Generator().manual_seed(1)
    h0 = torch.randn(pos.shape[0], 128, generator=gen).to(dev)
    h1 = torch.randn(pos.shape[0], 128, generator=gen).to(dev)
    p = pos.to(dev)
    graphed = GraphedBSGMP(model, ids, gs, h0, p)
    with torch.no_grad():
        for h in (h1, h0, h1):
            ref = model(h, ids, gs, p)
            out = graphed(h).clone()
            if mode == "fp32":
                assert torch.equal(out, ref)
            else:  # the segmented reduce uses red.add: summation order can differ between runs
                assert float((out - ref).abs().max() / ref.abs().max()) < 2e-6
        # rollout-style feedback: the output of one step is the input of the next
        x = h0
        y = h0
        for _ in range(5):
            x = model(x, ids, gs, p)
            y = graphed(y).clone()
        assert float((x - y).abs().max() / x.abs().max()) < (1e-5 if mode != "bf16" else 1e-2)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-6), ("bf16", 1e-2)])
def test_graphed_training_step_matches_eager(mode, tol):
    """A whole forward+backward step (the fused tcgen05 kernels, cudaMemsetAsync, weight packing) replayed
    from one CUDA graph gives the eager step's loss and gradients and follows in-place input updates.
    Tolerance: fp32 uses no atomics on the path (1e-6).  bf16 reduces with red.add, whose order varies from
    run to run; the 1e-7 differences in the aggregated messages flip bf16 roundings and ReLU masks further
    down, so two EAGER runs of the same step already differ by ~2e-3 of the largest gradient (printed below
    as the noise floor, measured 1.8e-3; graph vs eager measured 2.4e-3 .. 3.0e-3).  The bound is 1e-2."""
    from bsms_gnn_b200.graphed import GraphedStep
    from bsms_gnn_b200.ops import BSGMP
    dev = torch.device("cuda:0")
    m_gs, m_ids, pos, d = load_hier("grid44")
    gs, ids, p = [g.to(dev) for g in m_gs], [i.to(dev) for i in m_ids], pos.to(dev)
    model = BSGMP(d, 128, 3, 2, mode=mode).to(dev)
    model.load_state_dict(O.init_params(d, pos_dim=2, seed=4))
    params = list(model.parameters())
    gen = torch.Generator().manual_seed(3)
    h = torch.randn(2, pos.shape[0], 128, generator=gen).to(dev).requires_grad_(True)
    h_next = torch.randn(2, pos.shape[0], 128, generator=gen).to(dev)

    def step():
        for q in params:
            q.grad = None
        h.grad = None
        loss = model(h, ids, gs, p).square().mean()
        loss.backward()
        return loss

    graphed = GraphedStep(step, warmup=2)
    # the gradient tensors the captured backward writes into (an eager step re-points .grad elsewhere)
    static_h_grad, static_grads = h.grad, [q.grad for q in params]
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    for new_h in (None, h_next):
        if new_h is not None:
            with torch.no_grad():
                h.copy_(new_h)
        loss = graphed()
        torch.cuda.synchronize()
        g_loss, g_h, g_p = float(loss.detach()), static_h_grad.clone(), [g.clone() for g in static_grads]
        loss = step()  # eager, same inputs
        torch.cuda.synchronize()
        e_loss, e_h, e_p = float(loss.detach()), h.grad.clone(), [q.grad.clone() for q in params]
        step()  # a second eager run: the run-to-run noise floor of this mode
        torch.cuda.synchronize()
        floor = rel(h.grad, e_h)
        print(f"\n[{mode}] graph vs eager {rel(g_h, e_h):.2e}, eager vs eager {floor:.2e}")
        assert abs(g_loss - e_loss) <= tol * abs(e_loss)
        assert rel(g_h, e_h) < tol
        for a, b in zip(g_p, e_p):
            assert rel(a, b) < tol
