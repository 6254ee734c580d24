"""tests/util.py gmp_reference_kink_aware (test infrastructure of the GPU gradient checks) on the CPU: a gradient computed
with the OTHER one-sided ReLU derivative at the pre-activation closest to zero must be recognised and explained by
exactly that entry, and an unexplained error must stay an error."""
import torch

from oracle import bsms_oracle as O
from tests.util import KinkRelu, gmp_reference_kink_aware, load_hier, max_rel


def _setup():
    m_gs, m_ids, pos0, d = load_hier("grid12")
    g, N = m_gs[0], pos0.shape[0]
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(2, N, 128, generator=gen)
    pos = torch.randn(N, 2, generator=gen)
    params = {k[len("bottom_gmp."):]: v for k, v in O.init_params(0, pos_dim=2, seed=4).items()}
    w = torch.randn(2, N, 128, generator=gen).double()
    return x, g, pos, params, w


def _grads(x, g, pos, params, w, flips=()):
    pr = {"g." + k: v.double().clone().requires_grad_(True) for k, v in params.items()}
    xr = x.double().clone().requires_grad_(True)
    relu = KinkRelu(flips)
    out = O.gmp(xr, g, pos.double(), pr, "g", relu=relu)
    (out * w).sum().backward()
    return xr.grad, {k[2:]: v.grad for k, v in pr.items()}, relu.z


def test_kink_relu_is_plain_relu_without_flips():
    x, g, pos, params, w = _setup()
    gx, grads, _ = _grads(x, g, pos, params, w)
    pr = {"g." + k: v.double().clone().requires_grad_(True) for k, v in params.items()}
    xr = x.double().clone().requires_grad_(True)
    (O.gmp(xr, g, pos.double(), pr, "g") * w).sum().backward()
    assert torch.equal(gx, xr.grad)
    for k, v in grads.items():
        assert torch.equal(v, pr["g." + k].grad)


def test_single_flip_is_found_and_explains_everything():
    x, g, pos, params, w = _setup()
    _, _, zs = _grads(x, g, pos, params, w)
    for call in (1, 4):  # an edge-MLP layer and a node-MLP layer
        z = zs[call]
        i = int(z.abs().reshape(-1).argmin())
        margin = float(z.abs().min() / z.abs().max())
        got_gx, got_grads, _ = _grads(x, g, pos, params, w, [(call, i)])  # "the implementation" sits on the other side
        ref_gx, _, _ = _grads(x, g, pos, params, w)
        if max_rel(got_gx, ref_gx) < 5e-4:
            continue  # this entry's upstream gradient happens to be negligible: nothing to explain
        out, gx, grads, flips = gmp_reference_kink_aware(x, g, pos, params, w, got_gx.float(), 5e-4, delta=1.5 * margin)
        assert flips == [(call, i)]
        assert max_rel(got_gx, gx) < 1e-6
        for k, v in got_grads.items():
            assert max_rel(v, grads[k]) < 1e-6, k


def test_unexplained_error_is_not_accepted():
    x, g, pos, params, w = _setup()
    ref_gx, _, _ = _grads(x, g, pos, params, w)
    bad = ref_gx.clone().float()
    bad[1, 17] += 0.05 * ref_gx.abs().max().float()  # an arithmetic error, not a kink
    _, gx, _, flips = gmp_reference_kink_aware(x, g, pos, params, w, bad, 5e-4, delta=1e-5)
    assert max_rel(bad, gx) > 5e-4 and len(flips) <= 1
