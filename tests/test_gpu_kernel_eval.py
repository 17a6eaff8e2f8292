"""GPU: dense additive-kernel evaluation (hlvae_kernel_eval_fwd/bwd) against the golden kernel
matrices the reference produced and against oracle autograd gradients."""
import ast

import pytest
import torch

import helpers as h
from hlvae_b200 import kernels
from oracle import hlvae_oracle as orc

pytestmark = pytest.mark.gpu
DT = torch.float64


@pytest.mark.parametrize("name", ["kl_default_ragged", "kl_sweep_ragged", "kl_masked_bin", "kl_T32_M40"])
def test_dense_matrices_match_reference(name, device):
    g = h.load(name)
    kargs = ast.literal_eval(str(g["kargs"]))
    L = int(g["L"])
    k0, k1, _ = h.build_product_kernels(kargs, L, device, g["ros0"], g["rls0"], g["ros1"], g["rls1"])
    x, z = h.t(g["x"], device), h.t(g["z"], device)
    assert h.rel_err(k0(x, z).evaluate(), g["K0xz"]) < 1e-12          # [N,Q] x [L,M,Q]
    assert h.rel_err(k0(z, z).evaluate(), g["K0zz"]) < 1e-12          # [L,M,Q] x [L,M,Q]
    assert h.rel_err(k1(x, x).evaluate(), g["K1xx"]) < 1e-12          # [N,Q] x [N,Q]
    # stacked forms used by the reference's fixed-T path (elbo_functions.py:145-150)
    xs = x[:6].unsqueeze(0).expand(L, 6, x.shape[1])
    assert h.rel_err(k1(xs, xs).evaluate(), g["K1xx"][:, :6, :6]) < 1e-12
    x4 = torch.stack([xs, xs.flip(1)], 0)                              # [P=2, L, T, Q]
    out4 = k1(x4, x4).evaluate()
    assert out4.shape == (2, L, 6, 6)
    assert h.rel_err(out4[0], g["K1xx"][:, :6, :6]) < 1e-12


def test_dense_gradients_match_oracle(device):
    inp = h.make_kl_inputs(L=3, M=9, n_subj=4, T=5, seed=5, kargs=h.synth.MASKED_KERNEL_ARGS)
    k0, k1, _ = h.build_product_kernels(inp["kargs"], 3, device, inp["ros0"], inp["rls0"], inp["ros1"], inp["rls1"])
    gen = torch.Generator().manual_seed(1)
    x = inp["x"]
    z = inp["z"].clone().requires_grad_(True)
    spec0, _ = orc.compile_spec(**inp["kargs"])
    prm0 = orc.KernelParams(inp["ros0"].clone(), inp["rls0"].clone()).requires_grad_()
    Wt = torch.randn(3, x.shape[0], 9, generator=gen, dtype=DT)
    Wz = torch.randn(3, 9, 9, generator=gen, dtype=DT)
    ref = (orc.eval_additive(spec0, prm0, x, z) * Wt).sum() + (orc.eval_additive(spec0, prm0, z, z) * Wz).sum()
    ref.backward()
    zd = inp["z"].to(device).requires_grad_(True)
    got = (k0(x.to(device), zd).evaluate() * Wt.to(device)).sum() + (k0(zd, zd).evaluate() * Wz.to(device)).sum()
    got.backward()
    assert h.rel_err(got, ref) < 1e-12
    assert h.rel_err(zd.grad, z.grad) < 1e-11
    gos, gls = h.kernel_grads(k0)
    assert h.rel_err(gos, prm0.raw_outputscale.grad) < 1e-11
    assert h.rel_err(gls, prm0.raw_lengthscale.grad) < 1e-11


def test_empty_inputs(device):
    k0, _, _ = h.build_product_kernels(h.synth.DEFAULT_KERNEL_ARGS, 2, device, torch.zeros(3, 2), torch.ones(3, 2),
                                       torch.zeros(2, 2), torch.ones(1, 2))
    x = torch.zeros(0, 6, dtype=DT, device=device)
    z = torch.zeros(2, 4, 6, dtype=DT, device=device)
    assert k0(x, z).evaluate().shape == (2, 0, 4)
