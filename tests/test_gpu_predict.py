"""GPU parity of the GP posterior-mean predictors (hlvae_b200.predict; utils.py:99-271) against the frozen
outputs of the unmodified reference functions and against the oracle on larger fresh cases."""
import ast

import numpy as np
import pytest
import torch

import helpers as h
from hlvae_b200 import predict, synth, validation
from oracle import hlvae_oracle as orc

pytestmark = pytest.mark.gpu
DT = torch.float64


def _modules(g, dev):
    kargs = ast.literal_eval(str(g["kargs"]))
    L = int(g["L"])
    k0, k1, lik = h.build_product_kernels(kargs, L, dev, g["ros0"], g["rls0"], g["ros1"], g["rls1"], g["noise"])
    return kargs, L, k0.eval(), k1.eval(), lik.eval()


@pytest.mark.parametrize("name", ["predict_default_ragged", "predict_default_fixedT", "predict_sweep_ragged"])
def test_golden(name, device):
    g = h.load(name)
    kargs, L, k0, k1, lik = _modules(g, device)
    x, xt, mu, z = (h.t(g[k], device) for k in ("x", "test_x", "mu", "z"))
    zp = predict.batch_predict_varying_T(L, k0, k1, lik, x, xt, mu, z, kargs["id_covariate"], float(g["eps"]))
    assert zp.shape == (xt.shape[0], L)
    assert h.rel_err(zp, g["Z_pred"]) < 1e-7
    if not int(g["ragged"]):
        zf = predict.batch_predict(L, k0, k1, lik, x, xt, mu, z, int(g["n_subj"]), int(g["T"]), kargs["id_covariate"],
                                   float(g["eps"]))
        assert h.rel_err(zf, g["Z_pred"]) < 1e-7
    # zt_list may also be a list of per-latent [M, Q] tensors (utils.py:104)
    zl = predict.batch_predict_varying_T(L, k0, k1, lik, x, xt, mu, [z[i] for i in range(L)], kargs["id_covariate"],
                                         float(g["eps"]))
    # same arithmetic; only the order in which the chunks' partial sums reach the accumulators may differ
    assert h.rel_err(zl, zp) < 1e-12


@pytest.mark.parametrize("L,M,n_subj,T", [(8, 64, 60, 20), (3, 128, 25, 32), (4, 32, 40, 9), (3, 64, 9, 60), (2, 120, 7, 48)])
def test_against_oracle(L, M, n_subj, T, device):
    inp = h.make_kl_inputs(L, M, n_subj, T, seed=500 + M, ragged=True)
    k0, k1, lik = h.build_product_kernels(inp["kargs"], L, device, inp["ros0"], inp["rls0"], inp["ros1"], inp["rls1"],
                                          inp["noise"])
    x = inp["x"]
    rng = np.random.default_rng(1)
    sel = torch.from_numpy(rng.choice(x.shape[0], 37, replace=False))
    test_x = x[sel].clone()
    test_x[:, 0] += 0.25                                   # unseen time points of known subjects
    unseen, _ = synth.covariates(2, T, rng, first_id=99_000)
    test_x = torch.cat([test_x, unseen[:7]])
    spec0, spec1 = orc.compile_spec(**inp["kargs"])
    prm0, prm1 = orc.KernelParams(inp["ros0"], inp["rls0"]), orc.KernelParams(inp["ros1"], inp["rls1"])
    ref = orc.batch_predict(spec0, prm0, spec1, prm1, inp["noise"], x, test_x, inp["mu"], inp["z"],
                            orc.split_subjects_by_id(x, 2), 2, 1e-6)
    got = predict.batch_predict_varying_T(L, k0.eval(), k1.eval(), lik.eval(), x.to(device), test_x.to(device),
                                          inp["mu"].to(device), inp["z"].to(device), 2, 1e-6)
    # (K0zz + S) and K0zz + eps I are solved at cond ~ 1e7; both sides are float64
    assert h.rel_err(got, ref) < 1e-6
    # rows of a subject that is absent from prediction_x get the K0 term only; finite everywhere
    assert bool(torch.isfinite(got).all())


def test_cpu_tensors_fail_loudly():
    with pytest.raises((RuntimeError, NotImplementedError)):
        predict.batch_predict_varying_T(2, None, None, None, torch.zeros(4, 6, dtype=DT), torch.zeros(2, 6, dtype=DT),
                                        torch.zeros(4, 2, dtype=DT), torch.zeros(2, 3, 6, dtype=DT), 2, 1e-6)


@pytest.mark.parametrize("name", ["dubo_default", "dubo_sweep"])
def test_dubo_golden(name, device):
    """validation.validation_dubo (validation.py:16-76) against the unmodified reference's frozen output."""
    g = h.load(name)
    kargs, L, k0, k1, lik = _modules(g, device)
    d = validation.validation_dubo(L, k0, k1, lik, h.t(g["x"], device), h.t(g["mu"], device), h.t(g["log_v"], device),
                                   h.t(g["z"], device), int(g["n_subj"]), int(g["T"]), float(g["eps"]))
    assert d.shape == (1,) and h.rel_err(d, g["dubo"]) < 1e-7


def test_dubo_against_oracle_larger(device):
    L, M, n_subj, T = 6, 64, 50, 20
    inp = h.make_kl_inputs(L, M, n_subj, T, seed=611, ragged=False)
    k0, k1, lik = h.build_product_kernels(inp["kargs"], L, device, inp["ros0"], inp["rls0"], inp["ros1"], inp["rls1"],
                                          inp["noise"])
    spec0, spec1 = orc.compile_spec(**inp["kargs"])
    prm0, prm1 = orc.KernelParams(inp["ros0"], inp["rls0"]), orc.KernelParams(inp["ros1"], inp["rls1"])
    ref = orc.validation_dubo(spec0, prm0, spec1, prm1, inp["noise"], inp["x"], inp["mu"], inp["lv"], inp["z"], n_subj, T,
                              1e-6)
    got = validation.validation_dubo(L, k0.eval(), k1.eval(), lik.eval(), inp["x"].to(device), inp["mu"].to(device),
                                     inp["lv"].to(device), inp["z"].to(device), n_subj, T, 1e-6)
    assert h.rel_err(got, ref) < 1e-6


@pytest.mark.parametrize("name", ["legacy_bounds_default", "legacy_bounds_sweep"])
def test_unbatched_bounds_golden(name, device):
    """elbo_functions.deviance_upper_bound (:60-115) and elbo_functions.elbo (:9-57) on UN-BATCHED kernel objects
    (kernel_gen.generate_kernel_approx, :97-197) against the frozen outputs of the unmodified reference."""
    from hlvae_b200 import kernels, likelihoods
    g = h.load(name)
    kargs = ast.literal_eval(str(g["kargs"]))
    k0, k1 = kernels.generate_kernel_approx(kargs['cat_kernel'], kargs['bin_kernel'], kargs['sqexp_kernel'],
                                            kargs['cat_int_kernel'], kargs['bin_int_kernel'],
                                            kargs['covariate_missing_val'], kargs['id_covariate'])
    k0, k1 = k0.to(device).double(), k1.to(device).double()
    h.set_kernel_params(k0, h.t(g["ros0"], device), h.t(g["rls0"], device))
    h.set_kernel_params(k1, h.t(g["ros1"], device), h.t(g["rls1"], device))
    lik = likelihoods.GaussianLikelihood(noise_constraint=likelihoods.GreaterThan(1.0e-8)).to(device).double()
    lik.noise = h.t(g["noise"], device).reshape(1)
    x, mu, lv, z = (h.t(g[k], device) for k in ("x", "mu", "log_v", "z"))
    P, T, eps = int(g["n_subj"]), int(g["T"]), float(g["eps"])
    d = validation.deviance_upper_bound(k0.eval(), k1.eval(), lik.eval(), x, mu, lv, z, P, T, eps)
    e = validation.elbo(k0, k1, lik, x, mu, z, P, T, eps)
    assert d.shape == () and e.shape == ()
    assert h.rel_err(d, g["dubo"]) < 1e-8 and h.rel_err(e, g["elbo"]) < 1e-8
