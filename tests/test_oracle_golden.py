"""CPU: the oracle reproduces every golden fixture (which hold outputs of the unmodified reference,
written by oracle/make_goldens.py)."""
import ast

import numpy as np
import pytest
import torch

import helpers as h
from oracle import hlvae_oracle as orc

DT = torch.float64


@pytest.mark.parametrize("name", h.KL_CASES)
def test_oracle_kl_matches_golden(name):
    g = h.load(name)
    kargs = ast.literal_eval(str(g["kargs"]))
    fixed_T = int(g["T"]) if int(g["fixed_T_api"]) else None
    ref = h.oracle_kl(kargs, int(g["L"]), h.t(g["x"]), h.t(g["mu"]), h.t(g["log_v"]), h.t(g["z"]), h.t(g["m"]),
                      h.t(g["H"]), h.t(g["ros0"]), h.t(g["rls0"]), h.t(g["ros1"]), h.t(g["rls1"]), h.t(g["noise"]),
                      int(g["P_tot"]), int(g["n_subj"]), int(g["N_tot"]), float(g["eps"]), fixed_T=fixed_T)
    # fixtures hold the REFERENCE's outputs; tolerance = the float64 round-off floor between two
    # evaluation orders at cond(K0zz) ~ 1e7 (see oracle/make_goldens.py)
    assert h.rel_err(ref["kld"], g["kld"]) < 1e-6
    for key in ("d_mu", "d_logv", "d_z", "d_m", "d_H"):
        assert h.rel_err(ref[key], g[key]) < 2e-6, key
    if int(g["natural_gradient"]):
        assert h.rel_err(ref["grad_m"], g["grad_m"]) < 2e-6
        assert h.rel_err(ref["grad_H"], g["grad_H"]) < 2e-6
    for key in ("d_os0", "d_ls0", "d_os1", "d_ls1"):
        if g[key].size:
            assert h.rel_err(ref[key], g[key]) < 1e-4, key
    for key in ("A", "B", "C", "D", "E", "F", "kld_qu_pu"):
        assert h.rel_err(ref["terms"][key], g["term_" + key]) < 1e-12, key


@pytest.mark.parametrize("name", h.LOGLIK_CASES)
def test_oracle_loglik_matches_golden(name):
    g = h.load(name)
    types = h.parse_types(g)
    descs, E_x, P_th = orc.build_layout(types)
    nr, npos = h.golden_norm(g, "cpu")
    lvr = h.t(g["log_vy_real"]) if g["log_vy_real"].size else None
    lvp = h.t(g["log_vy_pos"]) if g["log_vy_pos"].size else None
    theta = h.t(g["theta"]).requires_grad_(True)
    lpx, lpm, params = orc.loglik_and_reconstruction(descs, h.t(g["data"]), h.t(g["mask"]), theta, lvr, lvp, nr, npos,
                                                     bool(int(g["conv"])))
    (lpx * h.t(g["g_up"])).sum().backward()
    mean, mode = orc.statistics(descs, params.detach(), lvp)
    dtr = orc.discrete_variables_transformation(descs, h.t(g["data"]))
    assert h.rel_err(lpx, g["log_p_x"]) < 1e-12
    assert h.rel_err(lpm, g["log_p_x_missing"]) < 1e-12
    assert h.rel_err(params, g["params"]) < 1e-12
    assert h.rel_err(theta.grad, g["d_theta"]) < 1e-12
    assert np.array_equal(mean.numpy(), g["recon_mean"])
    assert np.array_equal(mode.numpy(), g["recon_mode"])
    assert np.array_equal(dtr.numpy(), g["data_transformed"])


def test_fixed_T_and_iter_agree():
    """elbo_functions.py:118-193 and :196-285 are the same mathematics when every subject has T rows."""
    inp = h.make_kl_inputs(L=3, M=10, n_subj=4, T=5, seed=9)
    a = h.oracle_kl(inp["kargs"], 3, inp["x"], inp["mu"], inp["lv"], inp["z"], inp["m"], inp["H"], inp["ros0"],
                    inp["rls0"], inp["ros1"], inp["rls1"], inp["noise"], 200, 4, 200 * 5, 1e-6)
    b = h.oracle_kl(inp["kargs"], 3, inp["x"], inp["mu"], inp["lv"], inp["z"], inp["m"], inp["H"], inp["ros0"],
                    inp["rls0"], inp["ros1"], inp["rls1"], inp["noise"], 200, 4, 200 * 5, 1e-6, fixed_T=5)
    assert h.rel_err(a["kld"], b["kld"]) < 1e-9
    assert h.rel_err(a["grad_H"], b["grad_H"]) < 1e-9


@pytest.mark.parametrize("name", ["predict_default_ragged", "predict_default_fixedT", "predict_sweep_ragged"])
def test_oracle_predict_matches_reference_goldens(name):
    """GP posterior-mean prediction (utils.py:99-271): the oracle restatement against the frozen outputs of the
    unmodified reference functions."""
    import ast
    g = h.load(name)
    kargs = ast.literal_eval(str(g["kargs"]))
    spec0, spec1 = orc.compile_spec(**kargs)
    prm0 = orc.KernelParams(h.t(g["ros0"]), h.t(g["rls0"]))
    prm1 = orc.KernelParams(h.t(g["ros1"]), h.t(g["rls1"]))
    x = h.t(g["x"])
    zp = orc.batch_predict(spec0, prm0, spec1, prm1, h.t(g["noise"]), x, h.t(g["test_x"]), h.t(g["mu"]), h.t(g["z"]),
                           orc.split_subjects_by_id(x, kargs["id_covariate"]), kargs["id_covariate"], float(g["eps"]))
    assert h.rel_err(zp, g["Z_pred"]) < 1e-9


@pytest.mark.parametrize("name", ["dubo_default", "dubo_sweep"])
def test_oracle_dubo_matches_reference_goldens(name):
    """validation.validation_dubo (validation.py:16-76): oracle restatement vs the unmodified reference."""
    import ast
    g = h.load(name)
    kargs = ast.literal_eval(str(g["kargs"]))
    spec0, spec1 = orc.compile_spec(**kargs)
    prm0 = orc.KernelParams(h.t(g["ros0"]), h.t(g["rls0"]))
    prm1 = orc.KernelParams(h.t(g["ros1"]), h.t(g["rls1"]))
    d = orc.validation_dubo(spec0, prm0, spec1, prm1, h.t(g["noise"]), h.t(g["x"]), h.t(g["mu"]), h.t(g["log_v"]),
                            h.t(g["z"]), int(g["n_subj"]), int(g["T"]), float(g["eps"]))
    assert h.rel_err(d, g["dubo"]) < 1e-9


@pytest.mark.parametrize("name", ["legacy_bounds_default", "legacy_bounds_sweep"])
def test_oracle_unbatched_bounds_match_reference_goldens(name):
    """elbo_functions.deviance_upper_bound (:60-115) and elbo_functions.elbo (:9-57), one latent dimension,
    un-batched kernels: oracle restatement vs the unmodified reference."""
    import ast
    g = h.load(name)
    kargs = ast.literal_eval(str(g["kargs"]))
    spec0, spec1 = orc.compile_spec(**kargs)
    prm0 = orc.KernelParams(h.t(g["ros0"]), h.t(g["rls0"]))
    prm1 = orc.KernelParams(h.t(g["ros1"]), h.t(g["rls1"]))
    args = (spec0, prm0, spec1, prm1, h.t(g["noise"]), h.t(g["x"]))
    d = orc.deviance_upper_bound(*args, h.t(g["mu"]), h.t(g["log_v"]), h.t(g["z"]), int(g["n_subj"]), int(g["T"]),
                                 float(g["eps"]))
    e = orc.elbo(*args, h.t(g["mu"]), h.t(g["z"]), int(g["n_subj"]), int(g["T"]), float(g["eps"]))
    assert h.rel_err(d, g["dubo"]) < 1e-9 and h.rel_err(e, g["elbo"]) < 1e-9


def test_oracle_aux_likelihoods_match_reference_goldens():
    """loglik_real / loglik_pos with the variance network (HL_VAE/loglik.py:45-48,104-108) and loglik_beta
    (:216-256): oracle restatements vs the unmodified reference's outputs and gradients."""
    g = h.load("loglik_aux")
    mask, g_up = h.t(g["mask"]), h.t(g["g_up"])
    cases = [("real_norm", orc.loglik_real_rowvar, (h.t(g["real_norm_nm"]), h.t(g["real_norm_nv"])), False),
             ("real_plain", orc.loglik_real_rowvar, (), False),
             ("pos_norm", orc.loglik_pos_rowvar, (h.t(g["real_norm_nm"]), h.t(g["real_norm_nv"])), False)]
    for tag in ("beta_wide", "beta_fallback2", "beta_fallback3"):
        cases.append((tag, orc.loglik_beta, (h.t(g[tag + "_ranges"]).reshape(-1, 2),), True))
    for tag, fn, extra, beta in cases:
        data = h.t(g[tag + "_data"])
        D = data.shape[1]
        th = h.t(g[tag + "_theta"]).requires_grad_(True)
        args = list(extra)
        if beta:
            disp = h.t(g[tag + "_disp"]).requires_grad_(True)
            args.append(disp)
        lpx, lpm, pa, pb = fn(data, mask[:, :D], th, *args)
        (lpx * g_up[:, :D]).sum().backward()
        for key, val in (("log_p_x", lpx), ("log_p_x_missing", lpm), ("prm_a", pa), ("prm_b", pb), ("d_theta", th.grad)):
            assert h.rel_err(val, g[f"{tag}_{key}"]) < 1e-12, (tag, key)
        if beta:
            assert h.rel_err(disp.grad, g[tag + "_d_disp"]) < 1e-12, tag


@pytest.mark.parametrize("name", h.THETA_CASES + ["theta_logvar_mixed"])
def test_oracle_theta_matches_reference_goldens(name):
    """HLVAE.theta_estimation (HLVAE.py:416-453): oracle restatement vs the unmodified reference's outputs."""
    g = h.load(name)
    types, conv = h.parse_types(g), bool(int(g["conv"]))
    _, heads, kinds = h.golden_heads(g, "cpu", conv)
    y = h.t(g["y"]).requires_grad_(True)
    theta = orc.theta_estimation(types, heads, y, h.t(g["mask"]), conv=conv,
                                 logvar_network=bool(int(g.get("logvar_network", 0))))
    (theta * h.t(g["g_up"])).sum().backward()
    assert h.rel_err(theta, g["theta"]) < 1e-13
    assert h.rel_err(y.grad, g["d_y"]) < 1e-12
    for i, hd in enumerate(heads):
        for n, prm in hd.items():
            assert h.rel_err(prm.grad, g[f"d_g{i}_{n}"]) < 1e-11, (i, n)


@pytest.mark.parametrize("name", h.THETA_CASES)
def test_packed_head_form_equals_reference(name):
    """Host logic of hl-vae_b200/theta.py on CPU (no kernel): the per-column (weight, bias, mode) form that
    `pack_heads` builds reproduces the reference's theta when evaluated with plain torch ops."""
    from hlvae_b200 import theta as th, _lib
    g = h.load(name)
    types, conv = h.parse_types(g), bool(int(g["conv"]))
    obs_layer, _, _ = h.golden_heads(g, "cpu", conv)
    lay = th.HeadLayout(types, conv, "cpu")
    y = h.t(g["y"])
    W, b = th.pack_heads(obs_layer, lay, y.shape[2])
    z = b[None, :] + torch.einsum("npk,pk->np", y[:, lay.col_var.long(), :], W)
    mode = lay.col_mode
    z = torch.where(mode == _lib.HEAD_SIGMOID, torch.sigmoid(z), z)
    z = torch.where(mode == _lib.HEAD_ZERO, torch.zeros_like(z), z)
    z = torch.where(mode == _lib.HEAD_BIAS, b[None, :].expand_as(z), z)
    assert h.rel_err(z, g["theta"]) < 1e-13
    tv = lay.tile_var.tolist()
    vp = lay.var_pcol.tolist()
    assert tv[0] == 0 and tv[-1] == lay.D and all(a < c for a, c in zip(tv, tv[1:]))
    assert all(vp[c] - vp[a] <= th.MAX_TILE and c - a <= th.MAX_TILE for a, c in zip(tv, tv[1:]))


@pytest.mark.parametrize("name", h.NORM_CASES)
def test_oracle_batch_normalization_matches_reference_goldens(name):
    """HL_VAE.utils.batch_normalization (HL_VAE/utils.py:88-143): oracle restatement vs the unmodified reference."""
    g = h.load(name)
    types, conv = h.parse_types(g), bool(int(g["conv"]))
    descs, _, _ = orc.build_layout(types)
    X, nr, npos = orc.batch_normalization(descs, h.t(g["data"]), h.t(g["mask"]), conv)
    assert h.rel_err(X, g["X"]) < 1e-14
    for tag, mine in (("real", nr), ("pos", npos)):
        assert (mine is None) == (g[tag + "_mean"].size == 0)
        if mine is not None:
            assert h.rel_err(mine[0], g[tag + "_mean"]) < 1e-14 and h.rel_err(mine[1], g[tag + "_var"]) < 1e-14
