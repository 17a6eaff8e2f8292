"""Pin the oracle against the reference itself and freeze golden fixtures.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_goldens.py            # validate + (re)write tests/golden/*.npz

What it does, per case:
  1. imports the UNMODIFIED reference modules from /root/reference (kernel_gen, kernel_spec,
     elbo_functions, HLVAE, HL_VAE.loglik, HL_VAE.read_functions) behind the stand-ins in
     oracle/standins (gpytorch, matplotlib are not installable here);
  2. evaluates the reference on seeded synthetic inputs, with autograd gradients;
  3. evaluates oracle/hlvae_oracle.py on the same inputs and asserts agreement: 1e-6 relative
     on kld, 2e-6 on grad_m / grad_H and the gradients of mu, log_v, Z, m, H; 1e-4 on the
     kernel hyper-parameter gradients, whose float64 round-off floor between two
     mathematically equal evaluation orders is already 1e-5 here because cond(K0zz + eps I)
     ~ 1e7 when inducing points are sampled from data rows (hard part
     1; with well-separated inducing points the same checks agree to 1e-11); likelihood
     terms 1e-11; argmax maps exactly;
  4. writes inputs + reference outputs to tests/golden/<case>.npz.
The GPU box has no /root/reference: tests there read only the .npz files.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("HLVAE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "standins"))
sys.path.insert(1, REF)
sys.path.insert(2, ROOT)
warnings.filterwarnings("ignore")

import hlvae_b200  # noqa: E402  (only for the synthetic-input generators)
from hlvae_b200 import synth  # noqa: E402
from oracle import hlvae_oracle as orc  # noqa: E402

import elbo_functions as ref_elbo  # noqa: E402  (reference)
import kernel_gen as ref_kernel_gen  # noqa: E402
import gpytorch  # noqa: E402  (stand-in)

GOLD = os.path.join(ROOT, "tests", "golden")
DT = torch.float64


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=DT), torch.as_tensor(b, dtype=DT)
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def check(name, a, b, tol=1e-6):
    r = rel(a, b)
    if os.environ.get("GOLDEN_REPORT_ONLY"):
        print(f"    {name}: {r:.2e}")
        return r
    assert r <= tol, f"{name}: oracle vs reference rel diff {r:.3e} > {tol}"
    return r


# ------------------------------------------------------------------ kernels / KL
def ref_modules(L, kargs, gen):
    k0, k1 = ref_kernel_gen.generate_kernel_batched(L, kargs['cat_kernel'], kargs['bin_kernel'], kargs['sqexp_kernel'],
                                                    kargs['cat_int_kernel'], kargs['bin_int_kernel'],
                                                    kargs['covariate_missing_val'], kargs['id_covariate'])
    lik = gpytorch.likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]),
                                                  noise_constraint=gpytorch.constraints.GreaterThan(1.0e-8))
    lik.noise = 1                                    # HLVAE_main.py:211
    lik.raw_noise.requires_grad = False
    k0.train().double(); k1.train().double(); lik.train().double()   # HLVAE_main.py:235-237
    # move the hyper-parameters off their symmetric defaults (keeps the float32-initialised base)
    with torch.no_grad():
        for k in (k0, k1):
            for p in k.parameters():
                p.add_(0.3 * torch.randn(p.shape, generator=gen, dtype=DT))
    return k0, k1, lik


def extract_params(kmod):
    """raw_outputscale per component and raw_lengthscale per SE factor (depth-first order)."""
    ros = torch.stack([k.raw_outputscale.detach().clone() for k in kmod.kernels]) if len(kmod.kernels) else torch.zeros(0, 1, dtype=DT)
    rls = [mod.raw_lengthscale.detach().reshape(-1).clone() for mod in kmod.modules()
           if isinstance(mod, gpytorch.kernels.RBFKernel)]
    rls = torch.stack(rls) if rls else torch.zeros(0, ros.shape[1], dtype=DT)
    return ros, rls


def extract_grads(kmod):
    gos = torch.stack([k.raw_outputscale.grad.clone() for k in kmod.kernels])
    gls = [mod.raw_lengthscale.grad.reshape(-1).clone() for mod in kmod.modules()
           if isinstance(mod, gpytorch.kernels.RBFKernel)]
    gls = torch.stack(gls) if gls else torch.zeros(0, gos.shape[1], dtype=DT)
    return gos, gls


def kl_case(name, kargs, L, M, n_subj, T, ragged, fixed_T_api, seed, trained_like=False, continuous_age=False,
            natural_gradient=True, shuffle_rows=False):
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    x, lens = synth.covariates(n_subj, T, rng, ragged=ragged, t_min=3, continuous_age=continuous_age)
    pool, _ = synth.covariates(40, T, rng, continuous_age=continuous_age)
    z0 = synth.inducing_points(torch.cat([x, pool]), L, M, rng)
    N_b = x.shape[0]
    if shuffle_rows:                                  # _iter groups by id value, not by position
        x = x[torch.from_numpy(rng.permutation(N_b))]
    mu0 = torch.randn(N_b, L, generator=gen, dtype=DT)
    lv0 = -3.0 * torch.rand(N_b, L, generator=gen, dtype=DT)
    m, H = synth.variational_state(L, M, gen)
    k0, k1, lik = ref_modules(L, kargs, gen)
    eps = 1e-6
    P_tot, N_tot = 200, 200 * T
    idc = kargs['id_covariate']
    spec0, spec1 = orc.compile_spec(**kargs)
    ros0, rls0 = extract_params(k0)
    ros1, rls1 = extract_params(k1)
    noise = lik.noise_covar.noise.detach().reshape(-1).clone()

    def ref_call(m_, H_, mu_, lv_, z_):
        if fixed_T_api:
            return ref_elbo.minibatch_KLD_upper_bound(k0, k1, lik, L, m_, H_, x, mu_, lv_, z_, P_tot, n_subj, T,
                                                      natural_gradient, eps)
        return ref_elbo.minibatch_KLD_upper_bound_iter(k0, k1, lik, L, m_, H_, x, mu_, lv_, z_, P_tot, n_subj, N_tot,
                                                       natural_gradient, idc, eps)

    if trained_like:      # a few natural-gradient steps from the reference's own update rule
        for _ in range(8):
            with torch.no_grad():
                _, gm, gH = ref_call(m, H, mu0, lv0, z0)
            m, H = orc.natural_gradient_update(m, H, gm, gH, 0.3)

    mu = mu0.clone().requires_grad_(True)
    lv = lv0.clone().requires_grad_(True)
    z = z0.clone().requires_grad_(True)
    m_r = m.clone().requires_grad_(True)
    H_r = H.clone().requires_grad_(True)
    kld, gm, gH = ref_call(m_r, H_r, mu, lv, z)
    kld = kld.reshape(())
    kld.backward()
    g0os, g0ls = extract_grads(k0)
    g1os, g1ls = extract_grads(k1)

    # ---- oracle on the same inputs
    prm0 = orc.KernelParams(ros0.clone(), rls0.clone()).requires_grad_()
    prm1 = orc.KernelParams(ros1.clone(), rls1.clone()).requires_grad_()
    mu_o = mu0.clone().requires_grad_(True); lv_o = lv0.clone().requires_grad_(True)
    z_o = z0.clone().requires_grad_(True)
    m_o = m.clone().requires_grad_(True); H_o = H.clone().requires_grad_(True)
    if fixed_T_api:
        okld, ogm, ogH, terms = orc.minibatch_KLD_upper_bound(spec0, prm0, spec1, prm1, noise, m_o, H_o, x, mu_o, lv_o,
                                                              z_o, P_tot, n_subj, T, natural_gradient, eps, True)
    else:
        okld, ogm, ogH, terms = orc.minibatch_KLD_upper_bound_iter(spec0, prm0, spec1, prm1, noise, m_o, H_o, x, mu_o,
                                                                   lv_o, z_o, P_tot, n_subj, N_tot, natural_gradient,
                                                                   idc, eps, True)
    okld.backward()
    worst = {}
    worst['kld'] = check(name + ".kld", okld, kld)
    if natural_gradient:
        worst['grad_m'] = check(name + ".grad_m", ogm, gm, 2e-6)
        worst['grad_H'] = check(name + ".grad_H", ogH, gH, 2e-6)
    worst['d_mu'] = check(name + ".d_mu", mu_o.grad, mu.grad, 2e-6)
    worst['d_logv'] = check(name + ".d_logv", lv_o.grad, lv.grad, 2e-6)
    worst['d_z'] = check(name + ".d_z", z_o.grad, z.grad, 2e-6)
    worst['d_m'] = check(name + ".d_m", m_o.grad, m_r.grad, 2e-6)
    worst['d_H'] = check(name + ".d_H", H_o.grad, H_r.grad, 2e-6)
    worst['d_os0'] = check(name + ".d_os0", prm0.raw_outputscale.grad, g0os, 1e-4)
    if rls0.numel():
        worst['d_ls0'] = check(name + ".d_ls0", prm0.raw_lengthscale.grad, g0ls, 1e-4)
    worst['d_os1'] = check(name + ".d_os1", prm1.raw_outputscale.grad, g1os, 1e-4)
    if rls1.numel():
        worst['d_ls1'] = check(name + ".d_ls1", prm1.raw_lengthscale.grad, g1ls, 1e-4)

    # reference dense kernel matrices as extra goldens for the kernel-evaluation op
    with torch.no_grad():
        K0xz = k0(x, z0).evaluate()
        K0zz = k0(z0, z0).evaluate()
        K1xx = k1(x, x).evaluate()
        check(name + ".K0xz", orc.eval_additive(spec0, prm0, x, z0), K0xz, 1e-11)
        check(name + ".K1xx", orc.eval_additive(spec1, prm1, x, x), K1xx, 1e-11)

    npz = dict(
        kargs=repr(kargs), L=L, M=M, T=T, fixed_T_api=int(fixed_T_api), n_subj=n_subj, P_tot=P_tot, N_tot=N_tot,
        eps=eps, natural_gradient=int(natural_gradient),
        x=x.numpy(), lens=np.array(lens), z=z0.numpy(), mu=mu0.numpy(), log_v=lv0.numpy(), m=m.numpy(), H=H.numpy(),
        noise=noise.numpy(), ros0=ros0.numpy(), rls0=rls0.numpy(), ros1=ros1.numpy(), rls1=rls1.numpy(),
        kld=kld.detach().numpy(),
        grad_m=(gm.detach().numpy() if natural_gradient else np.zeros(0)),
        grad_H=(gH.detach().numpy() if natural_gradient else np.zeros(0)),
        d_mu=mu.grad.numpy(), d_logv=lv.grad.numpy(), d_z=z.grad.numpy(), d_m=m_r.grad.numpy(), d_H=H_r.grad.numpy(),
        d_os0=g0os.numpy(), d_ls0=g0ls.numpy(), d_os1=g1os.numpy(), d_ls1=g1ls.numpy(),
        K0xz=K0xz.numpy(), K0zz=K0zz.numpy(), K1xx=K1xx.numpy(),
        # individual terms are not returned by the reference; these come from the (just validated) oracle
        **{"term_" + k: terms[k].detach().numpy() for k in ('A', 'B', 'C', 'D', 'E', 'F', 'kld_qu_pu', 'S', 'p')},
    )
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **npz)
    print(f"  {name}: N_b={N_b} kld={float(kld):.6e} worst rel diff {max(worst.values()):.2e}")


# ------------------------------------------------------------------ likelihoods
def loglik_case(name, types, N, seed, conv=False, observed=0.7):
    import HLVAE as ref_hlvae                               # reference
    from HL_VAE import read_functions as ref_rf             # reference
    from HL_VAE.utils import batch_normalization as ref_bn  # reference
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    data, mask = synth.likelihood_batch(types, N, rng, observed=observed, pixel_like=conv)
    tinfo = orc.types_info_from_layout(types, conv=conv)
    descs, E_x, P_th = orc.build_layout(types)
    D = len(types)
    if conv:
        assert D == 1296
        dims = [D, [16], 4, [16], 5]
    else:
        dims = [E_x, [16], 4, [16], 3]
    model = ref_hlvae.HLVAE(dims, tinfo, D, vy_init=[1., .5], vy_fixed=False, logvar_network=False, conv=conv).double()
    with torch.no_grad():
        model._log_vy_real.add_(0.5 * torch.randn(model._log_vy_real.shape, generator=gen, dtype=DT))
        model._log_vy_pos.add_(0.5 * torch.randn(model._log_vy_pos.shape, generator=gen, dtype=DT))
    theta0 = torch.randn(N, P_th, generator=gen, dtype=DT) * 1.5
    if N >= 4 and any(k == 'cat' for k, _ in types):        # exact ties -> first index must win
        j = next(i for i, v in enumerate(descs) if v.kind == 'cat')
        theta0[0, descs[j].theta_col:descs[j].theta_col + descs[j].nclass] = 0.25
        theta0[1, descs[j].theta_col:descs[j].theta_col + descs[j].nclass] = 0.0
    param_mask = torch.ones(N, P_th, dtype=DT)
    _, norm = ref_bn(data, mask, param_mask, tinfo)
    theta = theta0.clone().requires_grad_(True)
    lpx, lpm, samples, params = model.loglik_and_reconstruction(theta, data, mask, param_mask, norm)
    g_up = torch.randn(N, D, generator=gen, dtype=DT)       # arbitrary upstream gradient
    (lpx * g_up).sum().backward()
    pcat = ref_rf.p_params_concatenation_by_key([{'x': params}], tinfo, N, data.device, 'x').detach()
    dtr = ref_rf.discrete_variables_transformation(data, tinfo)
    rmean, rmode = ref_rf.statistics(pcat, tinfo, data.device, conv, [model._log_vy_real, model._log_vy_pos])
    rmean, rmode = rmean.detach(), rmode.detach()

    # ---- oracle
    lvr = model._log_vy_real.detach().clone().requires_grad_(True)
    lvp = model._log_vy_pos.detach().clone().requires_grad_(True)
    nr = None if (conv or norm[0] == []) else (norm[0][0], norm[0][1])
    npos = None if norm[1] == [] else (norm[1][0], norm[1][1])
    th_o = theta0.clone().requires_grad_(True)
    olpx, olpm, oparams = orc.loglik_and_reconstruction(descs, data, mask, th_o, lvr, lvp, nr, npos, conv)
    (olpx * g_up).sum().backward()
    omean, omode = orc.statistics(descs, oparams.detach(), lvp.detach())
    odtr = orc.discrete_variables_transformation(descs, data)
    w = [check(name + ".log_p_x", olpx, lpx, 1e-11), check(name + ".log_p_x_missing", olpm, lpm, 1e-11),
         check(name + ".params", oparams, pcat, 1e-11), check(name + ".d_theta", th_o.grad, theta.grad, 1e-10)]
    disc = torch.tensor([v.kind in ('cat', 'ordinal') for v in descs])
    assert torch.equal(omean[:, disc], rmean[:, disc]), name + ": argmax imputation differs"
    assert torch.equal(odtr, dtr), name + ": discrete transform differs"
    w.append(check(name + ".mean", omean, rmean, 1e-11))
    w.append(check(name + ".mode", omode, rmode, 1e-11))
    if model._log_vy_real.grad is not None and lvr.grad is not None:
        w.append(check(name + ".d_log_vy_real", lvr.grad, model._log_vy_real.grad, 1e-10))
    if model._log_vy_pos.grad is not None and lvp.grad is not None:
        w.append(check(name + ".d_log_vy_pos", lvp.grad, model._log_vy_pos.grad, 1e-10))
    z = lambda t: np.zeros(0) if t is None else t.detach().numpy()
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        types=np.array([f"{k}:{c}" for k, c in types]), conv=int(conv),
        data=data.numpy(), mask=mask.numpy(), theta=theta0.numpy(), g_up=g_up.numpy(),
        log_vy_real=z(model._log_vy_real), log_vy_pos=z(model._log_vy_pos),
        norm_real_mean=z(nr[0] if nr else None), norm_real_var=z(nr[1] if nr else None),
        norm_pos_mean=z(npos[0] if npos else None), norm_pos_var=z(npos[1] if npos else None),
        log_p_x=lpx.detach().numpy(), log_p_x_missing=lpm.detach().numpy(), params=pcat.numpy(),
        d_theta=theta.grad.numpy(), d_log_vy_real=z(model._log_vy_real.grad), d_log_vy_pos=z(model._log_vy_pos.grad),
        recon_mean=rmean.numpy(), recon_mode=rmode.numpy(), data_transformed=dtr.numpy())
    print(f"  {name}: N={N} D={D} E_x={E_x} worst rel diff {max(w):.2e}")


def loglik_logvar_case(name, types, N, seed, observed=0.7):
    """HLVAE.loglik_and_reconstruction (HLVAE.py:381-414), p_params_concatenation_by_key and statistics
    (read_functions.py:206-218,268-302) of the unmodified reference for a model WITH the variance network
    (logvar_network=True): model-level fixture; the per-type arithmetic is pinned by loglik_aux_cases."""
    import HLVAE as ref_hlvae                               # reference
    from HL_VAE import read_functions as ref_rf             # reference
    from HL_VAE.utils import batch_normalization as ref_bn  # reference
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    data, mask = synth.likelihood_batch(types, N, rng, observed=observed)
    tinfo = orc.types_info_from_layout(types, conv=False, logvar_network=True)
    descs, E_x, _ = orc.build_layout(types)
    P_th, D = len(tinfo['param_indexes']), len(types)
    model = ref_hlvae.HLVAE([E_x, [16], 4, [16], 3], tinfo, D, vy_init=[1., .5], vy_fixed=False, logvar_network=True,
                            conv=False).double()
    theta0 = torch.randn(N, P_th, generator=gen, dtype=DT) * 1.2
    param_mask = torch.ones(N, P_th, dtype=DT)
    _, norm = ref_bn(data, mask, param_mask, tinfo)
    theta = theta0.clone().requires_grad_(True)
    lpx, lpm, samples, params = model.loglik_and_reconstruction(theta, data, mask, param_mask, norm)
    g_up = torch.randn(N, D, generator=gen, dtype=DT)
    (lpx * g_up).sum().backward()
    pcat = ref_rf.p_params_concatenation_by_key([{'x': params}], tinfo, N, data.device, 'x').detach()
    rmean, rmode = ref_rf.statistics(pcat, tinfo, data.device, False, [model._log_vy_real, model._log_vy_pos])
    z = lambda t: np.zeros(0) if (t is None or t == []) else t.detach().numpy()
    nr, npos = norm[0], norm[1]
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"), types=np.array([f"{k}:{c}" for k, c in types]), conv=0, logvar_network=1,
        data=data.numpy(), mask=mask.numpy(), theta=theta0.numpy(), g_up=g_up.numpy(),
        norm_real_mean=z(nr[0] if nr != [] else None), norm_real_var=z(nr[1] if nr != [] else None),
        norm_pos_mean=z(npos[0] if npos != [] else None), norm_pos_var=z(npos[1] if npos != [] else None),
        log_p_x=lpx.detach().numpy(), log_p_x_missing=lpm.detach().numpy(), params=pcat.numpy(),
        d_theta=theta.grad.numpy(), recon_mean=rmean.detach().numpy(), recon_mode=rmode.detach().numpy())
    print(f"  {name}: N={N} D={D} P_theta={P_th} (reference outputs frozen)")


def loglik_aux_cases():
    """The likelihood branches outside the default configuration, function level, unmodified reference:
    loglik_real / loglik_pos with extra_params = None (variance network, HL_VAE/loglik.py:45-48,104-108) and
    loglik_beta (:216-256, both for a theta wide enough for the first indexing and for the fallback)."""
    from HL_VAE import loglik as ref_ll                     # reference
    print("likelihood branches (variance network, beta): oracle vs unmodified reference")
    gen = torch.Generator().manual_seed(51)
    torch.manual_seed(51)
    out, w = {}, []
    N, D = 14, 5
    mask = (torch.rand(N, D, generator=gen) < 0.7).to(DT)
    g_up = torch.randn(N, D, generator=gen, dtype=DT)

    def run(tag, fn, ofn, data, theta0, norm, extra=None, oargs=()):
        th = theta0.clone().requires_grad_(True)
        ex = None if extra is None else extra.clone().requires_grad_(True)
        r = fn([data, mask[:, :data.shape[1]]], (tag.split("_")[0], '1'), th, norm, ex)
        (r['log_p_x'] * g_up[:, :data.shape[1]]).sum().backward()
        th_o = theta0.clone().requires_grad_(True)
        ex_o = None if extra is None else extra.clone().requires_grad_(True)
        o = ofn(data, mask[:, :data.shape[1]], th_o, *oargs) if extra is None else \
            ofn(data, mask[:, :data.shape[1]], th_o, *oargs, ex_o)
        (o[0] * g_up[:, :data.shape[1]]).sum().backward()
        w.append(check(tag + ".log_p_x", o[0], r['log_p_x'], 1e-12))
        w.append(check(tag + ".log_p_x_missing", o[1], r['log_p_x_missing'], 1e-12))
        w.append(check(tag + ".prm_a", o[2], r['params'][0], 1e-12))
        w.append(check(tag + ".prm_b", o[3], r['params'][1], 1e-12))
        w.append(check(tag + ".d_theta", th_o.grad, th.grad, 1e-11))
        out.update({tag + "_data": data.numpy(), tag + "_theta": theta0.numpy(), tag + "_log_p_x": r['log_p_x'].detach().numpy(),
                    tag + "_log_p_x_missing": r['log_p_x_missing'].detach().numpy(),
                    tag + "_prm_a": r['params'][0].detach().numpy(), tag + "_prm_b": r['params'][1].detach().numpy(),
                    tag + "_d_theta": th.grad.numpy()})
        if extra is not None:
            w.append(check(tag + ".d_disp", ex_o.grad, ex.grad, 1e-11))
            out[tag + "_disp"] = extra.numpy()
            out[tag + "_d_disp"] = ex.grad.numpy()

    data_r = torch.randn(N, D, generator=gen, dtype=DT)
    th2 = torch.randn(N, 2 * D, generator=gen, dtype=DT)
    th2[:, D:] = th2[:, D:] * 2.0 - 3.0
    nm, nv = torch.randn(D, generator=gen, dtype=DT), torch.rand(D, generator=gen, dtype=DT) + 0.2
    nv[0] = 1e-5                                             # below the clamp
    run("real_norm", ref_ll.loglik_real, orc.loglik_real_rowvar, data_r, th2, [nm, nv], oargs=(nm, nv))
    out["real_norm_nm"], out["real_norm_nv"] = nm.numpy(), nv.numpy()
    run("real_plain", ref_ll.loglik_real, orc.loglik_real_rowvar, data_r, th2, [])
    data_p = torch.exp(torch.randn(N, D, generator=gen, dtype=DT))
    run("pos_norm", ref_ll.loglik_pos, orc.loglik_pos_rowvar, data_p, th2, [nm, nv], oargs=(nm, nv))
    # beta: ranges as read_functions.py:119 builds them ([min, max + 1e-3]); dispersion parameter of HLVAE.py:226
    for tag, Db, width in (("beta_wide", 3, 6), ("beta_fallback2", 2, 2), ("beta_fallback3", 3, 3)):
        lo = torch.rand(Db, generator=gen, dtype=DT) * 2 - 1
        hi = lo + 0.5 + torch.rand(Db, generator=gen, dtype=DT) * 3
        xb = lo + (hi - lo) * (0.02 + 0.96 * torch.rand(N, Db, generator=gen, dtype=DT))
        rng_np = np.concatenate([[float(lo[i]), float(hi[i]) + 1e-3] for i in range(Db)])
        thb = torch.randn(N, width, generator=gen, dtype=DT) * 0.8
        disp = torch.tensor([1.0 + 0.7 * float(torch.randn((), generator=gen))], dtype=DT)
        run(tag, ref_ll.loglik_beta, orc.loglik_beta, xb, thb, rng_np, extra=disp,
            oargs=(torch.as_tensor(rng_np.reshape(Db, 2)),))
        out[tag + "_ranges"] = rng_np
    out["mask"], out["g_up"] = mask.numpy(), g_up.numpy()
    np.savez_compressed(os.path.join(GOLD, "loglik_aux.npz"), **out)
    print(f"  loglik_aux: worst rel diff {max(w):.2e}")
    loglik_logvar_case("loglik_logvar_mixed", synth.mixed_types(np.random.default_rng(53), 16), N=18, seed=53)


# ------------------------------------------------------------------ observation heads (y -> theta)
HEAD_PARAM_NAMES = {'count': ('weight', 'bias'), 'real': ('weight_mean', 'bias_mean'), 'pos': ('weight_mean', 'bias_mean'),
                    'cat': ('weight', 'bias'), 'ordinal': ('weight_thresholds', 'weight_region', 'bias_region')}


def theta_case(name, types, N, seed, conv=False, observed=0.7, logvar_network=False):
    """HLVAE.theta_estimation of the unmodified reference (HLVAE.py:416-453) on a seeded y, with an arbitrary
    upstream gradient: theta, d/dy and the gradient of every Observation_* parameter."""
    import HLVAE as ref_hlvae                               # reference
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    tinfo = orc.types_info_from_layout(types, conv=conv, logvar_network=logvar_network)
    descs, E_x, P_th = orc.build_layout(types)
    P_th = len(tinfo['param_indexes'])
    D = len(types)
    dims = [D, [16], 4, [16], 5] if conv else [E_x, [16], 4, [16], 3]
    Y = dims[4]
    model = ref_hlvae.HLVAE(dims, tinfo, D, vy_init=[1., .5], vy_fixed=False, logvar_network=logvar_network,
                            conv=conv).double()
    with torch.no_grad():                                   # move the heads away from their near-zero initial state
        for prm in model.obs_layer.parameters():
            prm.add_(0.4 * torch.randn(prm.shape, generator=gen, dtype=DT))
    if conv:                                                # y_grouped is a permuted view of [N, Y, D] (:341-342)
        y0 = (torch.randn(N, Y, D, generator=gen, dtype=DT) * 1.5).permute(0, 2, 1)
    else:
        y0 = torch.randn(N, D, Y, generator=gen, dtype=DT) * 1.5
    mask = (torch.rand(N, D, generator=gen, dtype=DT) < observed).to(DT)
    pm = torch.zeros(N, P_th, dtype=DT)
    for i, tpl in enumerate(tinfo['set_of_types']):        # read_functions.py:172-185
        mg = mask[:, torch.tensor(tinfo['data_types_indexes'] == i)]
        pc = torch.nonzero(torch.tensor(tinfo['param_indexes'] == i))[:, 0]
        if tpl[0] in ('real', 'pos') and logvar_network:
            pm[:, pc] = torch.cat([mg, mg], 1)
        else:
            pm[:, pc] = mg.repeat_interleave(int(tpl[1]) if tpl[0] in ('cat', 'ordinal') else 1, dim=1)
    g_up = torch.randn(N, P_th, generator=gen, dtype=DT)
    y = y0.clone().requires_grad_(True)
    theta = model.theta_estimation(y, mask, pm)
    (theta * g_up).sum().backward()

    # ---- oracle
    def names_of(kind):
        extra = ('weight_logvar', 'bias_logvar') if (logvar_network and kind in ('real', 'pos')) else ()
        return HEAD_PARAM_NAMES[kind] + extra

    heads, layer = [], 0
    for i, tpl in enumerate(tinfo['set_of_types']):
        mod = model.obs_layer[layer]
        heads.append({n: getattr(mod, n).detach().clone().requires_grad_(True) for n in names_of(tpl[0])})
        layer += 2 if (tpl[0] == 'real' and conv) else 1
    y_o = y0.clone().requires_grad_(True)
    th_o = orc.theta_estimation(types, heads, y_o, mask, conv=conv, logvar_network=logvar_network)
    (th_o * g_up).sum().backward()
    w = [check(name + ".theta", th_o, theta, 1e-13), check(name + ".d_y", y_o.grad, y.grad, 1e-12)]
    out = dict(types=np.array([f"{k}:{c}" for k, c in types]), conv=int(conv), y=y0.contiguous().numpy(),
               mask=mask.numpy(), g_up=g_up.numpy(), theta=theta.detach().numpy(), d_y=y.grad.contiguous().numpy(),
               n_groups=len(heads), logvar_network=int(logvar_network))
    layer = 0
    for i, tpl in enumerate(tinfo['set_of_types']):
        mod = model.obs_layer[layer]
        for n in names_of(tpl[0]):
            gref = getattr(mod, n).grad
            w.append(check(f"{name}.d_{tpl[0]}{tpl[1]}.{n}", heads[i][n].grad, gref, 1e-11))
            out[f"g{i}_{n}"] = getattr(mod, n).detach().numpy()
            out[f"d_g{i}_{n}"] = gref.numpy()
        out[f"g{i}_kind"] = np.array(f"{tpl[0]}:{tpl[1]}")
        layer += 2 if (tpl[0] == 'real' and conv) else 1
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"  {name}: N={N} D={D} P_theta={P_th} worst rel diff {max(w):.2e}")


def theta_cases():
    print("observation heads (theta_estimation): oracle vs unmodified reference")
    rng = np.random.default_rng(21)
    theta_case("theta_mixed", synth.mixed_types(rng, 24), N=24, seed=21)
    theta_case("theta_logvar_mixed", synth.mixed_types(np.random.default_rng(23), 18), N=20, seed=23, logvar_network=True)
    theta_case("theta_tabular_small", [('count', 1)] * 3 + [('ordinal', 5)] * 3 + [('cat', 5)] * 3 + [('real', 1)] * 2 +
               [('pos', 1)] * 2, N=16, seed=22)
    theta_case("theta_conv_d4", synth.HEALTHMNIST_D4_TYPES, N=3, seed=23, conv=True, observed=0.75)


# ------------------------------------------------------------------ batch normalisation
def norm_case(name, types, N, seed, conv=False, observed=0.7):
    """HL_VAE.utils.batch_normalization of the unmodified reference (HL_VAE/utils.py:88-143)."""
    from HL_VAE.utils import batch_normalization as ref_bn  # reference
    rng = np.random.default_rng(seed)
    data, mask = synth.likelihood_batch(types, N, rng, observed=observed, pixel_like=conv)
    tinfo = orc.types_info_from_layout(types, conv=conv)
    descs, E_x, P_th = orc.build_layout(types)
    X, norm = ref_bn(data, mask, torch.ones(N, P_th, dtype=DT), tinfo)
    Xo, nr, npos = orc.batch_normalization(descs, data, mask, conv)
    w = [check(name + ".X", Xo, X, 1e-14)]
    for tag, mine, ref in (("real", nr, norm[0]), ("pos", npos, norm[1])):
        assert (mine is None) == (ref == []), name + ": parameter presence differs for " + tag
        if mine is not None:
            w.append(check(f"{name}.{tag}_mean", mine[0], ref[0], 1e-14))
            w.append(check(f"{name}.{tag}_var", mine[1], ref[1], 1e-14))
    z = lambda t: np.zeros(0) if t is None else t.detach().numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), types=np.array([f"{k}:{c}" for k, c in types]), conv=int(conv),
                        data=data.numpy(), mask=mask.numpy(), X=X.numpy(),
                        real_mean=z(nr[0] if nr else None), real_var=z(nr[1] if nr else None),
                        pos_mean=z(npos[0] if npos else None), pos_var=z(npos[1] if npos else None))
    print(f"  {name}: N={N} D={len(types)} E_x={E_x} worst rel diff {max(w):.2e}")


def norm_cases():
    print("batch normalisation: oracle vs unmodified reference")
    rng = np.random.default_rng(31)
    norm_case("norm_mixed", synth.mixed_types(rng, 24), N=40, seed=31)
    norm_case("norm_tabular_small", [('count', 1)] * 3 + [('ordinal', 5)] * 3 + [('cat', 5)] * 3 + [('real', 1)] * 2 +
              [('pos', 1)] * 2, N=16, seed=32)
    norm_case("norm_conv_d4", synth.HEALTHMNIST_D4_TYPES, N=3, seed=33, conv=True, observed=0.75)


# ------------------------------------------------------------------ minibatch composition (samplers)
def sampler_cases():
    """Row indices of every minibatch of one epoch from the reference's own sampler classes (utils.py:36-97) under a
    fixed np.random seed.  torch >= 2.2 dropped Sampler.__init__(data_source): shimmed as in SURVEY.md 8(b)."""
    import torch.utils.data as tud
    from torch.utils.data import BatchSampler
    orig = tud.Sampler.__init__
    tud.Sampler.__init__ = lambda self, data_source=None: None
    try:
        import utils as ref_utils                               # reference
        print("samplers: oracle vs unmodified reference")
        flat = lambda bb: (np.concatenate([np.asarray(b, dtype=np.int64) for b in bb]),
                           np.cumsum([0] + [len(b) for b in bb]))
        P, T, bs, seed = 9, 5, 12, 41
        ds = [{'label': torch.tensor([0., 0., float(s), 0.])} for s in range(P) for _ in range(T)]
        np.random.seed(seed)
        ref = [list(b) for b in BatchSampler(ref_utils.SubjectSampler(ds, P, T), bs, False)]
        np.random.seed(seed)
        assert orc.fixed_T_batches(P, T, bs) == ref
        ids = [3] * 2 + [9] * 5 + [1] * 3 + [4] * 1 + [7] * 4 + [2] * 3 + [8] * 6
        ds2 = [{'label': torch.tensor([0., 0., float(s), 0.])} for s in ids]
        np.random.seed(seed + 1)
        ref2 = [list(b) for b in ref_utils.VaryingLengthBatchSampler(ref_utils.VaryingLengthSubjectSampler(ds2, 2), 3)]
        np.random.seed(seed + 1)
        assert orc.varying_T_batches(ids, 3) == ref2
        f1, f2 = flat(ref), flat(ref2)
        np.savez_compressed(os.path.join(GOLD, "samplers.npz"), fixed_P=P, fixed_T=T, fixed_batch=bs, fixed_seed=seed,
                            fixed_rows=f1[0], fixed_ptr=f1[1], var_ids=np.asarray(ids), var_batch=3, var_seed=seed + 1,
                            var_rows=f2[0], var_ptr=f2[1])
        print(f"  samplers: {len(ref)} fixed-T batches, {len(ref2)} varying-T batches identical")
    finally:
        tud.Sampler.__init__ = orig


# ------------------------------------------------------------------ GP posterior-mean prediction
def predict_case(name, kargs, L, M, n_subj, T, ragged, seed, n_test_subj=3, continuous_age=False):
    """utils.batch_predict_varying_T (and, with equal T, utils.batch_predict) of the unmodified reference.
    Both call `torch.solve(B, A)`, which current torch no longer has: it is shimmed as
    (torch.linalg.solve(A, B), None) for the duration of the call - the reference file itself is untouched."""
    import utils as ref_utils                                # reference (needs the matplotlib stand-in)
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    x, lens = synth.covariates(n_subj, T, rng, ragged=ragged, t_min=3, continuous_age=continuous_age)
    pool, _ = synth.covariates(40, T, rng, continuous_age=continuous_age)
    z = synth.inducing_points(torch.cat([x, pool]), L, M, rng)
    N = x.shape[0]
    mu = torch.randn(N, L, generator=gen, dtype=DT)
    k0, k1, lik = ref_modules(L, kargs, gen)
    k0.eval(); k1.eval(); lik.eval()
    idc = kargs['id_covariate']
    # test rows: unseen time points of the first n_test_subj subjects plus one subject absent from prediction_x
    starts = np.concatenate([[0], np.cumsum(lens)])
    rows = []
    for s_ in range(n_test_subj):
        xr = x[starts[s_]:starts[s_ + 1]].clone()
        xr[:, 0] += 0.5
        rows.append(xr[: max(2, len(xr) // 2)])
    extra, _ = synth.covariates(1, T, rng, first_id=10_000, continuous_age=continuous_age)
    rows.append(extra[:4])
    test_x = torch.cat(rows)
    eps = 1e-6
    had = hasattr(torch, "solve")
    old = getattr(torch, "solve", None)
    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)
    try:
        with torch.no_grad():
            zp = ref_utils.batch_predict_varying_T(L, k0, k1, lik, x, test_x, mu, z, idc, eps)
            zp_fixed = None
            if not ragged:
                zp_fixed = ref_utils.batch_predict(L, k0, k1, lik, x, test_x, mu, z, n_subj, T, idc, eps)
    finally:
        if had:
            torch.solve = old
        else:
            del torch.solve
    spec0, spec1 = orc.compile_spec(**kargs)
    ros0, rls0 = extract_params(k0)
    ros1, rls1 = extract_params(k1)
    noise = lik.noise_covar.noise.detach().reshape(-1).clone()
    prm0, prm1 = orc.KernelParams(ros0.clone(), rls0.clone()), orc.KernelParams(ros1.clone(), rls1.clone())
    with torch.no_grad():
        ozp = orc.batch_predict(spec0, prm0, spec1, prm1, noise, x, test_x, mu, z, orc.split_subjects_by_id(x, idc), idc, eps)
    w = [check(name + ".Z_pred", ozp, zp, 1e-7)]
    if zp_fixed is not None:
        w.append(check(name + ".Z_pred_fixedT", ozp, zp_fixed, 1e-7))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), kargs=repr(kargs), L=L, M=M, T=T, n_subj=n_subj, eps=eps,
                        ragged=int(ragged), x=x.numpy(), lens=np.array(lens), test_x=test_x.numpy(), mu=mu.numpy(),
                        z=z.numpy(), noise=noise.numpy(), ros0=ros0.numpy(), rls0=rls0.numpy(), ros1=ros1.numpy(),
                        rls1=rls1.numpy(), Z_pred=zp.numpy())
    print(f"  {name}: N={N} N_test={test_x.shape[0]} worst rel diff {max(w):.2e}")


def dubo_case(name, kargs, L, M, n_subj, T, seed, continuous_age=False):
    """validation.validation_dubo (validation.py:16-76) of the unmodified reference, `torch.solve` shimmed."""
    import validation as ref_validation                      # reference
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    x, lens = synth.covariates(n_subj, T, rng, continuous_age=continuous_age)
    pool, _ = synth.covariates(40, T, rng, continuous_age=continuous_age)
    z = synth.inducing_points(torch.cat([x, pool]), L, M, rng)
    N = x.shape[0]
    mu = torch.randn(N, L, generator=gen, dtype=DT)
    lv = -3.0 * torch.rand(N, L, generator=gen, dtype=DT)
    k0, k1, lik = ref_modules(L, kargs, gen)
    k0.eval(); k1.eval(); lik.eval()
    eps = 1e-6
    had, old = hasattr(torch, "solve"), getattr(torch, "solve", None)
    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)
    try:
        with torch.no_grad():
            d = ref_validation.validation_dubo(L, k0, k1, lik, x, mu, lv, z, n_subj, T, eps)
    finally:
        if had:
            torch.solve = old
        else:
            del torch.solve
    spec0, spec1 = orc.compile_spec(**kargs)
    ros0, rls0 = extract_params(k0)
    ros1, rls1 = extract_params(k1)
    noise = lik.noise_covar.noise.detach().reshape(-1).clone()
    prm0, prm1 = orc.KernelParams(ros0.clone(), rls0.clone()), orc.KernelParams(ros1.clone(), rls1.clone())
    with torch.no_grad():
        od = orc.validation_dubo(spec0, prm0, spec1, prm1, noise, x, mu, lv, z, n_subj, T, eps)
    w = check(name + ".dubo", od, d, 1e-8)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), kargs=repr(kargs), L=L, M=M, T=T, n_subj=n_subj, eps=eps,
                        x=x.numpy(), mu=mu.numpy(), log_v=lv.numpy(), z=z.numpy(), noise=noise.numpy(), ros0=ros0.numpy(),
                        rls0=rls0.numpy(), ros1=ros1.numpy(), rls1=rls1.numpy(), dubo=d.numpy())
    print(f"  {name}: N={N} dubo={float(d):.6e} rel diff {w:.2e}")


def legacy_bound_case(name, kargs, M, n_subj, T, seed, continuous_age=False):
    """elbo_functions.deviance_upper_bound (:60-115) and elbo_functions.elbo (:9-57) of the unmodified reference on
    UN-BATCHED kernel objects (kernel_gen.generate_kernel_approx, :97-197), one latent dimension; `torch.solve`
    shimmed."""
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    x, lens = synth.covariates(n_subj, T, rng, continuous_age=continuous_age)
    pool, _ = synth.covariates(40, T, rng, continuous_age=continuous_age)
    z = synth.inducing_points(torch.cat([x, pool]), 1, M, rng)[0]
    N = x.shape[0]
    mu = torch.randn(N, generator=gen, dtype=DT)
    lv = -3.0 * torch.rand(N, generator=gen, dtype=DT)
    k0, k1 = ref_kernel_gen.generate_kernel_approx(kargs['cat_kernel'], kargs['bin_kernel'], kargs['sqexp_kernel'],
                                                   kargs['cat_int_kernel'], kargs['bin_int_kernel'],
                                                   kargs['covariate_missing_val'], kargs['id_covariate'])
    lik = gpytorch.likelihoods.GaussianLikelihood(noise_constraint=gpytorch.constraints.GreaterThan(1.0e-8))
    lik.noise = 1
    k0.double(); k1.double(); lik.double()
    with torch.no_grad():
        for k in (k0, k1):
            for p in k.parameters():
                p.add_(0.3 * torch.randn(p.shape, generator=gen, dtype=DT))
    k0.eval(); k1.eval(); lik.eval()
    eps = 1e-6
    had, old = hasattr(torch, "solve"), getattr(torch, "solve", None)
    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)
    try:
        with torch.no_grad():
            d = ref_elbo.deviance_upper_bound(k0, k1, lik, x, mu, lv, z, n_subj, T, eps)
            e = ref_elbo.elbo(k0, k1, lik, x, mu, z, n_subj, T, eps)
    finally:
        if had:
            torch.solve = old
        else:
            del torch.solve
    spec0, spec1 = orc.compile_spec(**kargs)
    def params_1(kmod):              # raw parameters of un-batched modules as [n, 1] (one latent dimension)
        ros = torch.stack([k.raw_outputscale.detach().reshape(1).clone() for k in kmod.kernels])
        rls = [mod.raw_lengthscale.detach().reshape(1).clone() for mod in kmod.modules()
               if isinstance(mod, gpytorch.kernels.RBFKernel)]
        return ros, (torch.stack(rls) if rls else torch.zeros(0, 1, dtype=DT))

    ros0, rls0 = params_1(k0)
    ros1, rls1 = params_1(k1)
    noise = lik.noise_covar.noise.detach().reshape(-1).clone()
    prm0, prm1 = orc.KernelParams(ros0.clone(), rls0.clone()), orc.KernelParams(ros1.clone(), rls1.clone())
    with torch.no_grad():
        od = orc.deviance_upper_bound(spec0, prm0, spec1, prm1, noise, x, mu, lv, z, n_subj, T, eps)
        oe = orc.elbo(spec0, prm0, spec1, prm1, noise, x, mu, z, n_subj, T, eps)
    w = max(check(name + ".dubo", od, d, 1e-8), check(name + ".elbo", oe, e, 1e-8))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), kargs=repr(kargs), M=M, T=T, n_subj=n_subj, eps=eps,
                        x=x.numpy(), mu=mu.numpy(), log_v=lv.numpy(), z=z.numpy(), noise=noise.numpy(), ros0=ros0.numpy(),
                        rls0=rls0.numpy(), ros1=ros1.numpy(), rls1=rls1.numpy(), dubo=d.numpy(), elbo=e.numpy())
    print(f"  {name}: N={N} dubo={float(d):.6e} elbo={float(e):.6e} rel diff {w:.2e}")


def legacy_bound_cases():
    print("un-batched deviance_upper_bound / elbo: oracle vs unmodified reference (torch.solve shimmed)")
    legacy_bound_case("legacy_bounds_default", synth.DEFAULT_KERNEL_ARGS, M=12, n_subj=6, T=8, seed=41)
    legacy_bound_case("legacy_bounds_sweep", synth.SWEEP_KERNEL_ARGS, M=16, n_subj=5, T=10, seed=42, continuous_age=True)


def predict_cases():
    print("GP posterior-mean prediction: oracle vs unmodified reference (torch.solve shimmed)")
    predict_case("predict_default_ragged", synth.DEFAULT_KERNEL_ARGS, L=4, M=12, n_subj=6, T=8, ragged=True, seed=21)
    predict_case("predict_default_fixedT", synth.DEFAULT_KERNEL_ARGS, L=3, M=10, n_subj=5, T=6, ragged=False, seed=22)
    predict_case("predict_sweep_ragged", synth.SWEEP_KERNEL_ARGS, L=3, M=16, n_subj=7, T=10, ragged=True, seed=23,
                 continuous_age=True)
    print("validation deviance upper bound: oracle vs unmodified reference (torch.solve shimmed)")
    dubo_case("dubo_default", synth.DEFAULT_KERNEL_ARGS, L=4, M=12, n_subj=6, T=8, seed=31)
    dubo_case("dubo_sweep", synth.SWEEP_KERNEL_ARGS, L=3, M=16, n_subj=5, T=10, seed=32, continuous_age=True)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "predict":      # only the prediction fixtures
        predict_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "samplers":
        sampler_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "aux":          # only the variance-network / beta likelihood fixtures
        loglik_aux_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "legacy":       # only the un-batched bound fixtures
        legacy_bound_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "norm":         # only the batch-normalisation fixtures
        norm_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "theta":        # only the observation-head fixtures
        theta_cases()
        return
    print("KL upper bound: oracle vs unmodified reference (gpytorch stand-in)")
    kl_case("kl_default_ragged", synth.DEFAULT_KERNEL_ARGS, L=4, M=12, n_subj=6, T=8, ragged=True, fixed_T_api=False, seed=1)
    kl_case("kl_default_fixedT", synth.DEFAULT_KERNEL_ARGS, L=3, M=10, n_subj=5, T=6, ragged=False, fixed_T_api=True, seed=2)
    kl_case("kl_sweep_ragged", synth.SWEEP_KERNEL_ARGS, L=4, M=16, n_subj=7, T=10, ragged=True, fixed_T_api=False, seed=3,
            continuous_age=True)
    kl_case("kl_masked_bin", synth.MASKED_KERNEL_ARGS, L=2, M=9, n_subj=5, T=7, ragged=True, fixed_T_api=False, seed=4)
    kl_case("kl_trained_like", synth.DEFAULT_KERNEL_ARGS, L=4, M=12, n_subj=8, T=8, ragged=False, fixed_T_api=False, seed=5,
            trained_like=True)
    kl_case("kl_not_natgrad", synth.DEFAULT_KERNEL_ARGS, L=2, M=8, n_subj=4, T=5, ragged=True, fixed_T_api=False, seed=6,
            natural_gradient=False)
    kl_case("kl_shuffled_rows", synth.DEFAULT_KERNEL_ARGS, L=2, M=8, n_subj=5, T=6, ragged=True, fixed_T_api=False, seed=7,
            shuffle_rows=True)
    kl_case("kl_T32_M40", synth.DEFAULT_KERNEL_ARGS, L=2, M=40, n_subj=3, T=32, ragged=False, fixed_T_api=True, seed=8)
    print("likelihoods: oracle vs unmodified reference")
    rng = np.random.default_rng(11)
    loglik_case("loglik_mixed", synth.mixed_types(rng, 24), N=24, seed=11)
    loglik_case("loglik_tabular_small", [('count', 1)] * 3 + [('ordinal', 5)] * 3 + [('cat', 5)] * 3 + [('real', 1)] * 2 + [('pos', 1)] * 2,
                N=16, seed=12)
    loglik_case("loglik_conv_d4", synth.HEALTHMNIST_D4_TYPES, N=3, seed=13, conv=True, observed=0.75)
    loglik_aux_cases()
    predict_cases()
    legacy_bound_cases()
    theta_cases()
    norm_cases()
    sampler_cases()
    print("all oracle-vs-reference checks passed; goldens written to", GOLD)


if __name__ == "__main__":
    main()
