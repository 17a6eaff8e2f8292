"""Shared test helpers: load golden fixtures, run the CUDA path and the oracle on the same inputs.

Only tests/, __graft_entry__.smoke() and bench.py's checker legs import this (it imports oracle/).
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

import hlvae_b200  # noqa: F401
from hlvae_b200 import elbo, kernels, likelihoods, loglik, subjects, synth, theta as theta_mod
from oracle import hlvae_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
DT = torch.float64

KL_CASES = ["kl_default_ragged", "kl_default_fixedT", "kl_sweep_ragged", "kl_masked_bin", "kl_trained_like",
            "kl_not_natgrad", "kl_shuffled_rows", "kl_T32_M40"]
LOGLIK_CASES = ["loglik_mixed", "loglik_tabular_small", "loglik_conv_d4"]
THETA_CASES = ["theta_mixed", "theta_tabular_small", "theta_conv_d4"]
NORM_CASES = ["norm_mixed", "norm_tabular_small", "norm_conv_d4"]


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def t(a, dev="cpu", dtype=DT):
    return torch.as_tensor(np.asarray(a), dtype=dtype).to(dev)


def rel_err(a, b, scale=None):
    a, b = torch.as_tensor(a, dtype=DT).detach().cpu(), torch.as_tensor(b, dtype=DT).detach().cpu()
    den = float(b.abs().max()) if scale is None else float(scale)
    return float((a - b).abs().max()) / (den + 1e-300)


def elem_err(a, b, rtol, afloor=1e-7):
    """Element-wise check for per-row gradients (d_mu, d_logv, d_theta): the worst ratio
    |a_i - b_i| / (rtol |b_i| + afloor max|b|); <= 1 passes.  Unlike rel_err (max-norm over the whole tensor) this
    constrains the small entries too; the absolute floor only covers entries ~1e7 times smaller than the largest."""
    a, b = torch.as_tensor(a, dtype=DT).cpu(), torch.as_tensor(b, dtype=DT).cpu()
    if b.numel() == 0:
        return 0.0
    den = rtol * b.abs() + afloor * float(b.abs().max()) + 1e-300
    return float(((a - b).abs() / den).max())


# ---------------------------------------------------------------------------- KL
def kl_terms_from_product(lt):
    """ELBO terms of elbo_functions.py:166-181 / :256-277 assembled from what the kernels emit
    (hlvae_b200.elbo.last_terms, config.keep_terms): per-latent scalars scal = [A, B + sum(iB * K0_st), C, F]
    (hlvae_kl_subject / hlvae_kl_panel), the M x M pre-stage scalars pre = [log det K, log det H, tr(iK H),
    m^T iK m] (hlvae_mxm_pre) and the statistics S.  The reference's D is sum(iB * K0_st) - tr(iK S), so B and D are
    compared as B + D."""
    scal, pre, S, iK, H = (lt[k].double().cpu() for k in ("scal", "pre", "S", "iK", "H"))
    L, M = S.shape[0], S.shape[-1]
    Ss = 0.5 * (S + S.transpose(-1, -2))
    out = dict(A=scal[:, 0].sum(), C=scal[:, 2].sum(), F=scal[:, 3].sum())
    out["B+D"] = scal[:, 1].sum() - (iK * Ss).sum()
    out["E"] = ((iK @ H @ iK) * Ss).sum()
    out["kld_qu_pu"] = 0.5 * (pre[:, 2].sum() + pre[:, 3].sum() - L * M + pre[:, 0].sum() - pre[:, 1].sum())
    out["S"], out["p"] = S, lt["p"].double().cpu()
    return out


def trace_scales(iK, H, S):
    """sum of the absolute addends of tr(iK S) and tr(iK H iK S)"""
    iK, H, S = (torch.as_tensor(a, dtype=DT) for a in (iK, H, S))
    return float((iK.abs() * S.abs()).sum()), float(((iK @ H @ iK).abs() * S.abs()).sum())


def assert_terms_close(got, ref, tol, label=""):
    """Per-term gate of BASELINE.md section 4.  |delta| <= tol * max(|term|, |largest ELBO term|) would let the huge
    kld_qu_pu hide everything at the initial state, so each term is held to tol relative to ITSELF - except the two
    traces against S, tr(iK S) in D and E = tr(iK H iK S): with cond(K0zz + eps I) ~ 1e7 two float64 inverses of the
    same K0zz differ by ~1e-9 and the traces are sums of cancelling addends, so those two get the scale-aware bound of
    SURVEY.md section 7: max(|term|, sum of the absolute addends) (`ref["scale_D"]`, `ref["scale_E"]`)."""
    errs, bad = {}, {}
    for k in ("A", "B+D", "C", "E", "F", "kld_qu_pu"):
        g, r = float(got[k]), float(ref[k])
        scale = max(abs(r), float(ref.get({"B+D": "scale_D", "E": "scale_E"}.get(k, ""), 0.0)), 1e-300)
        errs[k] = abs(g - r) / scale
        if errs[k] > tol:
            bad[k] = errs[k]
    for k in ("S", "p"):
        errs[k] = rel_err(got[k].reshape(-1), torch.as_tensor(ref[k]).reshape(-1))
        if errs[k] > tol:
            bad[k] = errs[k]
    assert not bad, f"{label} ELBO terms outside tolerance: {bad} (all: {errs})"
    return errs


def set_kernel_params(kmod, ros, rls):
    """Write raw_outputscale [ncomp, L] / raw_lengthscale [n_se, L] (depth-first SE order, the
    order the golden generator extracted them in) into a product kernel module."""
    with torch.no_grad():
        for r, k in enumerate(kmod.kernels):
            k.raw_outputscale.copy_(ros[r].reshape(k.raw_outputscale.shape))
        rb = [mod for mod in kmod.modules() if isinstance(mod, kernels.RBFKernel)]
        assert len(rb) == rls.shape[0]
        for j, mod in enumerate(rb):
            mod.raw_lengthscale.copy_(rls[j].reshape(mod.raw_lengthscale.shape))


def kernel_grads(kmod):
    gos = torch.stack([k.raw_outputscale.grad.detach().reshape(-1) for k in kmod.kernels]) if len(kmod.kernels) else None
    rb = [mod for mod in kmod.modules() if isinstance(mod, kernels.RBFKernel)]
    gls = torch.stack([mod.raw_lengthscale.grad.detach().reshape(-1) for mod in rb]) if rb else None
    return gos, gls


def build_product_kernels(kargs, L, dev, ros0, rls0, ros1, rls1, noise_value=None):
    k0, k1 = kernels.generate_kernel_batched(L, **kargs)
    k0, k1 = k0.to(dev).double(), k1.to(dev).double()
    set_kernel_params(k0, t(ros0, dev), t(rls0, dev))
    set_kernel_params(k1, t(ros1, dev), t(rls1, dev))
    lik = likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]),
                                         noise_constraint=likelihoods.GreaterThan(1.0e-8)).to(dev).double()
    if noise_value is not None:
        lik.noise = t(noise_value, dev).reshape(L, 1)
    lik.raw_noise.requires_grad = False
    return k0, k1, lik


def run_kl_golden(name, dev, storage=torch.float64, layout="auto"):
    g = load(name)
    kargs = ast.literal_eval(str(g["kargs"]))
    L, M, T = int(g["L"]), int(g["M"]), int(g["T"])
    k0, k1, lik = build_product_kernels(kargs, L, dev, g["ros0"], g["rls0"], g["ros1"], g["rls1"], g["noise"])
    x = t(g["x"], dev)
    mu = t(g["mu"], dev).to(storage).requires_grad_(True)
    lv = t(g["log_v"], dev).to(storage).requires_grad_(True)
    z = t(g["z"], dev).requires_grad_(True)
    m = t(g["m"], dev).requires_grad_(True)
    H = t(g["H"], dev).requires_grad_(True)
    ng = bool(int(g["natural_gradient"]))
    n_subj, P_tot, N_tot, eps = int(g["n_subj"]), int(g["P_tot"]), int(g["N_tot"]), float(g["eps"])
    if int(g["fixed_T_api"]):
        kld, gm, gH = elbo.minibatch_KLD_upper_bound(k0, k1, lik, L, m, H, x, mu, lv, z, P_tot, n_subj, T, ng, eps)
    else:
        lay = None
        if layout == "lengths":
            lay = subjects.SubjectLayout.from_lengths(g["lens"].tolist(), dev)
        kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(k0, k1, lik, L, m, H, x, mu, lv, z, P_tot, n_subj, N_tot, ng,
                                                          kargs["id_covariate"], eps, layout=lay)
    kld.sum().backward()
    gos0, gls0 = kernel_grads(k0)
    gos1, gls1 = kernel_grads(k1)
    got = dict(kld=kld.detach().reshape(()), grad_m=gm, grad_H=gH, d_mu=mu.grad, d_logv=lv.grad, d_z=z.grad,
               d_m=m.grad, d_H=H.grad, d_os0=gos0, d_ls0=gls0, d_os1=gos1, d_ls1=gls1)
    return dict(golden=g, got=got, ng=ng)


def assert_kl_close(r, tol=5e-6, hyper_tol=1e-4, label=""):
    g, got = r["golden"], r["got"]
    errs = {}
    for key in ("kld", "d_mu", "d_logv", "d_z", "d_m", "d_H"):
        errs[key] = rel_err(got[key], g[key])
    if r["ng"]:
        for key in ("grad_m", "grad_H"):
            errs[key] = rel_err(got[key], g[key])
    bad = {k: v for k, v in errs.items() if v > tol}
    for key in ("d_mu", "d_logv"):                       # per-row gradients: element-wise as well
        errs[key + "/elem"] = elem_err(got[key], g[key], rtol=20 * tol)
        if errs[key + "/elem"] > 1.0:
            bad[key + "/elem"] = errs[key + "/elem"]
    for key in ("d_os0", "d_ls0", "d_os1", "d_ls1"):
        if g[key].size:
            errs[key] = rel_err(got[key], g[key])
            if not _hyper_ok(got[key], g[key], g["kld"], hyper_tol):
                bad[key] = errs[key]
    assert not bad, f"{label} KL terms outside tolerance: {bad} (all: {errs})"
    return errs


def oracle_kl(kargs, L, x, mu, lv, z, m, H, ros0, rls0, ros1, rls1, noise, P_tot, n_subj, N_tot, eps, fixed_T=None):
    """Oracle value + gradients on CPU for arbitrary inputs (float64)."""
    spec0, spec1 = orc.compile_spec(**kargs)
    c = lambda a: a.detach().cpu().to(DT).clone()
    prm0 = orc.KernelParams(c(ros0), c(rls0)).requires_grad_()
    prm1 = orc.KernelParams(c(ros1), c(rls1)).requires_grad_()
    mu_, lv_, z_, m_, H_ = (c(a).requires_grad_(True) for a in (mu, lv, z, m, H))
    if fixed_T is None:
        kld, gm, gH, terms = orc.minibatch_KLD_upper_bound_iter(spec0, prm0, spec1, prm1, c(noise), m_, H_, c(x), mu_, lv_,
                                                                z_, P_tot, n_subj, N_tot, True, kargs["id_covariate"],
                                                                eps, True)
    else:
        kld, gm, gH, terms = orc.minibatch_KLD_upper_bound(spec0, prm0, spec1, prm1, c(noise), m_, H_, c(x), mu_, lv_, z_,
                                                           P_tot, n_subj, fixed_T, True, eps, True)
    kld.backward()
    return dict(kld=kld.detach(), grad_m=gm.detach(), grad_H=gH.detach(), d_mu=mu_.grad, d_logv=lv_.grad, d_z=z_.grad,
                d_m=m_.grad, d_H=H_.grad, d_os0=prm0.raw_outputscale.grad, d_ls0=prm0.raw_lengthscale.grad,
                d_os1=prm1.raw_outputscale.grad, d_ls1=prm1.raw_lengthscale.grad, terms=terms)


def make_kl_inputs(L, M, n_subj, T, seed, ragged=False, kargs=None, continuous_age=False, distinct_z=False):
    kargs = kargs or synth.DEFAULT_KERNEL_ARGS
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    x, lens = synth.covariates(n_subj, T, rng, ragged=ragged, t_min=3, continuous_age=continuous_age)
    pool, _ = synth.covariates(max(40, 2 * M // T + 2), T, rng, continuous_age=continuous_age)
    z = synth.inducing_points(torch.cat([x, pool]), L, M, rng)
    if distinct_z:      # well-conditioned variant: jitter the continuous columns so no two K0 rows coincide
        z[:, :, 0] += torch.rand(L, M, generator=gen, dtype=DT)
        z[:, :, 1] += torch.rand(L, M, generator=gen, dtype=DT)
    N_b = x.shape[0]
    mu = torch.randn(N_b, L, generator=gen, dtype=DT)
    lv = -3.0 * torch.rand(N_b, L, generator=gen, dtype=DT)
    m, H = synth.variational_state(L, M, gen)
    spec0, spec1 = orc.compile_spec(**kargs)
    p0, p1 = orc.KernelParams.default(spec0, L), orc.KernelParams.default(spec1, L)
    ros0 = p0.raw_outputscale + 0.3 * torch.randn(p0.raw_outputscale.shape, generator=gen, dtype=DT)
    rls0 = p0.raw_lengthscale + 0.3 * torch.randn(p0.raw_lengthscale.shape, generator=gen, dtype=DT)
    ros1 = p1.raw_outputscale + 0.3 * torch.randn(p1.raw_outputscale.shape, generator=gen, dtype=DT)
    rls1 = p1.raw_lengthscale + 0.3 * torch.randn(p1.raw_lengthscale.shape, generator=gen, dtype=DT)
    noise = torch.ones(L, dtype=DT)
    return dict(kargs=kargs, x=x, lens=lens, z=z, mu=mu, lv=lv, m=m, H=H, ros0=ros0, rls0=rls0, ros1=ros1, rls1=rls1,
                noise=noise, L=L, M=M, T=T, n_subj=n_subj)


def run_kl_product(inp, dev, P_tot=200, eps=1e-6, storage=torch.float64, fixed_T=None, layout=None):
    L = inp["L"]
    k0, k1, lik = build_product_kernels(inp["kargs"], L, dev, inp["ros0"], inp["rls0"], inp["ros1"], inp["rls1"],
                                        inp["noise"])
    x = inp["x"].to(dev)
    mu = inp["mu"].to(dev).to(storage).requires_grad_(True)
    lv = inp["lv"].to(dev).to(storage).requires_grad_(True)
    z = inp["z"].to(dev).requires_grad_(True)
    m = inp["m"].to(dev).requires_grad_(True)
    H = inp["H"].to(dev).requires_grad_(True)
    N_tot = P_tot * inp["T"]
    if fixed_T is None:
        kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(k0, k1, lik, L, m, H, x, mu, lv, z, P_tot, inp["n_subj"], N_tot,
                                                          True, inp["kargs"]["id_covariate"], eps, layout=layout)
    else:
        kld, gm, gH = elbo.minibatch_KLD_upper_bound(k0, k1, lik, L, m, H, x, mu, lv, z, P_tot, inp["n_subj"], fixed_T,
                                                     True, eps, layout=layout)
    kld.sum().backward()
    gos0, gls0 = kernel_grads(k0)
    gos1, gls1 = kernel_grads(k1)
    return dict(kld=kld.detach().reshape(()), grad_m=gm, grad_H=gH, d_mu=mu.grad, d_logv=lv.grad, d_z=z.grad,
                d_m=m.grad, d_H=H.grad, d_os0=gos0, d_ls0=gls0, d_os1=gos1, d_ls1=gls1)


HYPER_ABS = 1e-8   # x |kld|: absolute floor for kernel hyper-parameter gradients, see _hyper_ok


def _hyper_ok(got, ref, kld_ref, hyper_tol):
    """Scale-aware bound for hyper-parameter gradients (hard part 1): they are
    differences of addends of the size of kld itself (1e6..1e9 at the reference's initial state,
    cond(K0zz + eps I) ~ 1e7), so two float64 evaluation orders of the REFERENCE already differ by
    ~1e-9 |kld| in absolute terms.  Accept relative error <= hyper_tol, or absolute error
    <= HYPER_ABS * |kld|."""
    a, b = torch.as_tensor(got, dtype=DT).cpu(), torch.as_tensor(ref, dtype=DT).cpu()
    return rel_err(a, b) <= hyper_tol or float((a - b).abs().max()) <= HYPER_ABS * abs(float(kld_ref))


def compare_kl(got, ref, tol, hyper_tol, label=""):
    errs, bad = {}, {}
    for key in ("kld", "grad_m", "grad_H", "d_mu", "d_logv", "d_z", "d_m", "d_H", "d_os0", "d_ls0", "d_os1", "d_ls1"):
        if got.get(key) is None or ref.get(key) is None or torch.as_tensor(ref[key]).numel() == 0:
            continue
        errs[key] = rel_err(got[key], ref[key])
        if key.startswith(("d_os", "d_ls")):
            if not _hyper_ok(got[key], ref[key], ref["kld"], hyper_tol):
                bad[key] = errs[key]
        elif errs[key] > tol:
            bad[key] = errs[key]
        if key in ("d_mu", "d_logv"):
            errs[key + "/elem"] = elem_err(got[key], ref[key], rtol=20 * tol)
            if errs[key + "/elem"] > 1.0:
                bad[key + "/elem"] = errs[key + "/elem"]
    assert not bad, f"{label} outside tolerance: {bad} (all: {errs})"
    return errs


def check_kl_vs_oracle(dev, L, M, n_subj, T, seed, tol, hyper_tol, ragged=True, storage=torch.float64, **kw):
    inp = make_kl_inputs(L, M, n_subj, T, seed, ragged=ragged, **kw)
    got = run_kl_product(inp, dev, storage=storage)
    ref = oracle_kl(inp["kargs"], L, inp["x"], inp["mu"].to(storage).to(DT), inp["lv"].to(storage).to(DT), inp["z"],
                    inp["m"], inp["H"], inp["ros0"], inp["rls0"], inp["ros1"], inp["rls1"], inp["noise"], 200,
                    n_subj, 200 * T, 1e-6)
    return compare_kl(got, ref, tol, hyper_tol, label=f"KL L={L} M={M} P_b={n_subj} T={T}")


# ---------------------------------------------------------------------------- likelihoods
def parse_types(g):
    return [(s.split(":")[0], int(s.split(":")[1])) for s in g["types"].tolist()]


def golden_norm(g, dev):
    nr = (t(g["norm_real_mean"], dev), t(g["norm_real_var"], dev)) if g["norm_real_mean"].size else None
    npos = (t(g["norm_pos_mean"], dev), t(g["norm_pos_var"], dev)) if g["norm_pos_mean"].size else None
    return nr, npos


def run_loglik_golden(name, dev, storage=torch.float64):
    g = load(name)
    types = parse_types(g)
    lay = loglik.VarLayout(types, dev)
    lvr = t(g["log_vy_real"], dev).requires_grad_(True) if g["log_vy_real"].size else None
    lvp = t(g["log_vy_pos"], dev).requires_grad_(True) if g["log_vy_pos"].size else None
    nr, npos = golden_norm(g, dev)
    vparam = lay.vparam(lvr, lvp, nr, npos, conv=bool(int(g["conv"])))
    theta = t(g["theta"], dev).to(storage).requires_grad_(True)
    out = loglik.fused_loglik(lay, t(g["data"], dev).to(storage), t(g["mask"], dev).to(storage), theta, vparam)
    (out["log_p_x"] * t(g["g_up"], dev).to(storage)).sum().backward()
    got = dict(out)
    got["d_theta"] = theta.grad
    got["d_log_vy_real"] = lvr.grad if lvr is not None else None
    got["d_log_vy_pos"] = lvp.grad if lvp is not None else None
    return dict(golden=g, got=got, types=types)


def assert_loglik_close(r, tol=1e-9, label=""):
    g, got = r["golden"], r["got"]
    errs = {}
    for key in ("log_p_x", "log_p_x_missing", "params", "d_theta", "recon_mean", "recon_mode"):
        errs[key] = rel_err(got[key], g[key])
    for key in ("d_log_vy_real", "d_log_vy_pos"):
        if g[key].size and got[key] is not None:
            errs[key] = rel_err(got[key], g[key])
    bad = {k: v for k, v in errs.items() if v > tol}
    errs["d_theta/elem"] = elem_err(got["d_theta"], g["d_theta"], rtol=20 * tol)
    if errs["d_theta/elem"] > 1.0:
        bad["d_theta/elem"] = errs["d_theta/elem"]
    assert not bad, f"{label} likelihood terms outside tolerance: {bad} (all: {errs})"
    disc = np.array([k in ("cat", "ordinal") for k, _ in r["types"]])
    gm = got["recon_mean"].detach().cpu().numpy()
    assert np.array_equal(gm[:, disc], g["recon_mean"][:, disc]), f"{label} categorical/ordinal argmax differs"
    assert np.array_equal(got["data_transformed"].detach().cpu().numpy(), g["data_transformed"]), \
        f"{label} discrete transform differs"
    return errs


# ---------------------------------------------------------------------------- observation heads
class _Head(torch.nn.Module):
    """Stand-in for one Observation_* module (HLVAE.py:11-89): same parameter names, no forward."""

    def __init__(self, params):
        super().__init__()
        for n, v in params.items():
            setattr(self, n, torch.nn.Parameter(v))


def golden_heads(g, dev, conv):
    """(obs_layer ModuleList as HLVAE.py:276-299 builds it, per-group oracle head dicts, group kinds)."""
    layers, heads, kinds = [], [], []
    for i in range(int(g["n_groups"])):
        kind = str(g[f"g{i}_kind"]).split(":")[0]
        names = [k[len(f"g{i}_"):] for k in g if k.startswith(f"g{i}_") and k != f"g{i}_kind"]
        layers.append(_Head({n: t(g[f"g{i}_{n}"], dev) for n in names}))
        heads.append({n: t(g[f"g{i}_{n}"]).requires_grad_(True) for n in names})
        kinds.append(kind)
        if kind == "real" and conv:
            layers.append(torch.nn.Sigmoid())
    return torch.nn.ModuleList(layers), heads, kinds


def run_theta_golden(name, dev, storage=torch.float64, mask_u8=False):
    g = load(name)
    types, conv = parse_types(g), bool(int(g["conv"]))
    obs_layer, _, kinds = golden_heads(g, dev, conv)
    lay = theta_mod.HeadLayout(types, conv, dev)
    y0 = t(g["y"], dev).to(storage)
    if conv:                                    # the convolutional decoder hands over a permuted view (HLVAE.py:341-342)
        y0 = y0.permute(0, 2, 1).contiguous().permute(0, 2, 1)
    y = y0.clone(memory_format=torch.preserve_format).requires_grad_(True)
    mask = t(g["mask"], dev)
    mask = mask.to(torch.uint8) if mask_u8 else mask.to(storage)
    W, b = theta_mod.pack_heads(obs_layer, lay, y.shape[2])
    theta = theta_mod.theta_heads(lay, y, mask, W, b)
    (theta.double() * t(g["g_up"], dev)).sum().backward()
    grads, layer = {}, 0
    for i, kind in enumerate(kinds):
        for n, prm in obs_layer[layer].named_parameters():
            grads[f"d_g{i}_{n}"] = prm.grad
        layer += 2 if (kind == "real" and conv) else 1
    return dict(g=g, theta=theta.detach(), d_y=y.grad, grads=grads)


def assert_theta_close(r, tol, label=""):
    g = r["g"]
    errs = {"theta": rel_err(r["theta"], g["theta"]), "d_y": rel_err(r["d_y"], g["d_y"])}
    for k, v in r["grads"].items():
        errs[k] = rel_err(v, g[k])
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, f"{label}: {bad}"
    return errs
