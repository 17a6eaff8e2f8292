"""Evaluation-time GP bounds built on the ELBO-path kernels, no autograd:

  * validation_dubo(latent_dim, ...)   - drop-in for validation.validation_dubo (validation.py:16-76): batched over
                                         the latent dimensions, returns a 1-element float64 tensor (sum over them);
  * deviance_upper_bound(...)          - drop-in for elbo_functions.deviance_upper_bound (elbo_functions.py:60-115):
                                         one latent dimension (un-batched kernels, z [M, Q]), returns a 0-dim tensor;
  * elbo(...)                          - drop-in for elbo_functions.elbo (elbo_functions.py:9-57): the collapsed GP
                                         evidence bound of one latent dimension for a sample y of the latent.

All three are the full-data variant of the training statistics, fixed T per subject as in the reference:
  * hlvae_kl_subject  - iB_st, log det B (:44-45,58), sum(iB * K0_st) + tr(iB D) (:66,68), log det D (:67)
  * hlvae_kl_panel    - pass 1 (w = 0, G = 0): S = K0zx iB K0xz (:53), p = K0zx iB m (:63), m^T iB m (:62);
                        pass 2 (G = W^-1, `qdiag` output): (iB K0xz)_r W^-1 (iB K0xz)_r^T per row, whose
                        v-weighted sum is tr(W^-1 K0zx iB D iB K0xz) (:69-71)
  * hlvae_mxm_aux     - the M x M part (:40-41,54-59,64): Cholesky of K0zz + eps I and of W = K0zz + S, the log
                        determinants, |L_W^-1 p|^2, sum(S * iK) and W^-1, float64, one CTA per latent dimension.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .elbo import _noise_vector
from .kernels import compile_spec
from .subjects import SubjectLayout

N_SM = 148
_NO_VARIANCE = -800.0        # log-variance whose exp underflows to exactly 0: drops the D = diag(v) terms


def _pieces(L, covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps, with_variance=True):
    """Per-latent pieces [L] of the bounds: tr_a = sum(iB * K0_st) + tr(iB D), tr2 = tr(W^-1 K0zx iB D iB K0xz),
    trS = sum(S * iK), qF1 = y^T iB y, qF2 = |L_W^-1 p|^2, logdet = log det Sigma, logdetD."""
    if not train_xt.is_cuda:
        raise RuntimeError("hlvae_b200: the validation bounds run on CUDA tensors only (no CPU fallback)")
    dev = train_xt.device
    f64 = dict(dtype=torch.float64, device=dev)
    x = train_xt.detach().to(torch.float64).contiguous()
    zc = z.detach().to(**f64).reshape(L, -1, z.shape[-1]).contiguous()
    mu = m.detach().to(**f64).reshape(-1, L).contiguous()
    lv = log_v.detach().to(**f64).reshape(-1, L).contiguous()
    N = x.shape[0]
    if N != P * T:
        raise ValueError("expected P * T subject-contiguous rows")
    M, Q = zc.shape[-2], zc.shape[-1]
    layout = SubjectLayout.fixed(N, T, dev)
    fs0, fs1 = compile_spec(covar_module0), compile_spec(covar_module1)
    os0, ls0 = (t.detach().contiguous() for t in fs0.constrained(L, dev))
    os1, ls1 = (t.detach().contiguous() for t in fs1.constrained(L, dev))
    noise = _noise_vector(likelihood, L, dev)
    off = _lib.acc_layout(L, M, Q)
    st = _lib.stream_ptr()
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    acc = torch.zeros(off["total"] + 1, **f64)
    binv = torch.empty(L, max(layout.tt_total, 1), **f64)
    scratch = torch.empty(N, L, **f64)
    _lib.call("hlvae_kl_subject", fs0.cspec, _lib.ptr(os0), _lib.ptr(ls0), fs1.cspec, _lib.ptr(os1), _lib.ptr(ls1),
              _lib.ptr(noise), L, Q, _lib.ptr(x), Q, _lib.ptr(layout.row_idx), _lib.ptr(layout.subj_ptr),
              _lib.ptr(layout.tt_ptr), layout.n_subj, max(layout.t_max, 1), _lib.ptr(lv), L, _lib.F64,
              _lib.ptr(binv), binv.shape[1], _lib.ptr(acc), M, _lib.ptr(scratch), 1.0, _lib.ptr(status), st)
    n_chunks = max(1, min((N_SM * 8 + L - 1) // L, (layout.n_subj + 2) // 3))
    spc = (layout.n_subj + n_chunks - 1) // n_chunks

    def panel(acc_, mu_, G_, qdiag):
        zero_w = torch.zeros(L, M, **f64)
        _lib.call("hlvae_kl_panel", fs0.cspec, _lib.ptr(os0), _lib.ptr(ls0), fs1.cspec, _lib.ptr(os1), _lib.ptr(ls1),
                  L, Q, M, _lib.ptr(x), Q, _lib.ptr(zc), _lib.ptr(layout.row_idx), _lib.ptr(layout.subj_ptr),
                  _lib.ptr(layout.tt_ptr), layout.n_subj, spc, _lib.ptr(mu_), L, _lib.F64, _lib.ptr(zero_w),
                  _lib.ptr(G_), _lib.ptr(binv), binv.shape[1], _lib.ptr(acc_), _lib.ptr(scratch), _lib.ptr(qdiag),
                  1.0, _lib.ptr(status), 0, st)

    panel(acc, mu, torch.zeros(L, M, M, **f64), None)
    S = acc[off["S"]:off["S"] + L * M * M].view(L, M, M)
    p = acc[off["p"]:off["p"] + L * M].view(L, M)
    scal = acc[off["scal"]:off["scal"] + L * _lib.NSCAL].view(L, _lib.NSCAL)
    # M x M part (:39-41,54-60,64,66): one launch
    mm = torch.zeros(L, 4, **f64)
    iW = torch.empty(L, M, M, **f64) if with_variance else None
    _lib.call("hlvae_mxm_aux", fs0.cspec, _lib.ptr(os0), _lib.ptr(ls0), L, Q, M, _lib.ptr(zc), float(eps), _lib.ptr(S),
              _lib.ptr(p), _lib.ptr(mm), None, None, _lib.ptr(iW), _lib.ptr(_lib.workspace(L, M, dev)),
              _lib.ptr(status), st)
    stt = status.tolist()
    if stt[0] == _lib.STATUS_NOT_PD:
        what = {-1: "K0zz + eps I", -5: "W = K0zz + K0zx iB K0xz"}.get(stt[2], "B_s")
        raise RuntimeError(f"hlvae_b200: cholesky: {what} is not positive-definite")
    if stt[0] == _lib.STATUS_T_TOO_LARGE:
        raise RuntimeError(f"hlvae_b200: a subject has more than {_lib.TMAX} rows")
    out = dict(qF1=scal[:, 0].clone(), tr_a=scal[:, 1].clone(), logdetD=scal[:, 3].clone(), qF2=mm[:, 2], trS=mm[:, 3],
               logdet=-mm[:, 0] + scal[:, 2] + mm[:, 1])                                           # :57-60
    if with_variance:
        # pass 2: per-row (iB K0xz)_r W^-1 (iB K0xz)_r^T
        q = torch.zeros(N, L, **f64)
        panel(torch.zeros_like(acc), torch.zeros_like(mu), iW, q)
        out["tr2"] = (torch.exp(lv) * q).sum(0)                                                    # :69-71
    return out


def validation_dubo(latent_dim, covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps):
    """validation.py:16-76."""
    with torch.no_grad():
        t = _pieces(latent_dim, covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps)
        dubo = 0.5 * (t["tr_a"] - t["tr2"] - t["trS"] + (t["qF1"] - t["qF2"]) - P * T + t["logdet"] - t["logdetD"])  # :72-74
        return dubo.sum().reshape(1)


def deviance_upper_bound(covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps):
    """elbo_functions.py:60-115: the same bound for ONE latent dimension (un-batched kernel objects, m and log_v
    [P T], z [M, Q]); returns a 0-dim tensor like the reference."""
    with torch.no_grad():
        t = _pieces(1, covar_module0, covar_module1, likelihood, train_xt, m.reshape(-1, 1), log_v.reshape(-1, 1),
                    z.reshape(1, *z.shape[-2:]), P, T, eps)
        # :113-114 with tr_iSigma_D = tr(iB D) - tr2 and tr = sum(iB * K0_st) - sum(S * iK): tr_a holds both first terms
        dubo = 0.5 * (t["tr_a"] - t["tr2"] - t["trS"] + (t["qF1"] - t["qF2"]) - P * T + t["logdet"] - t["logdetD"])
        return dubo.reshape(())


def elbo(covar_module0, covar_module1, likelihood, train_xt, train_yt, z, P, T, eps):
    """elbo_functions.py:9-57: collapsed evidence lower bound of one latent dimension for a sample train_yt [P T]."""
    with torch.no_grad():
        y = train_yt.reshape(-1, 1)
        t = _pieces(1, covar_module0, covar_module1, likelihood, train_xt, y, torch.full_like(y, _NO_VARIANCE),
                    z.reshape(1, *z.shape[-2:]), P, T, eps, with_variance=False)
        tr = t["tr_a"] - t["trS"]                                                                  # :53 (D = 0 here)
        log_like = -0.5 * T * P * math.log(2 * math.pi) - 0.5 * (t["logdet"] + t["qF1"] - t["qF2"])  # :54-55
        return (log_like - 0.5 * tr).reshape(())                                                   # :56
