"""Deviance upper bound of the validation pass: drop-in for validation.validation_dubo
(validation.py:16-76), same arguments, same return (1-element float64 tensor, summed over latent dims).

SURVEY.md section 8(f) row 3: the full-data variant of the training statistics.  Built on the ELBO-path
kernels, fixed T per subject as in the reference:
  * hlvae_kl_subject  - iB_st, log det B (:44-45,58), sum(iB * K0_st) + tr(iB D) (:66,68), log det D (:67)
  * hlvae_kl_panel    - pass 1 (w = 0, G = 0): S = K0zx iB K0xz (:53), p = K0zx iB m (:63), m^T iB m (:62);
                        pass 2 (G = W^-1, `qdiag` output): (iB K0xz)_r W^-1 (iB K0xz)_r^T per row, whose
                        v-weighted sum is tr(W^-1 K0zx iB D iB K0xz) (:69-71)
and float64 torch.linalg Cholesky factorisations for the M x M part (:40-41,54-59,64).  No autograd.
"""
from __future__ import annotations

import torch

from . import _lib
from .elbo import _noise_vector
from .kernels import compile_spec, evaluate_dense
from .subjects import SubjectLayout

N_SM = 148


def validation_dubo(latent_dim, covar_module0, covar_module1, likelihood, train_xt, m, log_v, z, P, T, eps):
    """validation.py:16-76."""
    if not train_xt.is_cuda:
        raise RuntimeError("hlvae_b200: validation_dubo runs on CUDA tensors only (no CPU fallback)")
    dev = train_xt.device
    L = latent_dim
    with torch.no_grad():
        f64 = dict(dtype=torch.float64, device=dev)
        x = train_xt.detach().to(torch.float64).contiguous()
        zc = z.detach().to(**f64).contiguous()
        mu = m.detach().to(**f64).contiguous()
        lv = log_v.detach().to(**f64).contiguous()
        N = x.shape[0]
        if N != P * T:
            raise ValueError("validation_dubo expects P * T subject-contiguous rows")
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
                      1.0, _lib.ptr(status), st)

        panel(acc, mu, torch.zeros(L, M, M, **f64), None)
        code = int(status[0])
        if code == _lib.STATUS_NOT_PD:
            raise RuntimeError("hlvae_b200: cholesky: B_s is not positive-definite")
        if code == _lib.STATUS_T_TOO_LARGE:
            raise RuntimeError(f"hlvae_b200: a subject has more than {_lib.TMAX} rows")
        S = acc[off["S"]:off["S"] + L * M * M].view(L, M, M)
        p = acc[off["p"]:off["p"] + L * M].view(L, M, 1)
        scal = acc[off["scal"]:off["scal"] + L * _lib.NSCAL].view(L, _lib.NSCAL)
        qF1, tr_a, logdetB, logdetD = scal[:, 0], scal[:, 1], scal[:, 2], scal[:, 3]
        # M x M part
        K0zz = evaluate_dense(covar_module0, zc, zc) + eps * torch.eye(M, **f64)               # :39
        LK = torch.linalg.cholesky(K0zz)                                                       # :40
        iK = torch.cholesky_inverse(LK)                                                        # :41
        W = K0zz + 0.5 * (S + S.transpose(1, 2))                                               # :54-55
        LW = torch.linalg.cholesky(W)                                                          # :56
        logdet = -2 * torch.log(torch.diagonal(LK, dim1=-2, dim2=-1)).sum(1) + logdetB \
            + 2 * torch.log(torch.diagonal(LW, dim1=-2, dim2=-1)).sum(1)                       # :57-60
        qF2 = (torch.linalg.solve_triangular(LW, p, upper=False) ** 2).sum((1, 2))             # :64
        trS = (S * iK).sum((1, 2))                                                             # :66, second term
        iW = torch.cholesky_inverse(LW).contiguous()
        # pass 2: per-row (iB K0xz)_r W^-1 (iB K0xz)_r^T
        q = torch.zeros(N, L, **f64)
        panel(torch.zeros_like(acc), torch.zeros_like(mu), iW, q)
        tr2 = (torch.exp(lv) * q).sum(0)                                                       # :69-71
        dubo = 0.5 * (tr_a - tr2 - trS + (qF1 - qF2) - P * T + logdet - logdetD)               # :72-74
        return dubo.sum().reshape(1)
