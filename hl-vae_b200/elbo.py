"""KL-divergence upper bound of the minibatch ELBO: drop-in for
elbo_functions.minibatch_KLD_upper_bound (elbo_functions.py:118-193) and
minibatch_KLD_upper_bound_iter (:196-285): same arguments, same returns
(kld_total attached to autograd, grad_m, grad_H).

Data flow (DESIGN.md has the derivation):
  1. M x M pre-stage, replicated, float64:  K0zz + eps I, Cholesky, iK, iH, w = iK m,
     G = iK H iK - iK.
  2. streaming stage, CUDA kernels behind the C ABI (hlvae_kl_subject, hlvae_kl_panel):
     per-latent accumulators S, p, dJ/dw, the scalars A, B + sum(iB*K0_st), C, F, and the
     gradients of mu, log_v, Z and kernel hyper-parameters, all in one pass.
  3. (data parallel) one all-reduce of the accumulator buffer.
  4. M x M post-stage: D, E, kld_qu_pu, kld_total, natural-gradient pieces.
"""
from __future__ import annotations

import torch

from . import _lib, config
from .kernels import _KernelEval, compile_spec
from .subjects import SubjectLayout

N_SM = 148


def _noise_vector(likelihood, L, device):
    nz = likelihood.noise_covar.noise if hasattr(likelihood, "noise_covar") else likelihood.noise
    return nz.detach().to(device=device, dtype=torch.float64).reshape(-1).expand(L).contiguous()


def _raise_status(status, info):
    st = status.tolist()
    if st[0] == _lib.STATUS_NOT_PD:
        raise RuntimeError(f"hlvae_b200: cholesky: B_s of subject {st[2]} (latent dim {st[1]}) is not "
                           "positive-definite")
    if st[0] == _lib.STATUS_T_TOO_LARGE:
        raise RuntimeError(f"hlvae_b200: subject {st[2]} has more than {_lib.TMAX} rows")
    if info is not None and bool((info != 0).any()):
        raise RuntimeError("hlvae_b200: cholesky: K0zz + eps I or H is not positive-definite")


class _KLD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, log_v, z, m, H, os0, ls0, os1, ls1, noise, x, fs0, fs1, layout, scale, const, eps,
                natural_gradient, out_shape):
        dev = x.device
        N, L = mu.shape
        M, Q = z.shape[-2], z.shape[-1]
        if mu.dtype != log_v.dtype:
            raise TypeError("mu and log_v must share a dtype")
        dcode = _lib.dtype_code(mu)
        mu_c, lv_c = mu.detach().contiguous(), log_v.detach().contiguous()
        x_c = x.detach().to(torch.float64).contiguous()
        f64 = dict(dtype=torch.float64, device=dev)

        # ---- 1. M x M pre-stage on a private graph (its gradient is taken below, inside forward)
        with torch.enable_grad():
            zl = z.detach().to(torch.float64).contiguous().requires_grad_(True)
            os0l = os0.detach().requires_grad_(True)
            ls0l = ls0.detach().requires_grad_(True)
            ml = m.detach().to(torch.float64).requires_grad_(True)
            Hl = H.detach().to(torch.float64).requires_grad_(True)
            eye = torch.eye(M, **f64)
            eyeL = eye.expand(L, M, M).contiguous()
            K0zz = _KernelEval.apply(fs0, os0l, ls0l, zl, zl) + eps * eye                 # :148,153 / :223-224
            LK, infoK = torch.linalg.cholesky_ex(K0zz)                                     # :154 / :225
            iK = torch.cholesky_solve(eyeL, LK)                            # :155 / :226
            LH, infoH = torch.linalg.cholesky_ex(Hl)                                       # :162 / :227
            iH = torch.cholesky_solve(eyeL, LH)                            # :163 / :228
            w = iK @ ml                                                                    # iK m in :166 / :230
            G = iK @ Hl @ iK - iK                                                          # :171 / :231 minus :170
        Gs = (0.5 * (G + G.transpose(-1, -2))).detach().contiguous()
        wd = w.detach().reshape(L, M).contiguous()

        # ---- 2. streaming stage
        off = _lib.acc_layout(L, M, Q)
        acc = torch.zeros(off["total"], **f64)
        g_mu = torch.zeros(N, L, **f64)
        g_lv = torch.zeros(N, L, **f64)
        binv = torch.empty(L, max(layout.tt_total, 1), **f64)
        status = torch.zeros(4, dtype=torch.int32, device=dev)
        os0c, ls0c = os0.detach().contiguous(), ls0.detach().contiguous()
        os1c, ls1c = os1.detach().contiguous(), ls1.detach().contiguous()
        lib, st = _lib.lib(), _lib.stream_ptr()
        if layout.n_subj > 0:
            _lib.call("hlvae_kl_subject", fs0.cspec, _lib.ptr(os0c), _lib.ptr(ls0c), fs1.cspec, _lib.ptr(os1c),
                                            _lib.ptr(ls1c), _lib.ptr(noise), L, Q, _lib.ptr(x_c), Q,
                                            _lib.ptr(layout.row_idx), _lib.ptr(layout.subj_ptr),
                                            _lib.ptr(layout.tt_ptr), layout.n_subj, max(layout.t_max, 1),
                                            _lib.ptr(lv_c), L, dcode, _lib.ptr(binv), binv.shape[1], _lib.ptr(acc), M,
                                            _lib.ptr(g_lv), _lib.ptr(status), st)
            n_chunks = max(1, min((N_SM * 8 + L - 1) // L, (layout.n_subj + 2) // 3))
            spc = (layout.n_subj + n_chunks - 1) // n_chunks
            _lib.call("hlvae_kl_panel", fs0.cspec, _lib.ptr(os0c), _lib.ptr(ls0c), fs1.cspec, _lib.ptr(os1c),
                                          _lib.ptr(ls1c), L, Q, M, _lib.ptr(x_c), Q, _lib.ptr(zl.detach()),
                                          _lib.ptr(layout.row_idx), _lib.ptr(layout.subj_ptr),
                                          _lib.ptr(layout.tt_ptr), layout.n_subj, spc, _lib.ptr(mu_c), L, dcode,
                                          _lib.ptr(wd), _lib.ptr(Gs), _lib.ptr(binv), binv.shape[1], _lib.ptr(acc),
                                          _lib.ptr(g_mu), _lib.ptr(status), st)
        # ---- 3. data parallel: one all-reduce of every accumulator (S, p, scalars, replicated-parameter grads)
        if config.process_group is not None:
            torch.distributed.all_reduce(acc, group=config.process_group)
        if config.check_errors:
            _raise_status(status, torch.cat([infoK, infoH]))

        def view(name, *shape):
            n = 1
            for s in shape:
                n *= s
            return acc[off[name]:off[name] + n].view(*shape)

        S, p, gw = view("S", L, M, M), view("p", L, M, 1), view("gw", L, M, 1)
        scal = view("scal", L, _lib.NSCAL).sum(0)
        nc0, nc1 = fs0.ncomp, fs1.ncomp
        gZ = view("gZ", L, M, Q)
        gos0, gls0 = view("gos0", _lib.MAX_COMPS, L)[:nc0], view("gls0", _lib.MAX_COMPS, L)[:nc0]
        gos1, gls1 = view("gos1", _lib.MAX_COMPS, L)[:nc1], view("gls1", _lib.MAX_COMPS, L)[:nc1]

        # ---- 4. M x M post-stage
        with torch.enable_grad():
            J_S = 0.5 * (G * S).sum()                                                      # S parts of D (:170) and E (:172)
            tr1 = (iK * Hl.transpose(-1, -2)).sum()                                        # :176 / :271
            qf1 = (ml * (iK @ ml)).sum()                                                   # :177 / :272
            logdetK = 2 * torch.log(torch.diagonal(LK, dim1=-1, dim2=-2)).sum()            # :178 / :273
            logdetH = 2 * torch.log(torch.diagonal(LH, dim1=-1, dim2=-2)).sum()            # :179 / :274
            kq = 0.5 * (tr1 + qf1 - L * M + logdetK - logdetH)                             # :180 / :275
            local = scale * (J_S + (w * gw).sum()) + kq
            gz_l, gos0_l, gls0_l, gm_l, gH_l = torch.autograd.grad(local, [zl, os0l, ls0l, ml, Hl])
        kld = scale * (0.5 * (scal[0] + scal[1] + scal[2] - scal[3]) + J_S.detach()) + kq.detach() - const  # :181 / :277

        grad_m = grad_H = None
        if natural_gradient:                                                               # :186-191 / :279-283
            iKd = iK.detach()
            Bm = iKd @ S @ iKd + iKd
            grad_m = -(iKd @ p) + Bm @ ml.detach()
            grad_H = 0.5 * (-iH.detach() + Bm)

        ctx.save_for_backward((scale * g_mu).to(mu.dtype), (scale * g_lv).to(log_v.dtype), gz_l + scale * gZ, gm_l,
                              gH_l, gos0_l + scale * gos0, gls0_l + scale * gls0, scale * gos1, scale * gls1)
        ctx.mark_non_differentiable(*[t for t in (grad_m, grad_H) if t is not None])
        return kld.reshape(out_shape), grad_m, grad_H

    @staticmethod
    def backward(ctx, g_kld, g_gm, g_gH):
        g = g_kld.reshape(())
        return tuple(g * t for t in ctx.saved_tensors) + (None,) * 10


def _kld(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z, layout, scale, const,
         natural_gradient, eps, out_shape):
    if not train_xt.is_cuda:
        raise RuntimeError("hlvae_b200: the KL upper bound runs on CUDA tensors only (no CPU fallback)")
    fs0, fs1 = compile_spec(covar_module0), compile_spec(covar_module1)
    L = latent_dim
    os0, ls0 = fs0.constrained(L, train_xt.device)
    os1, ls1 = fs1.constrained(L, train_xt.device)
    noise = _noise_vector(likelihood, L, train_xt.device)
    return _KLD.apply(mu, log_v, z, m, H, os0, ls0, os1, ls1, noise, train_xt, fs0, fs1, layout, float(scale),
                      float(const), float(eps), bool(natural_gradient), out_shape)


def minibatch_KLD_upper_bound(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z,
                              P_tot, P_batch, T, natural_gradient, eps, layout=None):
    """elbo_functions.py:118-193 (fixed T).  Rows must be subject-contiguous, T per subject."""
    if layout is None:
        layout = SubjectLayout.fixed(train_xt.shape[0], T, train_xt.device)
    return _kld(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z, layout,
                P_tot / P_batch, latent_dim * P_tot * T / 2, natural_gradient, eps, ())


def minibatch_KLD_upper_bound_iter(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z,
                                   P, P_in_current_batch, N, natural_gradient, id_covariate, eps, layout=None):
    """elbo_functions.py:196-285 (varying T).  `layout` (optional) skips the id grouping when the
    caller already knows the subject lengths (SubjectLayout.from_lengths)."""
    if layout is None:
        layout = SubjectLayout.from_ids(train_xt[:, id_covariate])
    return _kld(covar_module0, covar_module1, likelihood, latent_dim, m, H, train_xt, mu, log_v, z, layout,
                P / P_in_current_batch, latent_dim * N / 2, natural_gradient, eps, (1,))


def natural_gradient_update(m, H, grad_m, grad_H, lr):
    """training.py:130-137 (kept here so callers outside training.py can reuse it)."""
    eye = torch.eye(H.shape[-1], dtype=H.dtype, device=H.device).expand_as(H)
    iH = torch.cholesky_solve(eye, torch.linalg.cholesky(H))
    iH_new = iH + lr * (grad_H + grad_H.transpose(-1, -2))
    H_new = torch.cholesky_solve(eye, torch.linalg.cholesky(iH_new)).detach()
    m_new = (H_new @ (iH @ m - lr * (grad_m - 2 * (grad_H @ m)))).detach()
    return m_new, H_new
