"""GP posterior-mean prediction of the latent variables: drop-in for utils.batch_predict_varying_T
(utils.py:99-191) and utils.batch_predict (:193-271), same arguments, same return (Z_pred [N_test, L]).

SURVEY.md section 8(f) row 1, built on the kernels of the ELBO path:
  * hlvae_kl_subject  - B_s = K1(x_s, x_s) + noise I, Cholesky, explicit inverse      (:149-153 / :236-237)
  * hlvae_kl_panel    - with w = 0, G = 0:  S = sum_s K0zx_s iB_s K0xz_s (:156-158 / :242-244),
                        p = K0zx iB mu and, through its d/dmu output, iB mu itself       (:159 / :245);
                        a second call with w = a, mu = 0 gives -iB K0xz a                (:162-166 / :246-247)
  * hlvae_kernel_eval_fwd - the dense blocks K0zz, K0xz, K0Xz, K1Xx                     (:128-130,169,175-186)
  * hlvae_mxm_aux     - the two M x M solves (the reference's `torch.solve` calls at :162,169 / :246,249 no longer
                        exist in torch): a = (K0zz + S)^-1 p and c = K0zz^-1 K0zx mu_tilde = K0zz^-1 (p - S a)
  * hlvae_kernel_matvec - K0Xz c without materialising K0Xz                           (:169 / :249)
Evaluation-time code: no autograd, CUDA tensors only, no library (cuBLAS / cuSOLVER) call on the path.
"""
from __future__ import annotations

import torch

from . import _lib
from .elbo import _noise_vector
from .kernels import compile_spec, evaluate_dense
from .subjects import SubjectLayout

N_SM = 148


def _stream_pass(fs0, fs1, hp, L, Q, M, x, z, layout, mu, w, noise, binv=None):
    """One pass of the per-subject + panel kernels with G = 0: returns (S [L,M,M], p [L,M],
    r [N,L] = iB (mu - K0xz w), binv)."""
    dev = x.device
    os0, ls0, os1, ls1 = hp
    f64 = dict(dtype=torch.float64, device=dev)
    off = _lib.acc_layout(L, M, Q)
    acc = torch.zeros(off["total"] + 1, **f64)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    N = x.shape[0]
    g_mu = torch.zeros(N, L, **f64)
    st = _lib.stream_ptr()
    if binv is None:
        binv = torch.empty(L, max(layout.tt_total, 1), **f64)
        zeros_nl = torch.zeros(N, L, **f64)
        g_lv = torch.empty(N, L, **f64)
        _lib.call("hlvae_kl_subject", fs0.cspec, _lib.ptr(os0), _lib.ptr(ls0), fs1.cspec, _lib.ptr(os1), _lib.ptr(ls1),
                  _lib.ptr(noise), L, Q, _lib.ptr(x), Q, _lib.ptr(layout.row_idx), _lib.ptr(layout.subj_ptr),
                  _lib.ptr(layout.tt_ptr), layout.n_subj, max(layout.t_max, 1), _lib.ptr(zeros_nl), L, _lib.F64,
                  _lib.ptr(binv), binv.shape[1], _lib.ptr(acc), M, _lib.ptr(g_lv), 1.0, _lib.ptr(status), st)
        acc.zero_()                       # the per-subject scalars / K1 gradients are not used here
    G = torch.zeros(L, M, M, **f64)
    n_chunks = max(1, min((N_SM * 8 + L - 1) // L, (layout.n_subj + 2) // 3))
    spc = (layout.n_subj + n_chunks - 1) // n_chunks
    _lib.call("hlvae_kl_panel", fs0.cspec, _lib.ptr(os0), _lib.ptr(ls0), fs1.cspec, _lib.ptr(os1), _lib.ptr(ls1), L, Q, M,
              _lib.ptr(x), Q, _lib.ptr(z), _lib.ptr(layout.row_idx), _lib.ptr(layout.subj_ptr), _lib.ptr(layout.tt_ptr),
              layout.n_subj, spc, _lib.ptr(mu), L, _lib.F64, _lib.ptr(w), _lib.ptr(G), _lib.ptr(binv), binv.shape[1],
              _lib.ptr(acc), _lib.ptr(g_mu), None, 1.0, _lib.ptr(status), 0, st)
    code = int(status[0])
    if code == _lib.STATUS_NOT_PD:
        raise RuntimeError("hlvae_b200: cholesky: B_s is not positive-definite")
    if code == _lib.STATUS_T_TOO_LARGE:
        raise RuntimeError(f"hlvae_b200: a subject has more than {_lib.TMAX} rows")
    S = acc[off["S"]:off["S"] + L * M * M].view(L, M, M)
    p = acc[off["p"]:off["p"] + L * M].view(L, M)
    return S, p, g_mu, binv


def _predict(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list, layout,
             id_covariate, eps):
    if isinstance(covar_module0, (list, tuple)):
        raise NotImplementedError("per-latent kernel lists (the unbatched legacy path) are not supported")
    if not prediction_x.is_cuda:
        raise RuntimeError("hlvae_b200: prediction runs on CUDA tensors only (no CPU fallback)")
    dev = prediction_x.device
    L = latent_dim
    with torch.no_grad():
        x = prediction_x.detach().to(torch.float64).contiguous()
        xt = test_x.detach().to(device=dev, dtype=torch.float64).contiguous()
        z = (torch.stack(list(zt_list)) if isinstance(zt_list, (list, tuple)) else zt_list).detach().to(
            device=dev, dtype=torch.float64).contiguous()
        mu64 = mu.detach().to(device=dev, dtype=torch.float64).contiguous()
        M, Q = z.shape[-2], z.shape[-1]
        fs0, fs1 = compile_spec(covar_module0), compile_spec(covar_module1)
        hp = tuple(t.detach().contiguous() for pair in (fs0.constrained(L, dev), fs1.constrained(L, dev)) for t in pair)
        noise = _noise_vector(likelihoods, L, dev)
        zero_w = torch.zeros(L, M, dtype=torch.float64, device=dev)
        # S = sum_s K0zx_s iB_s K0xz_s, p = K0zx iB mu, iB mu  (:139-160 / :236-245)
        S, p, iB_mu, binv = _stream_pass(fs0, fs1, hp, L, Q, M, x, z, layout, mu64, zero_w, noise)
        # a = (K0zz + eps I + S)^-1 p (:162 / :246) and, since K0zx mu_tilde = K0zx iB (mu - K0xz a) = p - S a,
        # c = (K0zz + eps I)^-1 (p - S a) (:169 / :249) - one launch, no dense K0xz
        a = torch.empty(L, M, dtype=torch.float64, device=dev)
        c = torch.empty(L, M, dtype=torch.float64, device=dev)
        status = torch.zeros(4, dtype=torch.int32, device=dev)
        _lib.call("hlvae_mxm_aux", fs0.cspec, _lib.ptr(hp[0]), _lib.ptr(hp[1]), L, Q, M, _lib.ptr(z), float(eps),
                  _lib.ptr(S.contiguous()), _lib.ptr(p.contiguous()), None, _lib.ptr(a), _lib.ptr(c), None,
                  _lib.ptr(_lib.workspace(L, M, dev)), _lib.ptr(status), _lib.stream_ptr())
        if int(status[0]) == _lib.STATUS_NOT_PD:
            raise RuntimeError("hlvae_b200: cholesky: K0zz + eps I (+ K0zx iB K0xz) is not positive-definite")
        # -iB K0xz a  (:162-166 / :246-247)
        _, _, corr, _ = _stream_pass(fs0, fs1, hp, L, Q, M, x, z, layout, torch.zeros_like(mu64), a, noise, binv=binv)
        mu_tilde = iB_mu + corr                                                             # [N, L]  (:167 / :248)
        first = torch.empty(xt.shape[0], L, dtype=torch.float64, device=dev)                # K0Xz c  (:169 / :249)
        _lib.call("hlvae_kernel_matvec", fs0.cspec, _lib.ptr(hp[0]), _lib.ptr(hp[1]), L, Q, _lib.ptr(xt), xt.shape[0],
                  _lib.ptr(z), M, _lib.ptr(c), _lib.ptr(first), _lib.stream_ptr())
        # K1(X*, x) mu_tilde over the conditioning rows whose subject appears in test_x (:171-186 / :251-267)
        id_rows = all(any(fs1.cspec.comp[r].disc_kind[f] == _lib.KIND_CAT and fs1.cspec.comp[r].disc_col[f] == id_covariate
                          for f in range(fs1.cspec.comp[r].ndisc)) for r in range(fs1.ncomp))
        if id_rows and layout.n_subj > 0:
            # every K1 term carries the id kernel (kernel_gen.py:225-242,269-289 route exactly those terms to K1):
            # only the rows of a test row's own subject contribute -> one pass over [N_test, L], no dense block
            first_rows = layout.row_idx[layout.subj_ptr[:-1].long()].long()
            ids_sorted, perm = torch.sort(x[first_rows, id_covariate])     # id value of every subject of the layout
            tid = xt[:, id_covariate].contiguous()
            pos = torch.searchsorted(ids_sorted, tid).clamp_(max=ids_sorted.numel() - 1)
            sid = torch.where(ids_sorted[pos] == tid, perm[pos], torch.full_like(pos, -1)).to(torch.int32)
            out2 = torch.empty(xt.shape[0], L, dtype=torch.float64, device=dev)
            mt = mu_tilde.contiguous()
            _lib.call("hlvae_subject_matvec", fs1.cspec, _lib.ptr(hp[2]), _lib.ptr(hp[3]), L, Q, _lib.ptr(xt), xt.shape[0],
                      _lib.ptr(x), _lib.ptr(layout.row_idx), _lib.ptr(layout.subj_ptr), _lib.ptr(sid), _lib.ptr(mt),
                      _lib.ptr(out2), _lib.stream_ptr())
            return (first + out2).contiguous()                                              # :188 / :269
        test_ids = torch.unique(xt[:, id_covariate])
        mask = torch.isin(x[:, id_covariate], test_ids)
        second = torch.zeros_like(first)
        if bool(mask.any()):
            # a K1 term without the id kernel (not produced by kernel_gen.py, kept for API completeness): dense block
            xm = x[mask].contiguous()
            K1Xx = evaluate_dense(covar_module1, xt.unsqueeze(0).expand(L, *xt.shape).contiguous(),
                                  xm.unsqueeze(0).expand(L, *xm.shape).contiguous())
            second = (K1Xx * mu_tilde[mask].T.unsqueeze(1)).sum(2).T
        return (first + second).contiguous()                                                # :188 / :269


def batch_predict_varying_T(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list,
                            id_covariate, eps):
    """utils.py:99-191."""
    layout = SubjectLayout.from_ids(prediction_x[:, id_covariate])
    return _predict(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list, layout,
                    id_covariate, eps)


def batch_predict(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list, P, T,
                  id_covariate, eps):
    """utils.py:193-271 (rows subject-contiguous, T per subject)."""
    layout = SubjectLayout.fixed(prediction_x.shape[0], T, prediction_x.device)
    return _predict(latent_dim, covar_module0, covar_module1, likelihoods, prediction_x, test_x, mu, zt_list, layout,
                    id_covariate, eps)
