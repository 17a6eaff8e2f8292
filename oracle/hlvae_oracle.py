"""CPU oracle for the HL-VAE per-step ELBO hot path (float64, plain PyTorch ops).

TEST INFRASTRUCTURE ONLY.  This file restates the reference algorithm so that the
CUDA path can be checked against it.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product
package never imports anything from `oracle/`.

Parity status: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), and its kernel algebra lives in gpytorch, which is neither vendored
nor pinned nor installable here.  The pin is therefore "outputs of the reference itself
run here": `oracle/make_goldens.py` executes the unmodified reference modules from
/root/reference behind the stand-ins in `oracle/standins/` and asserts that this
restatement agrees before freezing fixtures into `tests/golden/`.  At the gpytorch
boundary itself parity is UNPINNED (the stand-in restates gpytorch's documented
semantics); the reference's own gpytorch-free statement of the same kernels,
GP_model.py:27-116, is the authority followed here.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64

# --------------------------------------------------------------------------------------
# Additive-kernel specification
# --------------------------------------------------------------------------------------
SE, CAT, BIN = 0, 1, 2


@dataclass
class Factor:
    """One base kernel inside a product: kind in {SE, CAT, BIN}, covariate column, and
    (SE only) the index of its lengthscale row."""
    kind: int
    col: int
    ls: int = -1


@dataclass
class Component:
    """One ScaleKernel term: outputscale[index] * prod(factors)."""
    factors: List[Factor] = field(default_factory=list)


@dataclass
class AdditiveSpec:
    """Flat description of one AdditiveKernel; `n_ls` counts SE factors."""
    comps: List[Component] = field(default_factory=list)
    n_ls: int = 0

    def add(self, factors: Sequence[Tuple[int, int]]):
        comp = Component()
        for kind, col in factors:
            if kind == SE:
                comp.factors.append(Factor(SE, col, self.n_ls))
                self.n_ls += 1
            else:
                comp.factors.append(Factor(kind, col))
        self.comps.append(comp)


def compile_spec(cat_kernel, bin_kernel, sqexp_kernel, cat_int_kernel, bin_int_kernel,
                 covariate_missing_val, id_covariate) -> Tuple[AdditiveSpec, AdditiveSpec]:
    """Component order and routing of kernel_gen.py:219-310 (generate_kernel_batched):
    categorical, squared-exponential, binary, cat x SE interactions, bin x SE interactions;
    terms whose categorical covariate is the id covariate go to spec1, the rest to spec0;
    a covariate listed in covariate_missing_val gets an extra BinKernel(mask) factor."""
    missing = {d['covariate']: d['mask'] for d in covariate_missing_val}
    k0, k1 = AdditiveSpec(), AdditiveSpec()

    def masked(kind, col):
        f = [(kind, col)]
        if col in missing:
            f.append((BIN, missing[col]))
        return f

    for idx in cat_kernel:                                   # kernel_gen.py:225-242
        (k1 if idx == id_covariate else k0).add(masked(CAT, idx))
    for idx in sqexp_kernel:                                 # kernel_gen.py:245-254
        k0.add(masked(SE, idx))
    for idx in bin_kernel:                                   # kernel_gen.py:257-266
        k0.add(masked(BIN, idx))
    for d in cat_int_kernel:                                 # kernel_gen.py:269-289
        f = masked(CAT, d['cat_covariate']) + masked(SE, d['cont_covariate'])
        (k1 if d['cat_covariate'] == id_covariate else k0).add(f)
    for d in bin_int_kernel:                                 # kernel_gen.py:292-308
        k0.add(masked(BIN, d['bin_covariate']) + masked(SE, d['cont_covariate']))
    return k0, k1


def softplus_inv(v: float) -> float:
    """gpytorch inv_softplus: raw such that softplus(raw) = v  [gpytorch]."""
    return v + math.log(-math.expm1(-v))


@dataclass
class KernelParams:
    """Raw (unconstrained) parameters of one AdditiveSpec, gpytorch parametrisation:
    outputscale = softplus(raw_outputscale) (ScaleKernel), lengthscale =
    softplus(raw_lengthscale) (RBFKernel).  Shapes [n_comp, L] and [n_ls, L]."""
    raw_outputscale: torch.Tensor
    raw_lengthscale: torch.Tensor

    @staticmethod
    def default(spec: AdditiveSpec, L: int) -> "KernelParams":
        # kernel_spec.py:68 initialises lengthscale 2.5 while the parameter is still
        # float32; HLVAE_main.py:235 casts to double afterwards, so the raw value is the
        # float32 rounding of inv_softplus(2.5).  raw_outputscale starts at 0 [gpytorch].
        raw_ls = float(np.float32(softplus_inv(2.5)))
        return KernelParams(torch.zeros(len(spec.comps), L, dtype=DT),
                            torch.full((spec.n_ls, L), raw_ls, dtype=DT))

    def requires_grad_(self, flag=True):
        self.raw_outputscale.requires_grad_(flag)
        self.raw_lengthscale.requires_grad_(flag)
        return self


def eval_additive(spec: AdditiveSpec, prm: KernelParams, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """Dense additive kernel K[..., L?, n1, n2] = sum_r s_r,l * prod(factors).

    x1, x2: [..., n, Q] with broadcastable leading dims; if neither carries the latent
    axis the result gets it from the parameters (result [L, n1, n2]); inputs shaped
    [L, n, Q] or [P, L, n, Q] keep their own leading dims.
    Arithmetic of GP_model.py:27-116: BinKernel 27-33 (x1 + x2 == 2), CatKernel 35-41
    (x1 - x2 == 0), RbfKernel 43-69 exp(-(x1-x2)^2 / (2 l^2)), ScaleKernel 71-97,
    ProductKernel 109-116, AdditiveKernel 99-107.  Same values as gpytorch's
    RBF/Scale/Product/Additive assembled by kernel_gen.py:219-310 [gpytorch]."""
    s_all = F.softplus(prm.raw_outputscale)       # [R, L]
    l_all = F.softplus(prm.raw_lengthscale) if spec.n_ls else None
    nd = max(x1.dim(), x2.dim(), 3)
    total = None
    for r, comp in enumerate(spec.comps):
        term = None
        for f in comp.factors:
            a = x1[..., f.col].unsqueeze(-1)
            b = x2[..., f.col].unsqueeze(-2)
            if f.kind == SE:
                ell = l_all[f.ls].reshape([-1] + [1, 1])    # [L,1,1]
                k = torch.exp(-((a - b) ** 2) / (2.0 * ell ** 2))
            elif f.kind == CAT:
                k = (a - b == 0).to(DT)
            else:
                k = (a + b == 2).to(DT)
            term = k if term is None else term * k
        s = s_all[r].reshape([-1] + [1, 1])
        term = s * term
        total = term if total is None else total + term
    if total is None:
        raise ValueError("empty additive kernel")
    return total


# --------------------------------------------------------------------------------------
# KL upper bound (elbo_functions.py)
# --------------------------------------------------------------------------------------
def _chol_inv(A: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    Lc = torch.linalg.cholesky(A)
    eye = torch.eye(A.shape[-1], dtype=A.dtype, device=A.device)
    return Lc, torch.cholesky_solve(eye.expand_as(A).contiguous(), Lc)


def kld_terms(spec0, prm0, spec1, prm1, noise, m, H, x, mu, log_v, z, subjects: List[torch.Tensor],
              eps: float) -> Dict[str, torch.Tensor]:
    """All per-minibatch quantities of elbo_functions.py:220-283 (the per-subject form,
    which the fixed-T form 143-191 equals when every subject has T rows).

    subjects: list of row-index tensors, one per subject.  noise: [L] likelihood noise.
    Returns the scalars A..F, kld_qu_pu pieces, and the sufficient statistics S (ng_P2)
    and p (ng_P1) used by the natural-gradient pieces."""
    L, M = H.shape[0], H.shape[-1]
    K0xz = eval_additive(spec0, prm0, x, z)                                   # :222 / :147
    K0zz = eval_additive(spec0, prm0, z, z) + eps * torch.eye(M, dtype=DT)    # :223-224
    LK, iK = _chol_inv(K0zz)                                                  # :225-226
    LH, iH = _chol_inv(H)                                                     # :227-228
    r_all = (K0xz @ iK @ m).squeeze(2) - mu.T                                 # :230  [L,N]
    E_part = iK @ H @ iK                                                      # :231
    A = B = C = D = E = torch.zeros((), dtype=DT)
    S = torch.zeros(L, M, M, dtype=DT)
    p = torch.zeros(L, M, 1, dtype=DT)
    for idx in subjects:                                                      # :243-266
        xs = x[idx]
        T = xs.shape[0]
        xs_l = xs.unsqueeze(0).expand(L, T, xs.shape[1])
        K0ss = eval_additive(spec0, prm0, xs_l, xs_l)                         # :248
        Bs = eval_additive(spec1, prm1, xs_l, xs_l) + torch.eye(T, dtype=DT) * noise.reshape(L, 1, 1)  # :249-250
        LB, iB = _chol_inv(Bs)                                                # :251-252
        Ks = K0xz[:, idx]                                                     # :253
        Ss = Ks.transpose(1, 2) @ iB @ Ks                                     # :254
        rs = r_all[:, idx].unsqueeze(2)
        A = A + (rs.transpose(1, 2) @ iB @ rs).sum()                          # :256
        B = B + (torch.diagonal(iB, dim1=-1, dim2=-2) * torch.exp(log_v[idx].T)).sum()   # :257
        C = C + 2 * torch.log(torch.diagonal(LB, dim1=-2, dim2=-1)).sum()     # :258
        D = D + (iB * K0ss).sum() - (Ss * iK).sum()                           # :259
        E = E + (E_part * Ss).sum()                                           # :260
        p = p + Ks.transpose(1, 2) @ (iB @ mu[idx].T.unsqueeze(2))            # :263-265
        S = S + Ss                                                            # :266
    Fv = log_v.sum()                                                          # :268
    tr1 = (iK * H.transpose(-1, -2)).sum()                                    # :271
    qf1 = (m * (iK @ m)).sum()                                                # :272
    logdetK = 2 * torch.log(torch.diagonal(LK, dim1=-1, dim2=-2)).sum()       # :273
    logdetH = 2 * torch.log(torch.diagonal(LH, dim1=-1, dim2=-2)).sum()       # :274
    kq = 0.5 * (tr1 + qf1 - L * M + logdetK - logdetH)                        # :275
    return dict(A=A, B=B, C=C, D=D, E=E, F=Fv, kld_qu_pu=kq, tr1=tr1, qf1=qf1, logdetK=logdetK,
                logdetH=logdetH, S=S, p=p, iK=iK, iH=iH)


def _finish(t, scale, const, m, natural_gradient):
    kld_total = scale * 0.5 * (t['A'] + t['B'] + t['C'] + t['D'] + t['E'] - t['F']) + t['kld_qu_pu'] - const
    grad_m = grad_H = None
    if natural_gradient:                                                      # :186-191 / :279-283
        Bm = t['iK'] @ t['S'] @ t['iK'] + t['iK']
        grad_m = -(t['iK'] @ t['p']) + Bm @ m
        grad_H = 0.5 * (-t['iH'] + Bm)
    return kld_total, grad_m, grad_H


def split_subjects_by_id(x: torch.Tensor, id_covariate: int) -> List[torch.Tensor]:
    """elbo_functions.py:242-244: sorted unique ids, boolean row masks."""
    ids = x[:, id_covariate]
    return [torch.nonzero(ids == s, as_tuple=False).squeeze(1) for s in torch.unique(ids).tolist()]


def split_subjects_fixed_T(n_rows: int, T: int) -> List[torch.Tensor]:
    """elbo_functions.py:144,159: rows are taken subject-contiguous, T per subject."""
    return [torch.arange(s * T, (s + 1) * T) for s in range(n_rows // T)]


def minibatch_KLD_upper_bound_iter(spec0, prm0, spec1, prm1, noise, m, H, x, mu, log_v, z, P,
                                   P_in_current_batch, N, natural_gradient, id_covariate, eps,
                                   return_terms=False):
    """elbo_functions.py:196-285 (ragged T_s; constant term uses N = len(dataset), :277)."""
    t = kld_terms(spec0, prm0, spec1, prm1, noise, m, H, x, mu, log_v, z,
                  split_subjects_by_id(x, id_covariate), eps)
    L = H.shape[0]
    out = _finish(t, P / P_in_current_batch, L * N / 2, m, natural_gradient)
    return out + (t,) if return_terms else out


def minibatch_KLD_upper_bound(spec0, prm0, spec1, prm1, noise, m, H, x, mu, log_v, z, P_tot, P_batch, T,
                              natural_gradient, eps, return_terms=False):
    """elbo_functions.py:118-193 (fixed T; constant term L * P_tot * T / 2, :181)."""
    t = kld_terms(spec0, prm0, spec1, prm1, noise, m, H, x, mu, log_v, z,
                  split_subjects_fixed_T(x.shape[0], T), eps)
    L = H.shape[0]
    out = _finish(t, P_tot / P_batch, L * P_tot * T / 2, m, natural_gradient)
    return out + (t,) if return_terms else out


def natural_gradient_update(m, H, grad_m, grad_H, lr):
    """training.py:130-137."""
    _, iH = _chol_inv(H)
    iH_new = iH + lr * (grad_H + grad_H.transpose(-1, -2))
    _, H_new = _chol_inv(iH_new)
    m_new = H_new @ (iH @ m - lr * (grad_m - 2 * (grad_H @ m)))
    return m_new.detach(), H_new.detach()


def batch_predict(spec0, prm0, spec1, prm1, noise, prediction_x, test_x, mu, z, subjects: List[torch.Tensor],
                  id_covariate: int, eps: float) -> torch.Tensor:
    """GP posterior-mean prediction of the latent variables at `test_x`: utils.batch_predict_varying_T
    (utils.py:99-191; with every subject T rows long it equals utils.batch_predict, :193-271).  The
    reference's `torch.solve(B, A)` (removed from torch) is `torch.linalg.solve(A, B)`.

    prediction_x [N, Q] / mu [N, L]: covariates and encoder means of the conditioning rows; `subjects`:
    row-index tensors of prediction_x, one per subject; returns Z_pred [N_test, L]."""
    L, M = z.shape[0], z.shape[1]
    N = prediction_x.shape[0]
    K0xz = eval_additive(spec0, prm0, prediction_x, z)                        # :128 / :229
    K0zz = eval_additive(spec0, prm0, z, z) + eps * torch.eye(M, dtype=DT)    # :129,132 / :230,234
    K0Xz = eval_additive(spec0, prm0, test_x, z)                              # :130 / :232
    K0zx = K0xz.transpose(-1, -2)
    Hm = K0zz.clone()
    iB_mu = torch.zeros(L, N, 1, dtype=DT)
    iBs = []
    for idx in subjects:                                                      # :139-160
        xs = prediction_x[idx]
        T = xs.shape[0]
        xs_l = xs.unsqueeze(0).expand(L, T, xs.shape[1])
        Bs = eval_additive(spec1, prm1, xs_l, xs_l) + torch.eye(T, dtype=DT) * noise.reshape(L, 1, 1)   # :149-150
        _, iB = _chol_inv(Bs)                                                 # :152-153
        Ks = K0xz[:, idx]
        Hm = Hm + Ks.transpose(1, 2) @ iB @ Ks                                # :156-158
        iB_mu[:, idx] = iB @ mu[idx].T.unsqueeze(2)                           # :159
        iBs.append(iB)
    t = K0xz @ torch.linalg.solve(Hm, K0zx @ iB_mu)                           # :162
    corr = torch.zeros(L, N, 1, dtype=DT)
    for idx, iB in zip(subjects, iBs):                                        # :164-166
        corr[:, idx] = iB @ t[:, idx]
    mu_tilde = iB_mu - corr                                                   # :167
    first = K0Xz @ torch.linalg.solve(K0zz, K0zx @ mu_tilde)                  # :169
    test_ids = torch.unique(test_x[:, id_covariate])                          # :171-172
    mask = torch.isin(prediction_x[:, id_covariate], test_ids)
    second = torch.zeros(L, test_x.shape[0], 1, dtype=DT)
    xm = prediction_x[mask]
    for sid in test_ids:                                                      # :175-186
        sel = test_x[:, id_covariate] == sid
        xt = test_x[sel]
        K1Xx = eval_additive(spec1, prm1, xt.unsqueeze(0).expand(L, *xt.shape), xm.unsqueeze(0).expand(L, *xm.shape))
        second[:, sel] = K1Xx @ mu_tilde[:, mask]
    return (first + second).squeeze(2).T                                      # :188


def validation_dubo(spec0, prm0, spec1, prm1, noise, x, m, log_v, z, P: int, T: int, eps: float) -> torch.Tensor:
    """Deviance upper bound of the validation pass, validation.validation_dubo (validation.py:16-76): fixed T,
    summed over the latent dimensions.  m, log_v [N, L] (encoder means / log-variances of the validation rows);
    `torch.solve(m_st, B_st)` (:59) is B_st^-1 m_st."""
    L, M = z.shape[0], z.shape[1]
    v = torch.exp(log_v)                                                      # :34
    xs = x.reshape(P, T, -1)
    K0xz = eval_additive(spec0, prm0, x, z)                                   # :38
    K0zz = eval_additive(spec0, prm0, z, z) + eps * torch.eye(M, dtype=DT)    # :39
    LK = torch.linalg.cholesky(K0zz)
    iK = torch.cholesky_inverse(LK)                                           # :40-41
    X4 = xs.unsqueeze(1).expand(P, L, T, xs.shape[-1])                        # :37 stacked_x_st
    K0_all = eval_additive(spec0, prm0, X4, X4)                               # :42  [P, L, T, T]
    B_all = eval_additive(spec1, prm1, X4, X4) + torch.eye(T, dtype=DT) * noise.reshape(1, L, 1, 1)   # :43
    total = torch.zeros((), dtype=DT)
    for l in range(L):                                                        # :48-75
        K0_st, B_st = K0_all[:, l], B_all[:, l]
        LB = torch.linalg.cholesky(B_st)
        iB = torch.cholesky_inverse(LB)                                       # :44-45
        m_st = m[:, l].reshape(P, T, 1)
        v_st = v[:, l].reshape(P, T)
        K_st = K0xz[l].reshape(P, T, M)
        iB_K = iB @ K_st                                                      # :52
        S = K0xz[l].T @ iB_K.reshape(P * T, M)                                # :53
        W = K0zz[l] + S
        W = (W + W.T) / 2                                                     # :54-55
        LW = torch.linalg.cholesky(W)
        logdet = -2 * torch.log(torch.diagonal(LK[l])).sum() + 2 * torch.log(torch.diagonal(LB, dim1=-2, dim2=-1)).sum() \
            + 2 * torch.log(torch.diagonal(LW)).sum()                         # :57-60
        iB_m = iB @ m_st                                                      # :61
        qF1 = (m_st * iB_m).sum()
        p = K0xz[l].T @ iB_m.reshape(P * T)
        qF2 = (torch.linalg.solve_triangular(LW, p[:, None], upper=False) ** 2).sum()    # :64
        tr = (iB * K0_st).sum() - (S * iK[l]).sum()                           # :66
        logDetD = torch.log(v[:, l]).sum()
        tr_iB_D = (torch.diagonal(iB, dim1=-2, dim2=-1) * v_st).sum()         # :68
        Dh = (iB_K * torch.sqrt(v_st)[:, :, None]).reshape(P * T, M)
        tr2 = torch.diagonal(torch.cholesky_solve(Dh.T @ Dh, LW)).sum()       # :69-71
        total = total + 0.5 * ((tr_iB_D - tr2) + (qF1 - qF2) - P * T + logdet - logDetD + tr)   # :72-74
    return total.reshape(1)


def deviance_upper_bound(spec0, prm0, spec1, prm1, noise, x, m, log_v, z, P: int, T: int, eps: float) -> torch.Tensor:
    """elbo_functions.deviance_upper_bound (elbo_functions.py:60-115): ONE latent dimension (parameters [n, 1], m and
    log_v [P T], z [M, Q]).  Line for line the l-th summand of validation_dubo above (validation.py:48-75 is the
    batched copy of :75-114), so it is evaluated through it."""
    return validation_dubo(spec0, prm0, spec1, prm1, noise, x, m.reshape(-1, 1), log_v.reshape(-1, 1),
                           z.reshape(1, *z.shape[-2:]), P, T, eps).reshape(())


def elbo(spec0, prm0, spec1, prm1, noise, x, y, z, P: int, T: int, eps: float) -> torch.Tensor:
    """elbo_functions.elbo (elbo_functions.py:9-57): collapsed evidence lower bound of one latent dimension for a
    sample y [P T] of the latent; parameters [n, 1], z [M, Q]; `torch.solve(y_st, B_st)` (:47) is B_st^-1 y_st."""
    M = z.shape[-2]
    z1 = z.reshape(1, M, -1)
    xs = x.reshape(P, T, -1)
    K0xz = eval_additive(spec0, prm0, x, z1)[0]                                # :29
    K0zz = eval_additive(spec0, prm0, z1, z1)[0] + eps * torch.eye(M, dtype=DT)   # :30
    LK = torch.linalg.cholesky(K0zz)
    iK = torch.cholesky_inverse(LK)                                           # :31-32
    X4 = xs.unsqueeze(1)                                                      # [P, 1, T, Q]
    K0_st = eval_additive(spec0, prm0, X4, X4)[:, 0]                          # :33
    B_st = eval_additive(spec1, prm1, X4, X4)[:, 0] + torch.eye(T, dtype=DT) * noise.reshape(())   # :34-35
    LB = torch.linalg.cholesky(B_st)
    iB = torch.cholesky_inverse(LB)                                           # :36-37
    iB_K = iB @ K0xz.reshape(P, T, M)                                         # :39
    S = K0xz.T @ iB_K.reshape(P * T, M)                                       # :40
    W = K0zz + S
    W = (W + W.T) / 2                                                         # :41-42
    LW = torch.linalg.cholesky(W)
    logdet = -2 * torch.log(torch.diagonal(LK)).sum() + 2 * torch.log(torch.diagonal(LB, dim1=-2, dim2=-1)).sum() \
        + 2 * torch.log(torch.diagonal(LW)).sum()                             # :44-47
    y_st = y.reshape(P, T, 1)
    iB_y = iB @ y_st                                                          # :48
    qF1 = (y_st * iB_y).sum()
    p = K0xz.T @ iB_y.reshape(P * T)
    qF2 = (torch.linalg.solve_triangular(LW, p[:, None], upper=False) ** 2).sum()   # :51
    tr = (iB * K0_st).sum() - (S * iK).sum()                                  # :53
    log_like = -0.5 * T * P * math.log(2 * math.pi) - 0.5 * (logdet + qF1 - qF2)    # :54-55
    return log_like - 0.5 * tr                                                # :56


# --------------------------------------------------------------------------------------
# Heterogeneous likelihoods (HL_VAE/loglik.py) on the packed [N, E_x] / [N, P_theta] layout
# --------------------------------------------------------------------------------------
@dataclass
class VarDesc:
    """One data variable: type, number of classes, first data column, first theta column,
    and index inside its type group (for per-variable extra / normalisation params)."""
    kind: str
    nclass: int
    data_col: int
    theta_col: int
    group_pos: int


def build_layout(types: Sequence[Tuple[str, int]]) -> Tuple[List[VarDesc], int, int]:
    """Column layout of read_functions.py:65-124,144-173 (logvar_network=False): cat and
    ordinal variables occupy nclass data columns and nclass theta columns; real, pos and
    count one of each.  Returns (descs, E_x, P_theta)."""
    descs, e, p = [], 0, 0
    counters: Dict[str, int] = {}
    for kind, nclass in types:
        w = nclass if kind in ('cat', 'ordinal') else 1
        g = counters.get(kind, 0)
        counters[kind] = g + 1
        descs.append(VarDesc(kind, nclass if kind in ('cat', 'ordinal') else 1, e, p, g))
        e += w
        p += w
    return descs, e, p


def types_info_from_layout(types: Sequence[Tuple[str, int]], conv=False, logvar_network=False) -> dict:
    """types_info dict as read_functions.py:141-195 builds it (keys used on the hot path).  With logvar_network the
    real / positive variables own two parameter columns each (:164-173)."""
    type_tuple = [(k, str(c if k in ('cat', 'ordinal') else 1)) for k, c in types]
    set_of_types = sorted(set(type_tuple))
    data_idx, exp_idx, par_idx = [], [], []
    for k, c in type_tuple:
        gid = set_of_types.index((k, c))
        w = int(c) if k in ('cat', 'ordinal') else 1
        data_idx.append(gid)
        exp_idx += [gid] * w
        par_idx += [gid] * (2 if (logvar_network and k in ('real', 'pos')) else w)
    return dict(types_dict=[dict(type=k, dim=1, nclass=int(c)) for k, c in type_tuple],
                set_of_types=set_of_types,
                data_types_indexes=np.array(data_idx, dtype=float),
                exp_types_indexes=np.array(exp_idx, dtype=float),
                param_indexes=np.array(par_idx, dtype=float),
                beta_ranges=[], conv=conv, use_ranges=False, conv_range=255)


def loglik_real(data, mask, theta, log_vy, norm_mean=None, norm_var=None):
    """loglik.py:27-70 with extra_params (per-variable log-variance).  norm_* None is the
    conv case (`normalization_params == []`, :36-41)."""
    if norm_var is None:
        dmean, dvar = torch.tensor(0., dtype=DT), torch.tensor(1., dtype=DT)
    else:
        dmean, dvar = norm_mean, torch.clamp(norm_var, min=3e-4)              # :38
    est_log_vy = -8.0 + F.softplus(log_vy + 8.0)                              # :51
    est_var = dvar * torch.exp(est_log_vy)                                    # :52,56
    est_mean = torch.sqrt(dvar) * theta + dmean                               # :55
    lp = -0.5 * (data - est_mean) ** 2 / est_var - 0.5 * math.log(2 * math.pi) - 0.5 * torch.log(est_var)   # :58
    return lp * mask, lp * (1.0 - mask), est_mean                             # :62-67


def loglik_pos(data, mask, theta, log_vy, norm_mean, norm_var):
    """loglik.py:73-121 (log-normal on log(1+x)) with extra_params."""
    lvar = torch.clamp(norm_var, min=1e-3)                                    # :80
    ld = torch.log(1.0 + data)                                                # :84
    est_mean = torch.sqrt(lvar) * theta + norm_mean                           # :96
    est_var = lvar * torch.exp(log_vy)                                        # :100
    lp = -0.5 * (ld - est_mean) ** 2 / est_var - 0.5 * torch.log(2 * math.pi * est_var) - ld   # :102
    return lp * mask, lp * (1.0 - mask), est_mean


def loglik_real_rowvar(data, mask, theta, norm_mean=None, norm_var=None):
    """loglik.py:27-70 with extra_params = None (variance network, :45-48): theta = [means, raw log-variances]."""
    D = data.shape[1]
    dvar = torch.clamp(norm_var, min=3e-4) if norm_var is not None else torch.tensor(1.0, dtype=DT)
    dmean = norm_mean if norm_mean is not None else torch.tensor(0.0, dtype=DT)
    est_mean, raw = theta[:, :D], theta[:, D:2 * D]                            # :46
    est_var = torch.exp(-8.0 + F.softplus(raw + 8.0))                          # :47-48
    est_mean = torch.sqrt(dvar) * est_mean + dmean                            # :55
    est_var = dvar * est_var                                                  # :56
    lp = -0.5 * (data - est_mean) ** 2 / est_var - 0.5 * math.log(2 * math.pi) - 0.5 * torch.log(est_var)   # :58
    return lp * mask, lp * (1.0 - mask), est_mean, est_var


def loglik_pos_rowvar(data, mask, theta, norm_mean, norm_var):
    """loglik.py:73-121 with extra_params = None (:89,104-108): per-row log-variances from theta."""
    D = data.shape[1]
    lvar = torch.clamp(norm_var, min=1e-3)
    ld = torch.log(1.0 + data)
    est_mean = torch.sqrt(lvar) * theta[:, :D] + norm_mean
    est_var = lvar * torch.exp(theta[:, D:2 * D])                             # :108
    lp = -0.5 * (ld - est_mean) ** 2 / est_var - 0.5 * torch.log(2 * math.pi * est_var) - ld
    return lp * mask, lp * (1.0 - mask), est_mean, est_var


def loglik_beta(data, mask, theta, ranges, disp):
    """loglik.py:216-256.  ranges [D, 2] = (min, max + 1e-3) per variable, disp the raw dispersion parameter.
    theta narrower than 2 D columns -> the reference's fallback (:232-235): one [N, 1] parameter, theta[:, 0],
    shared by the whole group."""
    D = data.shape[1]
    dmin, dmax = ranges[:, 0], ranges[:, 1]
    xc = (data - dmin) / (dmax - dmin) + 1e-6                                  # :224
    est = theta[:, :D] if theta.shape[1] >= 2 * D else theta[:, 0:1]           # :230-235
    phi = torch.clamp(F.softplus(disp.reshape(-1)[0]), 1e-6, 1e20)             # :240
    mean = 0.5 * (1.0 + torch.erf(est / math.sqrt(2.0)))                       # :241-242
    al, be = phi * mean, phi * (1.0 - mean)                                    # :244-245
    lp = (al - 1) * torch.log(xc) + (be - 1) * torch.log(1 - xc) - torch.lgamma(al) - torch.lgamma(be) \
        + torch.lgamma(al + be)                                                # :247-248
    return lp * mask, lp * (1.0 - mask), al, be


def loglik_cat(data, mask, theta, C):
    """loglik.py:124-146.  data one-hot [N, D*C], theta [N, D*C] -> params log pi [N,D,C]."""
    N, D = mask.shape
    log_pi = theta.reshape(N, D, C)
    log_pi = log_pi - torch.logsumexp(log_pi, 2).reshape(N, D, 1)             # :134
    lp = (data.reshape(N, D, C) * F.log_softmax(log_pi, 2)).sum(-1)           # :135
    return lp * mask, lp * (1.0 - mask), log_pi


def loglik_ordinal(data, mask, theta, C):
    """loglik.py:149-188.  data thermometer [N, D*C]; theta [N, D, C] = C-1 thresholds, 1 location."""
    N, D = mask.shape
    eps = 1e-6
    d = data.reshape(N, D, C)
    th = theta.reshape(N, D, C)
    part, loc = th[:, :, :-1], th[:, :, -1]
    loc = F.softplus(loc[:, :, None])                                         # :163
    thr = torch.cumsum(torch.clamp(F.softplus(part), eps, 1e20), 2)           # :164
    sg = torch.sigmoid(thr - loc)                                             # :165
    one = torch.ones(N, D, 1, dtype=DT)
    probs = torch.cat([sg, one], 2) - torch.cat([torch.zeros(N, D, 1, dtype=DT), sg], 2)   # :166-167
    probs = torch.clamp(probs, eps, 1.0)                                      # :169
    vals = d.detach().to(torch.int32).sum(2)                                  # :172
    vals[mask == 0] = 1                                                       # :173
    onehot = F.one_hot((vals - 1).long(), C).to(DT)                           # :174
    probs = probs / probs.sum(2).reshape(N, D, 1)                             # :178
    lp = (onehot * F.log_softmax(torch.log(probs), -1)).sum(-1)               # :179
    return lp * mask, lp * (1.0 - mask), probs


def loglik_count(data, mask, theta):
    """loglik.py:191-213: Poisson with rate clamp(softplus(theta), 1e-6, 1e20)."""
    lam = torch.clamp(F.softplus(theta), 1e-6, 1e20)                          # :203
    lp = data * torch.log(lam) - lam - torch.lgamma(data + 1.0)               # td.Poisson.log_prob, :205-206
    return lp * mask, lp * (1.0 - mask), lam


def loglik_and_reconstruction(descs: List[VarDesc], data, mask, theta, log_vy_real=None, log_vy_pos=None,
                              norm_real=None, norm_pos=None, conv=False):
    """HLVAE.py:381-414 on the packed layout: returns log_p_x [N,D], log_p_x_missing [N,D]
    and params [N,P_theta] in the column order of read_functions.py:206-218."""
    N = data.shape[0]
    D = len(descs)
    lpx = torch.zeros(N, D, dtype=DT)
    lpm = torch.zeros(N, D, dtype=DT)
    ptheta = sum(v.nclass for v in descs)
    params = torch.zeros(N, ptheta, dtype=DT)
    groups: Dict[Tuple[str, int], List[int]] = {}
    for j, v in enumerate(descs):
        groups.setdefault((v.kind, v.nclass), []).append(j)
    for (kind, C), js in groups.items():
        dcols = torch.tensor([descs[j].data_col + c for j in js for c in range(C)])
        pcols = torch.tensor([descs[j].theta_col + c for j in js for c in range(C)])
        jj = torch.tensor(js)
        gpos = torch.tensor([descs[j].group_pos for j in js])
        dd, th, mk = data[:, dcols], theta[:, pcols], mask[:, jj]
        if kind == 'real':
            if conv:
                dd = dd / 255                                                 # HLVAE.py:393-394
            nm, nv = (None, None) if norm_real is None else (norm_real[0][gpos], norm_real[1][gpos])
            a, b, prm = loglik_real(dd, mk, th, log_vy_real[gpos], nm, nv)
        elif kind == 'pos':
            a, b, prm = loglik_pos(dd, mk, th, log_vy_pos[gpos], norm_pos[0][gpos], norm_pos[1][gpos])
        elif kind == 'count':
            a, b, prm = loglik_count(dd, mk, th)
        elif kind == 'cat':
            a, b, prm = loglik_cat(dd, mk, th, C)
        elif kind == 'ordinal':
            a, b, prm = loglik_ordinal(dd, mk, th, C)
        else:
            raise ValueError(kind)
        lpx[:, jj] = a
        lpm[:, jj] = b
        params[:, pcols] = prm.reshape(N, -1)
    return lpx, lpm, params


def statistics(descs: List[VarDesc], params, log_vy_pos=None):
    """read_functions.py:268-302: mean and mode per variable; categorical and ordinal use
    argmax over the class axis of `params` (first index wins ties, as torch.argmax)."""
    N = params.shape[0]
    mean = torch.zeros(N, len(descs), dtype=DT)
    mode = torch.zeros(N, len(descs), dtype=DT)
    for j, v in enumerate(descs):
        pr = params[:, v.theta_col:v.theta_col + v.nclass]
        if v.kind == 'real':
            mean[:, j] = pr[:, 0]
            mode[:, j] = pr[:, 0]                                             # :275-278
        elif v.kind == 'pos':
            var = torch.exp(log_vy_pos[v.group_pos])                          # :284
            mean[:, j] = torch.exp(pr[:, 0] + 0.5 * var) - 1.0                # :287
            mode[:, j] = torch.exp(pr[:, 0] - var) - 1.0                      # :289
        elif v.kind == 'count':
            mean[:, j] = pr[:, 0]
            mode[:, j] = torch.floor(pr[:, 0])                                # :293-295
        else:
            am = torch.argmax(pr, 1).to(DT)                                   # :296-302
            mean[:, j] = am
            mode[:, j] = am
    return mean, mode


def discrete_variables_transformation(descs: List[VarDesc], data):
    """read_functions.py:221-235."""
    out = torch.zeros(data.shape[0], len(descs), dtype=DT)
    for j, v in enumerate(descs):
        d = data[:, v.data_col:v.data_col + v.nclass]
        if v.kind == 'cat':
            out[:, j] = torch.argmax(d, 1).to(DT)
        elif v.kind == 'ordinal':
            out[:, j] = d.sum(1) - 1
        else:
            out[:, j] = d[:, 0]
    return out


def batch_norm_params(descs: List[VarDesc], data, mask):
    """Normalisation parameters fed to loglik_real / loglik_pos, HL_VAE/utils.py:96-131
    (non-conv).  Returns (norm_real, norm_pos) each (mean [D_g], var [D_g]) or None."""
    def cols(kind):
        js = [j for j, v in enumerate(descs) if v.kind == kind]
        return js, [descs[j].data_col for j in js]
    out = []
    for kind in ('real', 'pos'):
        js, dc = cols(kind)
        if not js:
            out.append(None)
            continue
        mk = mask[:, js]
        obs = data[:, dc] * mk
        if kind == 'pos':
            obs = torch.log(1.0 + obs)                                        # utils.py:124
        mean = (obs * mk).sum(0) / mk.sum(0)                                  # :106 / :125
        var = (((obs - mean) * mk) ** 2).sum(0) / mk.sum(0)                   # :107 / :126
        if kind == 'pos':
            var = torch.clamp(var, 1e-6, 1e20)                                # :127
        out.append((mean, var))
    return out[0], out[1]


# --------------------------------------------------------------------------------------
# Observation heads: y -> theta (HLVAE.py:11-89, 416-453), SURVEY.md 8(f) row 2
# --------------------------------------------------------------------------------------
def head_forward(kind: str, prm: Dict[str, torch.Tensor], gamma: torch.Tensor, logvar_network=False) -> torch.Tensor:
    """One Observation_* module on gamma [N, d, Y] -> [N, d, dim].
    count: HLVAE.py:21-23; real / pos: :42-52 (mean head; with logvar_network the log-variance head's output is
    appended along the VARIABLE axis, :51, so the group holds all means, then all raw log-variances);
    cat: :63-68 (a zero logit in front); ordinal: :84-89 (thresholds repeated over the batch, then the region)."""
    lin = lambda w, b: torch.einsum("bdy,dya->bda", gamma, w) + b
    if kind == 'count':
        return lin(prm['weight'], prm['bias'])
    if kind in ('real', 'pos'):
        if logvar_network:
            return torch.cat([lin(prm['weight_mean'], prm['bias_mean']), lin(prm['weight_logvar'], prm['bias_logvar'])], 1)
        return lin(prm['weight_mean'], prm['bias_mean'])
    if kind == 'cat':
        th = lin(prm['weight'], prm['bias'])
        zero = torch.zeros(th.shape[0], th.shape[1], 1, dtype=DT)
        return torch.cat((zero, th), dim=-1)
    if kind == 'ordinal':
        thr = prm['weight_thresholds'].repeat((gamma.shape[0], 1, 1))
        return torch.cat((thr, lin(prm['weight_region'], prm['bias_region'])), dim=-1)
    raise NotImplementedError(kind)


def theta_estimation(types: Sequence[Tuple[str, int]], heads: List[Dict[str, torch.Tensor]], y: torch.Tensor,
                     mask: torch.Tensor, conv: bool = False, logvar_network: bool = False) -> torch.Tensor:
    """HLVAE.theta_estimation (HLVAE.py:416-453).  `heads[i]` holds the parameters of the i-th type group's
    module, groups ordered as types_info['set_of_types'].  Follows the reference literally: heads on y * mask
    (:419,424-427), Sigmoid on the real group of the convolutional model (:429-431), times the parameter mask
    (:433-434); the same on y * (1 - mask) under no_grad (:436-446); missing entries overwrite (:449-453)."""
    ti = types_info_from_layout(types, conv=conv, logvar_network=logvar_network)
    N = y.shape[0]
    P = len(ti['param_indexes'])
    pm = torch.zeros(N, P, dtype=DT)                        # read_functions.py:147,172-185: mask per parameter column
    for i, tpl in enumerate(ti['set_of_types']):
        vsel = torch.tensor(ti['data_types_indexes'] == i)
        pcols = torch.nonzero(torch.tensor(ti['param_indexes'] == i))[:, 0]
        mg = mask[:, vsel]
        if tpl[0] in ('real', 'pos') and logvar_network:
            pm[:, pcols] = torch.cat([mg, mg], 1)           # :176-180
        else:
            pm[:, pcols] = mg.repeat_interleave(int(tpl[1]) if tpl[0] in ('cat', 'ordinal') else 1, dim=1)
    theta = torch.zeros(N, P, dtype=DT)
    observed_y = y * mask[:, :, None]
    missing_y = y * (1 - mask)[:, :, None]
    for i, tpl in enumerate(ti['set_of_types']):
        vsel = torch.tensor(ti['data_types_indexes'] == i)
        psel = torch.tensor(ti['param_indexes'] == i)
        dim = int(tpl[1])
        pmi = pm[:, psel].reshape(N, -1, dim)
        cov = int(vsel.sum())

        def sig(o):                                         # Sigmoid on the first cov_dim entries only (:429-431)
            return torch.cat([torch.sigmoid(o[:, :cov]), o[:, cov:]], 1) if (tpl[0] == 'real' and conv) else o

        obs = sig(head_forward(tpl[0], heads[i], observed_y[:, vsel, :], logvar_network)) * pmi
        with torch.no_grad():
            mis = sig(head_forward(tpl[0], heads[i], missing_y[:, vsel, :], logvar_network)) * (1 - pmi)
        merged = torch.where(pmi == 0, mis, obs).reshape(N, -1)
        theta = theta.index_put((torch.arange(N)[:, None], torch.nonzero(psel)[:, 0][None, :]), merged)
    return theta


def batch_normalization(descs: List[VarDesc], data, mask, conv=False):
    """HL_VAE/utils.py:88-143 on the packed layout: the encoder's input X_list [N, E_x] and the normalisation
    parameters (norm_real, norm_pos), each (mean [D_g], var [D_g]) or None.
    real: conv -> observed / 255, no parameters (:99-103); else masked mean / variance and
    (observed - mean) / sqrt(var + 1e-5) * mask (:104-108).  count: log of the observed value, 0 where missing
    (:113-119).  pos: the same standardisation on log(1 + observed), variance clamped to [1e-6, 1e20] (:120-131).
    cat / ordinal: data times the mask repeated per class (:132-138)."""
    out = torch.zeros_like(data)
    params = {'real': None, 'pos': None}
    for kind in ('real', 'pos'):
        js = [j for j, v in enumerate(descs) if v.kind == kind]
        if not js:
            continue
        dc = [descs[j].data_col for j in js]
        mk = mask[:, js]
        obs = data[:, dc] * mk
        if kind == 'real' and conv:
            out[:, dc] = obs / 255
            continue
        if kind == 'pos':
            obs = torch.log(1.0 + obs)
        mean = (obs * mk).sum(0) / mk.sum(0)
        var = (((obs - mean) * mk) ** 2).sum(0) / mk.sum(0)
        if kind == 'pos':
            var = torch.clamp(var, 1e-6, 1e20)
        out[:, dc] = (obs - mean[None, :]) / torch.sqrt(var + 1e-5) * mk
        params[kind] = (mean, var)
    for j, v in enumerate(descs):
        c0 = v.data_col
        if v.kind == 'count':
            aux = torch.log(data[:, c0] * mask[:, j])
            out[:, c0] = torch.where(mask[:, j] == 0, torch.zeros_like(aux), aux)
        elif v.kind in ('cat', 'ordinal'):
            out[:, c0:c0 + v.nclass] = data[:, c0:c0 + v.nclass] * mask[:, j:j + 1]
    return out, params['real'], params['pos']


# --------------------------------------------------------------------------------------
# Minibatch composition of the reference's samplers (utils.py:36-97, training.py:38-47)
# --------------------------------------------------------------------------------------
def fixed_T_batches(P: int, T: int, batch_size: int) -> List[List[int]]:
    """BatchSampler(SubjectSampler(dataset, P, T), batch_size, drop_last=False): a np.random.shuffle permutation of
    the subjects, rows T s .. T (s+1) - 1 of each in turn (utils.py:43-48), cut into batches of `batch_size` rows."""
    import itertools
    r = np.arange(P)
    np.random.shuffle(r)
    order = list(itertools.chain.from_iterable([list(range(T * x, T * (x + 1))) for x in r]))
    return [order[b:b + batch_size] for b in range(0, len(order), batch_size)]


def varying_T_batches(subject_ids: Sequence[int], subjects_per_batch: int) -> List[List[int]]:
    """VaryingLengthBatchSampler(VaryingLengthSubjectSampler(dataset, id_covariate), subjects_per_batch): subjects own
    the rows from their first occurrence to the next subject's first occurrence (utils.py:63-65), permuted with
    np.random.shuffle (:68-69); a batch closes when a new subject arrives and it already holds
    `subjects_per_batch` subjects (:86-96)."""
    from collections import OrderedDict
    l = [int(v) for v in subject_ids]
    uniq = list(OrderedDict.fromkeys(l))
    starts = [l.index(x) for x in uniq]
    ends = starts[1:] + [len(l)]
    r = np.arange(len(uniq))
    np.random.shuffle(r)
    batches, batch, seen = [], [], set()
    for x in r:
        for i in range(starts[x], ends[x]):
            if x not in seen:
                if len(seen) == subjects_per_batch:
                    batches.append(batch)
                    batch, seen = [], set()
                seen.add(x)
            batch.append(i)
    batches.append(batch)
    return batches


# --------------------------------------------------------------------------------------
# One ELBO-path step (training.py:83,104-137): used by the CPU baseline in bench.py
# --------------------------------------------------------------------------------------
def elbo_path_step(state: dict, natural_gradient_lr=0.01) -> Dict[str, torch.Tensor]:
    """nll + KL forward, backward to (theta, mu, log_v, Z, kernel raw parameters, log_vy),
    natural-gradient update of (m, H).  `state` holds every tensor of one minibatch."""
    s = state
    for t in s['leaves']:
        t.grad = None
    lpx, _, params = loglik_and_reconstruction(s['descs'], s['data'], s['mask'], s['theta'], s.get('log_vy_real'),
                                               s.get('log_vy_pos'), s.get('norm_real'), s.get('norm_pos'),
                                               s.get('conv', False))
    nll = -(lpx.sum(1)).sum() * s['P'] / s['P_b']                             # training.py:83,104,122
    recon, _ = statistics(s['descs'], params.detach(), s.get('log_vy_pos'))
    kld, gm, gH = minibatch_KLD_upper_bound_iter(s['spec0'], s['prm0'], s['spec1'], s['prm1'], s['noise'], s['m'],
                                                 s['H'], s['x'], s['mu'], s['log_v'], s['z'], s['P'], s['P_b'],
                                                 s['N'], True, s['id_covariate'], s['eps'])
    loss = nll + kld                                                          # training.py:124
    loss.backward()                                                           # :127
    m_new, H_new = natural_gradient_update(s['m'], s['H'], gm.detach(), gH.detach(), natural_gradient_lr)
    return dict(loss=loss.detach(), nll=nll.detach(), kld=kld.detach(), m=m_new, H=H_new, recon=recon)
