"""Synthetic inputs shaped like the reference's data (no files, no network).

Covariate layout follows the HealthMNIST label order after dataset_def.py:47:
[time_age, disease_time, subject id, gender, disease, location]
(Heterogeneous_Health_MNIST_generate.py:105-106,152-188).  Likelihood inputs follow the
encodings of HL_VAE/read_functions.py:65-124 (one-hot, thermometer, count shift).
"""
from __future__ import annotations

import numpy as np
import torch

DEFAULT_KERNEL_ARGS = dict(            # config/hlvae_config_file.txt:41-46, id_covariate=2
    cat_kernel=[2], bin_kernel=[], sqexp_kernel=[0],
    cat_int_kernel=[{'cont_covariate': 0, 'cat_covariate': 2},
                    {'cont_covariate': 0, 'cat_covariate': 3},
                    {'cont_covariate': 1, 'cat_covariate': 4}],
    bin_int_kernel=[], covariate_missing_val=[], id_covariate=2)

SWEEP_KERNEL_ARGS = dict(              # BASELINE.json configs[2]: SE(time)+CA(id)+SE(age)xCA(sex)
    cat_kernel=[2], bin_kernel=[], sqexp_kernel=[0],
    cat_int_kernel=[{'cont_covariate': 1, 'cat_covariate': 3}],
    bin_int_kernel=[], covariate_missing_val=[], id_covariate=2)

MASKED_KERNEL_ARGS = dict(             # exercises BinKernel, mask factors and bin x SE
    cat_kernel=[2, 3], bin_kernel=[4], sqexp_kernel=[0, 1],
    cat_int_kernel=[{'cont_covariate': 1, 'cat_covariate': 2},
                    {'cont_covariate': 0, 'cat_covariate': 3}],
    bin_int_kernel=[{'cont_covariate': 1, 'bin_covariate': 5}],
    covariate_missing_val=[{'covariate': 1, 'mask': 4}], id_covariate=2)


def covariates(n_subjects, T, rng, ragged=False, t_min=5, first_id=0, continuous_age=False):
    """[N, 6] float64 covariates, subject-contiguous rows; returns (x, rows_per_subject)."""
    rows, lens = [], []
    for s in range(n_subjects):
        Ts = int(rng.integers(t_min, T + 1)) if ragged else T
        sid = first_id + s
        t = np.arange(Ts, dtype=np.float64)
        if ragged:
            t = np.sort(rng.choice(T, Ts, replace=False)).astype(np.float64)
        sick = rng.random() < 0.5
        onset = rng.integers(0, T)
        dis_t = (t - onset) if sick else np.zeros(Ts)
        if continuous_age:
            dis_t = rng.uniform(40, 80) + 0.5 * t
        loc = float(rng.random() < 0.5)
        x = np.stack([t, dis_t, np.full(Ts, sid, dtype=np.float64), np.full(Ts, float(sid % 2)),
                      np.full(Ts, float(sick)), np.full(Ts, loc)], 1)
        rows.append(x)
        lens.append(Ts)
    return torch.from_numpy(np.concatenate(rows, 0)), lens


def inducing_points(x_pool, L, M, rng):
    """HLVAE_main.py:224-226: per latent dimension, M rows of the covariates without replacement."""
    z = torch.zeros(L, M, x_pool.shape[1], dtype=torch.float64)
    for i in range(L):
        z[i] = x_pool[rng.choice(x_pool.shape[0], M, replace=False)]
    return z


def variational_state(L, M, gen):
    """HLVAE_main.py:259-263: m ~ N(0,1), H = (G/10)(G/10)^T."""
    m = torch.randn(L, M, 1, generator=gen).double()
    G = (torch.randn(L, M, M, generator=gen) / 10).double()
    return m, G @ G.transpose(-1, -2)


HEALTHMNIST_D4_TYPES = [('real', 1) if (i // 36 < 18 and i % 36 < 18) else ('cat', 5) for i in range(1296)]
TABULAR_TYPES = ([('count', 1)] * 64 + [('ordinal', 5)] * 64 + [('cat', 5)] * 64 + [('real', 1)] * 32 + [('pos', 1)] * 32)


def mixed_types(rng, n_vars=20, max_class=6):
    kinds = ['real', 'pos', 'count', 'cat', 'ordinal']
    out = []
    for i in range(n_vars):
        k = kinds[i % 5] if i < 10 else kinds[int(rng.integers(0, 5))]
        out.append((k, int(rng.integers(2, max_class + 1)) if k in ('cat', 'ordinal') else 1))
    return out


def likelihood_batch(types, N, rng, observed=0.7, pixel_like=False):
    """(data [N,E_x], mask [N,D]) in the reference's encodings."""
    cols = []
    for kind, C in types:
        if kind == 'cat':
            c = rng.integers(0, C, N)
            cols.append(np.eye(C)[c])
        elif kind == 'ordinal':
            c = rng.integers(0, C, N)
            cols.append((np.arange(C)[None, :] <= c[:, None]).astype(np.float64))
        elif kind == 'count':
            cols.append((rng.poisson(3.0, N) + 1).astype(np.float64)[:, None])
        elif kind == 'pos':
            cols.append(np.exp(rng.normal(0, 1, N))[:, None])
        else:
            v = rng.integers(0, 256, N).astype(np.float64) if pixel_like else rng.normal(0, 1, N)
            cols.append(v[:, None])
    data = np.concatenate(cols, 1)
    mask = (rng.random((N, len(types))) < observed).astype(np.float64)
    return torch.from_numpy(data), torch.from_numpy(mask)


def device_likelihood_batch(layout, N, device, gen, dtype=torch.float32, observed=0.75, pixel_like=True):
    """Same encodings as likelihood_batch, generated directly on `device` for large N (bench sizes).
    `layout` is a hlvae_b200.loglik.VarLayout; returns (data [N,E_x] dtype, mask [N,D] uint8)."""
    D = layout.D
    data = torch.zeros(N, layout.E_x, dtype=dtype, device=device)
    kinds = layout.var_kind.tolist()
    ncls = layout.var_nclass.to(device)
    dcol = layout.var_dcol.to(device).long()
    u = torch.rand(N, D, device=device, generator=gen)
    cls = (u * ncls[None, :]).floor().long().clamp_(min=0)
    cls = torch.minimum(cls, (ncls[None, :] - 1).long())
    kind_t = torch.tensor(kinds, device=device)
    for name, code in (("cat", 3), ("ordinal", 4)):
        idx = torch.nonzero(kind_t == code).squeeze(1)
        if idx.numel() == 0:
            continue
        if name == "cat":
            data.scatter_(1, dcol[idx][None, :] + cls[:, idx], 1.0)
        else:
            cmax = int(ncls[idx].max())
            for c in range(cmax):
                sel = idx[ncls[idx] > c]
                data[:, dcol[sel] + c] = (cls[:, sel] >= c).to(dtype)
    idx = torch.nonzero(kind_t == 0).squeeze(1)
    if idx.numel():
        v = (torch.rand(N, idx.numel(), device=device, generator=gen) * 256).floor() if pixel_like else \
            torch.randn(N, idx.numel(), device=device, generator=gen)
        data[:, dcol[idx]] = v.to(dtype)
    idx = torch.nonzero(kind_t == 1).squeeze(1)
    if idx.numel():
        data[:, dcol[idx]] = torch.exp(torch.randn(N, idx.numel(), device=device, generator=gen)).to(dtype)
    idx = torch.nonzero(kind_t == 2).squeeze(1)
    if idx.numel():
        lam = torch.full((N, idx.numel()), 3.0, device=device)
        data[:, dcol[idx]] = (torch.poisson(lam, generator=gen) + 1).to(dtype)
    mask = (torch.rand(N, D, device=device, generator=gen) < observed).to(torch.uint8)
    return data, mask
