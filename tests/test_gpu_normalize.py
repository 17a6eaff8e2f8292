"""GPU parity of the batch-normalisation kernels (hlvae_batch_norm_stats / _apply) against the
unmodified reference's HL_VAE.utils.batch_normalization outputs (tests/golden/norm_*.npz) and the oracle on fresh
seeds.  float64 storage: 1e-12 relative; float32 storage / uint8 inputs: 2e-6."""
import numpy as np
import pytest
import torch

import helpers as h
from hlvae_b200 import normalize as nz, synth
from oracle import hlvae_oracle as orc

pytestmark = pytest.mark.gpu
DT = torch.float64


def _check_params(params, g, tol):
    for i, tag in enumerate(("real", "pos")):
        if g[tag + "_mean"].size == 0:
            assert params[i] == []
        else:
            assert h.rel_err(params[i][0], g[tag + "_mean"]) < tol and h.rel_err(params[i][1], g[tag + "_var"]) < tol


@pytest.mark.parametrize("name", h.NORM_CASES)
def test_golden_f64(name, device):
    g = h.load(name)
    types, conv = h.parse_types(g), bool(int(g["conv"]))
    ti = orc.types_info_from_layout(types, conv=conv)
    X, params = nz.batch_normalization(h.t(g["data"], device), h.t(g["mask"], device), None, ti)
    assert X.dtype == torch.float64 and h.rel_err(X, g["X"]) < 1e-12
    _check_params(params, g, 1e-12)


@pytest.mark.parametrize("name", h.NORM_CASES)
def test_golden_f32_and_u8_mask(name, device):
    g = h.load(name)
    types, conv = h.parse_types(g), bool(int(g["conv"]))
    ti = orc.types_info_from_layout(types, conv=conv)
    X, params = nz.batch_normalization(h.t(g["data"], device).float(), h.t(g["mask"], device).to(torch.uint8), None, ti)
    assert X.dtype == torch.float32 and h.rel_err(X, g["X"]) < 2e-6
    _check_params(params, g, 2e-6)


@pytest.mark.parametrize("types,conv,N", [(synth.TABULAR_TYPES, False, 4097), (synth.HEALTHMNIST_D4_TYPES, True, 257),
                                          (synth.HEALTHMNIST_D4_TYPES, False, 130)])
def test_random_vs_oracle(types, conv, N, device):
    rng = np.random.default_rng(N)
    data, mask = synth.likelihood_batch(types, N, rng, observed=0.7, pixel_like=conv)
    descs, _, _ = orc.build_layout(types)
    Xo, nr, npos = orc.batch_normalization(descs, data, mask, conv)
    lay = nz.NormLayout(types, conv, device)
    X, mean, var = nz.normalize(lay, data.to(device), mask.to(device))
    assert h.rel_err(X, Xo) < 1e-11
    for tag, ref in (("real", nr), ("pos", npos)):
        if ref is not None:
            idx = lay.var.idx[tag]
            assert h.rel_err(mean[idx], ref[0]) < 1e-12 and h.rel_err(var[idx], ref[1]) < 1e-11
    if conv:                                         # pixel values and one-hot codes are exact in uint8
        X8, _, _ = nz.normalize(lay, data.to(torch.uint8).to(device), mask.to(torch.uint8).to(device), out_dtype=DT)
        assert torch.equal(X8, X)


def test_all_missing_column_is_nan_like_reference(device):
    """A variable with no observed entry: the reference divides by sum(mask) = 0 (utils.py:106) -> NaN column."""
    types = [('real', 1), ('real', 1), ('cat', 3)]
    rng = np.random.default_rng(0)
    data, mask = synth.likelihood_batch(types, 12, rng)
    mask[:, 1] = 0
    descs, _, _ = orc.build_layout(types)
    Xo, nr, _ = orc.batch_normalization(descs, data, mask, False)
    X, mean, var = nz.normalize(nz.NormLayout(types, False, device), data.to(device), mask.to(device))
    assert torch.isnan(Xo[:, 1]).all() and torch.isnan(X[:, 1]).all()
    keep = [0, 2, 3, 4]
    assert h.rel_err(X[:, keep], Xo[:, keep]) < 1e-12


def test_properties_full_size(device):
    """configs[3]-sized tabular batch (64 000 rows, 30 % missing), float32: standardised columns have masked mean 0 and
    masked variance var / (var + 1e-5); categorical blocks are data times mask exactly."""
    types = synth.TABULAR_TYPES
    N = 64000
    lay = nz.NormLayout(types, False, device)
    gen = torch.Generator(device=device).manual_seed(1)
    data, mask = synth.device_likelihood_batch(lay.var, N, device, gen, dtype=torch.float32, observed=0.7, pixel_like=False)
    X, mean, var = nz.normalize(lay, data, mask)
    m = mask.double()
    for tag in ("real", "pos"):
        idx = lay.var.idx[tag]
        cols = lay.var.var_dcol.long()[idx]
        xm = X[:, cols].double()
        mk = m[:, idx]
        mu = (xm * mk).sum(0) / mk.sum(0)
        v2 = ((xm * mk) ** 2).sum(0) / mk.sum(0)
        assert float(mu.abs().max()) < 1e-4
        assert h.rel_err(v2, var[idx] / (var[idx] + 1e-5)) < 1e-3
    ic = lay.var.idx["cat"]
    c0 = int(lay.var.var_dcol[ic[0]])
    assert torch.equal(X[:, c0:c0 + 5], data[:, c0:c0 + 5] * mask[:, ic[0]].unsqueeze(1).float())


def test_empty_batch(device):
    types = [('real', 1), ('cat', 3), ('pos', 1)]
    lay = nz.NormLayout(types, False, device)
    X, mean, var = nz.normalize(lay, torch.zeros(0, 5, dtype=DT, device=device), torch.zeros(0, 3, dtype=DT, device=device))
    assert X.shape == (0, 5) and mean.shape == (3,)
