"""GPU parity of the observation-head kernels (hlvae_theta_fwd / _bwd) against the
unmodified reference's HLVAE.theta_estimation outputs (tests/golden/theta_*.npz) and against the oracle on
fresh seeds.  float64 storage: 1e-12 relative; float32 storage: 1e-5 (north_star allows 1e-4)."""
import numpy as np
import pytest
import torch

import helpers as h
from hlvae_b200 import synth, theta as th
from oracle import hlvae_oracle as orc

pytestmark = pytest.mark.gpu
DT = torch.float64


@pytest.mark.parametrize("name", h.THETA_CASES)
@pytest.mark.parametrize("mask_u8", [False, True])
def test_golden_f64(name, mask_u8, device):
    r = h.run_theta_golden(name, device, mask_u8=mask_u8)
    errs = h.assert_theta_close(r, tol=1e-12, label=name)
    print(name, {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.parametrize("name", h.THETA_CASES)
def test_golden_f32_storage(name, device):
    r = h.run_theta_golden(name, device, storage=torch.float32, mask_u8=True)
    assert r["theta"].dtype == torch.float32 and r["d_y"].dtype == torch.float32
    h.assert_theta_close(r, tol=1e-5, label=name)


def _random_heads(types, conv, Y, gen):
    """Per-group head parameters in the reference's module layout (HLVAE.py:11-89), oracle and product copies."""
    ti = orc.types_info_from_layout(types, conv=conv)
    heads = []
    for i, tpl in enumerate(ti['set_of_types']):
        n = int((ti['data_types_indexes'] == i).sum())
        C = int(tpl[1])
        r = lambda *s: torch.randn(*s, generator=gen, dtype=DT) * 0.5
        if tpl[0] == 'count':
            heads.append(dict(weight=r(n, Y, 1), bias=r(n, 1)))
        elif tpl[0] in ('real', 'pos'):
            heads.append(dict(weight_mean=r(n, Y, 1), bias_mean=r(n, 1)))
        elif tpl[0] == 'cat':
            heads.append(dict(weight=r(n, Y, C - 1), bias=r(n, C - 1)))
        else:
            heads.append(dict(weight_thresholds=1.0 + r(n, C - 1), weight_region=r(n, Y, 1), bias_region=r(n, 1)))
    return ti, heads


def _product_layers(ti, heads, conv, device):
    layers = []
    for tpl, hd in zip(ti['set_of_types'], heads):
        layers.append(h._Head({n: v.detach().clone().to(device) for n, v in hd.items()}))
        if tpl[0] == 'real' and conv:
            layers.append(torch.nn.Sigmoid())
    return torch.nn.ModuleList(layers)


@pytest.mark.parametrize("types,conv,N,Y,layout", [
    (synth.TABULAR_TYPES, False, 301, 3, "dense"),          # several tiles, ragged last row batch
    (synth.TABULAR_TYPES, False, 64, 8, "dense"),           # even y_dim (bank-conflicting flat staging)
    (synth.TABULAR_TYPES, False, 37, 11, "dense"),          # y_dim > 8 instantiation
    (synth.TABULAR_TYPES, False, 33, 16, "dense"),          # float64 with y_dim 15 / 16: backward needs > 48 KB of shared memory
    (synth.HEALTHMNIST_D4_TYPES, True, 9, 15, "permuted"),
    (synth.HEALTHMNIST_D4_TYPES, True, 21, 5, "permuted"),  # conv layout [N, Y, D] viewed as [N, D, Y]
    (synth.HEALTHMNIST_D4_TYPES, True, 9, 5, "dense"),
    ([('cat', 16)] * 40 + [('ordinal', 9)] * 7 + [('real', 1)], False, 19, 4, "dense"),   # 16-class variables: 16 per tile
])
def test_random_vs_oracle(types, conv, N, Y, layout, device):
    for per_column in (False, True):
        errs = _random_case(types, conv, N, Y, layout, device, seed=N + Y, per_column=per_column)
        bad = {k: v for k, v in errs.items() if not v < (1e-12 if k in ("theta", "d_y") else 1e-11)}
        assert not bad, (per_column, bad)


def test_long_stripes_float32_parameter_gradients(device):
    """Many rows per CTA stripe: the thread-per-variable backward keeps its weight / bias gradient sums in float32 and
    flushes them to the float64 accumulators every 256 rows."""
    errs = _random_case(synth.TABULAR_TYPES[::8], False, 40000, 5, "dense", device, seed=5, storage=torch.float32)
    bad = {k: v for k, v in errs.items() if not v < 1e-5}
    assert not bad, bad


def _random_case(types, conv, N, Y, layout, device, seed, storage=DT, observed=0.7, scale=1.0, per_column=False):
    """One seeded comparison of theta_heads (forward, d y, d head parameters) with the oracle; returns the relative
    errors.  `storage` float32 = float32 y / theta in HBM with a uint8 mask (the benchmarked arithmetic)."""
    gen = torch.Generator().manual_seed(seed)
    ti, heads = _random_heads(types, conv, Y, gen)
    for hd in heads:
        for v in hd.values():
            v.requires_grad_(True)
    D = len(types)
    y0 = torch.randn(N, D, Y, generator=gen, dtype=DT) * scale
    if storage != DT:
        y0 = y0.to(storage).to(DT)                        # both sides start from the stored values
    mask = (torch.rand(N, D, generator=gen) < observed).to(DT)
    descs, _, P = orc.build_layout(types)
    g_up = torch.randn(N, P, generator=gen, dtype=DT)
    y_o = y0.clone().requires_grad_(True)
    th_o = orc.theta_estimation(types, heads, y_o, mask, conv=conv)
    (th_o * g_up).sum().backward()

    obs_layer = _product_layers(ti, heads, conv, device)
    lay = th.HeadLayout(types, conv, device)
    if per_column:                                        # keep the thread-per-column backward kernel (layouts with <= 5
        lay.max_cols = 0                                  # columns per variable and y_dim <= 8 take the thread-per-variable one)
    yd = y0.to(device).to(storage)
    if layout == "permuted":
        yd = yd.permute(0, 2, 1).contiguous().permute(0, 2, 1)
    yd = yd.clone(memory_format=torch.preserve_format).requires_grad_(True)
    assert (yd.stride(1) == 1) == (layout == "permuted") or D == 1 or Y == 1
    W, b = th.pack_heads(obs_layer, lay, Y)
    md = mask.to(device)
    theta = th.theta_heads(lay, yd, md.to(torch.uint8) if storage != DT else md, W, b)
    assert theta.dtype == storage
    (theta.double() * g_up.to(device)).sum().backward()
    assert yd.grad.stride() == yd.stride()
    errs = {"theta": h.rel_err(theta, th_o), "d_y": h.rel_err(yd.grad, y_o.grad)}
    layer = 0
    for i, (tpl, hd) in enumerate(zip(ti['set_of_types'], heads)):
        for n, prm in obs_layer[layer].named_parameters():
            errs[f"g{i}_{tpl[0]}_{n}"] = h.rel_err(prm.grad, hd[n].grad)
        layer += 2 if (tpl[0] == 'real' and conv) else 1
    return errs


def test_understated_max_cols_poisons_dy(device):
    """hlvae_theta_bwd's max_cols selects the thread-per-variable kernel (<= 5 columns per variable).  A caller that
    understates it gets NaN in d/dy of the wider variables, not a silently truncated gradient."""
    types = [('cat', 7)] * 3 + [('real', 1)] * 4
    gen = torch.Generator().manual_seed(3)
    ti, heads = _random_heads(types, False, 4, gen)
    lay = th.HeadLayout(types, False, device)
    assert lay.max_cols == 7
    W, b = th.pack_heads(_product_layers(ti, heads, False, device), lay, 4)
    y = torch.randn(9, len(types), 4, generator=gen, dtype=DT).to(device).requires_grad_(True)
    mask = torch.ones(9, len(types), dtype=DT, device=device)
    lay.max_cols = 5                                         # the caller's error
    th.theta_heads(lay, y, mask, W, b).sum().backward()
    assert bool(torch.isnan(y.grad[:, :3]).all()) and bool(torch.isfinite(y.grad[:, 3:]).all())
    # a whole tile wider than its staging buffer (128 variables x 7 columns against 128 x 5): poisoned, no overrun
    types = [('cat', 7)] * 130
    ti, heads = _random_heads(types, False, 4, gen)
    lay = th.HeadLayout(types, False, device)
    W, b = th.pack_heads(_product_layers(ti, heads, False, device), lay, 4)
    y = torch.randn(9, len(types), 4, generator=gen, dtype=DT).to(device).requires_grad_(True)
    lay.max_cols = 5
    th.theta_heads(lay, y, torch.ones(9, len(types), dtype=DT, device=device), W, b).sum().backward()
    assert bool(torch.isnan(y.grad).all())
