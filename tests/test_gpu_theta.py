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
    (synth.HEALTHMNIST_D4_TYPES, True, 21, 5, "permuted"),  # conv layout [N, Y, D] viewed as [N, D, Y]
    (synth.HEALTHMNIST_D4_TYPES, True, 9, 5, "dense"),
    ([('cat', 16)] * 40 + [('ordinal', 9)] * 7 + [('real', 1)], False, 19, 4, "dense"),   # 16-class variables: 16 per tile
])
def test_random_vs_oracle(types, conv, N, Y, layout, device):
    gen = torch.Generator().manual_seed(N + Y)
    ti, heads = _random_heads(types, conv, Y, gen)
    for hd in heads:
        for v in hd.values():
            v.requires_grad_(True)
    D = len(types)
    y0 = torch.randn(N, D, Y, generator=gen, dtype=DT)
    mask = (torch.rand(N, D, generator=gen) < 0.7).to(DT)
    descs, _, P = orc.build_layout(types)
    g_up = torch.randn(N, P, generator=gen, dtype=DT)
    y_o = y0.clone().requires_grad_(True)
    th_o = orc.theta_estimation(types, heads, y_o, mask, conv=conv)
    (th_o * g_up).sum().backward()

    obs_layer = _product_layers(ti, heads, conv, device)
    lay = th.HeadLayout(types, conv, device)
    yd = y0.to(device)
    if layout == "permuted":
        yd = yd.permute(0, 2, 1).contiguous().permute(0, 2, 1)
    yd = yd.clone(memory_format=torch.preserve_format).requires_grad_(True)
    assert (yd.stride(1) == 1) == (layout == "permuted")
    W, b = th.pack_heads(obs_layer, lay, Y)
    theta = th.theta_heads(lay, yd, mask.to(device), W, b)
    (theta * g_up.to(device)).sum().backward()
    assert h.rel_err(theta, th_o) < 1e-12
    assert h.rel_err(yd.grad, y_o.grad) < 1e-12
    assert yd.grad.stride() == yd.stride()
    layer = 0
    for tpl, hd in zip(ti['set_of_types'], heads):
        for n, prm in obs_layer[layer].named_parameters():
            assert h.rel_err(prm.grad, hd[n].grad) < 1e-11, (tpl, n)
        layer += 2 if (tpl[0] == 'real' and conv) else 1


def test_dropin_method_signature(device):
    """theta_estimation(self, y, miss_list, param_miss_list) bound to a model-like object (HLVAE.py:416)."""
    g = h.load("theta_tabular_small")
    types = h.parse_types(g)
    obs_layer, _, _ = h.golden_heads(g, device, False)

    class M:
        pass
    m = M()
    m.types_info = orc.types_info_from_layout(types)
    m.conv = False
    m.obs_layer = obs_layer
    y = h.t(g["y"], device)
    mask = h.t(g["mask"], device)
    out = th.theta_estimation(m, y, mask, None)
    assert h.rel_err(out, g["theta"]) < 1e-12
    assert out.dtype == torch.float64 and out.shape == g["theta"].shape


def test_properties_full_size(device):
    """configs[1] batch (16 000 rows, D4 conv layout, float32): forward is independent of the mask; d/dy is zero
    exactly where the mask is zero; theta of constant columns is exact; against plain torch ops on the packed form."""
    types = synth.HEALTHMNIST_D4_TYPES
    N, Y, D = 16000, 5, 1296
    gen = torch.Generator(device=device).manual_seed(3)
    lay = th.HeadLayout(types, True, device)
    y = torch.randn(N, Y, D, generator=gen, device=device, dtype=torch.float32).permute(0, 2, 1).requires_grad_(True)
    mask = (torch.rand(N, D, generator=gen, device=device) < 0.75).to(torch.uint8)
    W = (torch.randn(lay.P, Y, generator=gen, device=device, dtype=DT) * 0.3).requires_grad_(True)
    b = (torch.randn(lay.P, generator=gen, device=device, dtype=DT) * 0.3).requires_grad_(True)
    theta = th.theta_heads(lay, y, mask, W, b)
    theta2 = th.theta_heads(lay, y.detach(), torch.ones_like(mask), W, b)
    assert torch.equal(theta, theta2)
    mode = lay.col_mode
    assert torch.equal(theta[:, mode == 2], torch.zeros_like(theta[:, mode == 2]))
    g_up = torch.randn(N, lay.P, generator=gen, device=device, dtype=torch.float32)
    (theta * g_up).sum().backward()
    assert torch.equal(y.grad[mask == 0], torch.zeros_like(y.grad[mask == 0]))
    # plain torch float64 evaluation of the same op on a row sample
    rows = torch.arange(0, N, 97, device=device)
    ys = y.detach()[rows].double().requires_grad_(True)
    Wt, bt = W.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    z = bt[None, :] + torch.einsum("npk,pk->np", ys[:, lay.col_var.long(), :], Wt)
    z = torch.where(mode == 1, torch.sigmoid(z), z)
    z = torch.where(mode == 2, torch.zeros_like(z), z)
    assert h.rel_err(theta[rows], z) < 1e-5
    pm = mask[rows][:, lay.col_var.long()].double()
    (z * g_up[rows].double() * pm).sum().backward()
    assert h.rel_err(y.grad[rows], ys.grad) < 1e-5


def test_no_cpu_fallback():
    types = [('real', 1), ('cat', 3)]
    lay = th.HeadLayout(types, False, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        th.theta_heads(lay, torch.zeros(2, 2, 3, dtype=DT), torch.ones(2, 2, dtype=DT), torch.zeros(4, 3, dtype=DT),
                       torch.zeros(4, dtype=DT))


def test_edge_cases(device):
    """Empty batch, single row, an all-missing row (no gradient leaves it), a y view that is not dense, y_dim too large."""
    types = [('real', 1), ('cat', 4), ('ordinal', 3), ('count', 1), ('pos', 1)]
    gen = torch.Generator().manual_seed(0)
    ti, heads = _random_heads(types, False, 3, gen)
    obs_layer = _product_layers(ti, heads, False, device)
    lay = th.HeadLayout(types, False, device)
    W, b = th.pack_heads(obs_layer, lay, 3)
    empty = th.theta_heads(lay, torch.zeros(0, 5, 3, dtype=DT, device=device), torch.zeros(0, 5, dtype=DT, device=device), W, b)
    assert empty.shape == (0, lay.P)
    y1 = torch.randn(1, 5, 3, generator=gen, dtype=DT).to(device).requires_grad_(True)
    m1 = torch.zeros(1, 5, dtype=DT, device=device)
    t1 = th.theta_heads(lay, y1, m1, W, b)
    t1.sum().backward()
    ref = orc.theta_estimation(types, heads, y1.detach().cpu(), m1.cpu())
    assert h.rel_err(t1, ref) < 1e-12                         # forward does not depend on the mask
    assert torch.equal(y1.grad, torch.zeros_like(y1))         # nothing observed: no gradient
    big = torch.randn(6, 5, 8, generator=gen, dtype=DT).to(device)
    yv = big[:, :, ::2][:, :, :3]                             # gapped strides: made contiguous by the host layer
    tv = th.theta_heads(lay, yv, torch.ones(6, 5, dtype=DT, device=device), W, b)
    assert h.rel_err(tv, orc.theta_estimation(types, heads, yv.cpu().contiguous(), torch.ones(6, 5, dtype=DT))) < 1e-12
    lay17 = th.HeadLayout([('real', 1)], False, device)
    with pytest.raises(NotImplementedError):
        th.theta_heads(lay17, torch.zeros(2, 1, 17, dtype=DT, device=device), torch.ones(2, 1, dtype=DT, device=device),
                       torch.zeros(1, 17, dtype=DT, device=device), torch.zeros(1, dtype=DT, device=device))


def test_variance_network_heads_golden(device):
    """HLVAE.theta_estimation with logvar_network=True (HLVAE.py:26-57,416-453): every real / positive variable owns
    a mean and a raw log-variance column, laid out group-wise as [means..., log-variances...], against the
    unmodified reference's theta, d/dy and head-parameter gradients."""
    g = h.load("theta_logvar_mixed")
    types = h.parse_types(g)
    obs_layer, _, kinds = h.golden_heads(g, device, False)
    ti = orc.types_info_from_layout(types, conv=False, logvar_network=True)

    class Model:
        pass

    model = Model()
    model.obs_layer, model.types_info, model.conv, model.logvar_network = obs_layer, ti, False, True
    y = h.t(g["y"], device).requires_grad_(True)
    theta = th.theta_estimation(model, y, h.t(g["mask"], device), None)
    assert theta.shape == g["theta"].shape
    (theta * h.t(g["g_up"], device)).sum().backward()
    assert h.rel_err(theta, g["theta"]) < 1e-12 and h.rel_err(y.grad, g["d_y"]) < 1e-12
    for i, kind in enumerate(kinds):
        for n, prm in obs_layer[i].named_parameters():
            assert h.rel_err(prm.grad, g[f"d_g{i}_{n}"]) < 1e-11, (i, n)
