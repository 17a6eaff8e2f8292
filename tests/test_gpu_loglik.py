"""GPU parity of the fused masked log-likelihood (hlvae_loglik_fwd/bwd, hlvae_statistics,
hlvae_discrete_transform).  Float64 storage: 1e-9 relative against the reference's outputs;
categorical / ordinal argmax imputations and the discrete transform bit-exact."""
import numpy as np
import pytest
import torch

import helpers as h
from hlvae_b200 import loglik, synth
from oracle import hlvae_oracle as orc

pytestmark = pytest.mark.gpu
DT = torch.float64


@pytest.mark.parametrize("name", h.LOGLIK_CASES)
def test_golden(name, device):
    r = h.run_loglik_golden(name, device)
    errs = h.assert_loglik_close(r, tol=1e-9, label=name)
    print(name, {k: f"{v:.1e}" for k, v in errs.items()})


def test_per_type_functions_match_golden(device):
    """The HL_VAE/loglik.py-style entry points, called the way HLVAE.py:387-412 calls them."""
    g = h.load("loglik_mixed")
    types = h.parse_types(g)
    ti = orc.types_info_from_layout(types)
    data, mask, theta = h.t(g["data"], device), h.t(g["mask"], device), h.t(g["theta"], device)
    nr, npos = h.golden_norm(g, device)
    lvr, lvp = h.t(g["log_vy_real"], device), h.t(g["log_vy_pos"], device)
    lpx = torch.zeros_like(mask)
    for i, tpl in enumerate(ti['set_of_types']):
        fn = getattr(loglik, 'loglik_' + tpl[0])
        dsel = torch.tensor(ti['exp_types_indexes'] == i, device=device)
        vsel = torch.tensor(ti['data_types_indexes'] == i, device=device)
        psel = torch.tensor(ti['param_indexes'] == i, device=device)
        norm = list(nr) if tpl[0] == 'real' else list(npos) if tpl[0] == 'pos' else torch.tensor(0.)
        extra = lvr if tpl[0] == 'real' else lvp if tpl[0] == 'pos' else None
        out = fn([data[:, dsel], mask[:, vsel]], tpl, theta[:, psel], norm, extra)
        assert set(out) == {'log_p_x', 'log_p_x_missing', 'params', 'samples'}
        assert out['samples'] is not None
        lpx[:, vsel] = out['log_p_x']
        assert h.rel_err(out['params'].reshape(data.shape[0], -1), g["params"][:, ti['param_indexes'] == i]) < 1e-9
    assert h.rel_err(lpx, g["log_p_x"]) < 1e-9


def test_float32_storage_argmax_exact(device):
    """float32-stored inputs: arithmetic is float64 on the rounded inputs, so against the oracle fed
    the same rounded inputs the argmax maps stay bit-exact and values agree to float32 rounding."""
    rng = np.random.default_rng(5)
    types = synth.TABULAR_TYPES
    N = 2048
    data, mask = synth.likelihood_batch(types, N, rng)
    gen = torch.Generator().manual_seed(5)
    theta = (torch.randn(N, 768, generator=gen, dtype=DT) * 1.5).float()
    descs, E_x, P_th = orc.build_layout(types)
    nr, npos = orc.batch_norm_params(descs, data.float().double(), mask)
    lvr = torch.randn(32, generator=gen, dtype=DT) * 0.3
    lvp = torch.randn(32, generator=gen, dtype=DT) * 0.3
    olpx, olpm, oparams = orc.loglik_and_reconstruction(descs, data.float().double(), mask, theta.double(), lvr, lvp, nr, npos)
    omean, omode = orc.statistics(descs, oparams, lvp)
    lay = loglik.VarLayout(types, device)
    vparam = lay.vparam(lvr.to(device), lvp.to(device), [a.to(device) for a in nr], [a.to(device) for a in npos])
    out = loglik.fused_loglik(lay, data.float().to(device), mask.to(torch.uint8).to(device), theta.to(device), vparam)
    disc = np.array([k in ("cat", "ordinal") for k, _ in types])
    # float32 arithmetic, but every argmax decision float32 cannot prove is redone in float64
    assert np.array_equal(out["recon_mean"].cpu().numpy()[:, disc], omean.numpy()[:, disc].astype(np.float32))
    assert h.rel_err(out["log_p_x"], olpx) < 5e-6          # north_star: 1e-4 relative in fp32
    assert h.rel_err(out["params"], oparams) < 5e-6
    assert out["log_p_x"].dtype == torch.float32


def _fp32_case(types, N, seed, device, conv=False, observed=0.7, u8=False, theta_scale=1.5, storage=torch.float32):
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    data, mask = synth.likelihood_batch(types, N, rng, observed=observed, pixel_like=conv)
    data = data.to(storage).double()                   # the oracle sees the inputs as stored (float32-rounded or float64)
    descs, E_x, P_th = orc.build_layout(types)
    theta = (torch.randn(N, P_th, generator=gen, dtype=DT) * theta_scale).to(storage)
    n_real = sum(k == "real" for k, _ in types)
    n_pos = sum(k == "pos" for k, _ in types)
    lvr = torch.randn(n_real, generator=gen, dtype=DT) * 0.3
    lvp = torch.randn(n_pos, generator=gen, dtype=DT) * 0.3
    if conv:
        nr, npos = None, None
    else:
        nr, npos = orc.batch_norm_params(descs, data, mask)
    g_up = torch.randn(N, len(types), generator=gen, dtype=DT).to(storage)
    th64 = theta.double().clone().requires_grad_(True)
    lvr64, lvp64 = lvr.clone().requires_grad_(True), lvp.clone().requires_grad_(True)
    olpx, olpm, oparams = orc.loglik_and_reconstruction(descs, data, mask, th64, lvr64, lvp64, nr, npos, conv=conv)
    (olpx * g_up.double()).sum().backward()
    omean, omode = orc.statistics(descs, oparams.detach(), lvp)
    odtr = orc.discrete_variables_transformation(descs, data)
    lay = loglik.VarLayout(types, device)
    lvr_d, lvp_d = lvr.to(device).requires_grad_(True), lvp.to(device).requires_grad_(True)
    vparam = lay.vparam(lvr_d if n_real else None, lvp_d if n_pos else None,
                        None if nr is None else [a.to(device) for a in nr],
                        None if npos is None else [a.to(device) for a in npos], conv=conv)
    th = theta.detach().to(device).requires_grad_(True)
    d_in = data.to(torch.uint8) if u8 else data.to(storage)
    m_in = mask.to(torch.uint8) if u8 else mask.to(storage)
    out = loglik.fused_loglik(lay, d_in.to(device), m_in.to(device), th, vparam)
    (out["log_p_x"] * g_up.to(device)).sum().backward()
    disc = np.array([k in ("cat", "ordinal") for k, _ in types])
    ref = dict(log_p_x=olpx.detach(), log_p_x_missing=olpm.detach(), params=oparams.detach(), d_theta=th64.grad,
               recon_mean=omean, recon_mode=omode, data_transformed=odtr,
               d_lvr=lvr64.grad if n_real else None, d_lvp=lvp64.grad if n_pos else None)
    got = dict(out, d_theta=th.grad, d_lvr=lvr_d.grad if n_real else None, d_lvp=lvp_d.grad if n_pos else None)
    return got, ref, disc


@pytest.mark.parametrize("u8", [False, True])
def test_float32_fast_path_healthmnist_d4(u8, device):
    """BASELINE.json configs[1] variable layout (324 real + 972 cat x 5, conv scaling) with float32
    storage -> float32 SFU arithmetic; uint8 data / mask are exact for pixel values and one-hot codes."""
    got, ref, disc = _fp32_case(synth.HEALTHMNIST_D4_TYPES, 512, 31, device, conv=True, observed=0.75, u8=u8)
    for key in ("log_p_x", "log_p_x_missing", "params", "d_theta", "recon_mode"):
        assert h.rel_err(got[key], ref[key]) < 1e-5, key
    assert h.rel_err(got["d_lvr"], ref["d_lvr"]) < 1e-4
    assert np.array_equal(got["recon_mean"].cpu().numpy()[:, disc], ref["recon_mean"].numpy()[:, disc].astype(np.float32))
    assert np.array_equal(got["data_transformed"].cpu().numpy(), ref["data_transformed"].numpy().astype(np.float32))
    assert h.rel_err(got["log_p_x_sum"], ref["log_p_x"].sum()) < 1e-6


def test_float32_fast_path_tabular_all_types(device):
    """configs[3] layout: count / ordinal / cat / real / pos, 30 % missing, N = 16000: ~1e6 ordinal
    variables make float32 near-ties certain, so the exact float64 re-evaluation is exercised."""
    got, ref, disc = _fp32_case(synth.TABULAR_TYPES, 16000, 33, device)
    for key in ("log_p_x", "log_p_x_missing", "params", "d_theta"):
        assert h.rel_err(got[key], ref[key]) < 1e-5, key
    # Poisson mode = floor(lambda) may differ where lambda is within float32 rounding of an integer
    cnt = np.array([k == "count" for k, _ in synth.TABULAR_TYPES])
    gm, rm = got["recon_mode"].cpu().numpy().astype(np.float64), ref["recon_mode"].numpy()
    assert np.abs(gm[:, ~cnt] - rm[:, ~cnt]).max() < 1e-5 * np.abs(rm[:, ~cnt]).max()
    assert (gm[:, cnt] != rm[:, cnt]).mean() < 1e-4
    assert h.rel_err(got["d_lvr"], ref["d_lvr"]) < 1e-4 and h.rel_err(got["d_lvp"], ref["d_lvp"]) < 1e-4
    assert np.array_equal(got["recon_mean"].cpu().numpy()[:, disc], ref["recon_mean"].numpy()[:, disc].astype(np.float32))
    assert np.array_equal(got["data_transformed"].cpu().numpy(), ref["data_transformed"].numpy().astype(np.float32))


def test_float64_storage_wide_classes(device):
    """The drop-in case (every tensor float64) with 16-class variables in a full tile: 128 variables x 16 classes x 8 rows
    of float64 theta and data do not fit one CTA's shared memory - the tile must shrink, not the launch fail."""
    types = [('cat', 16)] * 70 + [('ordinal', 16)] * 50 + [('real', 1)] * 5 + [('count', 1)] * 5 + [('cat', 3)] * 10
    got, ref, disc = _fp32_case(types, 41, 77, device, storage=torch.float64)
    for k in ("log_p_x", "log_p_x_missing", "params", "d_theta"):
        assert got[k].dtype == torch.float64 and h.rel_err(got[k], ref[k]) < 1e-9, k
    assert np.array_equal(got["recon_mean"].cpu().numpy()[:, disc], ref["recon_mean"].numpy()[:, disc])


@pytest.mark.parametrize("C", [5, 11])
def test_float32_ordinal_gradient_at_large_logits(C, device):
    """Ordinal variables with logits ~ N(0, 16) (small class probabilities): the float32 backward takes the pair of
    1 / p_y terms in closed form, (sigma'(u_y) - sigma'(u_{y-1})) / p_y = sigma(-u_{y-1}) - sigma(u_y); the general
    form lost 2e-3 .. 4e-3 of the largest gradient to cancellation here."""
    got, ref, disc = _fp32_case([("ordinal", C)] * 40, 500, 77, device, observed=0.7, theta_scale=4.0)
    for key in ("log_p_x", "params", "d_theta"):
        assert h.rel_err(got[key], ref[key]) < 5e-6, key


def test_float32_argmax_adversarial_ties(device):
    """Logits that are distinct in float32 but collapse onto one double after `theta - lse`
    (reference: first index wins), exact ties, and ordinal thresholds giving equal class masses."""
    types = [("cat", 5)] * 6 + [("ordinal", 4)] * 2
    descs, E_x, P_th = orc.build_layout(types)
    th = torch.zeros(3, P_th, dtype=torch.float32)
    th[0, 0:5] = torch.tensor([0.0, 1e-20, -3.0, 1e-20, -1.0])          # collapses: reference picks 0
    th[0, 5:10] = torch.tensor([0.0, 2.0, 2.0, -1.0, 2.0])               # exact tie -> first of the ties
    th[0, 10:15] = torch.tensor([0.0, -1e-30, 1e-30, 0.0, 0.0])
    th[0, 15:20] = torch.tensor([0.0, 1.0000001, 1.0, 0.5, 1.0000001])
    th[1] = torch.randn(P_th) * 1e-8                                      # all logits nearly equal
    th[2] = torch.randn(P_th) * 3
    th[0, 30:34] = torch.tensor([0.0, 0.0, 0.0, 0.0])
    data = torch.zeros(3, E_x, dtype=DT)
    data[:, 0::5][:, :6] = 1.0
    data[:, 30] = 1.0
    data[:, 34] = 1.0
    mask = torch.ones(3, len(types), dtype=DT)
    _, _, oparams = orc.loglik_and_reconstruction(descs, data, mask, th.double())
    omean, _ = orc.statistics(descs, oparams)
    lay = loglik.VarLayout(types, device)
    out = loglik.fused_loglik(lay, data.float().to(device), mask.float().to(device), th.to(device), lay.vparam())
    assert np.array_equal(out["recon_mean"].cpu().numpy(), omean.numpy().astype(np.float32))
    assert omean[0, 0].item() == 0.0 and omean[0, 1].item() == 1.0


def test_sum_output_backward_matches_elementwise(device):
    """log_p_x_sum is accumulated in the kernel; its backward (device scalar, no [N, D] gradient)
    must equal the backward of log_p_x.sum(), also when both outputs are used."""
    rng = np.random.default_rng(2)
    types = synth.TABULAR_TYPES
    N = 300
    data, mask = synth.likelihood_batch(types, N, rng)
    descs, _, P_th = orc.build_layout(types)
    nr, npos = orc.batch_norm_params(descs, data, mask)
    lay = loglik.VarLayout(types, device)
    z32 = torch.zeros(32, dtype=DT, device=device)
    for storage in (torch.float64, torch.float32):
        grads = []
        for mode in ("elementwise", "sum", "both"):
            lv = z32.clone().requires_grad_(True)
            vparam = lay.vparam(lv, z32, [a.to(device) for a in nr], [a.to(device) for a in npos])
            th = torch.randn(N, P_th, dtype=DT, generator=torch.Generator().manual_seed(1)).to(storage).to(device)
            th.requires_grad_(True)
            out = loglik.fused_loglik(lay, data.to(storage).to(device), mask.to(torch.uint8).to(device), th, vparam)
            loss = {"elementwise": lambda: -2.5 * out["log_p_x"].sum(dtype=DT), "sum": lambda: -2.5 * out["log_p_x_sum"],
                    "both": lambda: -1.5 * out["log_p_x"].sum(dtype=DT) - out["log_p_x_sum"]}[mode]()
            loss.backward()
            grads.append((th.grad.double(), lv.grad, loss.detach()))
        tol = 1e-12 if storage == torch.float64 else 1e-5
        for g in grads[1:]:
            assert h.rel_err(g[0], grads[0][0]) < tol and h.rel_err(g[1], grads[0][1]) < tol
            assert h.rel_err(g[2], grads[0][2]) < (1e-12 if storage == torch.float64 else 1e-6)


def test_edge_cases(device):
    lay = loglik.VarLayout([("cat", 3), ("ordinal", 4), ("count", 1)], device)
    vparam = lay.vparam()
    # N = 0
    z = lambda *s: torch.zeros(*s, dtype=DT, device=device)
    out = loglik.fused_loglik(lay, z(0, 8), z(0, 3), z(0, 8), vparam)
    assert out["log_p_x"].shape == (0, 3)
    # all-missing rows give log_p_x = 0 and put everything in log_p_x_missing; missing ordinal -> class 0
    data = torch.tensor([[0., 1, 0, 1, 1, 0, 0, 4.]], dtype=DT, device=device)
    th = torch.tensor([[0.3, 0.3, 0.3, 0.5, -0.2, 0.1, 0.7, 1.2]], dtype=DT, device=device)
    out = loglik.fused_loglik(lay, data, z(1, 3), th, vparam)
    assert float(out["log_p_x"].abs().max()) == 0.0
    descs, _, _ = orc.build_layout([("cat", 3), ("ordinal", 4), ("count", 1)])
    _, olpm, oparams = orc.loglik_and_reconstruction(descs, data.cpu(), torch.zeros(1, 3, dtype=DT), th.cpu())
    assert h.rel_err(out["log_p_x_missing"], olpm) < 1e-12
    assert out["recon_mean"][0, 0].item() == 0.0          # exact tie -> first index


def test_full_size_properties(device):
    """BASELINE.json config-4 shape (D=256 tabular, 30 % missing) at N=16000: size-independent checks -
    categorical params are log-probabilities (logsumexp = 0), ordinal params sum to 1,
    log_p_x + log_p_x_missing does not depend on the mask, and gradients vanish on missing entries."""
    rng = np.random.default_rng(9)
    types = synth.TABULAR_TYPES
    N = 16000
    data, mask = synth.likelihood_batch(types, N, rng)
    gen = torch.Generator().manual_seed(9)
    theta = torch.randn(N, 768, generator=gen, dtype=DT)
    descs, _, _ = orc.build_layout(types)
    nr, npos = orc.batch_norm_params(descs, data, mask)
    lay = loglik.VarLayout(types, device)
    vparam = lay.vparam(torch.zeros(32, dtype=DT, device=device), torch.zeros(32, dtype=DT, device=device),
                        [a.to(device) for a in nr], [a.to(device) for a in npos])
    th = theta.to(device).requires_grad_(True)
    out = loglik.fused_loglik(lay, data.to(device), mask.to(device), th, vparam)
    out2 = loglik.fused_loglik(lay, data.to(device), torch.ones_like(mask).to(device), th, vparam)
    tot = out["log_p_x"] + out["log_p_x_missing"]
    obs = mask.to(device) == 1
    # ordinal likelihood of a missing entry is evaluated at class 0 (loglik.py:173), so compare observed entries
    assert h.rel_err(tot[obs], out2["log_p_x"][obs]) < 1e-12
    P = out["params"]
    cat_cols = slice(64 + 320, 64 + 320 + 320)
    assert float(torch.logsumexp(P[:, cat_cols].reshape(N, 64, 5), 2).abs().max()) < 1e-12
    assert float((P[:, 64:64 + 320].reshape(N, 64, 5).sum(2) - 1).abs().max()) < 1e-12
    out["log_p_x"].sum().backward()
    gmask = torch.repeat_interleave(mask.to(device), lay.var_nclass.long(), dim=1)
    assert float((th.grad * (1 - gmask)).abs().max()) == 0.0
    # oracle on a slice
    sl = slice(0, 256)
    olpx, _, _ = orc.loglik_and_reconstruction(descs, data[sl], mask[sl], theta[sl], torch.zeros(32, dtype=DT),
                                               torch.zeros(32, dtype=DT), nr, npos)
    assert h.rel_err(out["log_p_x"][sl], olpx) < 1e-10


def test_standalone_monitoring_kernels(device):
    g = h.load("loglik_mixed")
    types = h.parse_types(g)
    lay = loglik.VarLayout(types, device)
    vparam = torch.zeros(4, lay.D, dtype=DT, device=device)
    vparam[2, lay.idx["pos"]] = h.t(g["log_vy_pos"], device)[lay.gpos["pos"]]
    mean, mode = loglik.statistics(lay, h.t(g["params"], device), vparam)
    assert h.rel_err(mean, g["recon_mean"]) < 1e-12 and h.rel_err(mode, g["recon_mode"]) < 1e-12
    dtr = loglik.discrete_variables_transformation(lay, h.t(g["data"], device))
    assert np.array_equal(dtr.cpu().numpy(), g["data_transformed"])


def test_staging_paths_agree(device):
    """The kernels stage rows with TMA bulk copies when every array starts and ends on a 16-byte boundary and
    fall back to element-wise cp.async otherwise.  Same values through 16-byte aligned buffers and through
    views that start 4 (float32) / 1 (uint8) bytes into a larger allocation must give identical bits."""
    rng = np.random.default_rng(41)
    types = synth.TABULAR_TYPES
    N = 1000
    data, mask = synth.likelihood_batch(types, N, rng)
    lay = loglik.VarLayout(types, device)
    descs, E_x, P_th = orc.build_layout(types)
    nr, npos = orc.batch_norm_params(descs, data, mask)
    z32 = torch.zeros(32, dtype=DT, device=device)
    theta = torch.randn(N, P_th, generator=torch.Generator().manual_seed(41)).to(device)

    def shifted(t):                       # same values, base pointer off the 16-byte grid
        buf = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
        v = buf[1:].view(t.shape)
        v.copy_(t)
        assert v.data_ptr() % 16 != 0 and v.is_contiguous()
        return v

    outs = []
    for shift in (False, True):
        f = shifted if shift else (lambda t: t.clone())
        th = f(theta).requires_grad_(True)
        lvr, lvp = z32.clone().requires_grad_(True), z32.clone().requires_grad_(True)
        vparam = lay.vparam(lvr, lvp, [a.to(device) for a in nr], [a.to(device) for a in npos])
        out = loglik.fused_loglik(lay, f(data.float().to(device)), f(mask.to(torch.uint8).to(device)), th, vparam)
        (out["log_p_x_sum"] * -1.5).backward()
        outs.append((out, th.grad.clone(), lvr.grad.clone()))
    for key in ("log_p_x", "log_p_x_missing", "params", "recon_mean", "recon_mode", "data_transformed"):
        assert torch.equal(outs[0][0][key], outs[1][0][key]), key
    assert torch.equal(outs[0][1], outs[1][1])
    assert h.rel_err(outs[0][2], outs[1][2]) < 1e-12 and h.rel_err(outs[0][0]["log_p_x_sum"], outs[1][0]["log_p_x_sum"]) < 1e-12


@pytest.mark.parametrize("storage,tol", [(torch.float64, 1e-11), (torch.float32, 2e-5)])
def test_variance_network_and_beta_branches_golden(storage, tol, device):
    """loglik_real / loglik_pos with extra_params = None (logvar_network=True, HL_VAE/loglik.py:45-48,104-108) and
    loglik_beta (:216-256, incl. the reference's fallback indexing for a narrow theta) through the reference-shaped
    functions (hlvae_loglik_aux_fwd / _bwd) against the unmodified reference's outputs and gradients."""
    from hlvae_b200 import loglik as ll
    g = h.load("loglik_aux")
    mask, g_up = h.t(g["mask"], device), h.t(g["g_up"], device)
    nm, nv = h.t(g["real_norm_nm"], device), h.t(g["real_norm_nv"], device)
    cases = [("real_norm", ll.loglik_real, ('real', '1'), [nm, nv], False), ("real_plain", ll.loglik_real, ('real', '1'), [], False),
             ("pos_norm", ll.loglik_pos, ('pos', '1'), [nm, nv], False)]
    for tag in ("beta_wide", "beta_fallback2", "beta_fallback3"):
        cases.append((tag, ll.loglik_beta, ('beta', '1'), g[tag + "_ranges"], True))
    old, ll.PRODUCE_SAMPLES = ll.PRODUCE_SAMPLES, True
    try:
        for tag, fn, tpl, norm, beta in cases:
            data = h.t(g[tag + "_data"], device).to(storage)
            D = data.shape[1]
            th = h.t(g[tag + "_theta"], device).to(storage).requires_grad_(True)
            disp = h.t(g[tag + "_disp"], device).requires_grad_(True) if beta else None
            out = fn([data, mask[:, :D]], tpl, th, norm, disp)
            (out["log_p_x"].double() * g_up[:, :D]).sum().backward()
            got = dict(log_p_x=out["log_p_x"], log_p_x_missing=out["log_p_x_missing"], prm_a=out["params"][0],
                       prm_b=out["params"][1], d_theta=th.grad)
            for key, val in got.items():
                assert val.shape == g[f"{tag}_{key}"].shape, (tag, key, val.shape)
                assert h.rel_err(val, g[f"{tag}_{key}"]) < tol, (tag, key, h.rel_err(val, g[f"{tag}_{key}"]))
            if beta:
                assert h.rel_err(disp.grad, g[tag + "_d_disp"]) < tol, tag
            assert out["samples"].shape[0] == data.shape[0]
    finally:
        ll.PRODUCE_SAMPLES = old


def test_variance_network_model_level_golden(device):
    """HLVAE.loglik_and_reconstruction (HLVAE.py:381-414) + p_params_concatenation_by_key + statistics
    (read_functions.py:206-218,268-302) for a model with logvar_network=True against the unmodified reference:
    log_p_x, log_p_x_missing, the concatenated parameters, d/dtheta and the imputation statistics (cat / ordinal
    argmax bit-exact)."""
    from hlvae_b200 import loglik as ll
    g = h.load("loglik_logvar_mixed")
    types = h.parse_types(g)
    ti = orc.types_info_from_layout(types, conv=False, logvar_network=True)

    class Model:
        pass

    m = Model()
    m.types_info, m.conv, m.logvar_network, m._log_vy_real, m._log_vy_pos = ti, False, True, None, None
    nr = [h.t(g["norm_real_mean"], device), h.t(g["norm_real_var"], device)] if g["norm_real_mean"].size else []
    npos = [h.t(g["norm_pos_mean"], device), h.t(g["norm_pos_var"], device)] if g["norm_pos_mean"].size else []
    theta = h.t(g["theta"], device).requires_grad_(True)
    old, ll.PRODUCE_SAMPLES = ll.PRODUCE_SAMPLES, False
    try:
        lpx, lpm, _, params = ll.loglik_and_reconstruction(m, theta, h.t(g["data"], device), h.t(g["mask"], device),
                                                           None, [nr, npos])
    finally:
        ll.PRODUCE_SAMPLES = old
    (lpx * h.t(g["g_up"], device)).sum().backward()
    assert h.rel_err(lpx, g["log_p_x"]) < 1e-11 and h.rel_err(lpm, g["log_p_x_missing"]) < 1e-11
    assert h.rel_err(theta.grad, g["d_theta"]) < 1e-10
    # p_params_concatenation_by_key (read_functions.py:206-218): lists are concatenated, tensors flattened
    pcat = torch.zeros(theta.shape[0], len(ti['param_indexes']), dtype=DT, device=device)
    for i, prm in enumerate(params):
        flat = torch.cat(prm, 1) if isinstance(prm, list) else prm.reshape(prm.shape[0], -1)
        pcat[:, torch.as_tensor(np.nonzero(ti['param_indexes'] == i)[0], device=device)] = flat.detach().double()
    assert h.rel_err(pcat, g["params"]) < 1e-11
    mean, mode = ll.statistics_general(pcat, ti, False, [None, None])
    disc = np.array([k in ("cat", "ordinal") for k, _ in types])
    assert np.array_equal(mean.cpu().numpy()[:, disc], g["recon_mean"][:, disc])
    assert h.rel_err(mean, g["recon_mean"]) < 1e-11 and h.rel_err(mode, g["recon_mode"]) < 1e-11
