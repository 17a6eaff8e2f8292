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
    # decisions are made in float64 before rounding the outputs to float32
    assert np.array_equal(out["recon_mean"].cpu().numpy()[:, disc], omean.numpy()[:, disc].astype(np.float32))
    assert h.rel_err(out["log_p_x"], olpx) < 1e-6
    assert out["log_p_x"].dtype == torch.float32


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
