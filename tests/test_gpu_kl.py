"""GPU parity of the KL upper bound (hlvae_kl_subject + hlvae_kl_panel + M x M stage) through the
reference-facing functions minibatch_KLD_upper_bound[_iter].

Tolerances: north_star asks for 1e-4 relative on every ELBO term and gradient.  The CUDA path
computes in float64, so most terms are checked far tighter (1e-6); kernel hyper-parameter
gradients get 1e-4 because the float64 reference itself only agrees with an equivalent float64
evaluation order to ~3e-5 there (cond(K0zz + eps I) ~ 1e7, see oracle/make_goldens.py)."""
import numpy as np
import pytest
import torch

import helpers as h
from hlvae_b200 import config, elbo, subjects

pytestmark = pytest.mark.gpu
DT = torch.float64


@pytest.mark.parametrize("name", h.KL_CASES)
def test_golden(name, device):
    r = h.run_kl_golden(name, device)
    errs = h.assert_kl_close(r, tol=5e-6, hyper_tol=1e-4, label=name)
    print(name, {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.parametrize("name", h.KL_CASES)
def test_every_elbo_term_against_golden(name, device):
    """BASELINE.md section 4 gate: A, B + D, C, E, F, kld_qu_pu (elbo_functions.py:166-181 / :256-277) and the
    statistics S, p one by one - not only their sum kld_total - from what the kernels emit (scal, pre, S)."""
    g = h.load(name)
    config.keep_terms = True
    try:
        h.run_kl_golden(name, device)
        got = h.kl_terms_from_product(elbo.last_terms)
    finally:
        config.keep_terms = False
    ref = {k: g["term_" + k] for k in ("A", "C", "E", "F", "kld_qu_pu", "S", "p")}
    ref["B+D"] = g["term_B"] + g["term_D"]
    iK = torch.linalg.inv(torch.as_tensor(g["K0zz"]) + float(g["eps"]) * torch.eye(int(g["M"]), dtype=DT))
    ref["scale_D"], ref["scale_E"] = h.trace_scales(iK, g["H"], g["term_S"])
    errs = h.assert_terms_close(got, ref, tol=5e-6, label=name)
    print(name, {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.parametrize("L,M,n_subj,T,storage,tol", [(8, 64, 40, 20, torch.float64, 1e-6),
                                                      (8, 64, 40, 20, torch.float32, 1e-4)])
def test_every_elbo_term_against_oracle(L, M, n_subj, T, storage, tol, device):
    """Same gate on fresh seeded inputs at M = 64 (the benchmark's M), float64 and float32 storage."""
    from oracle import hlvae_oracle as orc
    inp = h.make_kl_inputs(L, M, n_subj, T, seed=41)
    config.keep_terms = True
    try:
        h.run_kl_product(inp, device, storage=storage)
        got = h.kl_terms_from_product(elbo.last_terms)
    finally:
        config.keep_terms = False
    ref = h.oracle_kl(inp["kargs"], L, inp["x"], inp["mu"].to(storage).to(DT), inp["lv"].to(storage).to(DT), inp["z"],
                      inp["m"], inp["H"], inp["ros0"], inp["rls0"], inp["ros1"], inp["rls1"], inp["noise"], 200,
                      n_subj, 200 * T, 1e-6)["terms"]
    ref = {k: v.detach() for k, v in ref.items()}
    ref["B+D"] = ref["B"] + ref["D"]
    ref["scale_D"], ref["scale_E"] = h.trace_scales(ref["iK"], inp["H"], ref["S"])
    errs = h.assert_terms_close(got, ref, tol=tol, label=f"terms M={M} {storage}")
    print({k: f"{v:.1e}" for k, v in errs.items()})


def test_golden_with_host_known_lengths(device):
    r = h.run_kl_golden("kl_default_ragged", device, layout="lengths")
    h.assert_kl_close(r, tol=5e-6, hyper_tol=1e-4)


def test_golden_float32_storage(device):
    """mu / log_v stored in float32 (arithmetic stays float64): compared with the float64 reference
    outputs at the north_star tolerance."""
    r = h.run_kl_golden("kl_trained_like", device, storage=torch.float32)
    h.assert_kl_close(r, tol=1e-4, hyper_tol=1e-4)


@pytest.mark.parametrize("L,M,n_subj,T,ragged", [(8, 64, 20, 20, False), (4, 32, 30, 20, True), (3, 120, 12, 20, True),
                                                 (2, 128, 9, 32, False), (5, 17, 40, 7, True), (2, 64, 1, 1, False)])
def test_against_oracle(L, M, n_subj, T, ragged, device):
    errs = h.check_kl_vs_oracle(device, L, M, n_subj, T, seed=100 + M, tol=1e-6, hyper_tol=1e-4, ragged=ragged)
    print((L, M, n_subj, T), {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.parametrize("L,M,n_subj,T,ragged", [(3, 64, 6, 64, False), (4, 32, 14, 64, True), (2, 120, 8, 48, True),
                                                 (3, 64, 9, 33, False)])
def test_long_subjects_against_oracle(L, M, n_subj, T, ragged, device):
    """Subjects of 33 .. 64 rows (HLVAE_TMAX): the CTA-per-pair kernel next to the warp-per-pair one (ragged batches
    mix both), 64-row panels.  The reference's loop (elbo_functions.py:243-266) takes any subject length."""
    errs = h.check_kl_vs_oracle(device, L, M, n_subj, T, seed=300 + M + T, tol=1e-6, hyper_tol=1e-4, ragged=ragged)
    print((L, M, n_subj, T), {k: f"{v:.1e}" for k, v in errs.items()})


def test_too_long_subject_raises(device):
    inp = h.make_kl_inputs(2, 16, 3, 70, seed=5, ragged=False)
    with pytest.raises((RuntimeError, ValueError), match="rows"):
        h.run_kl_product(inp, device)


def test_well_conditioned_is_tight(device):
    """With few, distinct inducing points (cond(K0zz) small) the same comparison holds to 1e-9 on
    every term and gradient, which separates implementation error from the conditioning floor."""
    errs = h.check_kl_vs_oracle(device, 4, 8, 16, 20, seed=7, tol=1e-9, hyper_tol=1e-9, ragged=True, distinct_z=True,
                                kargs=h.synth.SWEEP_KERNEL_ARGS, continuous_age=True)
    print({k: f"{v:.1e}" for k, v in errs.items()})


def test_subject_order_and_sharding_invariance_full_size(device):
    """Size-independent properties at a BASELINE.json config-2 shape (L=32, M=64, N_b=4000):
    (i) permuting subjects leaves every output unchanged up to summation order;
    (ii) the accumulators of two disjoint subject shards add up to the unsharded ones."""
    L, M, P_b, T = 32, 64, 200, 20
    inp = h.make_kl_inputs(L, M, P_b, T, seed=11)
    base = h.run_kl_product(inp, device)
    lens = inp["lens"]
    perm = np.random.default_rng(0).permutation(P_b)
    starts = np.concatenate([[0], np.cumsum(lens)])
    rows = np.concatenate([np.arange(starts[s], starts[s + 1]) for s in perm])
    inp2 = dict(inp)
    inp2["x"], inp2["mu"], inp2["lv"] = inp["x"][rows], inp["mu"][rows], inp["lv"][rows]
    other = h.run_kl_product(inp2, device)
    # only the summation order of the accumulators changes; cond(K0zz + eps I) ~ 1e7 amplifies that
    assert h.rel_err(other["kld"], base["kld"]) < 1e-7
    assert h.rel_err(other["grad_H"], base["grad_H"]) < 1e-7
    assert h.rel_err(other["d_z"], base["d_z"]) < 1e-6
    inv = np.argsort(rows)
    assert h.rel_err(other["d_mu"][inv], base["d_mu"]) < 1e-7
    # sharding: run each half with a sub-layout and add the raw streaming outputs through kld linearity:
    # kld(all) - kld_qu part is additive in subjects, so compare d_mu rows (local) and the sum of d_z.
    full_lay = subjects.SubjectLayout.from_lengths(lens, device)
    halves = [full_lay.shard(r, 2) for r in range(2)]
    outs = [h.run_kl_product(inp, device, layout=lay) for lay in halves]
    # rows outside a shard get zero gradient
    r0 = halves[0].row_idx.cpu().numpy()
    r1 = halves[1].row_idx.cpu().numpy()
    assert float(outs[0]["d_mu"][r1].abs().max()) == 0.0 and float(outs[1]["d_mu"][r0].abs().max()) == 0.0


def test_non_pd_block_raises(device):
    inp = h.make_kl_inputs(2, 8, 3, 4, seed=3)
    inp["noise"] = torch.full((2,), -5.0, dtype=DT)      # B_s = K1 + noise I no longer positive definite
    k0, k1, lik = h.build_product_kernels(inp["kargs"], 2, device, inp["ros0"], inp["rls0"], inp["ros1"], inp["rls1"])

    class Lik:                                           # any object with noise_covar.noise is accepted
        class noise_covar:
            noise = torch.full((2, 1), -5.0, dtype=DT, device=device)
    with pytest.raises(RuntimeError, match="positive-definite"):
        elbo.minibatch_KLD_upper_bound_iter(k0, k1, Lik, 2, inp["m"].to(device), inp["H"].to(device),
                                            inp["x"].to(device), inp["mu"].to(device), inp["lv"].to(device),
                                            inp["z"].to(device), 10, 3, 40, True, 2, 1e-6)


def test_natural_gradient_step_matches_oracle(device):
    from oracle import hlvae_oracle as orc
    inp = h.make_kl_inputs(3, 16, 6, 8, seed=17)
    got = h.run_kl_product(inp, device)
    m2, H2 = elbo.natural_gradient_update(inp["m"].to(device), inp["H"].to(device), got["grad_m"], got["grad_H"], 0.01)
    ref = h.oracle_kl(inp["kargs"], 3, inp["x"], inp["mu"], inp["lv"], inp["z"], inp["m"], inp["H"], inp["ros0"],
                      inp["rls0"], inp["ros1"], inp["rls1"], inp["noise"], 200, 6, 200 * 8, 1e-6)
    m3, H3 = orc.natural_gradient_update(inp["m"], inp["H"], ref["grad_m"], ref["grad_H"], 0.01)
    assert h.rel_err(m2, m3) < 1e-6 and h.rel_err(H2, H3) < 1e-6
    # the update normally reuses H^-1 of the KL call that produced grad_H; plain tensors (no such record) make it
    # factorise H itself (training.py:131-132) - same result
    m4, H4 = elbo.natural_gradient_update(inp["m"].to(device), inp["H"].to(device), got["grad_m"].clone(),
                                          got["grad_H"].clone(), 0.01)
    assert h.rel_err(m4, m2) < 1e-9 and h.rel_err(H4, H2) < 1e-9


@pytest.mark.parametrize("M", [32, 64, 128])
def test_kernel_sweep_config(M, device):
    """BASELINE.json configs[2]: SE(time) + CA(id) + SE(age) x CA(sex), continuous ages, ragged T_s,
    M in {32, 64, 128} (each M selects a different panel-kernel shape), against the oracle."""
    errs = h.check_kl_vs_oracle(device, 4, M, 24, 20, seed=300 + M, tol=1e-6, hyper_tol=1e-4, ragged=True,
                                kargs=h.synth.SWEEP_KERNEL_ARGS, continuous_age=True)
    print(M, {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.parametrize("M,rp", [(64, 40), (64, 64), (120, 32), (120, 48), (120, 64), (24, 64)])
def test_every_panel_shape_against_oracle(M, rp, device, monkeypatch):
    """hlvae_kl_panel's shapes are picked per minibatch by elbo._row_panel; here each one is forced in turn
    (HLVAE_PANEL_RP) on the same ragged batch: two-CTA 40-row and one-CTA 64-row panels at M = 64, 32 / 48 / 64-row
    panels at M = 120, the two-CTA 64-row shape at M <= 32."""
    monkeypatch.setenv("HLVAE_PANEL_RP", str(rp))
    errs = h.check_kl_vs_oracle(device, 3, M, 17, 20, seed=700 + M, tol=1e-6, hyper_tol=1e-4, ragged=True)
    print((M, rp), {k: f"{v:.1e}" for k, v in errs.items()})


def test_row_panel_choice():
    """elbo._row_panel: rows per panel from the subject packing, never below the longest subject."""
    from hlvae_b200 import elbo, subjects
    lay = subjects.SubjectLayout.from_lengths([20] * 40, "cpu")
    assert elbo._row_panel(lay, 16) == 64 and elbo._row_panel(lay, 64) == 40 and elbo._row_panel(lay, 120) == 64
    lay = subjects.SubjectLayout.from_lengths([30] * 40, "cpu")          # one subject per 40 rows, two per 64
    assert elbo._row_panel(lay, 64) == 64
    lay = subjects.SubjectLayout.from_lengths([45, 12, 7, 60], "cpu")
    assert elbo._row_panel(lay, 64) == 64 and elbo._row_panel(lay, 120) == 64
    lay = subjects.SubjectLayout.from_lengths([33] * 6, "cpu")
    assert elbo._row_panel(lay, 120) in (48, 64)


def test_masked_bin_spec_against_oracle(device):
    """BinKernel factors, missing-covariate masks and bin x SE interactions (5 components in K0, two
    beyond the cached three) on fresh seeded inputs."""
    errs = h.check_kl_vs_oracle(device, 3, 40, 16, 12, seed=77, tol=1e-6, hyper_tol=1e-4, ragged=True,
                                kargs=h.synth.MASKED_KERNEL_ARGS)
    print({k: f"{v:.1e}" for k, v in errs.items()})


def test_full_size_minibatch_properties(device):
    """BASELINE.json configs[2] upper size: 64 000 rows (3200 subjects x T = 20), L = 32, M = 64, float32
    storage.  Size-independent checks: (i) kld and every replicated gradient are additive over a
    partition of the subjects for fixed (m, H, Z, hyper-parameters) - the streaming part is a sum over
    subjects and kld_qu_pu a constant; (ii) gradients w.r.t. mu / log_v of a subject do not depend on
    which other subjects are in the batch."""
    L, M, P_b, T = 32, 64, 3200, 20
    inp = h.make_kl_inputs(L, M, P_b, T, seed=13)
    lay = subjects.SubjectLayout.from_lengths(inp["lens"], device)
    parts = [lay.shard(r, 4) for r in range(4)]
    kw = dict(storage=torch.float32, P_tot=P_b)          # scale P / P_b = 1 for every run
    full = h.run_kl_product(inp, device, layout=lay, **kw)
    outs = []
    for part in parts:
        inp_p = dict(inp)
        inp_p["n_subj"] = P_b                             # keep the P / P_b scale of the full batch
        outs.append(h.run_kl_product(inp_p, device, layout=part, **kw))
    # kld(part) = J(part) + kq - const ; kld(full) = sum J(part) + kq - const
    empty = subjects.SubjectLayout.from_lengths([], device)
    inp_e = dict(inp)
    base = h.run_kl_product(inp_e, device, layout=empty, **kw)       # kq - const (no subjects)
    total = sum(o["kld"] - base["kld"] for o in outs) + base["kld"]
    # tr(G S) with |G| ~ |iK|^2 ~ 1e12 amplifies the float64 summation-order difference of S (measured 1e-7)
    assert h.rel_err(total, full["kld"]) < 1e-6
    d_mu = sum(o["d_mu"] for o in outs)
    assert h.rel_err(d_mu, full["d_mu"]) < 1e-5           # float32 storage of the gradient
    d_z = sum(o["d_z"] - base["d_z"] for o in outs) + base["d_z"]
    assert h.rel_err(d_z, full["d_z"]) < 5e-5
    gH = sum(o["grad_H"] - base["grad_H"] for o in outs) + base["grad_H"]
    assert h.rel_err(gH, full["grad_H"]) < 5e-5


def test_device_loader_layout_feeds_the_kl_kernels(device):
    """A minibatch from hlvae_b200.data.DeviceSubjectLoader (device gathers + CSR from known subject lengths) gives the
    same KL bound as grouping the same rows by their id column the way the reference does (torch.unique)."""
    import numpy as np
    from hlvae_b200 import data as D, subjects
    inp = h.make_kl_inputs(L=3, M=12, n_subj=9, T=7, seed=21, ragged=True)
    x = inp["x"]
    n = x.shape[0]
    ds = D.DeviceDataset(np.zeros((n, 4)), np.ones((n, 2)), x.numpy(), id_covariate=2, device=device)
    np.random.seed(3)
    batch = next(iter(D.DeviceSubjectLoader(ds, 5, varying_T=True)))
    rows = batch["idx"]
    assert batch["label"].is_cuda and batch["digit"].dtype == torch.uint8
    sub = dict(inp)
    sub["x"], sub["mu"], sub["lv"] = x[rows.cpu()], inp["mu"][rows.cpu()], inp["lv"][rows.cpu()]
    sub["n_subj"] = batch["layout"].n_subj
    a = h.run_kl_product(sub, device, layout=batch["layout"])
    b = h.run_kl_product(sub, device, layout=subjects.SubjectLayout.from_ids(sub["x"][:, 2].to(device)))
    assert h.rel_err(a["kld"], b["kld"]) < 1e-9 and h.rel_err(a["d_mu"], b["d_mu"]) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [True, False])
def test_fused_hyper_constraint_matches_stock_formulation(batched, device):
    """hlvae_hyper_constrain (one launch for softplus + lower bound of every outputscale / lengthscale of both
    kernels, one for the chain rule back) against the stock-PyTorch formulation FlatSpec.constrained: values and
    gradients of the raw parameters, for parameters batched over the latent dimensions and for scalar ones that are
    broadcast (their gradient is the sum over the latent dimensions)."""
    from hlvae_b200 import kernels
    L = 5
    if batched:
        k0, k1 = kernels.generate_kernel_batched(L, **h.synth.MASKED_KERNEL_ARGS)
    else:
        k0, k1 = kernels.generate_kernel_approx(**h.synth.MASKED_KERNEL_ARGS)
    k0, k1 = k0.to(device).double(), k1.to(device).double()
    gen = torch.Generator().manual_seed(3)
    for prm in list(k0.parameters()) + list(k1.parameters()):
        prm.data += (torch.randn(prm.shape, generator=gen, dtype=torch.float64) * 0.7).to(device)
    fs0, fs1 = kernels.compile_spec(k0), kernels.compile_spec(k1)
    ws = [torch.randn(n, L, generator=gen, dtype=torch.float64).to(device) for n in (fs0.ncomp, fs0.ncomp, fs1.ncomp, fs1.ncomp)]

    def run(outs):
        for prm in list(k0.parameters()) + list(k1.parameters()):
            prm.grad = None
        sum((o * w_).sum() for o, w_ in zip(outs, ws)).backward()
        return [o.detach().clone() for o in outs], [prm.grad.clone() for prm in list(k0.parameters()) + list(k1.parameters())]

    got_v, got_g = run(kernels.constrained_pair(fs0, fs1, L, device))
    ref_v, ref_g = run(fs0.constrained(L, device) + fs1.constrained(L, device))
    for a, b in zip(got_v, ref_v):
        assert a.shape == b.shape and h.rel_err(a, b) < 1e-14
    for a, b in zip(got_g, ref_g):
        assert a.shape == b.shape and h.rel_err(a, b) < 1e-13


@pytest.mark.gpu
def test_many_tiny_subjects_per_chunk(device, monkeypatch):
    """Chunks of more than 256 subjects (the part of the subject CSR a CTA copies to shared memory): 1500 subjects of
    three or four rows in four chunks per latent dimension."""
    monkeypatch.setenv("HLVAE_PANEL_WAVES", "0")
    # 5200 rows against 16 inducing points drawn from them: cond(K0zz + eps I) puts the float64 floor of grad_m / grad_H
    # at ~5e-6 (the same with chunks below 256 subjects: 3e-6 / 6e-6); every streamed quantity agrees to 1e-7
    errs = h.check_kl_vs_oracle(device, 2, 16, 1500, 4, seed=41, tol=5e-5, hyper_tol=1e-3, ragged=True)
    assert errs["kld"] < 1e-6 and errs["d_mu"] < 1e-8 and errs["d_z"] < 1e-6
    print({k: f"{v:.1e}" for k, v in errs.items()})
