"""One kl_panel CTA per latent dimension (HLVAE_PANEL_WAVES=0, <= 30 panels), several runs: everything a CTA sums in a
fixed order (S, p, d_mu, d_logv) must come out bit-identical from run to run - a shared-memory race in the panel
pipeline would show up here."""
import os, sys
os.environ["HLVAE_PANEL_WAVES"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from hlvae_b200 import config, elbo, kernels, likelihoods, subjects, synth
config.check_errors = False
config.keep_terms = True
dev = torch.device("cuda:0")
bad = 0
for M, n_subj, T, ragged, L in ((64, 60, 20, False, 8), (64, 45, 20, True, 8), (32, 90, 20, True, 4), (120, 30, 20, False, 4), (16, 7, 10, True, 3)):
    rng = np.random.default_rng(0); gen = torch.Generator().manual_seed(0)
    x, lens = synth.covariates(n_subj, T, rng, ragged=ragged)
    pool, _ = synth.covariates(400, T, np.random.default_rng(1))
    z = synth.inducing_points(pool, L, M, np.random.default_rng(1)).to(dev).requires_grad_(True)
    m, H = synth.variational_state(L, M, gen)
    k0, k1 = kernels.generate_kernel_batched(L, **synth.DEFAULT_KERNEL_ARGS)
    k0, k1 = k0.to(dev).double(), k1.to(dev).double()
    lik = likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=likelihoods.GreaterThan(1e-8)); lik.noise = 1
    lik = lik.to(dev).double()
    N_b = x.shape[0]
    mu = torch.randn(N_b, L, generator=gen, dtype=torch.float64).to(dev).requires_grad_(True)
    lv = (-3.0 * torch.rand(N_b, L, generator=gen, dtype=torch.float64)).to(dev).requires_grad_(True)
    lay = subjects.SubjectLayout.from_lengths(lens, dev)
    xd, md, Hd = x.to(dev), m.to(dev), H.to(dev)
    outs = []
    for it in range(6):
        for t_ in (mu, lv, z, *k0.parameters(), *k1.parameters()): t_.grad = None
        kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(k0, k1, lik, L, md, Hd, xd, mu, lv, z, 200, n_subj, 4000, True, 2, 1e-6, layout=lay)
        kld.sum().backward()
        torch.cuda.synchronize()
        t = elbo.last_terms
        outs.append([t["S"].clone(), t["p"].clone(), mu.grad.clone(), lv.grad.clone()])
    for it in range(1, 6):
        for name, a, b in zip(("S", "p", "d_mu", "d_logv"), outs[0], outs[it]):
            if not torch.equal(a, b):
                bad += 1
                print("MISMATCH", M, n_subj, name, it, float((a - b).abs().max() / a.abs().max()))
    print("case", M, n_subj, T, ragged, "rows", N_b, "done", flush=True)
print("determinism:", "OK" if bad == 0 else f"{bad} mismatches")
