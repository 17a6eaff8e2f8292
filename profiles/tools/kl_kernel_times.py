import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
import __graft_entry__ as g
g.build()
from hlvae_b200 import _lib, config, elbo, kernels, likelihoods, subjects, synth
config.check_errors = False
config.overlap = False
dev = torch.device("cuda:0")
L, T = 32, 20
for M, n_subj in ((64, 800), (120, 800), (32, 800), (64, 20)):
    rng = np.random.default_rng(0); gen = torch.Generator().manual_seed(0)
    x, lens = synth.covariates(n_subj, T, rng)
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
    def step():
        for t_ in (mu, lv, z, *k0.parameters(), *k1.parameters()): t_.grad = None
        kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(k0, k1, lik, L, md, Hd, xd, mu, lv, z, 200, n_subj, 4000, True, 2, 1e-6, layout=lay)
        kld.sum().backward()
        return elbo.natural_gradient_update(md, Hd, gm, gH, 0.01)
    for _ in range(3): step()
    torch.cuda.synchronize()
    _lib.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    per = {}
    for name, a, b in _lib.PROFILE: per.setdefault(name, []).append(a.elapsed_time(b))
    _lib.PROFILE = None
    print(M, N_b, "step ms", round(e0.elapsed_time(e1) / 10, 3), {k[6:]: round(float(np.mean(v)), 3) for k, v in per.items()}, flush=True)
