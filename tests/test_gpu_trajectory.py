"""k-step trajectory parity (the third numerical regime, "after k real training steps").

tests/golden/trajectory_small.npz holds six steps of the UNMODIFIED reference `training.hensman_training`
(training.py:70-143, recorded by oracle/make_trajectory_golden.py): per step the state before the step, the NN outputs
(theta, mu, log_v - the trunk is outside this repo's path) and the reference's results.  The test replays the same
loop on the GPU through the drop-in surface - fused likelihoods, minibatch_KLD_upper_bound_iter, backward, Adam on
(kernel hyper-parameters, Z, likelihood log-variances), natural-gradient update of (m, H) - carrying ITS OWN state
from step to step, so errors would compound, and compares every step's loss terms, gradients and state."""
import ast

import numpy as np
import pytest
import torch

import helpers as h
from hlvae_b200 import elbo, loglik

pytestmark = pytest.mark.gpu
DT = torch.float64


def test_six_steps_of_hensman_training(device):
    g = h.load("trajectory_small")
    types = h.parse_types(g)
    kargs = ast.literal_eval(str(g["kargs"]))
    L, n_steps = int(g["L"]), int(g["n_steps"])
    T = lambda key: h.t(g[key], device)
    k0, k1, lik = h.build_product_kernels(kargs, L, device, g["s0_ros0"], g["s0_rls0"], g["s0_ros1"], g["s0_rls1"],
                                          g["noise"])
    z = T("s0_z").requires_grad_(True)
    m, H = T("s0_m"), T("s0_H")
    lvr, lvp = T("s0_log_vy_real").requires_grad_(True), T("s0_log_vy_pos").requires_grad_(True)
    opt = torch.optim.Adam([{'params': k0.parameters()}, {'params': k1.parameters()}, {'params': z},
                            {'params': [lvr, lvp]}], lr=float(g["lr_adam"]))          # HLVAE_main.py:231-233,277-278
    lay = loglik.VarLayout(types, device)
    worst = {}

    def chk(name, got, ref, tol):
        e = h.rel_err(got, ref)
        worst[name] = max(worst.get(name, 0.0), e)
        assert e <= tol, f"step {t}: {name} differs by {e:.2e} (> {tol})"

    def state_check(prefix, tol, tol_adam):
        """m, H follow the (deterministic) natural-gradient rule; Z, the kernel hyper-parameters and the likelihood
        log-variances are moved by Adam, whose step lr * g / (|g| + eps) turns the float64 noise floor of the
        hyper-parameter gradients (~1e-8 |kld|; kld ~ 1e6 here) into errors of up to ~1e-3 of a
        step: their tolerance is a fraction of the distance Adam has moved them (lr per step)."""
        chk("m", m, g[prefix + "m"], tol)
        chk("H", H, g[prefix + "H"], tol)
        chk("z", z, g[prefix + "z"], tol)
        lr = float(g["lr_adam"])
        for (gos, gls), key in ((h_params(k0), "0"), (h_params(k1), "1")):
            # raw_outputscale starts at 0 and has moved t * lr at most: measured against that distance
            d = float((gos.cpu() - torch.as_tensor(g[f"{prefix}ros{key}"])).abs().max()) / (lr * max(t, 1))
            worst["raw_outputscale" + key + " / (lr t)"] = max(worst.get("raw_outputscale" + key + " / (lr t)", 0.0), d)
            assert d <= tol_adam, f"step {t}: raw_outputscale{key} is off by {d:.2e} Adam steps"
            chk("raw_lengthscale" + key, gls, g[f"{prefix}rls{key}"], tol)
        chk("log_vy_real", lvr, g[prefix + "log_vy_real"], tol)
        chk("log_vy_pos", lvp, g[prefix + "log_vy_pos"], tol)

    def h_params(kmod):
        from hlvae_b200 import kernels
        ros = torch.stack([k.raw_outputscale.detach().reshape(-1) for k in kmod.kernels])
        rls = torch.stack([mod.raw_lengthscale.detach().reshape(-1) for mod in kmod.modules()
                           if isinstance(mod, kernels.RBFKernel)])
        return ros, rls

    def sync_noise_entries(t_next):
        """Adam divides by |g|: an inducing-point coordinate whose gradient is mathematically zero moves by a full
        +-lr step in the direction of the REFERENCE'S OWN float64 rounding noise (|g| ~ 1e-7 against gradients of
        1e5..1e6 in the same tensor).  Those coordinates are not reproducible by any other evaluation order -
        including another run of the reference - so they are taken from the recording; every coordinate with a
        gradient above 1e-6 of the tensor's largest is carried by this test's own Adam and compared."""
        gref = torch.as_tensor(g[f"s{t_next - 1}_d_z"], device=device)
        noise = gref.abs() < 1e-6 * gref.abs().max()
        znext = T(f"s{t_next}_z") if t_next < n_steps else T("end_z")
        with torch.no_grad():
            z[noise] = znext[noise]
        return int(noise.sum()), int((gref != 0).sum())

    for t in range(n_steps):
        s = f"s{t}_"
        if t > 0:
            n_noise, n_nz = sync_noise_entries(t)
        state_check(s, 1e-7, 1e-3)                               # our own carried state vs the reference's, before step t
        opt.zero_grad()
        theta = T(s + "theta").requires_grad_(True)
        mu, lv = T(s + "mu").requires_grad_(True), T(s + "log_v").requires_grad_(True)
        vparam = lay.vparam(lvr, lvp, (T(s + "norm_real_mean"), T(s + "norm_real_var")),
                            (T(s + "norm_pos_mean"), T(s + "norm_pos_var")), conv=False)
        out = loglik.fused_loglik(lay, T(s + "data"), T(s + "mask"), theta, vparam)
        chk("log_p_x", out["log_p_x"], g[s + "log_p_x"], 1e-10)
        P, P_b, N = int(g[s + "P"]), int(g[s + "P_b"]), int(g[s + "N"])
        nll = -out["log_p_x_sum"] * P / P_b                                            # training.py:83,104,122
        kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(k0, k1, lik, L, m, H, T(s + "x"), mu, lv, z, P, P_b, N, True,
                                                          kargs["id_covariate"], float(g["eps"]))   # :110-113
        (nll + kld).backward()                                                         # :124-127
        opt.step()                                                                     # :128
        chk("kld", kld, g[s + "kld"], 1e-6)
        chk("grad_m", gm, g[s + "grad_m"], 5e-6)
        chk("grad_H", gH, g[s + "grad_H"], 5e-6)
        chk("d_theta", theta.grad, g[s + "d_theta"], 1e-12)
        chk("d_mu", mu.grad, g[s + "d_mu"], 1e-6)
        chk("d_logv", lv.grad, g[s + "d_logv"], 1e-6)
        chk("d_z", z.grad, g[s + "d_z"], 1e-6)
        m, H = elbo.natural_gradient_update(m, H, gm, gH, float(g["lr_natgrad"]))     # :130-137
    t = n_steps
    sync_noise_entries(n_steps)
    state_check("end_", 1e-7, 1e-3)
    print({k: f"{v:.1e}" for k, v in worst.items()})
