"""Trajectory fixture: k real training steps of the UNMODIFIED reference `training.hensman_training`
(training.py:22-143) on a small synthetic heterogeneous longitudinal data set, recorded step by step.

Run in the build container (needs /root/reference):   python oracle/make_trajectory_golden.py
Writes tests/golden/trajectory_small.npz.  Test infrastructure, like everything under oracle/.

What runs: the reference's own dataset class (dataset_def.py, from CSV files written to a temporary directory),
its samplers / data loader (utils.py:24-97), HLVAE model (HLVAE.py), kernel_gen.generate_kernel_batched,
elbo_functions.minibatch_KLD_upper_bound_iter, the Adam step and the natural-gradient update, exactly as
HLVAE_main.py:200-300 wires them - behind the stand-ins for the absent gpytorch / matplotlib packages
(oracle/standins) and the `Sampler.__init__` shim of SURVEY.md section 8(b).  Two call sites are wrapped by RECORDERS
that pass everything through unchanged:
  * nnet_model.loglik_and_reconstruction  -> theta, data, mask, normalisation parameters, log_p_x (+ d loss / d theta)
  * training.minibatch_KLD_upper_bound_iter -> m, H, Z, kernel hyper-parameters, covariates, mu, log_v, and
    kld, grad_m, grad_H (+ d loss / d mu, d log_v)
so the fixture holds, for every step t: the state BEFORE the step (m_t, H_t, Z_t, raw kernel parameters, the
likelihood log-variances), the NN outputs of that step (theta_t, mu_t, log_v_t - the trunk is outside the path this
repo replaces), and the reference's results.  The state after the last step closes the trajectory.
tests/test_gpu_trajectory.py replays the same loop on the GPU through the drop-in surface, carrying ITS OWN state
from step to step, and compares loss / m / H / Z / hyper-parameters per step.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("HLVAE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "standins"))
sys.path.insert(1, REF)
sys.path.insert(2, ROOT)

import torch.utils.data as tud  # noqa: E402

tud.Sampler.__init__ = lambda self, data_source=None: None          # torch >= 2.2 (shim 3)

import gpytorch  # noqa: E402  (stand-in)
import dataset_def  # noqa: E402  (reference)
import kernel_gen as ref_kernel_gen  # noqa: E402
import training as ref_training  # noqa: E402
import HLVAE as ref_hlvae  # noqa: E402

DT = torch.float64
GOLD = os.path.join(ROOT, "tests", "golden")

TYPES = [('real', 1)] * 4 + [('pos', 1)] * 2 + [('count', 1)] * 2 + [('cat', 3)] * 2 + [('ordinal', 4)] * 2
KARGS = dict(cat_kernel=[2], bin_kernel=[], sqexp_kernel=[0],
             cat_int_kernel=[{'cont_covariate': 0, 'cat_covariate': 2}, {'cont_covariate': 0, 'cat_covariate': 3},
                             {'cont_covariate': 1, 'cat_covariate': 4}],
             bin_int_kernel=[], covariate_missing_val=[], id_covariate=2)


def write_csvs(tmp, rng, P, t_lo, t_hi):
    rows, raw, mask = [], [], []
    for s in range(P):
        Ts = int(rng.integers(t_lo, t_hi + 1))
        sick = rng.random() < 0.5
        onset = int(rng.integers(0, t_hi))
        for t in range(Ts):
            rows.append([float(t), float(t - onset) if sick else 0.0, float(s), float(s % 2), float(sick),
                         float(rng.random() < 0.5)])
    N = len(rows)
    cols = []
    for kind, C in TYPES:
        if kind == 'real':
            cols.append(rng.normal(0, 1, N))
        elif kind == 'pos':
            cols.append(np.exp(rng.normal(0, 0.7, N)))
        elif kind == 'count':
            cols.append(rng.poisson(3.0, N).astype(float) + 1)          # minimum > 0: no shift (read_functions.py:102-107)
        else:
            c = rng.integers(0, C, N)
            c[:C] = np.arange(C)                                        # every class present
            cols.append(c.astype(float))
    data = np.stack(cols, 1)
    m = (rng.random((N, len(TYPES))) < 0.8).astype(int)
    np.savetxt(os.path.join(tmp, "data.csv"), data, delimiter=",", fmt="%.10g")
    np.savetxt(os.path.join(tmp, "mask.csv"), m, delimiter=",", fmt="%d")
    with open(os.path.join(tmp, "types.csv"), "w") as f:
        f.write("type,dim,nclass\n")
        for kind, C in TYPES:
            f.write(f"{kind},1,{C}\n")
    with open(os.path.join(tmp, "labels.csv"), "w") as f:
        f.write("time_age,disease_time,subject,gender,disease,location\n")
        for r in rows:
            f.write(",".join(f"{v:.10g}" for v in r) + "\n")
    return N


def params_of(kmod):
    ros = torch.stack([k.raw_outputscale.detach().clone() for k in kmod.kernels])
    rls = torch.stack([mod.raw_lengthscale.detach().reshape(-1).clone() for mod in kmod.modules()
                       if isinstance(mod, gpytorch.kernels.RBFKernel)])
    return ros, rls


def main():
    seed, L, M, P, epochs, spb = 7, 3, 10, 9, 2, 3
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    with tempfile.TemporaryDirectory() as tmp:
        N = write_csvs(tmp, rng, P, 4, 6)
        ds = dataset_def.HeterogeneousHealthMNISTDataset("data.csv", "labels.csv", "mask.csv", "types.csv", "", tmp,
                                                         logvar_network=False)
        ti = ds.types_info
        ti['conv'], ti['use_ranges'], ti['conv_range'] = False, False, 255          # HLVAE_main.py:98-100
        E_x = ds.cov_dim_ext
        model = ref_hlvae.HLVAE([E_x, [16], L, [16], 3], ti, ds.n_variables, vy_init=[1., .5], logvar_network=False,
                                conv=False).double()
        Q = 6
        lik = gpytorch.likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]),
                                                      noise_constraint=gpytorch.constraints.GreaterThan(1.0e-8))
        lik.noise = 1
        lik.raw_noise.requires_grad = False
        k0, k1 = ref_kernel_gen.generate_kernel_batched(L, KARGS['cat_kernel'], KARGS['bin_kernel'], KARGS['sqexp_kernel'],
                                                        KARGS['cat_int_kernel'], KARGS['bin_int_kernel'],
                                                        KARGS['covariate_missing_val'], KARGS['id_covariate'])
        train_x = torch.tensor(np.nan_to_num(ds.label_source.values), dtype=DT)
        z = torch.zeros(L, M, Q, dtype=DT)
        for i in range(L):
            z[i] = train_x[np.random.choice(N, M, replace=False)].clone()            # HLVAE_main.py:224-226
        z.requires_grad_(True)
        k0.train().double(); k1.train().double(); lik.train().double()
        m = torch.randn(L, M, 1).double()
        H = (torch.randn(L, M, M) / 10).double()
        H = (H @ H.transpose(-1, -2)).detach()
        opt = torch.optim.Adam([{'params': k0.parameters()}, {'params': k1.parameters()}, {'params': z},
                                {'params': model.parameters()}], lr=1e-3)              # HLVAE_main.py:231-233,277-278
        model.train()

        steps = []
        cur = {}

        orig_ll = model.loglik_and_reconstruction

        def rec_ll(theta, batch_data_list, miss_list, param_miss_list, normalization_params, s=None):
            out = orig_ll(theta, batch_data_list, miss_list, param_miss_list, normalization_params, s)
            cur.clear()
            nr, npos = normalization_params[0], normalization_params[1]
            cur.update(theta=theta.detach().clone(), data=batch_data_list.detach().clone(), mask=miss_list.detach().clone(),
                       norm_real_mean=nr[0].detach().clone(), norm_real_var=nr[1].detach().clone(),
                       norm_pos_mean=npos[0].detach().clone(), norm_pos_var=npos[1].detach().clone(),
                       log_vy_real=model._log_vy_real.detach().clone(), log_vy_pos=model._log_vy_pos.detach().clone(),
                       log_p_x=out[0].detach().clone())
            # backward runs after this step's KL call has appended its record: the gradient lands in that record
            theta.register_hook(lambda g: steps[-1].__setitem__("d_theta", g.detach().clone()))
            return out

        model.loglik_and_reconstruction = rec_ll
        orig_kl = ref_training.minibatch_KLD_upper_bound_iter

        def rec_kl(c0, c1, lk, latent_dim, m_, H_, x, mu, log_v, z_, P_, P_b, N_, ng, idc, eps):
            out = orig_kl(c0, c1, lk, latent_dim, m_, H_, x, mu, log_v, z_, P_, P_b, N_, ng, idc, eps)
            ros0, rls0 = params_of(c0)
            ros1, rls1 = params_of(c1)
            st = dict(cur)
            st.update(m=m_.detach().clone(), H=H_.detach().clone(), z=z_.detach().clone(), ros0=ros0, rls0=rls0,
                      ros1=ros1, rls1=rls1, x=x.detach().clone(), mu=mu.detach().clone(), log_v=log_v.detach().clone(),
                      P=P_, P_b=P_b, N=N_, kld=out[0].detach().clone(), grad_m=out[1].detach().clone(),
                      grad_H=out[2].detach().clone())
            # d kld / d(mu, log_v) alone (in training the two also receive gradient through the decoder, which is
            # outside the path): an extra, purely observing autograd pass over the reference's own graph
            gk = torch.autograd.grad(out[0].sum(), [mu, log_v], retain_graph=True)
            st.update(d_mu=gk[0].detach().clone(), d_logv=gk[1].detach().clone())
            steps.append(st)
            return out

        # gradients of the replicated parameters as the reference's backward() leaves them (leaf hooks; z_ above is
        # zt_list itself, the kernel parameters are leaves of c0 / c1): what Adam consumes
        z.register_hook(lambda g: steps[-1].__setitem__("d_z", g.detach().clone()))

        ref_training.minibatch_KLD_upper_bound_iter = rec_kl
        try:
            res = ref_training.hensman_training(model, epochs, ds, opt, 'GPapprox_closed', 1, L, k0, k1, lik, m, H, z, P,
                                                6, True, Q, KARGS['id_covariate'], tmp, natural_gradient=True,
                                                natural_gradient_lr=0.01, subjects_per_batch=spb, eps=1e-6,
                                                results_path=tmp, validation_dataset=None, generation_dataset=None,
                                                prediction_dataset=None, save_interval=1000)
        finally:
            ref_training.minibatch_KLD_upper_bound_iter = orig_kl
        m_end, H_end = res[5], res[6]

    out = dict(types=np.array([f"{k}:{c}" for k, c in TYPES]), kargs=repr(KARGS), L=L, M=M, n_steps=len(steps),
               lr_adam=1e-3, lr_natgrad=0.01, eps=1e-6, noise=np.ones(L))
    for t, st in enumerate(steps):
        for key, val in st.items():
            if key in ("d_theta", "d_mu", "d_logv") and val is None:
                continue
            out[f"s{t}_{key}"] = val.numpy() if torch.is_tensor(val) else np.asarray(val)
    ros0, rls0 = params_of(k0)
    ros1, rls1 = params_of(k1)
    out.update(end_m=m_end.detach().numpy(), end_H=H_end.detach().numpy(), end_z=z.detach().numpy(), end_ros0=ros0.numpy(),
               end_rls0=rls0.numpy(), end_ros1=ros1.numpy(), end_rls1=rls1.numpy(),
               end_log_vy_real=model._log_vy_real.detach().numpy(), end_log_vy_pos=model._log_vy_pos.detach().numpy())
    np.savez_compressed(os.path.join(GOLD, "trajectory_small.npz"), **out)
    klds = [float(st["kld"]) for st in steps]
    print(f"trajectory_small: {len(steps)} steps of hensman_training, rows per step "
          f"{[int(st['x'].shape[0]) for st in steps]}, kld {klds[0]:.4e} -> {klds[-1]:.4e}")


if __name__ == "__main__":
    main()
