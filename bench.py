#!/usr/bin/env python
"""Benchmark of the HL-VAE per-step ELBO hot path (BASELINE.json metric: ELBO train steps/sec,
forward + backward).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): synthetic HealthMNIST-shaped longitudinal minibatch, L=32
latent dimensions, M=64 inducing points, 800 subjects x T=20 = 16000 rows per step, D4 variable
layout (324 real + 972 categorical x 5 -> E_x = P_theta = 5184), default additive kernel.
One step = fused masked log-likelihood forward+backward (theta given) + KL upper bound
forward+backward (mu, log_v given) + natural-gradient update of (m, H): the part of
training.py:82-137 this repo owns (the NN trunk stays stock PyTorch and is not timed).

`value`  : steps/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`    : steps/s through the same public functions with HOST (pinned) inputs: every step copies
           data, mask, theta, mu, log_v, covariates host->device and reads the loss back.
`roofline`: dominant kernel of the step, algorithmic work / CUDA-event time of that kernel measured
           inside the timed region (events bracket each C-ABI call on the launching stream).
`cpu_baseline`: the oracle port of the reference path (float64 PyTorch, all host cores) on a bounded
           sample of the same workload; `--impl reference` runs only that and prints it as the line.
N > 1: weak scaling, every rank owns 800 subjects of a global N x 800-subject minibatch and the
per-latent accumulators are all-reduced once per step (NCCL); value counts 16000-row step
equivalents per second over all ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics as pystat
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L, M, Q, T = 32, 64, 6, 20
SUBJ_PER_RANK = 800
P_TOTAL, N_TOTAL = 5000, 100000            # ~100k-sample dataset of configs[1]
EPS, NG_LR = 1e-6, 0.01
WORKLOAD = "configs[1]: synthetic HealthMNIST-shaped, L=32, M=64, 800 subjects x T=20 = 16000 rows/step, D4 (324 real + 972 cat x5)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--subjects", type=int, default=SUBJ_PER_RANK)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the compact configs[2]/[3]/[4] figures of the default line")
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-overlap", action="store_true", help="single stream: no likelihood / KL overlap")
    ap.add_argument("--no-split-backward", action="store_true", help="one backward() of nll + kld instead of one per branch")
    ap.add_argument("--no-natgrad-stream", action="store_true", help="natural-gradient update on the KL stream, not its own")
    ap.add_argument("--no-kl-priority", action="store_true", help="KL stream at default priority")
    ap.add_argument("--workload", default="elbo", choices=["elbo", "predict", "sweep", "theta", "full", "norm", "contraction"],
                    help="elbo: the BASELINE.json metric (default); predict: SURVEY 8(f) row 1, GP posterior-mean "
                         "prediction (utils.batch_predict_varying_T), its own JSON line")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------
def cpu_sample_state(sample_subjects, Lx=L, Mx=M, Tx=T, types=None, conv=True, seed=0):
    """One minibatch of the benchmark workload as float64 CPU tensors for oracle.elbo_path_step: `sample_subjects`
    subjects x Tx rows, default additive kernel, D4 variable layout (or `types`)."""
    from hlvae_b200 import synth
    from oracle import hlvae_oracle as orc
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    x, lens = synth.covariates(sample_subjects, Tx, rng)
    pool, _ = synth.covariates(200, Tx, rng)
    z = synth.inducing_points(pool, Lx, Mx, rng).requires_grad_(True)
    N_b = x.shape[0]
    types = synth.HEALTHMNIST_D4_TYPES if types is None else types
    descs, E_x, P_th = orc.build_layout(types)
    data, mask = synth.likelihood_batch(types, N_b, rng, observed=0.75, pixel_like=True)
    spec0, spec1 = orc.compile_spec(**synth.DEFAULT_KERNEL_ARGS)
    prm0 = orc.KernelParams.default(spec0, Lx).requires_grad_()
    prm1 = orc.KernelParams.default(spec1, Lx).requires_grad_()
    m, H = synth.variational_state(Lx, Mx, gen)
    # theta, mu, log_v as the benchmark stores them (float32 values), held in float64 for the oracle
    f32 = lambda t_: t_.float().double()
    theta = f32(torch.randn(N_b, P_th, generator=gen, dtype=torch.float64)).requires_grad_(True)
    mu = f32(torch.randn(N_b, Lx, generator=gen, dtype=torch.float64)).requires_grad_(True)
    lv = f32(-3.0 * torch.rand(N_b, Lx, generator=gen, dtype=torch.float64)).requires_grad_(True)
    n_real = sum(1 for k, _ in types if k == "real")
    lvr = torch.zeros(n_real, dtype=torch.float64, requires_grad=True)
    return dict(descs=descs, data=data, mask=mask, theta=theta, log_vy_real=lvr, conv=conv, spec0=spec0, prm0=prm0,
                spec1=spec1, prm1=prm1, noise=torch.ones(Lx, dtype=torch.float64), m=m, H=H, x=x, mu=mu, log_v=lv,
                z=z, P=P_TOTAL, P_b=sample_subjects, N=N_TOTAL, id_covariate=2, eps=EPS, lens=lens,
                leaves=[theta, mu, lv, z, lvr, prm0.raw_outputscale, prm0.raw_lengthscale, prm1.raw_outputscale,
                        prm1.raw_lengthscale])


def cpu_reference_arm(steps, warmup, sample_subjects=100, budget_s=25.0, keep_first=False, **kw):
    """Times oracle.elbo_path_step (float64 PyTorch restatement of training.py:82-137 for this path) on all host
    cores, on `sample_subjects` x T rows of the benchmark workload.  `value` is the MEASURED rate on that sample
    converted to 16000-row step equivalents in proportion to the rows; with sample_subjects = 800 (the --impl reference
    arm) the sample IS the 16000-row step and nothing is scaled.  (Measured: the oracle's cost grows faster than
    linearly in the rows - 2000 rows x8 under-states the 16000-row step time about 4x - so a scaled small sample
    flatters the CPU.)  Stops after `steps` timed steps or when `budget_s` is used up (at least one timed step)."""
    from oracle import hlvae_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    state = cpu_sample_state(sample_subjects, **kw)
    N_b = state["x"].shape[0]
    times, first = [], None
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = orc.elbo_path_step(state, NG_LR)
        if i == 0 and keep_first:
            first = dict(out, d_theta=state["theta"].grad.clone(), d_mu=state["mu"].grad.clone(),
                         d_logv=state["log_v"].grad.clone(), d_z=state["z"].grad.clone(), m0=state["m"].clone(),
                         H0=state["H"].clone())
        state["m"], state["H"] = out["m"], out["H"]
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        else:
            t_start = time.perf_counter()            # the budget covers the timed steps
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    ms = 1e3 * pystat.median(times)
    scale = (SUBJ_PER_RANK * T) / N_b
    note = "" if scale == 1 else f"; rate converted x{scale:.0f} in rows to 16000-row step equivalents"
    return dict(value=1e3 / (ms * scale), ms_sample_step=ms, steps_timed=len(times), cores=torch.get_num_threads(),
                sample=f"{N_b} rows ({sample_subjects} subjects x T={T}) of the same workload per step, float64 oracle "
                       f"port, {len(times)} timed steps" + note,
                state=state, first=first)


def run_reference(args):
    """Reference arm: the oracle port of the reference's CPU path at the benchmark's OWN configuration - 16000 rows
    per step, nothing scaled.  One such step takes tens of seconds on the host cores, so the arm does 1 warm-up step
    and then as many timed steps as fit a 60 s budget (at least one) and reports the true count."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_arm(max(args.steps, 1), 1, sample_subjects=SUBJ_PER_RANK, budget_s=60.0)
    ms = r["ms_sample_step"]
    small = cpu_reference_arm(5, 1, sample_subjects=100, budget_s=15.0)
    c1 = cpu_reference_arm(5, 1, sample_subjects=20, budget_s=15.0, Mx=120)
    line = dict(metric="ELBO train steps/sec (fwd+bwd)", value=1e3 / ms, unit="steps/s", impl="reference",
                n_gpus=args.gpus, steps=r["steps_timed"], warmup=1, ms_per_step=ms, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD, rows_per_step=SUBJ_PER_RANK * T,
                            arm="oracle port of the reference CPU path (the shipped reference needs gpytorch, absent "
                                "on the box), measured at the full 16000-row step; 1 warm-up, timed steps within a 60 s budget"),
                cpu_baseline=dict(value=1e3 / ms, unit="steps/s", cores=r["cores"], kind="port", sample=r["sample"],
                                  small_sample=dict(value=small["value"], sample=small["sample"]),
                                  config1=dict(value=1e3 / c1["ms_sample_step"], unit="steps/s",
                                               sample="BASELINE.json configs[0] shape: L=32, M=120, 400 rows (20 subjects x T=20), D4; "
                                                      f"{c1['steps_timed']} timed steps")),
                e2e=dict(value=1e3 / ms, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region, through NVML (10 ms period; an
    nvidia-smi subprocess per sample would be slower than the whole timed region)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self.stop_flag, self.err = index, [], None, set(), False, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[0].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self.stop_flag:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                bits = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for n, b in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(n)
                time.sleep(0.01)
        except Exception as e:      # keep the benchmark alive; the JSON line then shows samples = 0
            self.err = repr(e)

    def summary(self):
        d = dict(sm_mhz=(pystat.median(self.sm) if self.sm else None), sm_max_mhz=self.mx, reasons=sorted(self.reasons),
                 samples=len(self.sm))
        if self.err:
            d["error"] = self.err
        return d


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure_fp64_peak(dev):
    """FP64 GEMM throughput of this GPU (cuBLAS DGEMM 4096^3, best of 5): the denominator for the
    FP64-tensor-pipe kernel.  MEASURED_PEAKS.json only has bf16 and HBM."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def build_gpu_state(dev, n_subj, rank, seed=0):
    from hlvae_b200 import kernels, likelihoods, loglik, subjects, synth
    rng = np.random.default_rng(seed + rank)
    gen_cpu = torch.Generator().manual_seed(seed)               # replicated state: same on every rank
    gen_dev = torch.Generator(device=dev).manual_seed(seed + rank)
    x, lens = synth.covariates(n_subj, T, rng, first_id=rank * n_subj)
    pool, _ = synth.covariates(400, T, np.random.default_rng(seed))
    z = synth.inducing_points(pool, L, M, np.random.default_rng(seed)).to(dev).requires_grad_(True)
    m, H = synth.variational_state(L, M, gen_cpu)
    k0, k1 = kernels.generate_kernel_batched(L, **synth.DEFAULT_KERNEL_ARGS)
    k0, k1 = k0.to(dev).double(), k1.to(dev).double()
    lik = likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=likelihoods.GreaterThan(1e-8))
    lik.noise = 1
    lik = lik.to(dev).double()
    lik.raw_noise.requires_grad = False
    lay = loglik.VarLayout(synth.HEALTHMNIST_D4_TYPES, dev)
    N_b = x.shape[0]
    data, mask = synth.device_likelihood_batch(lay, N_b, dev, gen_dev, dtype=torch.uint8)   # pixel values / one-hot codes: exact in uint8
    theta = torch.randn(N_b, lay.P_theta, device=dev, generator=gen_dev, dtype=torch.float32)
    mu = torch.randn(N_b, L, device=dev, generator=gen_dev, dtype=torch.float32)
    lv = -3.0 * torch.rand(N_b, L, device=dev, generator=gen_dev, dtype=torch.float32)
    log_vy_real = torch.zeros(324, dtype=torch.float64, device=dev, requires_grad=True)
    return dict(k0=k0, k1=k1, lik=lik, z=z, m=m.to(dev), H=H.to(dev), lay=lay, x=x.to(dev), data=data, mask=mask,
                theta=theta.requires_grad_(True), mu=mu.requires_grad_(True), lv=lv.requires_grad_(True),
                log_vy_real=log_vy_real, layout=subjects.SubjectLayout.from_lengths(lens, dev), n_subj=n_subj, N_b=N_b)


INPUT_KEYS = ("data", "mask", "x", "theta", "mu", "lv")


def elbo_step(s, world, inp=None):
    """One ELBO-path step through the public (reference-shaped) functions: training.py:82-137 minus the
    NN trunk and Adam.  `inp` (default: s) holds the step's inputs data / mask / x / theta / mu / lv; the
    replicated state (kernels, Z, m, H, log-variances) lives in `s` and is updated in place, so the
    function can be captured into a CUDA graph (hlvae_b200.graph.StepGraph)."""
    from hlvae_b200 import elbo, loglik
    inp = s if inp is None else inp
    for t_ in (inp["theta"], inp["mu"], inp["lv"], s["z"], s["log_vy_real"], *s["k0"].parameters(),
               *s["k1"].parameters()):
        t_.grad = None
    P_b = s["n_subj"] * world
    # The KL branch (and the natural-gradient update, which needs only its forward outputs) runs on a second
    # stream next to the HBM-bound likelihood kernels; autograd replays each branch on its own stream.
    cur = torch.cuda.current_stream()
    side = s.get("side")
    # net_loss = nll + kld (training.py:124) has two branches with disjoint leaves: backward() per branch gives the same
    # gradients and lets the likelihood backward kernel run next to the KL forward kernels instead of after them
    split = side is not None and s.get("split_backward", True)
    if side is not None:
        side.wait_stream(cur)
    with torch.cuda.stream(side if side is not None else cur):
        kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(s["k0"], s["k1"], s["lik"], L, s["m"], s["H"], inp["x"],
                                                          inp["mu"], inp["lv"], s["z"], P_TOTAL, P_b, N_TOTAL, True,
                                                          2, EPS, layout=s["layout"])            # training.py:110-113
        side2 = s.get("side2") if split else None
        if side2 is not None:
            # the natural-gradient update (32 CTAs, latency-bound) needs only grad_m / grad_H: it runs on a third stream
            # next to the KL backward (kernel_eval_bwd + the hyper-parameter transforms' backward kernels)
            side2.wait_stream(side)
            with torch.cuda.stream(side2):
                m_new, H_new = elbo.natural_gradient_update(s["m"], s["H"], gm, gH, NG_LR)     # :130-137
            for t_ in (gm, gH):
                t_.record_stream(side2)
        else:
            m_new, H_new = elbo.natural_gradient_update(s["m"], s["H"], gm, gH, NG_LR)         # :130-137
        if split:
            kld.backward()
    vparam = s["lay"].vparam(log_vy_real=s["log_vy_real"], conv=True)
    out = loglik.fused_loglik(s["lay"], inp["data"], inp["mask"], inp["theta"], vparam, monitor=True)
    nll = -out["log_p_x_sum"] * (P_TOTAL / P_b)                                             # training.py:83,104,122
    if split:
        nll.backward()
    if side is not None:
        cur.wait_stream(side)
        if split and s.get("side2") is not None:
            cur.wait_stream(s["side2"])
        for t_ in (kld, m_new, H_new):      # allocated on the side stream, consumed on this one
            t_.record_stream(cur)
    loss = nll + kld                                                                           # :124
    if not split:
        loss.backward()                                                                        # :127
    s["m"].copy_(m_new)
    s["H"].copy_(H_new)
    s["last"] = dict(nll=nll.detach(), kld=kld.detach())
    return loss.detach()


def gpu_state_from_cpu(dev, st, m0, H0):
    """The oracle's sample minibatch (cpu_sample_state) in the benchmark's storage configuration: theta, mu, log_v
    float32, data and mask uint8, covariates / Z / m / H float64, default kernel parameters."""
    from hlvae_b200 import kernels, likelihoods, loglik, subjects, synth
    k0, k1 = kernels.generate_kernel_batched(L, **synth.DEFAULT_KERNEL_ARGS)
    k0, k1 = k0.to(dev).double(), k1.to(dev).double()
    lik = likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=likelihoods.GreaterThan(1e-8))
    lik.noise = 1
    lik = lik.to(dev).double()
    lik.raw_noise.requires_grad = False
    lay = loglik.VarLayout(synth.HEALTHMNIST_D4_TYPES, dev)
    f32 = lambda a: a.detach().float().to(dev).requires_grad_(True)
    return dict(k0=k0, k1=k1, lik=lik, z=st["z"].detach().to(dev).requires_grad_(True), m=m0.to(dev).clone(),
                H=H0.to(dev).clone(), lay=lay, x=st["x"].to(dev), data=st["data"].to(torch.uint8).to(dev),
                mask=st["mask"].to(torch.uint8).to(dev), theta=f32(st["theta"]), mu=f32(st["mu"]), lv=f32(st["log_v"]),
                log_vy_real=torch.zeros(324, dtype=torch.float64, device=dev, requires_grad=True),
                layout=subjects.SubjectLayout.from_lengths(st["lens"], dev), n_subj=len(st["lens"]),
                N_b=st["x"].shape[0], side=None, side2=None)


def parity_vs_oracle(dev, r, streams):
    """The benchmarked step (float32 theta / mu / log_v, uint8 data and mask, the same streams and split backward)
    against the float64 oracle on the cpu_baseline sample's first step: every value the step produces."""
    st, first = r["state"], r["first"]
    g = gpu_state_from_cpu(dev, st, first["m0"], first["H0"])
    g.update(streams)
    loss = elbo_step(g, 1)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double().cpu() - b).abs().max() / (b.abs().max() + 1e-300))

    def elem(a, b, rtol=1e-4, afloor=1e-7):
        a = a.double().cpu()
        return float(((a - b).abs() / (rtol * b.abs() + afloor * b.abs().max() + 1e-300)).max())

    errs = dict(loss=rel(loss, first["loss"]), nll=rel(g["last"]["nll"], first["nll"]), kld=rel(g["last"]["kld"], first["kld"]),
                m_new=rel(g["m"].reshape(-1), first["m"].reshape(-1)), H_new=rel(g["H"], first["H"]),
                d_theta=rel(g["theta"].grad, first["d_theta"]), d_mu=rel(g["mu"].grad, first["d_mu"]),
                d_logv=rel(g["lv"].grad, first["d_logv"]), d_z=rel(g["z"].grad, first["d_z"]))
    worst_elem = dict(d_theta=elem(g["theta"].grad, first["d_theta"]), d_mu=elem(g["mu"].grad, first["d_mu"]),
                      d_logv=elem(g["lv"].grad, first["d_logv"]))
    return dict(against="float64 oracle port, first step of the cpu_baseline sample (same seeded inputs, float32-rounded "
                        "theta / mu / log_v)", rows=int(st["x"].shape[0]), tol=1e-4, rel_err=errs,
                elementwise_ratio_to_tol=worst_elem,
                ok=bool(max(errs.values()) <= 1e-4 and max(worst_elem.values()) <= 1.0))


def parity_sharded(dev, rank, world, n_subj_rank=64):
    """Multi-GPU correctness inside the scaling run: the KL bound of a small global minibatch computed sharded over
    the ranks (one all-reduce of the accumulators) against the same minibatch computed unsharded on rank 0."""
    import torch.distributed as dist
    from hlvae_b200 import config, elbo, kernels, likelihoods, subjects, synth
    rng = np.random.default_rng(123)
    gen = torch.Generator().manual_seed(123)
    P_b = n_subj_rank * world
    x, lens = synth.covariates(P_b, T, rng, ragged=True, t_min=5)
    pool, _ = synth.covariates(200, T, rng)
    z = synth.inducing_points(pool, L, M, rng).to(dev)
    m, H = synth.variational_state(L, M, gen)
    N_b = x.shape[0]
    mu = torch.randn(N_b, L, generator=gen, dtype=torch.float64).to(dev)
    lv = (-3.0 * torch.rand(N_b, L, generator=gen, dtype=torch.float64)).to(dev)
    k0, k1 = kernels.generate_kernel_batched(L, **synth.DEFAULT_KERNEL_ARGS)
    k0, k1 = k0.to(dev).double(), k1.to(dev).double()
    lik = likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=likelihoods.GreaterThan(1e-8))
    lik.noise = 1
    lik = lik.to(dev).double()
    full = subjects.SubjectLayout.from_lengths(lens, dev)

    def run(layout):
        zz = z.clone().requires_grad_(True)
        mm = mu.clone().requires_grad_(True)
        kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(k0, k1, lik, L, m.to(dev), H.to(dev), x.to(dev), mm, lv, zz,
                                                          P_TOTAL, P_b, N_TOTAL, True, 2, EPS, layout=layout)
        kld.sum().backward()
        return kld.detach().reshape(()), gm, gH, zz.grad, mm.grad

    sh = run(full.shard(rank, world))
    d_mu = sh[4].clone()
    dist.all_reduce(d_mu)                    # rows are disjoint across ranks: the sum is the full gradient
    out = None
    if rank == 0:
        pg, config.process_group = config.process_group, None
        un = run(full)
        config.process_group = pg
        rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-300))
        errs = dict(kld=rel(sh[0], un[0]), grad_m=rel(sh[1], un[1]), grad_H=rel(sh[2], un[2]), d_z=rel(sh[3], un[3]),
                    d_mu=rel(d_mu, un[4]))
        out = dict(against=f"the same {P_b}-subject ({N_b}-row) ragged minibatch unsharded on rank 0", world=world,
                   rel_err=errs, tol=1e-6, ok=bool(max(errs.values()) <= 1e-6))
    dist.barrier()
    return out


def bind_to_gpu_numa_node(dev):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity for the device) before the pinned host
    buffers of the end-to-end loop are allocated: first touch then places them in the memory of that socket, and the
    8 ranks of a box stop pulling their 440 MB per step through one socket's memory controllers and the inter-socket
    link.  Returns a short description for the JSON line (None when NVML or the call is not available)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev.index if dev.index is not None else torch.cuda.current_device())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"{len(cpus)} cpus ({cpus[0]}-{cpus[-1]}) next to GPU {dev.index}"
    except Exception as e:                                     # keep going unbound
        return f"unbound ({type(e).__name__})"


def allreduce_latency_us(dev, n_doubles, reps=50):
    """Device time of one NCCL all-reduce of the accumulator buffer (CUDA events, average of `reps` back to back)."""
    import torch.distributed as dist
    buf = torch.zeros(n_doubles, dtype=torch.float64, device=dev)
    for _ in range(5):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


ALGO = {}


def algorithmic_work(n_rows, n_subj):
    """Algorithmic work per launch (DESIGN.md section 'Kernels and rooflines')."""
    E_x = P_th = 5184
    D = 1296
    sz = 4                                   # float32 storage of theta and outputs; data and mask are uint8
    ALGO["hlvae_loglik_fwd"] = ("hbm", n_rows * (E_x + sz * P_th + D + sz * (5 * D + P_th)))
    ALGO["hlvae_loglik_bwd"] = ("hbm", n_rows * (E_x + sz * P_th + D + sz * P_th))   # upstream gradient is a device scalar
    # FP64 contraction flops: S = K^T V and W = V G (2 L N M^2 each) + B^-1 K and W V^T (2 L N T M each)
    ALGO["hlvae_kl_panel"] = ("tensor", 2.0 * L * n_rows * M * M * 2 + 2.0 * L * n_rows * T * M * 2)
    ALGO["hlvae_kl_subject"] = ("fp64", L * n_subj * (T ** 3 / 3 + T ** 3 / 3 + T ** 3 / 3 + 2 * 2.0 * T ** 3))


def run_gpu(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep this rank's threads (and therefore its pinned host buffers, first-touch) on the NUMA node next to
        # its GPU: with 8 ranks the e2e host->device copies otherwise cross sockets
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local]) if vis and vis.split(",")[0].isdigit() else local
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx))
        except Exception as e:
            print(f"[bench] rank {rank}: NUMA affinity not set ({e!r})", file=sys.stderr, flush=True)
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    from hlvae_b200 import _lib, config, parallel
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        parallel.enable()
    _lib.lib()
    config.check_errors = False              # keep the timed step free of host syncs
    s = build_gpu_state(dev, args.subjects, rank)
    # the KL branch is the critical path (2.2 of 2.8 ms): its stream gets the higher priority, so the block scheduler
    # places pending KL CTAs first and the likelihood CTAs fill what is left (measured: 2.85 -> 2.76 ms per step; the
    # opposite assignment - likelihood stream at high priority - costs 2.98 ms)
    s["side"] = None if args.no_overlap else torch.cuda.Stream(priority=0 if args.no_kl_priority else -1)
    s["split_backward"] = not args.no_split_backward
    s["side2"] = None if (args.no_overlap or args.no_natgrad_stream) else torch.cuda.Stream()
    config.overlap = not args.no_overlap
    n_rows = s["N_b"]
    algorithmic_work(n_rows, args.subjects)
    fp64_peak = measure_fp64_peak(dev)       # before any graph capture (uses the RNG)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from hlvae_b200.graph import StepGraph
    m0, H0 = s["m"].clone(), s["H"].clone()

    def reset_state():
        s["m"].copy_(m0)
        s["H"].copy_(H0)

    # ---- timed region: the step captured once into a CUDA graph and replayed (hlvae_b200.graph)
    warm = max(args.warmup, 3)
    use_graph = not args.eager
    graph = None
    if use_graph:
        try:
            graph = StepGraph(lambda: elbo_step(s, world), warmup=warm)
        except Exception as e:                    # report, then measure the eager path instead
            print(f"[bench] CUDA-graph capture failed ({e!r}); timing the eager path", file=sys.stderr, flush=True)
            use_graph = False
    run_step = (graph.replay if use_graph else (lambda: elbo_step(s, world)))
    for _ in range(warm):
        run_step()
    barrier()
    reset_state()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    e0.record()
    for i in range(args.steps):
        run_step()
        marks[i].record()
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    per_step = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(args.steps)]
    spread = dict(min=min(per_step), median=pystat.median(per_step), max=max(per_step))
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = world * (n_rows / (SUBJ_PER_RANK * T)) * 1e3 / ms_per_step

    # ---- the same step launched eagerly, every C-ABI call bracketed by CUDA events on its stream
    # (events cannot be read back from inside a replayed graph): per-kernel durations + eager step time
    reset_state()
    n_prof = max(3, min(args.steps, 20))
    side_saved, s["side"], config.overlap = s["side"], None, False      # one stream: kernels timed one at a time
    for _ in range(2):
        elbo_step(s, world)
    barrier()
    _lib.PROFILE = []
    launches0 = _lib.LAUNCHES
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(n_prof):
        elbo_step(s, world)
    p1.record()
    barrier()
    sampler.stop_flag = True
    eager_ms = p0.elapsed_time(p1) / n_prof
    prof = _lib.PROFILE
    _lib.PROFILE = None
    s["side"], config.overlap = side_saved, not args.no_overlap
    launches_per_step = (_lib.LAUNCHES - launches0) // n_prof
    launches = launches_per_step * args.steps

    # per-kernel device time inside the timed region
    per = {}
    for name, a, b in prof:
        per.setdefault(name, []).append(a.elapsed_time(b))
    kern = {k: dict(ms_avg=float(np.mean(v)), launches_per_step=len(v) / n_prof) for k, v in per.items()}
    hbm_peak, peak_src = measured_peaks()
    for k, d in kern.items():
        if k in ALGO:
            bound, work = ALGO[k]
            d["bound"] = bound
            if bound == "hbm":
                d["achieved"] = work / (d["ms_avg"] * 1e-3) / 1e9
                d["peak"], d["unit"] = hbm_peak, "GB/s"
            else:
                d["achieved"] = work / (d["ms_avg"] * 1e-3) / 1e12
                d["peak"], d["unit"] = fp64_peak, "TFLOP/s"
            d["frac"] = d["achieved"] / d["peak"]
        d["share_of_step"] = d["ms_avg"] * d["launches_per_step"] / ms_per_step
    top = max((k for k in kern if k in ALGO and ALGO[k][0] in ("hbm", "tensor")), key=lambda k: kern[k]["share_of_step"])
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(tp) and n_rows == SUBJ_PER_RANK * T:      # captured at this workload size only
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get(top), tj.get("source")
        for k_, d_ in kern.items():
            if k_ in tj:
                d_["traffic"] = tj[k_]
    roof = dict(kernel=top, bound=kern[top]["bound"], achieved=kern[top]["achieved"], peak=kern[top]["peak"],
                unit=kern[top]["unit"], frac=kern[top]["frac"], traffic=traffic, traffic_source=traffic_src,
                peak_source=(peak_src if kern[top]["bound"] == "hbm" else
                             "FP64 DGEMM 4096^3 measured in this run (cuBLAS; MEASURED_PEAKS.json has no FP64 entry)"),
                ms_avg=kern[top]["ms_avg"])

    # ---- end to end: every step's inputs travel from pinned host memory and its loss is read back.
    # Two device input sets: the copy stream fills one while the graph of the other runs.
    e2e = None
    if not args.no_e2e:
        reset_state()
        cpus_before = os.sched_getaffinity(0)
        numa = bind_to_gpu_numa_node(dev)          # pinned pages are placed where the allocating thread runs
        host = {k: s[k].detach().cpu().pin_memory() for k in INPUT_KEYS}
        os.sched_setaffinity(0, cpus_before)       # only the allocation is bound: the CPU legs keep every core
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        sets = []
        for _ in range(2):
            d = {k: torch.empty_like(s[k]) for k in ("data", "mask", "x")}
            d.update({k: torch.empty_like(s[k]).requires_grad_(True) for k in ("theta", "mu", "lv")})
            sets.append(d)
        copy_stream = torch.cuda.Stream()
        loss_host = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)]

        def upload(b):
            with torch.no_grad():
                for k in INPUT_KEYS:
                    sets[b][k].copy_(host[k], non_blocking=True)

        upload(0), upload(1)
        torch.cuda.synchronize()
        steps_e = []
        for b in range(2):
            fn = (lambda bb: (lambda: elbo_step(s, world, sets[bb])))(b)
            steps_e.append(StepGraph(fn, warmup=1).replay if use_graph else fn)
        n_e2e = max(4, min(args.steps, 20))
        copied = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        losses = []

        def e2e_loop(n):
            cur = torch.cuda.current_stream()
            with torch.cuda.stream(copy_stream):
                upload(0)
                copied[0].record(copy_stream)
            for i in range(n):
                b = i & 1
                cur.wait_event(copied[b])
                loss = steps_e[b]()
                loss_host[b].copy_(loss.reshape(1), non_blocking=True)
                done[b].record(cur)
                if i + 1 < n:
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(done[1 - b])       # that buffer's previous step has finished
                        upload(1 - b)
                        copied[1 - b].record(copy_stream)
                if i >= 1:                                            # read the previous step's loss (host side)
                    done[1 - b].synchronize()
                    losses.append(float(loss_host[1 - b]))
            done[(n - 1) & 1].synchronize()
            losses.append(float(loss_host[(n - 1) & 1]))

        e2e_loop(3)
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        x0.record()
        e2e_loop(n_e2e)
        x1.record()
        barrier()
        dt = max(time.perf_counter() - t0, x0.elapsed_time(x1) * 1e-3)
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = dict(value=world * (n_rows / (SUBJ_PER_RANK * T)) * n_e2e / dt, unit="steps/s", h2d_bytes_per_step=h2d,
                   d2h_bytes_per_step=8, steps=n_e2e, h2d_gb_per_s_per_rank=h2d * n_e2e / dt / 1e9, host_binding=numa,
                   bound="host -> device copies (PCIe 5 x16, ~55 GB/s per GPU; the ranks of one box share the host's "
                         "memory and PCIe root bandwidth): the step itself needs a third of this time",
                   how="pinned host -> device copy of data, mask, covariates, theta, mu, log_v every step on a copy "
                       "stream (double-buffered against the running step), loss copied back and read every step; the "
                       "step's other products (g_theta, g_mu, g_logv, parameter gradients, m, H) stay on the device "
                       "for the optimiser, as in training.py - and in training theta, mu, log_v would not come from "
                       "the host either (the NN trunk produces them on the device): this is the worst case for the path")

    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_arm(steps=10, warmup=1, budget_s=20.0, keep_first=True)
        cpu = dict(value=r["value"], unit="steps/s", cores=r["cores"], kind="port", sample=r["sample"],
                   note="bounded sample for the default run; `bench.py --impl reference` measures the full 16000-row step")
        config.check_errors = True
        parity = parity_vs_oracle(dev, r, dict(side=s["side"], side2=s["side2"], split_backward=s["split_backward"]))
        config.check_errors = False
        del r
    allreduce_us = None
    if world > 1:
        parity = parity_sharded(dev, rank, world)
        off = _lib.acc_layout(L, M, Q)
        allreduce_us = allreduce_latency_us(dev, off["total"] + 2 + L)
    others = None
    if rank == 0 and world == 1 and not args.no_sweep:
        others = other_configs(dev, s, fp64_peak, hbm_peak)

    if rank == 0:
        line = dict(metric="ELBO train steps/sec (fwd+bwd)", value=value, unit="steps/s", n_gpus=world,
                    steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f64 (KL bound) + f32 (likelihoods)", data="synthetic",
                    config=dict(workload=WORKLOAD, rows_per_rank=n_rows, global_rows_per_step=n_rows * world,
                                storage="theta, mu, log_v and outputs f32; data and mask u8; likelihood arithmetic f32 (SFU) with exact f64 argmax re-evaluation; KL stream and M x M stage f64",
                                l2="inputs larger than L2 (data + theta = 415 MB per step, 126 MB L2)",
                                parallelism=f"dp{world}: subjects sharded, one all-reduce of accumulators" if world > 1 else "single GPU"),
                    clocks=sampler.summary(), e2e=e2e, gpu_launches=launches, roofline=roof, cpu_baseline=cpu,
                    parity=parity, step_ms_spread=spread, allreduce_us=allreduce_us, other_configs=others,
                    kernels=kern, fp64_peak_tflops=fp64_peak, cuda_graph=use_graph, eager_ms_per_step=eager_ms,
                    overlap=not args.no_overlap, split_backward=not args.no_split_backward,
                    kl_stream_priority=not args.no_kl_priority,
                    kernel_timing="CUDA events around every C-ABI call in an eager, single-stream pass of the same "
                                  "step run right after the timed region (events are not readable inside a replayed "
                                  "graph; the timed region overlaps the KL branch with the likelihood kernels)")
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tearing down a NCCL communicator whose collectives were captured into still-live CUDA graphs can
        # block; the result line is out, so synchronise, meet the other ranks and leave.
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


def run_predict(args):
    """utils.batch_predict_varying_T (utils.py:99-191) at the configs[1] shape:
    16000 conditioning rows (800 subjects x T=20), 4000 test rows, L=32, M=64.  One call = one "step"."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    import __graft_entry__ as g
    g.build()
    from hlvae_b200 import kernels, likelihoods, predict, synth
    from oracle import hlvae_oracle as orc
    rng = np.random.default_rng(0)
    gen = torch.Generator().manual_seed(0)
    x, lens = synth.covariates(args.subjects, T, rng)
    pool, _ = synth.covariates(400, T, np.random.default_rng(1))
    z = synth.inducing_points(pool, L, M, np.random.default_rng(1))
    sel = torch.from_numpy(rng.choice(x.shape[0], x.shape[0] // 4, replace=False))
    test_x = x[sel].clone()
    test_x[:, 0] += 0.5
    mu = torch.randn(x.shape[0], L, generator=gen, dtype=torch.float64)
    k0, k1 = kernels.generate_kernel_batched(L, **synth.DEFAULT_KERNEL_ARGS)
    k0, k1 = k0.to(dev).double().eval(), k1.to(dev).double().eval()
    lik = likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=likelihoods.GreaterThan(1e-8))
    lik.noise = 1
    lik = lik.to(dev).double().eval()
    xd, xtd, mud, zd = x.to(dev), test_x.to(dev), mu.to(dev), z.to(dev)
    call = lambda: predict.batch_predict_varying_T(L, k0, k1, lik, xd, xtd, mud, zd, 2, EPS)
    for _ in range(max(args.warmup, 3)):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # CPU arm: the oracle restatement on a bounded sample (100 subjects, 500 test rows), scaled in rows
    ns = 100
    xs_, ls_ = synth.covariates(ns, T, np.random.default_rng(0))
    ts_ = xs_[:500].clone()
    ts_[:, 0] += 0.5
    spec0, spec1 = orc.compile_spec(**synth.DEFAULT_KERNEL_ARGS)
    p0, p1 = orc.KernelParams.default(spec0, L), orc.KernelParams.default(spec1, L)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    with torch.no_grad():
        orc.batch_predict(spec0, p0, spec1, p1, torch.ones(L, dtype=torch.float64), xs_, ts_, mu[:xs_.shape[0]], z,
                          orc.split_subjects_by_id(xs_, 2), 2, EPS)
    cpu_s = (time.perf_counter() - t0) * (args.subjects / ns)
    line = dict(metric="GP posterior-mean prediction calls/sec (utils.batch_predict_varying_T)", value=1e3 / ms,
                unit="calls/s", n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms,
                higher_is_better=True, dtype="f64", data="synthetic",
                config=dict(workload=f"SURVEY 8(f).1: {x.shape[0]} conditioning rows ({args.subjects} subjects x T={T}), "
                                     f"{test_x.shape[0]} test rows, L={L}, M={M}, default kernel"),
                cpu_baseline=dict(value=1.0 / cpu_s, unit="calls/s", cores=torch.get_num_threads(), kind="port",
                                  sample=f"{ns} subjects / 500 test rows, float64 oracle port, scaled x{args.subjects / ns:.0f} in rows"),
                checksum=float(out.abs().sum()))
    print(json.dumps(line), flush=True)


def run_theta(args):
    """HLVAE.theta_estimation (HLVAE.py:416-453) at the configs[1] batch: 16000 rows, D4
    (324 real + 972 cat x5 -> P_theta = 5184), y_dim = 5, convolutional layout (y_grouped = permuted view of
    [N, y_dim, D]), float32 storage, uint8 mask.  One forward + backward = one "step"; inputs (y 415 MB, upstream
    gradient 332 MB) exceed the 126 MB L2."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    import __graft_entry__ as g
    g.build()
    from hlvae_b200 import _lib, synth, theta as th
    from oracle import hlvae_oracle as orc
    hbm_peak, _ = measured_peaks()
    types = synth.HEALTHMNIST_D4_TYPES
    N, Y, D = args.subjects * T, 5, len(types)
    gen = torch.Generator(device=dev).manual_seed(0)
    lay = th.HeadLayout(types, True, dev)
    P = lay.P
    y = torch.randn(N, Y, D, generator=gen, device=dev, dtype=torch.float32).permute(0, 2, 1)
    mask = (torch.rand(N, D, generator=gen, device=dev) < 0.75).to(torch.uint8)
    W = torch.randn(P, Y, generator=gen, device=dev, dtype=torch.float64) * 0.3
    b = torch.randn(P, generator=gen, device=dev, dtype=torch.float64) * 0.3
    g_up = torch.randn(N, P, generator=gen, device=dev, dtype=torch.float32)
    yq = y.detach().requires_grad_(True)
    Wq, bq = W.requires_grad_(True), b.requires_grad_(True)
    _lib.PROFILE = []

    def step():
        yq.grad = Wq.grad = bq.grad = None
        theta = th.theta_heads(lay, yq, mask, Wq, bq)
        theta.backward(g_up)
        return theta

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    _lib.PROFILE.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tc0 = time.perf_counter()
    for _ in range(args.steps):
        out = step()
    enqueue_ms = (time.perf_counter() - tc0) * 1e3 / args.steps
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    per = {}
    for name, a, c in _lib.PROFILE:
        per.setdefault(name, []).append(a.elapsed_time(c))
    _lib.PROFILE = None
    bytes_fwd = N * (4 * D * Y + 4 * P)
    bytes_bwd = N * (4 * D * Y + D + 4 * P + 4 * D * Y)
    kern = {}
    for name, nbytes in (("hlvae_theta_fwd", bytes_fwd), ("hlvae_theta_bwd", bytes_bwd)):
        t_ms = float(np.mean(per[name]))
        kern[name] = dict(ms_avg=t_ms, bound="hbm", achieved=nbytes / t_ms / 1e6, peak=hbm_peak, unit="GB/s",
                          frac=nbytes / t_ms / 1e6 / hbm_peak, algorithmic_bytes=nbytes)
    # CPU arm: the oracle restatement of the reference method on a bounded row sample, scaled in rows
    ns = 400
    cg = torch.Generator().manual_seed(0)
    ti = orc.types_info_from_layout(types, conv=True)
    heads = []
    for i, tpl in enumerate(ti['set_of_types']):
        n, C = int((ti['data_types_indexes'] == i).sum()), int(tpl[1])
        r = lambda *s_: (torch.randn(*s_, generator=cg, dtype=torch.float64) * 0.3).requires_grad_(True)
        heads.append(dict(weight=r(n, Y, C - 1), bias=r(n, C - 1)) if tpl[0] == 'cat' else
                     dict(weight_mean=r(n, Y, 1), bias_mean=r(n, 1)))
    yc = torch.randn(ns, D, Y, generator=cg, dtype=torch.float64).requires_grad_(True)
    mc = (torch.rand(ns, D, generator=cg) < 0.75).double()
    gc = torch.randn(ns, P, generator=cg, dtype=torch.float64)
    torch.set_num_threads(os.cpu_count() or 1)
    orc.theta_estimation(types, heads, yc, mc, conv=True).backward(gc)
    t0 = time.perf_counter()
    reps_c = 3
    for _ in range(reps_c):
        orc.theta_estimation(types, heads, yc, mc, conv=True).backward(gc)
    cpu_s = (time.perf_counter() - t0) / reps_c * (N / ns)
    line = dict(metric="observation-head (theta_estimation) fwd+bwd passes/sec", value=1e3 / ms, unit="steps/s", n_gpus=1,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms, higher_is_better=True, dtype="f32",
                data="synthetic",
                config=dict(workload=f"SURVEY 8(f).2: {N} rows, D4 (324 real + 972 cat x5), y_dim={Y}, conv layout, "
                                     "float32 storage, uint8 mask", l2="inputs larger than L2 (y 415 MB + upstream 332 MB)"),
                gpu_launches=2 * args.steps, kernels=kern, host_enqueue_ms_per_step=enqueue_ms,
                roofline=dict(kernel="hlvae_theta_bwd", bound="hbm", achieved=kern["hlvae_theta_bwd"]["achieved"],
                              peak=hbm_peak, unit="GB/s", frac=kern["hlvae_theta_bwd"]["frac"], traffic=None),
                cpu_baseline=dict(value=1.0 / cpu_s, unit="steps/s", cores=torch.get_num_threads(), kind="port",
                                  sample=f"{ns} rows of the same workload, float64 oracle port, scaled x{N / ns:.0f} in rows"),
                checksum=float(out.double().abs().sum()))
    print(json.dumps(line), flush=True)


def run_full(args):
    """SURVEY.md 8(d) step rate (ii): the FULL training step of training.py:78-137 at the configs[1] batch - the
    stock-PyTorch NN trunk (convolutional encoder / decoder with the shapes of HLVAE.py:146-165, 244-275, hidden [500],
    y_dim 5; cuDNN / cuBLAS, float32 - the reference runs it in float64), then this repo's path: observation heads
    (hlvae_theta_*), fused likelihoods (hlvae_loglik_*), KL upper bound + natural-gradient update (hlvae_kl_*,
    hlvae_mxm_*), backward through everything, and one torch.optim.Adam step over trunk, heads, kernel
    hyper-parameters, inducing points and log-variances.  Eager launches, CUDA events."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    import __graft_entry__ as g
    g.build()
    import torch.nn as nn
    import torch.nn.functional as F
    from hlvae_b200 import _lib, config, elbo, loglik, synth, theta as th
    config.check_errors = False
    torch.manual_seed(0)
    s = build_gpu_state(dev, args.subjects, 0)
    lay, N_b, Y, H1 = s["lay"], s["N_b"], 5, 500
    hlay = th.HeadLayout(synth.HEALTHMNIST_D4_TYPES, True, dev)
    real_d, cat_d = lay.idx["real"], lay.idx["cat"]
    real_col = lay.var_dcol.long()[real_d]
    cat_col = (lay.var_dcol.long()[cat_d][:, None] + torch.arange(5, device=dev)[None, :]).reshape(-1)

    class Head(nn.Module):
        def __init__(self, **shapes):
            super().__init__()
            for n, shp in shapes.items():
                setattr(self, n, nn.Parameter(torch.randn(*shp, dtype=torch.float64) * 0.05))

    class Trunk(nn.Module):
        def __init__(self):
            super().__init__()
            self.rep_w = nn.Parameter(torch.randn(cat_d.numel(), 5) * 0.05)        # Representation_One_Hot, HLVAE.py:91-102
            self.rep_b = nn.Parameter(torch.randn(cat_d.numel()) * 0.05)
            self.conv1, self.conv2 = nn.Conv2d(1, 16, 3, 1, 1), nn.Conv2d(16, 32, 3, 1, 1)   # :146-152
            self.enc = nn.Linear(32 * 9 * 9, H1)
            self.mean, self.logvar = nn.Linear(H1, L), nn.Linear(H1, L)            # :167-177
            self.dec, self.y_layer = nn.Linear(L, H1), nn.Linear(H1, 32 * 9 * 9)   # :244-266
            self.deconv = nn.Sequential(nn.ConvTranspose2d(32, 16, 4, 2, 1), nn.ReLU(), nn.ConvTranspose2d(16, Y, 4, 2, 1))
            # obs_layer as HLVAE.py:276-299 builds it for the sorted type groups [('cat','5'), ('real','1')]
            self.obs_layer = nn.ModuleList([Head(weight=(cat_d.numel(), Y, 4), bias=(cat_d.numel(), 4)),
                                            Head(weight_mean=(real_d.numel(), Y, 1), bias_mean=(real_d.numel(), 1)),
                                            nn.Sigmoid()])

        def encode(self, data, mask):                                              # :302-335
            n = data.shape[0]
            img = torch.zeros(n, lay.D, device=dev)
            img[:, real_d] = data[:, real_col].float() / 255
            img[:, cat_d] = torch.einsum("bdc,dc->bd", data[:, cat_col].float().view(n, -1, 5), self.rep_w) + self.rep_b
            h = (img * mask.float()).view(n, 1, 36, 36)
            h = F.max_pool2d(F.relu(self.conv1(h)), 2)
            h = F.max_pool2d(F.relu(self.conv2(h)), 2).reshape(n, -1)
            h = F.relu(self.enc(h))
            return self.mean(h), torch.clamp(self.logvar(h), -15.0, 15.0)

        def decode(self, z):                                                       # :337-343
            yv = self.deconv(self.y_layer(F.relu(self.dec(z))).view(-1, 32, 9, 9))
            return yv.view(yv.shape[0], Y, -1).permute(0, 2, 1)                     # y_grouped: permuted view

    net = Trunk().to(dev)
    params = list(net.parameters()) + [s["z"], s["log_vy_real"]] + list(s["k0"].parameters()) + list(s["k1"].parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
    P_b = s["n_subj"]

    def step():
        opt.zero_grad(set_to_none=True)
        mu, lv = net.encode(s["data"], s["mask"])
        zs = mu + torch.randn_like(mu) * torch.exp(0.5 * lv)                        # HLVAE.py:360-365
        y = net.decode(zs)
        W, b = th.pack_heads(net.obs_layer, hlay, Y)
        theta = th.theta_heads(hlay, y, s["mask"], W, b)                            # HLVAE.py:344
        vparam = lay.vparam(log_vy_real=s["log_vy_real"], conv=True)
        out = loglik.fused_loglik(lay, s["data"], s["mask"], theta, vparam, monitor=True)
        nll = -out["log_p_x_sum"] * (P_TOTAL / P_b)
        kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(s["k0"], s["k1"], s["lik"], L, s["m"], s["H"], s["x"], mu, lv,
                                                          s["z"], P_TOTAL, P_b, N_TOTAL, True, 2, EPS, layout=s["layout"])
        loss = nll + kld
        loss.backward()
        opt.step()
        m_new, H_new = elbo.natural_gradient_update(s["m"], s["H"], gm, gH, NG_LR)
        s["m"].copy_(m_new)
        s["H"].copy_(H_new)
        return loss.detach()

    for _ in range(max(args.warmup, 3)):
        l0 = step()
    torch.cuda.synchronize()
    _lib.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tc0 = time.perf_counter()
    for _ in range(args.steps):
        l1 = step()
    enqueue_ms = (time.perf_counter() - tc0) * 1e3 / args.steps
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    per = {}
    for name, a, c in _lib.PROFILE:
        per[name] = per.get(name, 0.0) + a.elapsed_time(c) / args.steps
    _lib.PROFILE = None
    own = sum(per.values())
    line = dict(metric="full train steps/sec (NN trunk + ELBO path + Adam)", value=1e3 / ms, unit="steps/s", n_gpus=1,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms, higher_is_better=True,
                dtype="f32 trunk (stock PyTorch) + f64 KL + f32 likelihoods / heads", data="synthetic",
                config=dict(workload=WORKLOAD + "; SURVEY 8(d) step rate (ii): conv encoder/decoder trunk (hidden 500, y_dim 5), "
                                                "observation heads, fused likelihoods, KL + natural gradient, Adam; eager launches",
                            rows_per_step=N_b),
                host_enqueue_ms_per_step=enqueue_ms, own_kernels_ms_per_step=own, own_kernels_share=own / ms,
                kernels_ms={k: round(v, 4) for k, v in per.items()}, loss_first=float(l0), loss_last=float(l1))
    print(json.dumps(line), flush=True)


def run_norm(args):
    """HL_VAE.utils.batch_normalization (HL_VAE/utils.py:88-143) on (a) the configs[3]
    likelihood-heavy tabular batch (64 count + 64 ordinal(5) + 64 cat(5) + 32 real + 32 pos, 64 000 rows, float32 data,
    uint8 mask) and (b) the configs[1] batch (16 000 rows, D4, convolutional: uint8 data and mask).  Device time of
    the three launches and GB/s of algorithmic bytes (apply: data + mask in, X out; stats: the real / positive
    columns + their mask columns, twice)."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    import __graft_entry__ as g
    g.build()
    from hlvae_b200 import _lib, normalize as nz, synth
    hbm_peak, _ = measured_peaks()
    rows = []
    for tag, types, conv, N, ddt in (("configs[3] tabular", synth.TABULAR_TYPES, False, 64000, torch.float32),
                                     ("configs[1] D4 conv", synth.HEALTHMNIST_D4_TYPES, True, 16000, torch.uint8)):
        lay = nz.NormLayout(types, conv, dev)
        gen = torch.Generator(device=dev).manual_seed(0)
        data, mask = synth.device_likelihood_batch(lay.var, N, dev, gen, dtype=ddt, observed=0.7, pixel_like=conv)
        for _ in range(max(args.warmup, 3)):
            nz.normalize(lay, data, mask)
        torch.cuda.synchronize()
        _lib.PROFILE = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            X, mean, var = nz.normalize(lay, data, mask)
        e1.record()
        torch.cuda.synchronize()
        per = {}
        for name, a, c in _lib.PROFILE:
            per.setdefault(name, []).append(a.elapsed_time(c))
        _lib.PROFILE = None
        dsz = data.element_size()
        apply_bytes = N * (lay.var.E_x * dsz + lay.var.D + lay.var.E_x * 4)
        stats_bytes = 2 * N * lay.n_stat * (dsz + 1)
        t_apply = float(np.mean(per["hlvae_batch_norm_apply"]))
        t_stats = float(np.sum(per.get("hlvae_batch_norm_stats", [0.0]))) / args.steps
        rows.append(dict(case=tag, rows=N, E_x=lay.var.E_x, stat_vars=lay.n_stat, call_ms=e0.elapsed_time(e1) / args.steps,
                         apply_ms=t_apply, apply_gbs=apply_bytes / t_apply / 1e6, apply_frac=apply_bytes / t_apply / 1e6 / hbm_peak,
                         stats_ms_both_passes=t_stats, stats_gbs=(stats_bytes / t_stats / 1e6) if t_stats else None,
                         checksum=float(torch.nan_to_num(X.double()).abs().sum())))
    print(json.dumps(dict(workload="SURVEY 8(f).4: HL_VAE.utils.batch_normalization (hlvae_batch_norm_stats x2 + hlvae_batch_norm_apply)",
                          hbm_peak_gbs=hbm_peak, steps=args.steps, cases=rows)), flush=True)


def _timed_profile(fn, reps):
    from hlvae_b200 import _lib
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    _lib.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    per = {}
    for name, a, b in _lib.PROFILE:
        per.setdefault(name, []).append(a.elapsed_time(b))
    _lib.PROFILE = None
    return e0.elapsed_time(e1) / reps, {k: float(np.mean(v)) for k, v in per.items()}


def sweep_kl(dev, cases, reps, fp64_peak=None):
    """BASELINE.json configs[2]: SE(time) + CA(id) + SE(age) x CA(sex), continuous ages, L = 32, T = 20; one row per
    (M, subjects, ragged): device time of one KL forward + backward (eager, single stream) and of its kernels."""
    from hlvae_b200 import elbo, kernels, likelihoods, subjects, synth
    rows = []
    for Mx, n_subj, ragged in cases:
        rng = np.random.default_rng(7)
        gen = torch.Generator().manual_seed(7)
        x, lens = synth.covariates(n_subj, T, rng, ragged=ragged, t_min=5, continuous_age=True)
        pool, _ = synth.covariates(400, T, np.random.default_rng(8), continuous_age=True)
        z = synth.inducing_points(pool, L, Mx, np.random.default_rng(8)).to(dev).requires_grad_(True)
        m, H = synth.variational_state(L, Mx, gen)
        k0, k1 = kernels.generate_kernel_batched(L, **synth.SWEEP_KERNEL_ARGS)
        k0, k1 = k0.to(dev).double(), k1.to(dev).double()
        lik = likelihoods.GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=likelihoods.GreaterThan(1e-8))
        lik.noise = 1
        lik = lik.to(dev).double()
        N_b = x.shape[0]
        mu = torch.randn(N_b, L, generator=gen).to(dev).requires_grad_(True)
        lv = (-3.0 * torch.rand(N_b, L, generator=gen)).to(dev).requires_grad_(True)
        lay = subjects.SubjectLayout.from_lengths(lens, dev)
        xd, md, Hd = x.to(dev), m.to(dev), H.to(dev)

        def step():
            for t_ in (mu, lv, z, *k0.parameters(), *k1.parameters()):
                t_.grad = None
            kld, gm, gH = elbo.minibatch_KLD_upper_bound_iter(k0, k1, lik, L, md, Hd, xd, mu, lv, z, P_TOTAL, n_subj,
                                                              N_TOTAL, True, 2, EPS, layout=lay)
            kld.sum().backward()

        ms, per = _timed_profile(step, reps)
        stream = per.get("hlvae_kl_subject", 0.0) + per.get("hlvae_kl_panel", 0.0)
        flops = 2.0 * L * N_b * Mx * Mx * 2 + 2.0 * L * N_b * T * Mx * 2
        row = dict(M=Mx, rows=N_b, subjects=n_subj, ragged=ragged, kl_fwd_bwd_ms=round(ms, 3),
                   kl_subject_ms=round(per.get("hlvae_kl_subject", 0.0), 3),
                   kl_panel_ms=round(per.get("hlvae_kl_panel", 0.0), 3),
                   panel_tflops=round(flops / (per["hlvae_kl_panel"] * 1e-3) / 1e12, 2),
                   rows_per_s=round(N_b / (stream * 1e-3)))
        if fp64_peak:
            row["panel_frac_fp64_peak"] = round(row["panel_tflops"] / fp64_peak, 3)
        rows.append(row)
    return rows


def sweep_tabular(dev, sizes, reps, hbm_peak):
    """BASELINE.json configs[3]: 64 count + 64 ordinal(5) + 64 cat(5) + 32 real + 32 pos, 30 % missing, float32
    storage: fused likelihood forward / backward device time and GB/s of algorithmic bytes."""
    from hlvae_b200 import loglik, synth
    rows = []
    types = synth.TABULAR_TYPES
    layt = loglik.VarLayout(types, dev)
    E_x, P_th = layt.E_x, layt.P_theta
    for N_b in sizes:
        rng = np.random.default_rng(9)
        gen = torch.Generator(device=dev).manual_seed(9)
        data, mask = synth.likelihood_batch(types, 4000, rng)
        reps_ = N_b // 4000
        data = data.float().repeat(reps_, 1).to(dev)
        mask = mask.to(torch.uint8).repeat(reps_, 1).to(dev)
        theta = torch.randn(N_b, P_th, device=dev, generator=gen).requires_grad_(True)
        z32 = torch.zeros(32, dtype=torch.float64, device=dev)
        lvr, lvp = z32.clone().requires_grad_(True), z32.clone().requires_grad_(True)

        def step():
            theta.grad = lvr.grad = lvp.grad = None
            vparam = layt.vparam(lvr, lvp, [z32, torch.ones_like(z32)], [z32, torch.ones_like(z32)])
            out = loglik.fused_loglik(layt, data, mask, theta, vparam, monitor=True)
            (-out["log_p_x_sum"]).backward()

        ms, per = _timed_profile(step, reps)
        D = len(types)
        bf = N_b * (4 * E_x + 4 * P_th + D + 4 * (5 * D + P_th))
        bb = N_b * (4 * E_x + 4 * P_th + D + 4 * P_th)
        rows.append(dict(rows=N_b, D=D, fwd_ms=round(per["hlvae_loglik_fwd"], 3), bwd_ms=round(per["hlvae_loglik_bwd"], 3),
                         fwd_gbs=round(bf / (per["hlvae_loglik_fwd"] * 1e-3) / 1e9), fwd_frac=round(bf / (per["hlvae_loglik_fwd"] * 1e-3) / 1e9 / hbm_peak, 3),
                         bwd_gbs=round(bb / (per["hlvae_loglik_bwd"] * 1e-3) / 1e9), bwd_frac=round(bb / (per["hlvae_loglik_bwd"] * 1e-3) / 1e9 / hbm_peak, 3)))
    return rows


def run_sweep(args):
    """BASELINE.json configs[2] and configs[3] as tables (not the headline line): the kernel sweep at M in
    {32, 64, 128} x {4k, 16k, 16k ragged, 64k} rows and the tabular likelihood batch at 16k / 64k rows."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    import __graft_entry__ as g
    g.build()
    from hlvae_b200 import config
    config.check_errors = False
    config.overlap = False
    hbm_peak, _ = measured_peaks()
    reps = max(args.steps, 5)
    cases = [(Mx, n, r) for Mx in (32, 64, 128) for n, r in ((200, False), (800, False), (800, True), (3200, False))]
    kl_rows = sweep_kl(dev, cases, reps)
    ll_rows = sweep_tabular(dev, (16000, 64000), reps, hbm_peak)
    print(json.dumps(dict(workload="sweep: BASELINE.json configs[2] (kernel sweep, L=32, T=20, SE(time)+CA(id)+SE(age)xCA(sex)) "
                                   "and configs[3] (tabular likelihoods, 30 % missing, f32 storage)",
                          timing="CUDA events, eager single-stream launches, 3 warm-ups", kernel_sweep=kl_rows,
                          tabular_loglik=ll_rows, hbm_peak_gbs=hbm_peak)), flush=True)


def sweep_theta(dev, N, reps, hbm_peak):
    """SURVEY 8(f) row 2 in the default line: the observation-head kernels (hlvae_theta_fwd / _bwd) at the configs[1]
    batch (D4, y_dim 5, conv layout, float32 storage, uint8 mask) - same set-up as `--workload theta`."""
    from hlvae_b200 import _lib, synth, theta as th
    types = synth.HEALTHMNIST_D4_TYPES
    Y, D = 5, len(types)
    gen = torch.Generator(device=dev).manual_seed(0)
    lay = th.HeadLayout(types, True, dev)
    P = lay.P
    y = torch.randn(N, Y, D, generator=gen, device=dev, dtype=torch.float32).permute(0, 2, 1).requires_grad_(True)
    mask = (torch.rand(N, D, generator=gen, device=dev) < 0.75).to(torch.uint8)
    W = (torch.randn(P, Y, generator=gen, device=dev, dtype=torch.float64) * 0.3).requires_grad_(True)
    b = (torch.randn(P, generator=gen, device=dev, dtype=torch.float64) * 0.3).requires_grad_(True)
    g_up = torch.randn(N, P, generator=gen, device=dev, dtype=torch.float32)

    def step():
        y.grad = W.grad = b.grad = None
        th.theta_heads(lay, y, mask, W, b).backward(g_up)

    _, per = _timed_profile(step, reps)
    bytes_fwd = N * (4 * D * Y + 4 * P)
    bytes_bwd = N * (4 * D * Y + D + 4 * P + 4 * D * Y)
    f, bw = per["hlvae_theta_fwd"], per["hlvae_theta_bwd"]
    return dict(rows=N, fwd_ms=round(f, 3), bwd_ms=round(bw, 3), fwd_gbs=round(bytes_fwd / f / 1e6),
                fwd_frac=round(bytes_fwd / f / 1e6 / hbm_peak, 3), bwd_gbs=round(bytes_bwd / bw / 1e6),
                bwd_frac=round(bytes_bwd / bw / 1e6 / hbm_peak, 3))


def other_configs(dev, s_small, fp64_peak, hbm_peak):
    """Compact, driver-visible figures for the BASELINE.json configurations that are not the headline line
    (configs[2], [3], [4]) and for float64 storage; eager single-stream launches, CUDA events, 5 repetitions."""
    from hlvae_b200 import config
    ov, config.overlap = config.overlap, False
    out = {}
    out["configs2_kernel_sweep_16k_rows"] = sweep_kl(dev, [(32, 800, False), (64, 800, False), (128, 800, False)], 5,
                                                     fp64_peak)
    out["configs3_tabular_64k_rows"] = sweep_tabular(dev, (64000,), 5, hbm_peak)[0]
    out["survey_8f_observation_heads_16k_rows"] = sweep_theta(dev, SUBJ_PER_RANK * T, 5, hbm_peak)
    torch.cuda.empty_cache()

    def eager_ms(st, reps=5):
        for _ in range(2):
            elbo_step(st, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            elbo_step(st, 1)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    big = build_gpu_state(dev, 3200, 0)
    big.update(side=None, side2=None)
    ms_big = eager_ms(big)
    out["configs4_64000_rows_per_rank_step"] = dict(ms_per_step=round(ms_big, 3), rows=big["N_b"],
                                                    rows_per_s=round(big["N_b"] / (ms_big * 1e-3)),
                                                    step_equivalents_per_s=round(big["N_b"] / (SUBJ_PER_RANK * T) * 1e3 / ms_big, 1))
    del big
    torch.cuda.empty_cache()
    f64 = dict(s_small)
    f64.update(side=None, side2=None)
    for k in ("theta", "mu", "lv"):
        f64[k] = s_small[k].detach().double().requires_grad_(True)
    for k in ("data", "mask"):
        f64[k] = s_small[k].double()
    f64["m"], f64["H"] = s_small["m"].clone(), s_small["H"].clone()
    ms64 = eager_ms(f64)
    out["float64_storage_step"] = dict(ms_per_step=round(ms64, 3), steps_per_s=round(1e3 / ms64, 1),
                                       note="configs[1] with theta, mu, log_v, data, mask and every output float64 "
                                            "(the reference's storage); eager single-stream launches")
    config.overlap = ov
    return out


def run_contraction(args):
    """A/B of the sufficient-statistics contraction S = K^T V (elbo_functions.py:161 / :254,266) on the two tensor
    pipes at the configs[1] shape (L = 32, M = 64, 16000 rows): FP64 mma.sync against tcgen05.mma kind::i8 with an
    error-free 8-bit splitting (csrc/contraction_probe.cu).  Device time (CUDA events), error against the float64
    product, algorithmic and executed operations."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    import math
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    import __graft_entry__ as g
    g.build()
    from hlvae_b200 import _lib
    N = args.subjects * T
    gen = torch.Generator(device=dev).manual_seed(0)
    K = torch.rand(L, N, M, generator=gen, device=dev, dtype=torch.float64) * 2.0
    V = torch.randn(L, N, M, generator=gen, device=dev, dtype=torch.float64)
    ref = K.transpose(1, 2) @ V
    scale = float((K.abs().transpose(1, 2) @ V.abs()).max())
    ks = 2.0 ** math.ceil(math.log2(float(K.max()) * (1 + 1e-12)))
    vs = 2.0 ** math.ceil(math.log2(float(V.abs().max()) * (1 + 1e-12)))
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    rows = []
    for mode, ns, rpc in [(0, 0, 1920), (0, 0, 960)] + [(1, ns, rpc) for ns in (4, 5, 6, 7) for rpc in (1920, 960)]:
        S = torch.zeros(L, M, M, dtype=torch.float64, device=dev)

        def call():
            _lib.call("hlvae_contraction_probe", mode, ns, L, N, M, _lib.ptr(K), _lib.ptr(V), ks, vs, rpc, _lib.ptr(S),
                      _lib.ptr(status), _lib.stream_ptr())

        call()
        torch.cuda.synchronize()
        err = float((S - ref).abs().max()) / scale
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        algo = 2.0 * L * N * M * M
        pairs = sum(1 for g_ in range(ns) for p in range((ns + 1) // 2) if 0 <= g_ - 2 * p < ns) if mode else 1
        rows.append(dict(pipe="fp64 mma.sync" if mode == 0 else f"tcgen05 kind::i8, {ns} slices", rows_per_cta=rpc,
                         ms=round(ms, 4), rel_err=err, algorithmic_tflops=round(algo / (ms * 1e-3) / 1e12, 2),
                         executed_mma_ops=(2.0 * L * N * 128 * M * pairs) if mode else algo,
                         hbm_gbs=round(2 * L * N * M * 8 / (ms * 1e-3) / 1e9)))
    assert int(status[0]) == 0, status.tolist()
    print(json.dumps(dict(workload=f"contraction A/B: S = K^T V, L={L}, M={M}, {N} rows (operands read from HBM: "
                                   f"{2 * L * N * M * 8 / 1e6:.0f} MB)", steps=args.steps, cases=rows)), flush=True)


def main():
    args = parse()
    if args.workload == "contraction" and args.impl != "reference":
        run_contraction(args)
        return
    if args.workload == "predict" and args.impl != "reference":
        run_predict(args)
        return
    if args.workload == "norm" and args.impl != "reference":
        run_norm(args)
        return
    if args.workload == "full" and args.impl != "reference":
        run_full(args)
        return
    if args.workload == "theta" and args.impl != "reference":
        run_theta(args)
        return
    if args.workload == "sweep" and args.impl != "reference":
        run_sweep(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
