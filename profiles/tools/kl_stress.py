"""Randomised parity sweep of the KL path against the float64 oracle: random latent / inducing-point / subject counts,
subject lengths up to HLVAE_TMAX, ragged or fixed, the three kernel specifications of synth.py, float64 and float32
storage, random panel shapes.  Prints one line per case and a summary; exit code 1 on any failure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import helpers as h
from oracle import hlvae_oracle as orc
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(os.environ.get("SEED", "0")))
n_cases = int(os.environ.get("CASES", "40"))
bad = 0


def random_kargs(rng):
    """Random subsets of every component family of kernel_gen.generate_kernel_batched (columns: 0 time, 1 age, 2 id,
    3 categorical, 4 / 5 binary; age optionally masked by column 4), at least one component on either side of the
    id split, at most 8 per side."""
    sq = [c for c in (0, 1) if rng.integers(0, 2)]
    cat = [c for c in (2, 3) if rng.integers(0, 2)]
    bn = [c for c in (4, 5) if rng.integers(0, 2)]
    ci = [{'cont_covariate': int(rng.integers(0, 2)), 'cat_covariate': int(rng.choice([2, 3]))}
          for _ in range(int(rng.integers(0, 4)))]
    bi = [{'cont_covariate': int(rng.integers(0, 2)), 'bin_covariate': int(rng.choice([4, 5]))}
          for _ in range(int(rng.integers(0, 3)))]
    miss = [{'covariate': 1, 'mask': 4}] if rng.integers(0, 2) else []
    if 2 not in cat and not any(d['cat_covariate'] == 2 for d in ci):
        cat = [2] + cat
    if not (sq or bn or bi or [c for c in cat if c != 2] or [d for d in ci if d['cat_covariate'] != 2]):
        sq = [0]
    return dict(cat_kernel=cat, bin_kernel=bn, sqexp_kernel=sq, cat_int_kernel=ci, bin_int_kernel=bi,
                covariate_missing_val=miss, id_covariate=2)


for case in range(n_cases):
    L = int(rng.integers(1, 7))
    M = int(rng.choice([5, 8, 16, 17, 31, 32, 33, 40, 64, 65, 96, 120, 128]))
    T = int(rng.choice([1, 2, 3, 5, 8, 9, 16, 20, 24, 25, 32, 33, 40, 47, 64]))
    ragged = bool(rng.integers(0, 2)) and T >= 4
    n_subj = int(rng.integers(1, max(2, min(60, 1200 // T))))
    kargs = [h.synth.DEFAULT_KERNEL_ARGS, h.synth.SWEEP_KERNEL_ARGS, h.synth.MASKED_KERNEL_ARGS, None][int(rng.integers(0, 4))]
    if kargs is None:                                      # a random additive structure over the six synthetic covariates
        while True:
            kargs = random_kargs(rng)
            if all(len(sp.comps) <= 8 for sp in orc.compile_spec(**kargs)):      # HLVAE_MAX_COMPS per side
                break
    storage = torch.float64 if rng.integers(0, 2) else torch.float32
    rp_choices = [None]
    if M <= 64 and M > 32 and T <= 40: rp_choices += ["40", "64"]
    if M > 64: rp_choices += [r for r in ("32", "48", "64") if int(r) >= T]
    rp = rp_choices[int(rng.integers(0, len(rp_choices)))]
    if rp is None: os.environ.pop("HLVAE_PANEL_RP", None)
    else: os.environ["HLVAE_PANEL_RP"] = rp
    waves = [None, "0", "1", "16"][int(rng.integers(0, 4))]
    if waves is None: os.environ.pop("HLVAE_PANEL_WAVES", None)
    else: os.environ["HLVAE_PANEL_WAVES"] = waves
    tol = 1e-4 if storage == torch.float32 else 2e-5
    try:
        errs = h.check_kl_vs_oracle(dev, L, M, n_subj, T, seed=1000 + case, tol=tol, hyper_tol=2e-3, ragged=ragged,
                                    storage=storage, kargs=kargs, continuous_age=kargs is h.synth.SWEEP_KERNEL_ARGS)
        spec = 'fixed' if kargs in (h.synth.DEFAULT_KERNEL_ARGS, h.synth.SWEEP_KERNEL_ARGS, h.synth.MASKED_KERNEL_ARGS) else str({k: v for k, v in kargs.items() if v and k != 'id_covariate'})
        worst = max(errs.items(), key=lambda kv: kv[1] if "elem" not in kv[0] else 0.0)
        print(f"ok   L={L} M={M} subj={n_subj} T={T} ragged={ragged} rp={rp} waves={waves} {str(storage)[6:]} worst {worst[0]} {worst[1]:.1e} spec {spec}", flush=True)
    except Exception as e:
        bad += 1
        print(f"FAIL L={L} M={M} subj={n_subj} T={T} ragged={ragged} rp={rp} waves={waves} {str(storage)[6:]} {kargs if kargs not in (h.synth.DEFAULT_KERNEL_ARGS, h.synth.SWEEP_KERNEL_ARGS, h.synth.MASKED_KERNEL_ARGS) else 'fixed'}: {str(e)[:300]}", flush=True)
print("stress:", "OK" if bad == 0 else f"{bad} failures", "of", n_cases)
sys.exit(1 if bad else 0)
