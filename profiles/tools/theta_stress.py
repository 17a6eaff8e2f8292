"""Randomised parity sweep of the observation-head kernels (hlvae_theta_fwd / _bwd) against the float64 oracle:
random variable layouts (runs of every type, 2..16 classes), code widths 1..16, row counts that are not multiples of
the row batch, dense and permuted (convolutional) y, float64 and float32 storage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import test_gpu_theta as tt
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(os.environ.get("SEED", "0")))
n_cases = int(os.environ.get("CASES", "40"))
kinds = ['real', 'pos', 'count', 'cat', 'ordinal']
bad = 0
for case in range(n_cases):
    D = int(rng.choice([1, 2, 5, 17, 64, 129, 200]))
    types = []
    while len(types) < D:
        k = kinds[int(rng.integers(0, 5))]
        run = int(rng.integers(1, 40))
        C = int(rng.integers(2, 17)) if k in ('cat', 'ordinal') else 1
        types += [(k, C)] * min(run, D - len(types))
    N = int(rng.choice([1, 7, 8, 33, 250, 1001]))
    Y = int(rng.choice([1, 2, 3, 5, 8, 11, 16]))
    conv = bool(rng.integers(0, 2))
    layout = "permuted" if (conv and rng.integers(0, 2)) else "dense"
    f32 = bool(rng.integers(0, 2))
    observed = float(rng.choice([0.0, 0.3, 0.7, 1.0]))
    scale = float(rng.choice([0.5, 1.5, 4.0]))
    per_column = bool(rng.integers(0, 2))
    tag = f"D={D} N={N} Y={Y} conv={conv} {layout} f32={f32} obs={observed} scale={scale} per_column={per_column}"
    try:
        errs = tt._random_case(types, conv, N, Y, layout, dev, 900 + case, storage=torch.float32 if f32 else torch.float64,
                               observed=observed, scale=scale, per_column=per_column)
        tol = 1e-5 if f32 else 1e-11
        worst = {k: v for k, v in errs.items() if not v < tol}
        if worst:
            raise AssertionError(str(worst))
        print(f"ok   {tag} worst {max(errs.values()):.1e}", flush=True)
    except Exception as e:
        bad += 1
        print(f"FAIL {tag}: {type(e).__name__} {str(e)[:300]}", flush=True)
print("stress:", "OK" if bad == 0 else f"{bad} failures", "of", n_cases)
sys.exit(1 if bad else 0)
