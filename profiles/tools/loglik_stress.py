"""Randomised parity sweep of the fused likelihood kernels (float32 storage, the benchmarked arithmetic) against the
float64 oracle: random variable layouts (all five types, 2..16 classes, runs that make warps of one or of several
types), row counts that are not multiples of the batch size, uint8 or float data / mask, missing rates."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import helpers as h
import test_gpu_loglik as tl
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(os.environ.get("SEED", "0")))
n_cases = int(os.environ.get("CASES", "30"))
kinds = ['real', 'pos', 'count', 'cat', 'ordinal']
bad = 0
for case in range(n_cases):
    D = int(rng.choice([1, 3, 17, 64, 129, 200, 300]))
    types = []
    while len(types) < D:                                  # runs of one type, of random length
        k = kinds[int(rng.integers(0, 5))]
        run = int(rng.integers(1, 40))
        C = int(rng.integers(2, 17)) if k in ('cat', 'ordinal') else 1
        types += [(k, C)] * min(run, D - len(types))
    N = int(rng.choice([1, 7, 8, 33, 250, 1001]))
    # uint8 storage is exact for one-hot / thermometer codes and pixels only: use it when no real / pos / count column
    u8 = bool(rng.integers(0, 2)) and all(k in ('cat', 'ordinal') for k, _ in types)
    observed = float(rng.choice([0.3, 0.7, 1.0]))
    scale = float(rng.choice([0.5, 1.5, 4.0]))
    f64 = bool(rng.integers(0, 3) == 0)                    # the drop-in case: every tensor float64
    try:
        got, ref, disc = tl._fp32_case(types, N, 500 + case, dev, observed=observed, u8=u8, theta_scale=scale,
                                       storage=torch.float64 if f64 else torch.float32)
        if not all(bool(torch.isfinite(ref[k]).all()) for k in ("log_p_x", "params", "d_theta")):
            print(f"skip D={D} N={N}: the oracle itself is not finite here (normalisation of a column without two "
                  "observed values)", flush=True)
            continue
        errs = {k: h.rel_err(got[k], ref[k]) for k in ("log_p_x", "log_p_x_missing", "params", "d_theta")}
        ok = all(v < (1e-9 if f64 else 2e-5) for v in errs.values())
        ok = ok and np.array_equal(got["recon_mean"].cpu().numpy()[:, disc], ref["recon_mean"].numpy()[:, disc].astype(np.float64 if f64 else np.float32))
        ok = ok and np.array_equal(got["data_transformed"].cpu().numpy(), ref["data_transformed"].numpy().astype(np.float64 if f64 else np.float32))
        if not ok:
            raise AssertionError(str(errs))
        print(f"ok   D={D} N={N} u8={u8} f64={f64} obs={observed} scale={scale} worst {max(errs.values()):.1e}", flush=True)
    except Exception as e:
        bad += 1
        print(f"FAIL D={D} N={N} u8={u8} f64={f64} obs={observed} scale={scale}: {str(e)[:300]}", flush=True)
print("stress:", "OK" if bad == 0 else f"{bad} failures", "of", n_cases)
sys.exit(1 if bad else 0)
