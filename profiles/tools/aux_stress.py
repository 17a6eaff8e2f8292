"""Randomised parity sweeps of the remaining SURVEY 8(f) rows against the float64 oracle:
  norm     hlvae_batch_norm_stats / _apply: random variable layouts, row counts, missing rates, float64 / float32 / uint8
  predict  batch_predict_varying_T: random L, M, subject counts and lengths (to 64 rows), the three kernel specifications
  dubo     validation_dubo: the same with fixed-length subjects
Prints one line per case; exit code 1 on any failure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import helpers as h
from hlvae_b200 import normalize as nz, predict, synth, validation
from oracle import hlvae_oracle as orc
dev = torch.device("cuda:0")


def kl_random_kargs(rng):
    """The random kernel structures of kl_stress.py (that file runs its sweep on import, so the generator is read out of
    its source)."""
    src = open(os.path.join(ROOT, "profiles", "tools", "kl_stress.py")).read()
    ns = {}
    exec(src[src.index("def random_kargs"):src.index("for case in range(n_cases):")], {"np": np}, ns)
    return ns["random_kargs"](rng)


rng = np.random.default_rng(int(os.environ.get("SEED", "0")))
n_cases = int(os.environ.get("CASES", "30"))
which = os.environ.get("WHICH", "norm,predict,dubo").split(",")
kinds = ['real', 'pos', 'count', 'cat', 'ordinal']
bad = 0


def report(tag, fn):
    global bad
    try:
        print(f"ok   {tag} worst {fn():.1e}", flush=True)
    except Exception as e:
        bad += 1
        print(f"FAIL {tag}: {type(e).__name__} {str(e)[:300]}", flush=True)


def norm_case(types, conv, N, observed, seed, f32):
    data, mask = synth.likelihood_batch(types, N, np.random.default_rng(seed), observed=observed, pixel_like=conv)
    descs, _, _ = orc.build_layout(types)
    Xo, nr, npos = orc.batch_normalization(descs, data, mask, conv)
    if not bool(torch.isfinite(Xo).all()):
        return float("nan")                                # a column without two observed values: the reference divides by 0
    lay = nz.NormLayout(types, conv, dev)
    d = data.to(dev)
    X, mean, var = nz.normalize(lay, d.float() if f32 else d, mask.to(dev).to(torch.uint8) if f32 else mask.to(dev))
    tol = 2e-6 if f32 else 1e-11
    errs = [h.rel_err(X, Xo)]
    for tagk, ref in (("real", nr), ("pos", npos)):
        if ref is not None:
            idx = lay.var.idx[tagk]
            errs += [h.rel_err(mean[idx], ref[0]), h.rel_err(var[idx], ref[1])]
    assert max(errs) < tol, errs
    return max(errs)


def gp_case(kind, L, M, n_subj, T, ragged, kargs, seed):
    inp = h.make_kl_inputs(L, M, n_subj, T, seed=seed, ragged=ragged, kargs=kargs,
                           continuous_age=kargs is synth.SWEEP_KERNEL_ARGS)
    k0, k1, lik = h.build_product_kernels(inp["kargs"], L, dev, inp["ros0"], inp["rls0"], inp["ros1"], inp["rls1"],
                                          inp["noise"])
    spec0, spec1 = orc.compile_spec(**inp["kargs"])
    prm0, prm1 = orc.KernelParams(inp["ros0"], inp["rls0"]), orc.KernelParams(inp["ros1"], inp["rls1"])
    x = inp["x"]
    idc = inp["kargs"]["id_covariate"]
    if kind == "predict":
        r = np.random.default_rng(seed)
        sel = torch.from_numpy(r.choice(x.shape[0], min(x.shape[0], 23), replace=False))
        test_x = x[sel].clone()
        test_x[:, 0] += 0.25
        unseen, _ = synth.covariates(2, max(T, 4), r, first_id=99_000)
        test_x = torch.cat([test_x, unseen[:5]])
        ref = orc.batch_predict(spec0, prm0, spec1, prm1, inp["noise"], x, test_x, inp["mu"], inp["z"],
                                orc.split_subjects_by_id(x, idc), idc, 1e-6)
        got = predict.batch_predict_varying_T(L, k0.eval(), k1.eval(), lik.eval(), x.to(dev), test_x.to(dev),
                                              inp["mu"].to(dev), inp["z"].to(dev), idc, 1e-6)
    else:
        ref = orc.validation_dubo(spec0, prm0, spec1, prm1, inp["noise"], x, inp["mu"], inp["lv"], inp["z"], n_subj, T,
                                  1e-6)
        got = validation.validation_dubo(L, k0.eval(), k1.eval(), lik.eval(), x.to(dev), inp["mu"].to(dev),
                                         inp["lv"].to(dev), inp["z"].to(dev), n_subj, T, 1e-6)
    e = h.rel_err(got, ref)
    assert e < 2e-5 and bool(torch.isfinite(got).all()), e
    return e


for case in range(n_cases):
    if "norm" in which:
        D = int(rng.choice([1, 3, 17, 64, 129, 300]))
        types = []
        while len(types) < D:
            k = kinds[int(rng.integers(0, 5))]
            C = int(rng.integers(2, 17)) if k in ('cat', 'ordinal') else 1
            types += [(k, C)] * min(int(rng.integers(1, 40)), D - len(types))
        N = int(rng.choice([2, 7, 33, 250, 1001, 5000]))
        conv, f32 = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        observed = float(rng.choice([0.3, 0.7, 1.0]))
        report(f"norm D={D} N={N} conv={conv} f32={f32} obs={observed}",
               lambda: norm_case(types, conv, N, observed, 300 + case, f32))
    for kind in ("predict", "dubo"):
        if kind not in which:
            continue
        L = int(rng.integers(1, 7))
        M = int(rng.choice([5, 8, 17, 32, 33, 64, 65, 96, 120, 128]))
        T = int(rng.choice([1, 2, 3, 5, 9, 16, 20, 25, 32, 33, 47, 64]))
        ragged = kind == "predict" and bool(rng.integers(0, 2)) and T >= 4
        n_subj = int(rng.integers(1, max(2, min(50, 1000 // T))))
        ki = int(rng.integers(0, 4))
        if ki == 3:                                        # a random additive structure (see kl_stress.random_kargs)
            while True:
                kargs = kl_random_kargs(rng)
                if all(len(sp.comps) <= 8 for sp in orc.compile_spec(**kargs)):
                    break
        else:
            kargs = [synth.DEFAULT_KERNEL_ARGS, synth.SWEEP_KERNEL_ARGS, synth.MASKED_KERNEL_ARGS][ki]
        report(f"{kind} L={L} M={M} subj={n_subj} T={T} ragged={ragged} kernel={ki}",
               lambda: gp_case(kind, L, M, n_subj, T, ragged, kargs, 700 + case))
print("stress:", "OK" if bad == 0 else f"{bad} failures", "of", n_cases)
sys.exit(1 if bad else 0)
