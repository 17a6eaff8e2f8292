"""CPU emulation of tensor-core (tcgen05) arithmetic for the sufficient-statistics contraction of the KL upper
bound (SURVEY.md section 7 hard-part 1).  Test infrastructure, like everything under oracle/.

For the reference's three numerical regimes (random initial state of HLVAE_main.py:259-263, a trained-like state
after natural-gradient steps, and the state after k real steps from the initial one) it evaluates, in float64,
the quantities the streaming kernel hands to the M x M stage (elbo_functions.py:147-172, 222-266)

    S = sum_s K0xz_s^T B_s^-1 K0xz_s          W = V G,  V = B^-1 K0xz,  G = iK H iK - iK  (gradient side)

and then re-computes S and W under emulated tensor-core arithmetic:

    fp32K      K0xz rounded to float32 (what an ex2.approx evaluation delivers at best), float64 elsewhere
    tf32x3     operands split into 3 TF32 terms, products hi*hi + hi*lo + lo*hi (+ optional 6-product variant),
               float32 accumulation per 64-row panel (TMEM), float64 across panels
    bf16x6     operands split into 3 bf16 terms, 6 products, float32 accumulation per panel
    white-*    the whitened two-GEMM form: A_ = K0xz L_K^-T, Gw = A_^T B^-1 A_ under the same arithmetic
    i8xN       error-free splitting (Ozaki): operands scaled to fixed point, cut into N 8-bit slices, all slice
               pairs with i + j < N multiplied exactly (kind::i8, int32 accumulators) and recombined in float64

and reports the error of every consumer of S / W against the float64 values with the max-norm relative measure of
tests/helpers.rel_err:  D = -tr(iK S), E = tr(iK H iK S), grad_H, grad_m (natural-gradient pieces,
elbo_functions.py:186-191) and the K0 hyper-parameter / inducing-point gradients contracted from W.

Run:  python oracle/emulate_tensor_contraction.py > profiles/r02_contraction_emulation.txt
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hlvae_b200  # noqa: E402,F401
from hlvae_b200 import synth  # noqa: E402
from oracle import hlvae_oracle as orc  # noqa: E402

DT = torch.float64
PANEL = 64          # rows per TMEM accumulation (one row panel of hlvae_kl_panel)


# ----------------------------------------------------------------------------- operand formats
def to_tf32(a):
    """round-to-nearest-even to 10 explicit mantissa bits (float32 container)"""
    a32 = np.ascontiguousarray(a, dtype=np.float32)
    u = a32.view(np.uint32).astype(np.uint64)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def to_bf16(a):
    a32 = np.ascontiguousarray(a, dtype=np.float32)
    u = a32.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split(a, rnd, n):
    out, rem = [], np.asarray(a, dtype=np.float64).copy()
    for _ in range(n):
        h = rnd(rem).astype(np.float64)
        out.append(h)
        rem = rem - h
    return out


def mm_split(A, B, rnd, n, pairs, panel_axis_len=None):
    """sum over the listed (i, j) slice pairs of A_i @ B_j with float32 accumulation (numpy float32 matmul)"""
    As, Bs = split(A, rnd, n), split(B, rnd, n)
    acc = np.zeros(A.shape[:-1] + (B.shape[-1],), dtype=np.float32)
    for i, j in pairs:
        acc = acc + (As[i].astype(np.float32) @ Bs[j].astype(np.float32))
    return acc.astype(np.float64)


PAIRS3 = [(0, 0), (0, 1), (1, 0)]
PAIRS6 = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]


def mm_i8(A, B, nslice):
    """Error-free product: A, B scaled by powers of two to |.| < 1, fixed point with 8 * nslice bits, slice pairs
    with i + j < nslice (exact integer arithmetic, emulated in int64)."""
    sa = 2.0 ** np.ceil(np.log2(np.abs(A).max() + 1e-300))
    sb = 2.0 ** np.ceil(np.log2(np.abs(B).max() + 1e-300))
    bits = 8 * nslice

    def slices(X, s):
        q = np.floor(X / s * 2.0 ** (bits - 1)).astype(np.int64)        # signed fixed point, |q| < 2^(bits-1)
        out = []
        for k in range(nslice):                                          # least significant first, unsigned bytes
            out.append(q & 0xFF if k < nslice - 1 else q)                # top slice keeps the sign (signed byte)
            q = q >> 8
        return out[::-1]                                                 # most significant first

    As, Bs = slices(A, sa), slices(B, sb)
    tot = np.zeros(A.shape[:-1] + (B.shape[-1],), dtype=np.float64)
    for i in range(nslice):
        for j in range(nslice - i):
            tot += (As[i] @ Bs[j]).astype(np.float64) * 2.0 ** (-8 * (i + j))
    return tot * sa * sb * 2.0 ** (-2 * (bits - 1) + 16 * (nslice - 1))


# ----------------------------------------------------------------------------- the case
def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def build_case(L, M, n_subj, T, seed, regime):
    kargs = synth.DEFAULT_KERNEL_ARGS
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    x, _ = synth.covariates(n_subj, T, rng)
    pool, _ = synth.covariates(40, T, rng)
    z = synth.inducing_points(torch.cat([x, pool]), L, M, rng)
    N = x.shape[0]
    mu = torch.randn(N, L, generator=gen, dtype=DT)
    lv = -3.0 * torch.rand(N, L, generator=gen, dtype=DT)
    m, H = synth.variational_state(L, M, gen)
    spec0, spec1 = orc.compile_spec(**kargs)
    prm0, prm1 = orc.KernelParams.default(spec0, L), orc.KernelParams.default(spec1, L)
    noise = torch.ones(L, dtype=DT)
    steps = {"init": 0, "after-3-steps": 3, "trained-like": 8}[regime]
    lr = 0.3 if regime == "trained-like" else 0.01           # make_goldens.kl_case / config natural-gradient lr
    for _ in range(steps):
        with torch.no_grad():
            _, gm, gH = orc.minibatch_KLD_upper_bound(spec0, prm0, spec1, prm1, noise, m, H, x, mu, lv, z, n_subj,
                                                      n_subj, T, True, 1e-6)
        m, H = orc.natural_gradient_update(m, H, gm, gH, lr)
    return dict(spec0=spec0, spec1=spec1, prm0=prm0, prm1=prm1, x=x, z=z, mu=mu, lv=lv, m=m, H=H, noise=noise,
                L=L, M=M, T=T, P=n_subj)


def exact_parts(c):
    L, M, T, P = c["L"], c["M"], c["T"], c["P"]
    x, z = c["x"], c["z"]
    K = orc.eval_additive(c["spec0"], c["prm0"], x, z).numpy()                       # [L, N, M]
    Kzz = (orc.eval_additive(c["spec0"], c["prm0"], z, z) + 1e-6 * torch.eye(M, dtype=DT))
    LK = torch.linalg.cholesky(Kzz)
    iK = torch.cholesky_inverse(LK).numpy()
    iLK = torch.linalg.inv(LK).numpy()                                               # L_K^-1
    xs = x.reshape(P, T, -1).unsqueeze(1).expand(P, L, T, x.shape[1])
    B = orc.eval_additive(c["spec1"], c["prm1"], xs, xs) + torch.eye(T, dtype=DT) * c["noise"].reshape(1, L, 1, 1)
    iB = torch.linalg.inv(B).permute(1, 0, 2, 3).numpy()                             # [L, P, T, T]
    H = c["H"].numpy()
    m = c["m"].numpy()
    iH = np.linalg.inv(H)
    G = iK @ H @ iK - iK
    w = iK @ m                                                                       # [L, M, 1]
    return dict(K=K, iK=iK, iLK=iLK, iB=iB, H=H, iH=iH, m=m, G=G, w=w, cond=np.linalg.cond(Kzz.numpy()))


def block_apply(iB, X):
    """B^-1 X for block-diagonal B^-1 [L,P,T,T], X [L,N,M]"""
    L, P, T, _ = iB.shape
    return (iB @ X.reshape(L, P, T, -1)).reshape(L, P * T, -1)


def panel_sum(fn, K, V):
    """sum over 64-row panels of fn(K_panel^T, V_panel) in float64 (TMEM drained per panel)"""
    L, N, M = K.shape
    tot = np.zeros((L, M, V.shape[-1]))
    for r0 in range(0, N, PANEL):
        tot += fn(np.swapaxes(K[:, r0:r0 + PANEL], 1, 2), V[:, r0:r0 + PANEL])
    return tot


def consumers(e, S, W, c, comps):
    """Everything downstream of S and W."""
    iK, H, iH, m, G = e["iK"], e["H"], e["iH"], e["m"], e["G"]
    out = {}
    out["D=-tr(iK S)"] = -(iK * S).sum()
    out["E=tr(iKHiK S)"] = ((iK @ H @ iK) * S).sum()
    out["D+E=tr(G S)"] = (G * S).sum()
    Bm = iK @ S @ iK + iK
    out["grad_H"] = 0.5 * (-iH + Bm)
    out["grad_m(S part)"] = Bm @ m
    dJdK = W + e["rho"][:, :, None] * np.swapaxes(e["w"], 1, 2)                  # [L,N,M]
    gos, gls, gz = [], [], []
    for v, d in comps:
        gv = dJdK * v
        gos.append(gv.sum((1, 2)))
        gls.append((gv * d * d).sum((1, 2)))
        gz.append((gv * d).sum(1))
    out["d_os0"] = np.stack(gos)
    out["d_ls0"] = np.stack(gls)
    out["d_z"] = np.stack(gz)
    return out


def component_values(c):
    """unscaled per-component values v_r [L,N,M] and SE differences d_r (0 without SE factor)"""
    x, z = c["x"], c["z"]
    l_all = torch.nn.functional.softplus(c["prm0"].raw_lengthscale)
    s_all = torch.nn.functional.softplus(c["prm0"].raw_outputscale)
    comps = []
    for r, comp in enumerate(c["spec0"].comps):
        v = torch.ones(c["L"], x.shape[0], z.shape[1], dtype=DT)
        d = torch.zeros_like(v)
        for f in comp.factors:
            a = x[:, f.col].reshape(1, -1, 1)
            b = z[:, :, f.col].unsqueeze(1)
            if f.kind == orc.SE:
                ell = l_all[f.ls].reshape(-1, 1, 1)
                d = a - b
                v = v * torch.exp(-(d ** 2) / (2 * ell ** 2))
            elif f.kind == orc.CAT:
                v = v * (a - b == 0).to(DT)
            else:
                v = v * (a + b == 2).to(DT)
        comps.append(((s_all[r].reshape(-1, 1, 1) * v).numpy(), d.numpy()))
    return comps


def run(regime, L=4, M=64, n_subj=200, T=20, seed=0):
    c = build_case(L, M, n_subj, T, seed, regime)
    e = exact_parts(c)
    K, iB, G = e["K"], e["iB"], e["G"]
    mu = c["mu"].numpy().T[:, :, None]                                              # [L,N,1]
    V = block_apply(iB, K)
    S = np.swapaxes(K, 1, 2) @ V
    r = (K @ e["w"]) - mu
    e["rho"] = block_apply(iB, r)[:, :, 0]
    W = V @ G
    comps = component_values(c)
    ref = consumers(e, S, W, c, comps)
    kq_scale = abs(ref["D+E=tr(G S)"])
    with torch.no_grad():
        kld_ref = float(orc.minibatch_KLD_upper_bound(c["spec0"], c["prm0"], c["spec1"], c["prm1"], c["noise"], c["m"],
                                                      c["H"], c["x"], c["mu"], c["lv"], c["z"], n_subj, n_subj, T,
                                                      True, 1e-6)[0])
    print(f"\n=== regime {regime}: L={L} M={M} rows={K.shape[1]} cond(K0zz+eps I) = {e['cond'].max():.2e}, "
          f"|w|max = {np.abs(e['w']).max():.2e}, |G|max = {np.abs(G).max():.2e}, tr(G S) = {ref['D+E=tr(G S)']:.4e}, "
          f"kld_total = {kld_ref:.4e}")
    # float64 noise floor: the same S summed in another order (whitened form, per panel)
    iLK = e["iLK"]

    def scheme_direct(mmfn):
        return panel_sum(mmfn, K, V), mmfn(V, G)

    def scheme_white(mmfn):
        iLKt = np.swapaxes(iLK, 1, 2)
        A_ = mmfn(K, iLKt)                                                          # K0xz L_K^-T
        VA = block_apply(iB, A_)
        Gw = panel_sum(mmfn, A_, VA)                                                # L_K^-1 S L_K^-T
        LKm = np.linalg.inv(iLK)
        S_ = LKm @ Gw @ np.swapaxes(LKm, 1, 2)      # back to S in float64 (costs ~cond * 1e-16, see the noise-floor row)
        # gradient side: W = V G = (B^-1 A_) (L_K^T G L_K) L_K^-1
        Gwh = np.swapaxes(LKm, 1, 2) @ G @ LKm
        W_ = mmfn(mmfn(VA, Gwh), iLK)
        return S_, W_

    f64 = lambda A, B: A @ B
    K32 = K.astype(np.float32).astype(np.float64)
    schemes = {
        "float64, whitened order (noise floor)": lambda: scheme_white(f64),
        "fp32K  (K0xz rounded to float32, float64 arithmetic)": lambda: (
            np.swapaxes(K32, 1, 2) @ block_apply(iB, K32), block_apply(iB, K32) @ G),
        "tf32x3 direct (fp32 accum / panel)": lambda: scheme_direct(lambda A, B: mm_split(A, B, to_tf32, 2, PAIRS3)),
        "tf32x6 direct (3-way split, 6 products)": lambda: scheme_direct(lambda A, B: mm_split(A, B, to_tf32, 3, PAIRS6)),
        "bf16x6 direct": lambda: scheme_direct(lambda A, B: mm_split(A, B, to_bf16, 3, PAIRS6)),
        "white-tf32x3": lambda: scheme_white(lambda A, B: mm_split(A, B, to_tf32, 2, PAIRS3)),
        "white-tf32x6": lambda: scheme_white(lambda A, B: mm_split(A, B, to_tf32, 3, PAIRS6)),
        "white-bf16x6": lambda: scheme_white(lambda A, B: mm_split(A, B, to_bf16, 3, PAIRS6)),
        "i8x4 (32-bit fixed point)": lambda: scheme_direct(lambda A, B: mm_i8(A, B, 4)),
        "i8x5 (40-bit)": lambda: scheme_direct(lambda A, B: mm_i8(A, B, 5)),
        "i8x6 (48-bit)": lambda: scheme_direct(lambda A, B: mm_i8(A, B, 6)),
        "i8x7 (56-bit)": lambda: scheme_direct(lambda A, B: mm_i8(A, B, 7)),
    }
    keys = list(ref.keys())
    print(f"{'scheme':52s} " + " ".join(f"{k[:13]:>13s}" for k in keys) + "   gate(1e-4)")
    for name, fn in schemes.items():
        S_, W_ = fn()
        got = consumers(e, S_, W_, c, comps)
        errs = []
        for k in keys:
            if np.ndim(ref[k]) == 0:
                # scalar ELBO terms: scale-aware, against the larger of the term and the S-part of D + E
                errs.append(abs(got[k] - ref[k]) / max(abs(ref[k]), kq_scale))
            else:
                errs.append(rel(got[k], ref[k]))
                # kernel hyper-parameter gradients: the scale-aware bound of tests/helpers._hyper_ok
                if k in ("d_os0", "d_ls0") and np.abs(got[k] - ref[k]).max() <= 1e-8 * abs(kld_ref):
                    errs[-1] = min(errs[-1], 1e-4)
        ok = all(v <= 1e-4 for v in errs)
        print(f"{name:52s} " + " ".join(f"{v:13.2e}" for v in errs) + ("   pass" if ok else "   FAIL"))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for regime in ("init", "after-3-steps", "trained-like"):
        run(regime)
