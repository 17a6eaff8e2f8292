// Fused masked heterogeneous log-likelihood: HLVAE.loglik_and_reconstruction
// (HLVAE.py:381-414) over the per-type functions of HL_VAE/loglik.py:27-213, plus
// read_functions.statistics (:268-302) and discrete_variables_transformation (:221-235).
//
// Layout of the work: a CTA owns a TILE of consecutive variables (<= LL_THREADS, one per
// thread) and walks a stripe of row batches (LL_ROWS rows at a time).  Per batch it stages the
// tile's contiguous span of `theta` (cp.async, no register round trip) and `data` of every row
// of the batch into shared memory - one coalesced burst that keeps tens of KB in flight per CTA -
// then every thread evaluates its variable for each row from shared memory, the per-class
// outputs (params / g_theta) go back through shared memory as coalesced spans, and the [N, D]
// outputs are written directly (consecutive threads = consecutive columns).  Per-variable
// constants (variances, normalisation) are computed once per thread in float64 and reused for
// every row of the stripe.
//
// Arithmetic type R follows the storage type of theta: float64 storage -> float64 math (the
// drop-in case, reference tensors are float64); float32 storage -> float32 math on the SFU
// (ex2/lg2.approx), which makes the kernel HBM-bound.  In the float32 path the categorical /
// ordinal argmax imputations (which must be bit-exact against the float64 reference) re-evaluate
// the variable in float64 whenever the float32 decision is not provably the float64 one.
#include <type_traits>
#include "common.cuh"

using namespace hlvae;

namespace {

constexpr int LL_THREADS = 128;                          // threads per CTA = max variables per tile
#ifndef HLVAE_LL_ROWS
#define HLVAE_LL_ROWS 8
#endif
#ifndef HLVAE_LL_STAGES
#define HLVAE_LL_STAGES 1
#endif
constexpr int LL_ROWS = HLVAE_LL_ROWS;                   // rows staged per batch
constexpr int LL_CAPM = LL_THREADS + 16;                  // staged mask / upstream-gradient row (with alignment shift)
static_assert(4 * LL_ROWS <= 32, "one warp issues every bulk copy of a row batch");
constexpr int LL_STAGES = HLVAE_LL_STAGES;               // 2: prefetch the next row batch while this one is evaluated.  Measured
                                                         // again after the r02 instruction diet (fwd / bwd ms at configs[1]):
                                                         // 1 stage x 8 rows 0.270 / 0.179, 2 x 8 0.356 / 0.237 (4 CTAs per SM),
                                                         // 2 x 4 0.329 / 0.207, 1 x 4 0.305 / 0.213 - the resident CTAs of an
                                                         // SM already cover each other's loads
constexpr double LOG_2PI = 1.8378770664093454835606594728112;

// ------------------------------------------------------------------------------------
template <typename R> struct Mth;
template <> struct Mth<double> {
    static __device__ __forceinline__ double ex(double x) { return exp(x); }
    static __device__ __forceinline__ double lg(double x) { return log(x); }
    static __device__ __forceinline__ double lg1p(double x) { return log1p(x); }
    static __device__ __forceinline__ double lg1p_nn(double x) { return log1p(x); }
    static __device__ __forceinline__ double exm1(double x) { return expm1(x); }
    static __device__ __forceinline__ double lgam(double x) { return lgamma(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
};
template <> struct Mth<float> {
    static __device__ __forceinline__ float ex(float x) { return __expf(x); }
    static __device__ __forceinline__ float lg(float x) { return __logf(x); }
    static __device__ __forceinline__ float lg1p(float x) { return log1pf(x); }
    // log1p of a NON-NEGATIVE argument (the softplus sites: x = e^t).  [0, 0.25): x * degree-5 minimax polynomial of
    // log1p(x) / x (1.1e-7 relative); above: lg2.approx of 1 + x, whose 2^-22 absolute error is at most 1e-6 of
    // log(1.25).  11 instructions against log1pf's 28 (17 % of the tabular backward kernel's instructions).
    static __device__ __forceinline__ float lg1p_nn(float x) {
        float p = fmaf(-0.09238353371620178f, x, 0.1817312091588974f);
        p = fmaf(p, x, -0.2477860152721405f);
        p = fmaf(p, x, 0.33320868015289307f);
        p = fmaf(p, x, -0.49999740719795227f);
        p = fmaf(p, x, 1.0f);
        return x < 0.25f ? p * x : __logf(1.0f + x);
    }
    static __device__ __forceinline__ float exm1(float x) { return expm1f(x); }
    static __device__ __forceinline__ float lgam(float x) { return lgammaf(x); }
    static __device__ __forceinline__ float rcp(float x) { return __fdividef(1.0f, x); }
};

// torch.nn.functional.softplus (beta = 1, threshold = 20) and its derivative
template <typename R> __device__ __forceinline__ R softplus_(R x) {
    return x > R(20) ? x : Mth<R>::lg1p_nn(Mth<R>::ex(x));
}
template <typename R> __device__ __forceinline__ R sigmoid_(R x) { return Mth<R>::rcp(R(1) + Mth<R>::ex(-x)); }
template <typename R> __device__ __forceinline__ R dsoftplus_(R x) { return x > R(20) ? R(1) : sigmoid_<R>(x); }
template <typename R> __device__ __forceinline__ R clamp_(R x, R lo, R hi) { return fmin(fmax(x, lo), hi); }

template <typename T> __device__ __forceinline__ double ldd(const void* p, int64_t i) {
    return (double)reinterpret_cast<const T*>(p)[i];
}
template <typename TS> __device__ __forceinline__ void st(void* p, int64_t i, double v) {
    if (p) reinterpret_cast<TS*>(p)[i] = (TS)v;
}

__device__ __forceinline__ int d0_(int tile, int tile_vars) { return tile * tile_vars; }

// Thread -> variable of the tile.  A warp that holds variables of several types executes every type's code path for
// every row (the D4 layout alternates runs of 18 real and 18 categorical variables: half of the warps paid for both).
// The tile's variables are therefore handed out sorted by type (stable: runs stay contiguous, so the accesses of a
// warp stay on neighbouring columns): thread t takes the variable of rank t.  Returns its index inside the tile
// (>= n_vars: no variable).  One ballot per type and warp, once per CTA.
__device__ __forceinline__ int tile_variable_by_type(int tid, int d0, int n_vars, const int32_t* __restrict__ var_kind) {
    constexpr int NW = LL_THREADS / 32, NK = 6;               // 5 types + "no variable"
    __shared__ int s_cnt[NW][NK], s_perm[LL_THREADS];
    const int lane = tid & 31, w = tid >> 5;
    int kind = NK - 1;
    if (tid < n_vars) {
        const int k = var_kind[d0 + tid];
        kind = (k >= 0 && k < NK - 1) ? k : NK - 2;
    }
    int below = 0;                                           // same type, lower lane of this warp
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, kind == k);
        if (lane == 0) s_cnt[w][k] = __popc(m);
        if (kind == k) below = __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    int rank = below;
#pragma unroll
    for (int k = 0; k < NK; k++)
#pragma unroll
        for (int w2 = 0; w2 < NW; w2++)
            if (k < kind || (k == kind && w2 < w)) rank += s_cnt[w2][k];
    s_perm[rank] = tid;
    __syncthreads();
    return s_perm[tid];
}

// Per-variable constants, computed once per thread (float64, then rounded to R).
template <typename R> struct VarC {
    int kind, C, xo, po;      // type, classes, offsets of the variable inside the staged spans
    R nm, snv, ivar, lconst, idiv, ev, sg8;
    bool ok;
};

template <typename R>
__device__ __forceinline__ VarC<R> load_var(int d, int D, bool active, const int32_t* __restrict__ var_kind,
                                            const int32_t* __restrict__ var_nclass,
                                            const int32_t* __restrict__ var_dcol,
                                            const int32_t* __restrict__ var_pcol, const double* __restrict__ vparam,
                                            int xs0, int ps0, int cap) {
    VarC<R> v;
    v.kind = -1; v.C = 1; v.xo = 0; v.po = 0; v.ok = false;
    v.nm = v.snv = v.ivar = v.lconst = v.idiv = v.ev = v.sg8 = R(0);
    if (!active) return v;
    v.kind = var_kind[d];
    v.C = var_nclass[d];
    v.xo = var_dcol[d] - xs0;
    v.po = var_pcol[d] - ps0;
    // variables of a tile must be packed in order (header: "packed layout"); otherwise flag it
    v.ok = v.C >= 1 && v.C <= HLVAE_MAX_CLASS && v.xo >= 0 && v.po >= 0 && v.xo + v.C <= cap && v.po + v.C <= cap;
    const double nm = vparam[d], nv = vparam[D + d], e = vparam[2 * D + d], div = vparam[3 * D + d];
    v.nm = (R)nm;
    v.snv = (R)sqrt(nv);
    v.idiv = (R)(1.0 / div);
    if (v.kind == HLVAE_VAR_REAL) {
        const double t8 = e + 8.0;
        const double sp = t8 > 20.0 ? t8 : log1p(exp(t8));
        const double var = nv * exp(-8.0 + sp);                              // loglik.py:51-52,56
        v.ivar = (R)(1.0 / var);
        v.lconst = (R)(0.5 * LOG_2PI + 0.5 * log(var));                      // :58
        v.sg8 = (R)(t8 > 20.0 ? 1.0 : 1.0 / (1.0 + exp(-t8)));
    } else if (v.kind == HLVAE_VAR_POS) {
        const double var = nv * exp(e);                                      // :100
        v.ivar = (R)(1.0 / var);
        v.lconst = (R)(0.5 * log(2.0 * 3.14159265358979323846 * var));       // :102
        v.ev = (R)exp(e);                                                    // read_functions.py:284
    }
    return v;
}

// ------------------------------------------------------------------------------------
// Exact (float64) argmax decisions for the float32 path: same operation order as the float64
// reference (torch.logsumexp, then theta - lse; first index wins ties).
__device__ __noinline__ int cat_argmax_f64(const float* t, int C) {
    double mx = -INFINITY;
    for (int c = 0; c < C; c++) mx = fmax(mx, (double)t[c]);
    double se = 0.0;
    for (int c = 0; c < C; c++) se += exp((double)t[c] - mx);
    const double lse = mx + log(se);
    int am = 0;
    double best = (double)t[0] - lse;
    for (int c = 1; c < C; c++) {
        const double v = (double)t[c] - lse;
        if (v > best) { best = v; am = c; }
    }
    return am;
}

__device__ __forceinline__ double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }

// ordinal class probabilities exactly as loglik.py:163-178 (float64); returns the argmax
template <typename TI>
__device__ __forceinline__ int ordinal_probs_f64(const TI* th, int C, double* p, double& tot_out) {
    const double eps = 1e-6;
    const double loc = softplus_d((double)th[C - 1]);                        // :163
    double cum = 0.0, prev = 0.0, tot = 0.0;
    for (int c = 0; c < C; c++) {
        double sg = 1.0;
        if (c < C - 1) {
            cum += fmin(fmax(softplus_d((double)th[c]), eps), 1e20);         // :164
            sg = 1.0 / (1.0 + exp(-(cum - loc)));                            // :165
        }
        p[c] = fmin(fmax(sg - prev, eps), 1.0);                              // :166-169
        prev = sg;
        tot += p[c];
    }
    int am = 0;
    double best = -1.0;
    for (int c = 0; c < C; c++) {
        const double ph = p[c] / tot;                                        // :178
        if (ph > best) { best = ph; am = c; }
    }
    tot_out = tot;
    return am;
}

__device__ __noinline__ int ord_argmax_f64(const float* th, int C) {
    double p[HLVAE_MAX_CLASS], tot;
    return ordinal_probs_f64<float>(th, C, p, tot);
}

// ------------------------------------------------------------------------------------
// Categorical variable, forward (loglik.py:124-146): params = theta - logsumexp(theta) written over
// theta, log_p_x = sum_c x_c params_c (the reference's second log_softmax of the already normalised
// params is the identity up to 1 ulp and is dropped), argmax of params (read_functions.py:296-302)
// and of the one-hot data (:226-227).  CMAX is the unroll bound: when it equals C the loops are
// straight-line code on registers.
template <typename R, typename XT, int CMAX>
__device__ __forceinline__ void cat_forward(const XT* __restrict__ x, R* __restrict__ t, int C, R& lp, R& rmean, R& dtr) {
    R tv[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; c++) tv[c] = (c < C) ? t[c] : R(-INFINITY);
    R mx = tv[0];
    int am = 0;
#pragma unroll
    for (int c = 1; c < CMAX; c++) {
        const R tc = tv[c];
        if (tc > mx) { mx = tc; am = c; }                                    // first index wins ties
    }
    R se = R(0);
#pragma unroll
    for (int c = 0; c < CMAX; c++) se += Mth<R>::ex(tv[c] - mx);             // padded logits are -inf: exp = 0
    const R lse = mx + Mth<R>::lg(se);                                       // torch.logsumexp
    if constexpr (sizeof(R) == 4) {
        // The reference takes argmax of fl64(theta_c - lse); it can differ from argmax(theta) only when two
        // DISTINCT logits collapse onto one double after the subtraction: gap = mx - (largest logit below mx) <
        // 1e-13 (|mx| + |lse|).  Distinct float32 logits are at least 2^-24 |mx| apart and |lse| <= |mx| + ln 16, so
        // that needs |mx| < 4.7e-6: only then is the runner-up looked for at all.
        if (fabs(mx) < R(1e-5)) {
            R second = -INFINITY;
#pragma unroll
            for (int c = 0; c < CMAX; c++)
                if (tv[c] < mx) second = fmax(second, tv[c]);
            const R gap = mx - second;
            if (gap > R(0) && gap < R(1e-13) * (fabs(mx) + fabs(lse)) + R(1e-37))
                am = cat_argmax_f64(reinterpret_cast<const float*>(t), C);
        }
    } else {
        am = 0;
        R best = tv[0] - lse;
#pragma unroll
        for (int c = 1; c < CMAX; c++) {
            const R val = tv[c] - lse;
            if (c < C && val > best) { best = val; am = c; }
        }
    }
    R acc = R(0), dbest = (R)x[0];
    int dam = 0;
#pragma unroll
    for (int c = 0; c < CMAX; c++) {
        if (c < C) {
            const R pc = tv[c] - lse;
            const R xv = (R)x[c];
            acc += xv * pc;
            t[c] = pc;
            if (xv > dbest) { dbest = xv; dam = c; }
        }
    }
    lp = acc;
    rmean = (R)am;
    dtr = (R)dam;
}

// Categorical variable, backward: g (x_c - softmax(theta)_c sum(x)) over theta.
template <typename R, typename XT, int CMAX>
__device__ __forceinline__ void cat_backward(const XT* __restrict__ x, R* __restrict__ t, int C, R g) {
    R tv[CMAX], xv[CMAX];
    R mx = -INFINITY, sx = R(0);
#pragma unroll
    for (int c = 0; c < CMAX; c++) {
        tv[c] = (c < C) ? t[c] : R(-INFINITY);
        xv[c] = (c < C) ? (R)x[c] : R(0);
        mx = fmax(mx, tv[c]);
        sx += xv[c];
    }
    R se = R(0);
#pragma unroll
    for (int c = 0; c < CMAX; c++) { tv[c] = Mth<R>::ex(tv[c] - mx); se += tv[c]; }
    const R k = sx * Mth<R>::rcp(se);
#pragma unroll
    for (int c = 0; c < CMAX; c++)
        if (c < C) t[c] = g * (xv[c] - tv[c] * k);
}

// ------------------------------------------------------------------------------------
// One variable, forward.  x / t point into the staged spans (t is overwritten with `params`).  KIND >= 0 / CFIX > 0:
// the type / class count as compile-time constants (the callers' per-type row loops), else read from v.
template <typename R, typename XT, int KIND = -1, int CFIX = 0>
__device__ __forceinline__ void var_forward(const VarC<R>& v, const XT* __restrict__ x_, R* __restrict__ t, bool observed,
                                            R& lp, R& rmean, R& rmode, R& dtr) {
    struct { const XT* p; __device__ __forceinline__ R operator[](int c) const { return (R)p[c]; } } x{x_};
    const int C = CFIX > 0 ? CFIX : v.C;
    const int kind = KIND >= 0 ? KIND : v.kind;
    if (kind == HLVAE_VAR_REAL) {                                          // loglik.py:27-70
        const R xv = x[0] * v.idiv;                                          // HLVAE.py:393-394
        const R mean = v.snv * t[0] + v.nm;                                  // :55
        const R r = xv - mean;
        lp = R(-0.5) * r * r * v.ivar - v.lconst;                            // :58
        t[0] = mean;
        rmean = mean; rmode = mean;                                          // read_functions.py:275-278
        dtr = x[0];                                                          // read_functions.py:233
    } else if (kind == HLVAE_VAR_POS) {                                    // loglik.py:73-121
        const R ld = Mth<R>::lg1p(x[0]);                                     // :84
        const R mean = v.snv * t[0] + v.nm;                                  // :96
        const R r = ld - mean;
        lp = R(-0.5) * r * r * v.ivar - v.lconst - ld;                       // :102
        t[0] = mean;
        rmean = Mth<R>::ex(mean + R(0.5) * v.ev) - R(1);                     // read_functions.py:287
        rmode = Mth<R>::ex(mean - v.ev) - R(1);                              // :289
        dtr = x[0];
    } else if (kind == HLVAE_VAR_COUNT) {                                  // loglik.py:191-213
        const R lam = clamp_<R>(softplus_<R>(t[0]), R(1e-6), R(1e20));       // :203
        const R xv = x[0];
        lp = xv * Mth<R>::lg(lam) - lam - Mth<R>::lgam(xv + R(1));           // Poisson.log_prob
        t[0] = lam;
        rmean = lam; rmode = floor(lam);                                     // read_functions.py:293-295
        dtr = xv;
    } else if (kind == HLVAE_VAR_CAT) {                                    // loglik.py:124-146
        switch (C) {     // common class counts run fully unrolled from registers
            case 2: cat_forward<R, XT, 2>(x_, t, 2, lp, rmean, dtr); break;
            case 3: cat_forward<R, XT, 3>(x_, t, 3, lp, rmean, dtr); break;
            case 4: cat_forward<R, XT, 4>(x_, t, 4, lp, rmean, dtr); break;
            case 5: cat_forward<R, XT, 5>(x_, t, 5, lp, rmean, dtr); break;
            case 6: cat_forward<R, XT, 6>(x_, t, 6, lp, rmean, dtr); break;
            case 8: cat_forward<R, XT, 8>(x_, t, 8, lp, rmean, dtr); break;
            case 10: cat_forward<R, XT, 10>(x_, t, 10, lp, rmean, dtr); break;
            default: cat_forward<R, XT, HLVAE_MAX_CLASS>(x_, t, C, lp, rmean, dtr); break;
        }
        rmode = rmean;                                                       // read_functions.py:296-302
    } else {                                                                 // ordinal, loglik.py:149-188
        int vals = 0;
        R sx = R(0);
        for (int c = 0; c < C; c++) { vals += (int)x[c]; sx += x[c]; }       // :172
        if (!observed) vals = 1;                                             // :173
        int am = 0;
        R py = R(1), lsum = R(1);
        if constexpr (sizeof(R) == 8) {
            double p[HLVAE_MAX_CLASS], tot;
            am = ordinal_probs_f64<R>(t, C, p, tot);
            double ls = 0.0;
            for (int c = 0; c < C; c++) {
                const double ph = p[c] / tot;                                // :178
                t[c] = (R)ph;
                ls += ph;
                if (c == vals - 1) py = (R)ph;
            }
            lsum = (R)ls;
        } else {
            // float32: class probabilities as products of factors in (0, 1] so that small differences of
            // sigmoids keep their relative accuracy:
            //   sigma(u_c) - sigma(u_{c-1}) = sigma(u_c) sigma(-u_{c-1}) (1 - exp(-(u_c - u_{c-1})))
            // The unnormalised p_c overwrite theta_c in place (theta_c is consumed first; theta_{C-1} is `loc`).
            const R eps = R(1e-6);
            const R loc = softplus_<R>(t[C - 1]);
            R cum = R(0), sneg_prev = R(1), tot = R(0);
            for (int c = 0; c < C; c++) {
                R pc;
                if (c < C - 1) {
                    // delta = softplus(t) and 1 - exp(-delta) = sigmoid(t) from one exponential (below the clamp:
                    // 1 - exp(-eps)); the expm1 this replaces was the most expensive call of the path
                    const R tc = t[c];
                    const R et = Mth<R>::ex(fmin(tc, R(30)));
                    const R sp = tc > R(20) ? tc : Mth<R>::lg1p_nn(et);
                    const R omx = sp < eps ? R(9.999995e-7) : (tc > R(20) ? R(1) : et * Mth<R>::rcp(R(1) + et));
                    const R delta = clamp_<R>(sp, eps, R(1e20));
                    cum += delta;
                    const R e = Mth<R>::ex(loc - cum);                       // exp(-u_c)
                    const R sg = Mth<R>::rcp(R(1) + e);                      // sigma(u_c)
                    pc = (c == 0) ? sg : sg * sneg_prev * omx;
                    sneg_prev = (e > R(1e30)) ? R(1) : e * sg;               // sigma(-u_c)
                } else {
                    pc = sneg_prev;                                          // 1 - sigma(u_{C-2})
                }
                pc = clamp_<R>(pc, eps, R(1));
                tot += pc;
                t[c] = pc;
            }
            const R itot = Mth<R>::rcp(tot);
            R best = R(-1), second = R(-1), ls = R(0);
            for (int c = 0; c < C; c++) {
                const R ph = t[c] * itot;
                t[c] = ph;
                ls += ph;
                if (ph > best) { second = best; best = ph; am = c; }
                else second = fmax(second, ph);
                if (c == vals - 1) py = ph;
            }
            lsum = ls;
            // not provably the float64 decision -> the caller redoes this variable in float64 (rare)
            if (best - second < R(4e-6) * (R(4) + cum + loc) * best) am = -1;
        }
        lp = Mth<R>::lg(py) - Mth<R>::lg(lsum);                              // :179 log_softmax(log p)
        rmean = (R)am; rmode = (R)am;
        dtr = sx - R(1);                                                     // read_functions.py:229-230
    }
}

// ------------------------------------------------------------------------------------
// Staging helpers.  cp.async moves 4 / 8-byte elements global -> shared without a register round
// trip, so a CTA can put a whole row batch in flight before it waits.
template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N_) : "memory"); }

// ---- TMA (bulk asynchronous copy) versions of the same staging: ONE thread issues one
// cp.async.bulk per row and array (16-byte aligned start, size a multiple of 16 bytes - exactly the
// aligned chunk ranges used below), completion is counted on an mbarrier; the other 127 threads issue
// nothing.  Results go back the same way (cp.async.bulk shared -> global) for the whole chunks of a row.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// One row of one array by TMA (called by a single thread): returns the alignment shift, adds the bytes to `tx`.
template <typename T>
__device__ __forceinline__ int tma_row_bytes(const T* src, int span, unsigned& bytes) {
    constexpr int E = 16 / (int)sizeof(T);
    const int shift = (int)((reinterpret_cast<uintptr_t>(src) & 15) / sizeof(T));
    bytes = (unsigned)(((shift + span + E - 1) / E) * 16);
    return shift;
}

// Whole 16-byte chunks of a staged row go out as one bulk store issued by thread `issuer` (one row per thread,
// so the rows of a batch are issued in parallel); the partial chunks at the two ends are written element-wise
// by the first threads of the CTA.
template <typename T>
__device__ __forceinline__ void tma_unstage_row(T* __restrict__ dst, const T* __restrict__ src, int shift, int span, int tid,
                                                int issuer) {
    constexpr int E = 16 / (int)sizeof(T);
    const int i0 = shift > 0 ? 1 : 0;                         // first whole chunk
    const int i1 = (shift + span) / E;                        // one past the last whole chunk
    const int head = min(span, E * i0 - shift);               // elements before the first whole chunk
    const int tail0 = max(head, E * i1 - shift);              // first element after the last whole chunk
    if (tid == issuer && i1 > i0) bulk_s2g(dst + (E * i0 - shift), src + E * i0, (unsigned)((i1 - i0) * 16));
    if (tid < head) dst[tid] = src[shift + tid];
    if (tid >= E && tid - E < span - tail0) dst[tail0 + tid - E] = src[shift + tail0 + tid - E];
}

// The same for a whole batch of rows in one pass when every row has the same alignment shift (row stride a multiple
// of 16 bytes): thread = (row tid / 16, job tid % 16) - job 0 issues the row's bulk store, the jobs after it write
// the at most E - 1 elements of the partial chunk at either end.  (Row by row, every thread redid the chunk
// arithmetic for every row: 13 % of the forward kernel's instructions.)
template <typename T, int ROWS = LL_ROWS>
__device__ __forceinline__ void tma_unstage_batch(T* __restrict__ dst0, int64_t ld, const T* __restrict__ src0, int lds,
                                                  int shift, int span, int nr, int tid) {
    constexpr int E = 16 / (int)sizeof(T);
    static_assert(LL_THREADS / ROWS >= 2 * E - 1 || E > 8, "jobs per row");
    const int r = tid / (LL_THREADS / ROWS), j = tid % (LL_THREADS / ROWS);
    if (r >= nr) return;
    const int i0 = shift > 0 ? 1 : 0;
    const int i1 = (shift + span) / E;
    const int head = min(span, E * i0 - shift);
    const int tail0 = max(head, E * i1 - shift);
    T* dst = dst0 + (int64_t)r * ld;
    const T* src = src0 + r * lds;
    if (j == 0) {
        if (i1 > i0) bulk_s2g(dst + (E * i0 - shift), src + E * i0, (unsigned)((i1 - i0) * 16));
    } else {
        for (int e = j - 1; e < head; e += LL_THREADS / ROWS - 1) dst[e] = src[shift + e];
        for (int e = tail0 + j - 1; e < span; e += LL_THREADS / ROWS - 1) dst[e] = src[shift + e];
    }
}

// Stage `span` elements of one row as aligned 16-byte chunks: the copy starts at the 16-byte boundary
// below `src` (`shift` elements early; the caller finds element j at dst[shift + j]) and ends at the
// boundary above the last element.  The over-read stays inside the array as long as the array itself
// starts and ends on 16-byte boundaries (`vec_ok`); otherwise elements are copied one by one.
template <typename T>
__device__ __forceinline__ int stage_row(T* __restrict__ dst, const T* __restrict__ src, int span, int tid, bool vec_ok) {
    constexpr int E = 16 / (int)sizeof(T);
    if (vec_ok) {
        const int shift = (int)((reinterpret_cast<uintptr_t>(src) & 15) / sizeof(T));
        const T* s0 = src - shift;
        const int chunks = (shift + span + E - 1) / E;
        for (int i = tid; i < chunks; i += LL_THREADS) cp_async_elem<16>(dst + E * i, s0 + E * i);
        return shift;
    }
    if constexpr (sizeof(T) >= 4) {
        for (int i = tid; i < span; i += LL_THREADS) cp_async_elem<sizeof(T)>(dst + i, src + i);
    } else {
        for (int i = tid; i < span; i += LL_THREADS) dst[i] = src[i];
    }
    return 0;
}

// Write `span` elements staged at src[shift + j] to dst[j]; dst has the same 16-byte phase as the staged
// row (same column offset, same leading dimension), so whole chunks go out as 16-byte stores.
template <typename T>
__device__ __forceinline__ void unstage_row(T* __restrict__ dst, const T* __restrict__ src, int shift, int span,
                                            int tid, bool vec_ok) {
    constexpr int E = 16 / (int)sizeof(T);
    if (vec_ok && (int)((reinterpret_cast<uintptr_t>(dst) & 15) / sizeof(T)) == shift) {
        const int chunks = (shift + span + E - 1) / E;
        for (int i = tid; i < chunks; i += LL_THREADS) {
            const int j0 = E * i - shift;                    // first element of the chunk, in span coordinates
            if (j0 >= 0 && j0 + E <= span) {
                *reinterpret_cast<int4*>(dst + j0) = *reinterpret_cast<const int4*>(src + E * i);
            } else {
#pragma unroll
                for (int e = 0; e < E; e++)
                    if (j0 + e >= 0 && j0 + e < span) dst[j0 + e] = src[E * i + e];
            }
        }
    } else {
        for (int i = tid; i < span; i += LL_THREADS) dst[i] = src[shift + i];
    }
}

// grid: (variable tiles, row stripes); dynamic smem: 2 * LL_ROWS * cap elements of R
template <typename TS, typename TD, typename TM>
__global__ void __launch_bounds__(LL_THREADS, sizeof(TS) == 4 ? 6 : 3)
loglik_fwd_k(int64_t N, int D, int tile_vars, int cap, int64_t ld_data, int64_t ld_theta,
             const int32_t* __restrict__ var_kind, const int32_t* __restrict__ var_nclass,
             const int32_t* __restrict__ var_dcol, const int32_t* __restrict__ var_pcol,
             const double* __restrict__ vparam, const TD* __restrict__ data, const TS* __restrict__ theta,
             const TM* __restrict__ mask, TS* __restrict__ log_p_x, TS* __restrict__ log_p_x_missing,
             TS* __restrict__ params, TS* __restrict__ recon_mean, TS* __restrict__ recon_mode,
             TS* __restrict__ data_tr, double* __restrict__ ll_total) {
    using R = TS;
    extern __shared__ __align__(16) unsigned char ll_smem[];
    const int capt = cap + 8, capx = cap + 16;   // row strides: room for the alignment shift, multiples of 16 bytes
    // per stage: sT [LL_ROWS][capt] theta (overwritten with params), sX [LL_ROWS][capx] data and
    // sK [LL_ROWS][LL_CAPM] mask of the tile's variables, both in their storage types
    const size_t stage_bytes = (size_t)LL_ROWS * (capt * sizeof(R) + capx * sizeof(TD) + LL_CAPM * sizeof(TM));
    __shared__ int sShift[LL_STAGES][LL_ROWS], sShiftT[LL_STAGES][LL_ROWS], sShiftM[LL_STAGES][LL_ROWS];
    const bool m_ok = (((uintptr_t)mask | (uintptr_t)(N * D * (int64_t)sizeof(TM))) & 15) == 0;
    const int span_m = min(D, d0_(blockIdx.x, tile_vars) + tile_vars) - d0_(blockIdx.x, tile_vars);
    const bool x_ok = (((uintptr_t)data | (uintptr_t)(N * ld_data * (int64_t)sizeof(TD))) & 15) == 0;
    const bool t_ok = (((uintptr_t)theta | (uintptr_t)params | (uintptr_t)(N * ld_theta * (int64_t)sizeof(TS))) & 15) == 0;
    __shared__ double red[LL_THREADS / 32];
    const int tid = threadIdx.x;
    const int d0 = blockIdx.x * tile_vars;
    const int d1 = min(D, d0 + tile_vars);
    const int lt = tile_variable_by_type(tid, d0, d1 - d0, var_kind);   // this thread's variable inside the tile
    const int d = d0 + lt;
    const bool active = d < d1;
    const int xs0 = var_dcol[d0], ps0 = var_pcol[d0];
    const int span_x = max(0, min(cap, var_dcol[d1 - 1] + var_nclass[d1 - 1] - xs0));
    const int span_p = max(0, min(cap, var_pcol[d1 - 1] + var_nclass[d1 - 1] - ps0));
    const VarC<R> v = load_var<R>(d, D, active, var_kind, var_nclass, var_dcol, var_pcol, vparam, xs0, ps0, cap);
    const int64_t stride = (int64_t)gridDim.y * LL_ROWS;
    __shared__ __align__(8) unsigned long long mbar[LL_STAGES];
    const bool use_tma = x_ok && t_ok && m_ok;   // every staged array starts and ends on a 16-byte boundary
    // row stride of theta / params a multiple of 16 bytes: every row of the tile has the same alignment shift
    const bool uni_t = use_tma && ld_theta % (16 / (int)sizeof(TS)) == 0;
    const int shift_t = (int)((reinterpret_cast<uintptr_t>(theta + ps0) & 15) / sizeof(TS));
    const bool uni = uni_t && ld_data % (16 / (int)sizeof(TD)) == 0 && D % (16 / (int)sizeof(TM)) == 0;
    const int shift_x = (int)((reinterpret_cast<uintptr_t>(data + xs0) & 15) / sizeof(TD));
    const int shift_m = (int)((reinterpret_cast<uintptr_t>(mask + d0) & 15) / sizeof(TM));
    unsigned phase[LL_STAGES];
#pragma unroll
    for (int i = 0; i < LL_STAGES; i++) phase[i] = 0;
    if (use_tma) {
        if (tid == 0)
            for (int i = 0; i < LL_STAGES; i++) mbar_init(&mbar[i], 1);
        __syncthreads();
    }

    auto prefetch = [&](int64_t n0, int stg) {   // put one row batch in flight (nothing if n0 is past the end)
        if (n0 < N && use_tma) {
            if (tid < 32) {        // warp 0: lane = (array, row); every lane issues its own bulk copy
                R* sT = reinterpret_cast<R*>(ll_smem + stg * stage_bytes);
                TD* sX = reinterpret_cast<TD*>(sT + LL_ROWS * capt);
                TM* sK = reinterpret_cast<TM*>(sX + LL_ROWS * capx);
                const int nr = (int)min((int64_t)LL_ROWS, N - n0);
                const int r = tid % LL_ROWS, arr = tid / LL_ROWS;
                const bool job = r < nr && arr < 3;
                unsigned bytes = 0;
                int shift = 0;
                const void* src = nullptr;
                void* dst = nullptr;
                if (job) {
                    if (arr == 0) {
                        const TS* p_ = theta + (n0 + r) * ld_theta + ps0;
                        shift = tma_row_bytes<TS>(p_, span_p, bytes);
                        src = p_ - shift; dst = sT + r * capt; sShiftT[stg][r] = shift;
                    } else if (arr == 1) {
                        const TD* p_ = data + (n0 + r) * ld_data + xs0;
                        shift = tma_row_bytes<TD>(p_, span_x, bytes);
                        src = p_ - shift; dst = sX + r * capx; sShift[stg][r] = shift;
                    } else {
                        const TM* p_ = mask + (n0 + r) * D + d0;
                        shift = tma_row_bytes<TM>(p_, span_m, bytes);
                        src = p_ - shift; dst = sK + r * LL_CAPM; sShiftM[stg][r] = shift;
                    }
                }
                const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
                if (tid == 0) mbar_expect_tx(&mbar[stg], total);
                __syncwarp();
                if (job) bulk_g2s(dst, src, bytes, &mbar[stg]);
            }
        } else if (n0 < N) {
            R* sT = reinterpret_cast<R*>(ll_smem + stg * stage_bytes);
            TD* sX = reinterpret_cast<TD*>(sT + LL_ROWS * capt);
            TM* sK = reinterpret_cast<TM*>(sX + LL_ROWS * capx);
            const int nr = (int)min((int64_t)LL_ROWS, N - n0);
            for (int r = 0; r < nr; r++) {
                const int st_ = stage_row<TS>(sT + r * capt, theta + (n0 + r) * ld_theta + ps0, span_p, tid, t_ok);
                const int sh = stage_row<TD>(sX + r * capx, data + (n0 + r) * ld_data + xs0, span_x, tid, x_ok);
                const int sm = stage_row<TM>(sK + r * LL_CAPM, mask + (n0 + r) * D + d0, span_m, tid, m_ok);
                if (tid == 0) { sShift[stg][r] = sh; sShiftT[stg][r] = st_; sShiftM[stg][r] = sm; }
            }
        }
        cp_async_commit();
    };

    double ll = 0.0;
    int stg = 0;
    if (LL_STAGES == 2) prefetch((int64_t)blockIdx.y * LL_ROWS, 0);
    for (int64_t n0 = (int64_t)blockIdx.y * LL_ROWS; n0 < N; n0 += stride, stg ^= (LL_STAGES - 1)) {
        if (LL_STAGES == 2) {
            prefetch(n0 + stride, stg ^ 1);      // next batch travels while this one is evaluated
            cp_async_wait_group<1>();
        } else {
            prefetch(n0, 0);
            cp_async_wait_all();
        }
        __syncthreads();                         // shifts written by thread 0 / cp.async data of all threads
        if (use_tma) {
            mbar_wait(&mbar[stg], phase[stg]);
            phase[stg] ^= 1;
        }
        R* sT = reinterpret_cast<R*>(ll_smem + stg * stage_bytes);
        TD* sX = reinterpret_cast<TD*>(sT + LL_ROWS * capt);
        TM* sK = reinterpret_cast<TM*>(sX + LL_ROWS * capx);
        const int nr = (int)min((int64_t)LL_ROWS, N - n0);
        if (active) {
            int64_t o = n0 * D + d;
            // The row loop exists once per evaluator: the variable's type (and its validity) is fixed per thread, so
            // the dispatch is made once per batch instead of once per row - real variables and 5-class categorical
            // ones get loops of their own, everything else shares the general one.
            auto rows = [&](auto&& eval) {
#pragma unroll 1
                for (int r = 0; r < nr; r++, o += D) {
                    const int sm_ = uni ? shift_m : sShiftM[stg][r];
                    const int sx_ = uni ? shift_x : sShift[stg][r];
                    const int st_ = uni ? shift_t : sShiftT[stg][r];
                    const R m_ = (R)sK[r * LL_CAPM + sm_ + lt];
                    R lp = R(0), rmean = R(0), rmode = R(0), dtr = R(0);
                    eval(sX + r * capx + sx_ + v.xo, sT + r * capt + st_ + v.po, m_ != R(0), lp, rmean, rmode, dtr, r);
                    const R lpo = lp * m_;
                    ll += (double)lpo;
                    if (log_p_x) log_p_x[o] = lpo;
                    if (log_p_x_missing) log_p_x_missing[o] = lp * (R(1) - m_);
                    if (recon_mean) recon_mean[o] = rmean;
                    if (recon_mode) recon_mode[o] = rmode;
                    if (data_tr) data_tr[o] = dtr;
                }
            };
            if (!v.ok) {
                rows([&](const TD*, R*, bool, R& lp, R& rmean, R& rmode, R& dtr, int) { lp = rmean = rmode = dtr = (R)NAN; });
            } else if (v.kind == HLVAE_VAR_REAL) {
                rows([&](const TD* x, R* t, bool, R& lp, R& rmean, R& rmode, R& dtr, int) {     // loglik.py:27-70
                    const R x0 = (R)x[0];
                    const R mean = v.snv * t[0] + v.nm;
                    const R rr = x0 * v.idiv - mean;
                    lp = R(-0.5) * rr * rr * v.ivar - v.lconst;
                    t[0] = mean;
                    rmean = mean; rmode = mean;
                    dtr = x0;
                });
            } else if (v.kind == HLVAE_VAR_CAT && v.C == 5) {
                rows([&](const TD* x, R* t, bool, R& lp, R& rmean, R& rmode, R& dtr, int) {
                    cat_forward<R, TD, 5>(x, t, 5, lp, rmean, dtr);
                    rmode = rmean;
                });
            } else if (v.kind == HLVAE_VAR_POS) {
                rows([&](const TD* x, R* t, bool obs, R& lp, R& rmean, R& rmode, R& dtr, int) {
                    var_forward<R, TD, HLVAE_VAR_POS>(v, x, t, obs, lp, rmean, rmode, dtr);
                });
            } else if (v.kind == HLVAE_VAR_COUNT) {
                rows([&](const TD* x, R* t, bool obs, R& lp, R& rmean, R& rmode, R& dtr, int) {
                    var_forward<R, TD, HLVAE_VAR_COUNT>(v, x, t, obs, lp, rmean, rmode, dtr);
                });
            } else if (v.kind == HLVAE_VAR_CAT) {
                rows([&](const TD* x, R* t, bool obs, R& lp, R& rmean, R& rmode, R& dtr, int) {
                    var_forward<R, TD, HLVAE_VAR_CAT>(v, x, t, obs, lp, rmean, rmode, dtr);
                });
            } else {
                auto ordinal = [&](auto cfix) {
                    rows([&](const TD* x, R* t, bool obs, R& lp, R& rmean, R& rmode, R& dtr, int r) {
                        var_forward<R, TD, HLVAE_VAR_ORDINAL, decltype(cfix)::value>(v, x, t, obs, lp, rmean, rmode, dtr);
                        if constexpr (sizeof(R) == 4) {
                            if (rmean < R(0)) {              // float32 could not prove the argmax: redo from theta
                                const R am = (R)ord_argmax_f64(reinterpret_cast<const float*>(theta) +
                                                               (n0 + r) * ld_theta + ps0 + v.po, v.C);
                                rmean = am; rmode = am;
                            }
                        }
                    });
                };
                if (v.C == 5) ordinal(std::integral_constant<int, 5>{});
                else ordinal(std::integral_constant<int, 0>{});
            }
        }
        if (use_tma) fence_async_smem();         // params written through the generic proxy, read by the bulk store
        __syncthreads();
        if (params) {
            if (uni_t) {
                tma_unstage_batch<TS>(params + n0 * ld_theta + ps0, ld_theta, sT, capt, shift_t, span_p, nr, tid);
                if (tid % (LL_THREADS / LL_ROWS) == 0) {     // the issuing threads
                    bulk_commit();
                    bulk_wait_read();            // the staged rows have been read; the stage may be refilled
                }
            } else if (use_tma) {
                for (int r = 0; r < nr; r++)
                    tma_unstage_row<TS>(params + (n0 + r) * ld_theta + ps0, sT + r * capt, sShiftT[stg][r], span_p, tid, r);
                if (tid < nr) {                  // thread r issued row r
                    bulk_commit();
                    bulk_wait_read();            // the staged rows have been read; the stage may be refilled
                }
            } else {
                for (int r = 0; r < nr; r++)
                    unstage_row<TS>(params + (n0 + r) * ld_theta + ps0, sT + r * capt, sShiftT[stg][r], span_p, tid, t_ok);
            }
        }
        __syncthreads();                         // this stage is refilled by the prefetch of the next iteration
    }
    cp_async_wait_all();
    if (ll_total) {
        ll = warp_sum(ll);
        if ((tid & 31) == 0) red[tid >> 5] = ll;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < LL_THREADS / 32; w++) s += red[w];
            if (s != 0.0) atomicAdd(ll_total, s);
        }
    }
}

// Ordinal variable, backward (loglik.py:149-188 differentiated): theta is overwritten with
// g * d log_p_x / d theta.  CMAX is the unroll bound; when it equals C everything stays in registers.
template <typename R, typename XT, int CMAX>
__device__ __forceinline__ void ord_backward(const XT* __restrict__ x_, R* __restrict__ t, int C, bool observed, R g) {
    struct { const XT* p; __device__ __forceinline__ R operator[](int c) const { return (R)p[c]; } } x{x_};
    const R eps = R(1e-6);
    // softplus and its derivative of every logit from ONE exponential each (the second loop below used to evaluate
    // both again): softplus = log1p(e^t), softplus' = e^t / (1 + e^t), t > 20 -> (t, 1)
    auto sp_dsp = [](R tc, R& sp, R& dsp) {
        const R et = Mth<R>::ex(fmin(tc, R(30)));
        sp = tc > R(20) ? tc : Mth<R>::lg1p_nn(et);
        dsp = tc > R(20) ? R(1) : et * Mth<R>::rcp(R(1) + et);
    };
    const R t_loc = t[C - 1];
    R loc, dloc;
    sp_dsp(t_loc, loc, dloc);
    R dsg[CMAX], q[CMAX];      // dsg_c = sigma'(u_c) = sigma(u_c) sigma(-u_c)
    R spv[CMAX], dspv[CMAX];
    R cum = R(0), prev = R(0), tot = R(0), sneg_prev = R(1);
    int vals = 0;
#pragma unroll
    for (int c = 0; c < CMAX; c++)
        if (c < C) vals += (int)x[c];
    if (!observed) vals = 1;
    const int y = vals - 1;
    // float32: the derivative of log p_y needs sigma'(u_y) / p_y and -sigma'(u_{y-1}) / p_y - two terms of size 1 / p_y
    // whose SUM (what every delta_c, c < y, and loc see) cancels to O(1): with p_y = sigma(u_y) - sigma(u_{y-1}),
    //   (sigma'(u_y) - sigma'(u_{y-1})) / p_y = sigma(-u_{y-1}) - sigma(u_y)      (u_{-1} = -inf, u_{C-1} = +inf),
    // so the pair enters the running sums in this closed form (measured: relative error of d theta 2e-3 -> 1e-6 at
    // logits ~ N(0, 16)); the normalisation term -log(sum p) and the float64 path keep the general form below
    R sg_y = R(1), ds_y = R(0), sneg_ym1 = R(1), q_y = R(1);
#pragma unroll
    for (int c = 0; c < CMAX; c++) {
        spv[c] = R(0);
        dspv[c] = R(0);
        if (c < C) {
            R sg = R(1), qc, ds = R(0);
            if (c < C - 1) {
                sp_dsp(t[c], spv[c], dspv[c]);
                const R delta = clamp_<R>(spv[c], eps, R(1e20));
                cum += delta;
                const R e = Mth<R>::ex(loc - cum);                      // exp(-u_c)
                sg = Mth<R>::rcp(R(1) + e);
                const R sneg = (e > R(1e30)) ? R(1) : e * sg;           // sigma(-u_c), no cancellation
                ds = sg * sneg;
                if constexpr (sizeof(R) == 4) {
                    qc = (c == 0) ? sg : sg * sneg_prev * (spv[c] < eps ? R(9.999995e-7) : dspv[c]);   // 1 - exp(-delta)
                } else {
                    qc = sg - prev;
                }
                sneg_prev = sneg;
            } else {
                if constexpr (sizeof(R) == 4) qc = sneg_prev; else qc = sg - prev;
            }
            dsg[c] = ds;
            q[c] = qc;
            prev = sg;
            tot += clamp_<R>(qc, eps, R(1));
            if (c == y) { sg_y = sg; ds_y = ds; q_y = qc; }
            if (c == y - 1) sneg_ym1 = sneg_prev;                        // (updated above: sigma(-u_c))
        }
    }
    const bool y_in = q_y >= eps && q_y <= R(1);
    const R a_y = y_in ? Mth<R>::rcp(q_y) * ds_y : R(0);                 // d log p_y / d u_y
    const R kappa = y_in ? sneg_ym1 - sg_y : R(0);                       // d log p_y / d u_y + d log p_y / d u_{y-1}
    constexpr bool F32 = sizeof(R) == 4;
    // lp = log p_y - log tot ; the clamp passes gradient inside [eps, 1]
    // q_c = sg_c - sg_{c-1}: d/dsg_c = gq_c - gq_{c+1}, c < C-1; u_c = cum_c - loc
    const R itot = Mth<R>::rcp(tot);
    R g_loc = R(0), run = R(0), gq_next = R(0);
    {
        const R qc = q[C - 1];
        const R pc = clamp_<R>(qc, eps, R(1));
        const R gp = ((!F32 && C - 1 == y) ? Mth<R>::rcp(pc) : R(0)) - itot;
        gq_next = (qc >= eps && qc <= R(1)) ? gp : R(0);
    }
#pragma unroll
    for (int c = CMAX - 2; c >= 0; c--) {
        if (c <= C - 2) {
            const R qc = q[c];
            const R pc = clamp_<R>(qc, eps, R(1));
            const R gp = ((!F32 && c == y) ? Mth<R>::rcp(pc) : R(0)) - itot;
            const R gqc = (qc >= eps && qc <= R(1)) ? gp : R(0);
            const R gu = (gqc - gq_next) * dsg[c];
            gq_next = gqc;
            g_loc -= gu;
            run += gu;                                   // reverse cumulative sum -> d/d delta_c
            R run_c = run;
            if (F32) run_c += (c == y) ? a_y : (c < y ? kappa : R(0));
            const R sp = spv[c];
            const R ga = (sp >= eps && sp <= R(1e20)) ? run_c * dspv[c] : R(0);
            t[c] = g * ga;
        }
    }
    if (F32) g_loc -= kappa;
    t[C - 1] = g * g_loc * dloc;

}

// ------------------------------------------------------------------------------------
// One variable, backward: d log_p_x / d theta (times g) into t (in place of theta), and the
// derivative w.r.t. the raw log-variance parameter (real / pos).
template <typename R, typename XT, int KIND = -1, int CFIX = 0>
__device__ __forceinline__ void var_backward(const VarC<R>& v, const XT* __restrict__ x_, R* __restrict__ t, bool observed,
                                             R g, R& ge) {
    struct { const XT* p; __device__ __forceinline__ R operator[](int c) const { return (R)p[c]; } } x{x_};
    const int C = CFIX > 0 ? CFIX : v.C;
    const int kind = KIND >= 0 ? KIND : v.kind;
    if (kind == HLVAE_VAR_REAL) {
        const R r = x[0] * v.idiv - (v.snv * t[0] + v.nm);
        t[0] = g * v.snv * r * v.ivar;
        ge = g * (R(0.5) * r * r * v.ivar - R(0.5)) * v.sg8;
    } else if (kind == HLVAE_VAR_POS) {
        const R r = Mth<R>::lg1p(x[0]) - (v.snv * t[0] + v.nm);
        t[0] = g * v.snv * r * v.ivar;
        ge = g * (R(0.5) * r * r * v.ivar - R(0.5));
    } else if (kind == HLVAE_VAR_COUNT) {
        const R t0 = t[0];
        const R sp = softplus_<R>(t0);
        R gl = R(0);
        if (sp >= R(1e-6) && sp <= R(1e20)) gl = (x[0] * Mth<R>::rcp(sp) - R(1)) * dsoftplus_<R>(t0);
        t[0] = g * gl;
    } else if (kind == HLVAE_VAR_CAT) {
        switch (C) {
            case 2: cat_backward<R, XT, 2>(x_, t, 2, g); break;
            case 3: cat_backward<R, XT, 3>(x_, t, 3, g); break;
            case 4: cat_backward<R, XT, 4>(x_, t, 4, g); break;
            case 5: cat_backward<R, XT, 5>(x_, t, 5, g); break;
            case 6: cat_backward<R, XT, 6>(x_, t, 6, g); break;
            case 8: cat_backward<R, XT, 8>(x_, t, 8, g); break;
            case 10: cat_backward<R, XT, 10>(x_, t, 10, g); break;
            default: cat_backward<R, XT, HLVAE_MAX_CLASS>(x_, t, C, g); break;
        }
    } else {
        switch (C) {
            case 2: ord_backward<R, XT, 2>(x_, t, 2, observed, g); break;
            case 3: ord_backward<R, XT, 3>(x_, t, 3, observed, g); break;
            case 4: ord_backward<R, XT, 4>(x_, t, 4, observed, g); break;
            case 5: ord_backward<R, XT, 5>(x_, t, 5, observed, g); break;
            case 6: ord_backward<R, XT, 6>(x_, t, 6, observed, g); break;
            case 8: ord_backward<R, XT, 8>(x_, t, 8, observed, g); break;
            case 10: ord_backward<R, XT, 10>(x_, t, 10, observed, g); break;
            default: ord_backward<R, XT, HLVAE_MAX_CLASS>(x_, t, C, observed, g); break;
        }
    }
}

// (8 resident CTAs at 64 registers: 0.178 -> 0.172 ms at configs[1]; the forward kernel loses at that setting,
// 0.273 -> 0.292 ms, and stays at 6 x 80)
// Rows per batch of the backward kernel.  With float data the stage of 8 rows is 47 KB (4 CTAs = 16 warps per SM, and the
// single-stage pipeline leaves the SM idle while a CTA waits for its copies): 4 rows -> 8 CTAs per SM, tabular
// 64 000 rows 0.253 -> 0.213 ms.  uint8 data (D4: 30 KB, 6+ CTAs): 8 rows stay faster (0.178 vs 0.192 ms); so does the
// forward kernel on both layouts (0.250 vs 0.266, 0.276 vs 0.308 ms).
template <typename TS, typename TD>
struct BwdRows {
    static constexpr int value = (sizeof(TD) >= 4 && LL_ROWS > 4) ? 4 : LL_ROWS;
};

template <typename TS, typename TD, typename TM>
__global__ void __launch_bounds__(LL_THREADS, sizeof(TS) == 4 ? 8 : 3)
loglik_bwd_k(int64_t N, int D, int tile_vars, int cap, int64_t ld_data, int64_t ld_theta,
             const int32_t* __restrict__ var_kind, const int32_t* __restrict__ var_nclass,
             const int32_t* __restrict__ var_dcol, const int32_t* __restrict__ var_pcol,
             const double* __restrict__ vparam, const TD* __restrict__ data, const TS* __restrict__ theta,
             const TM* __restrict__ mask, const TS* __restrict__ g_lp, const double* __restrict__ g_scalar,
             TS* __restrict__ g_theta, double* __restrict__ g_lvy) {
    using R = TS;
    constexpr int LR = BwdRows<TS, TD>::value;
    extern __shared__ __align__(16) unsigned char ll_smem[];
    const int capt = cap + 8, capx = cap + 16;
    // per stage: sT [LR][capt] theta (overwritten with g_theta), sG [LR][LL_CAPM] upstream gradient,
    // sX [LR][capx] data, sK [LR][LL_CAPM] mask
    const size_t stage_bytes = (size_t)LR * ((capt + LL_CAPM) * sizeof(R) + capx * sizeof(TD) + LL_CAPM * sizeof(TM));
    __shared__ int sShift[LL_STAGES][LR], sShiftT[LL_STAGES][LR], sShiftM[LL_STAGES][LR],
        sShiftG[LL_STAGES][LR];
    const bool m_ok = (((uintptr_t)mask | (uintptr_t)(N * D * (int64_t)sizeof(TM))) & 15) == 0;
    const bool g_ok = (((uintptr_t)g_lp | (uintptr_t)(N * D * (int64_t)sizeof(TS))) & 15) == 0;
    const int span_m = min(D, d0_(blockIdx.x, tile_vars) + tile_vars) - d0_(blockIdx.x, tile_vars);
    const bool x_ok = (((uintptr_t)data | (uintptr_t)(N * ld_data * (int64_t)sizeof(TD))) & 15) == 0;
    const bool t_ok = (((uintptr_t)theta | (uintptr_t)g_theta | (uintptr_t)(N * ld_theta * (int64_t)sizeof(TS))) & 15) == 0;
    const int tid = threadIdx.x;
    const int d0 = blockIdx.x * tile_vars;
    const int d1 = min(D, d0 + tile_vars);
    const int lt = tile_variable_by_type(tid, d0, d1 - d0, var_kind);   // this thread's variable inside the tile
    const int d = d0 + lt;
    const bool active = d < d1;
    const int xs0 = var_dcol[d0], ps0 = var_pcol[d0];
    const int span_x = max(0, min(cap, var_dcol[d1 - 1] + var_nclass[d1 - 1] - xs0));
    const int span_p = max(0, min(cap, var_pcol[d1 - 1] + var_nclass[d1 - 1] - ps0));
    const VarC<R> v = load_var<R>(d, D, active, var_kind, var_nclass, var_dcol, var_pcol, vparam, xs0, ps0, cap);
    const R gs = g_scalar ? (R)(*g_scalar) : R(0);
    const int64_t stride = (int64_t)gridDim.y * LR;
    __shared__ __align__(8) unsigned long long mbar[LL_STAGES];
    const bool use_tma = x_ok && t_ok && m_ok && (!g_lp || g_ok);
    const bool uni_t = use_tma && ld_theta % (16 / (int)sizeof(TS)) == 0;
    const int shift_t = (int)((reinterpret_cast<uintptr_t>(theta + ps0) & 15) / sizeof(TS));
    const bool uni = uni_t && ld_data % (16 / (int)sizeof(TD)) == 0 && D % (16 / (int)sizeof(TM)) == 0 &&
                     (!g_lp || D % (16 / (int)sizeof(TS)) == 0);
    const int shift_x = (int)((reinterpret_cast<uintptr_t>(data + xs0) & 15) / sizeof(TD));
    const int shift_m = (int)((reinterpret_cast<uintptr_t>(mask + d0) & 15) / sizeof(TM));
    const int shift_g = g_lp ? (int)((reinterpret_cast<uintptr_t>(g_lp + d0) & 15) / sizeof(TS)) : 0;
    unsigned phase[LL_STAGES];
#pragma unroll
    for (int i = 0; i < LL_STAGES; i++) phase[i] = 0;
    if (use_tma) {
        if (tid == 0)
            for (int i = 0; i < LL_STAGES; i++) mbar_init(&mbar[i], 1);
        __syncthreads();
    }

    auto prefetch = [&](int64_t n0, int stg) {
        if (n0 < N && use_tma) {
            if (tid < 32) {        // warp 0: lane = (array, row); every lane issues its own bulk copy
                R* sT = reinterpret_cast<R*>(ll_smem + stg * stage_bytes);
                R* sG = sT + LR * capt;
                TD* sX = reinterpret_cast<TD*>(sG + LR * LL_CAPM);
                TM* sK = reinterpret_cast<TM*>(sX + LR * capx);
                const int nr = (int)min((int64_t)LR, N - n0);
                const int r = tid % LR, arr = tid / LR;
                const bool job = r < nr && (arr < 3 || (arr == 3 && g_lp != nullptr));
                unsigned bytes = 0;
                int shift = 0;
                const void* src = nullptr;
                void* dst = nullptr;
                if (r < nr && arr == 3 && !g_lp) sShiftG[stg][r] = 0;
                if (job) {
                    if (arr == 0) {
                        const TS* p_ = theta + (n0 + r) * ld_theta + ps0;
                        shift = tma_row_bytes<TS>(p_, span_p, bytes);
                        src = p_ - shift; dst = sT + r * capt; sShiftT[stg][r] = shift;
                    } else if (arr == 1) {
                        const TD* p_ = data + (n0 + r) * ld_data + xs0;
                        shift = tma_row_bytes<TD>(p_, span_x, bytes);
                        src = p_ - shift; dst = sX + r * capx; sShift[stg][r] = shift;
                    } else if (arr == 2) {
                        const TM* p_ = mask + (n0 + r) * D + d0;
                        shift = tma_row_bytes<TM>(p_, span_m, bytes);
                        src = p_ - shift; dst = sK + r * LL_CAPM; sShiftM[stg][r] = shift;
                    } else {
                        const TS* p_ = g_lp + (n0 + r) * D + d0;
                        shift = tma_row_bytes<TS>(p_, span_m, bytes);
                        src = p_ - shift; dst = sG + r * LL_CAPM; sShiftG[stg][r] = shift;
                    }
                }
                const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
                if (tid == 0) mbar_expect_tx(&mbar[stg], total);
                __syncwarp();
                if (job) bulk_g2s(dst, src, bytes, &mbar[stg]);
            }
        } else if (n0 < N) {
            R* sT = reinterpret_cast<R*>(ll_smem + stg * stage_bytes);
            R* sG = sT + LR * capt;
            TD* sX = reinterpret_cast<TD*>(sG + LR * LL_CAPM);
            TM* sK = reinterpret_cast<TM*>(sX + LR * capx);
            const int nr = (int)min((int64_t)LR, N - n0);
            for (int r = 0; r < nr; r++) {
                const int st_ = stage_row<TS>(sT + r * capt, theta + (n0 + r) * ld_theta + ps0, span_p, tid, t_ok);
                const int sh = stage_row<TD>(sX + r * capx, data + (n0 + r) * ld_data + xs0, span_x, tid, x_ok);
                const int sm = stage_row<TM>(sK + r * LL_CAPM, mask + (n0 + r) * D + d0, span_m, tid, m_ok);
                int sg = 0;
                if (g_lp) sg = stage_row<TS>(sG + r * LL_CAPM, g_lp + (n0 + r) * D + d0, span_m, tid, g_ok);
                if (tid == 0) { sShift[stg][r] = sh; sShiftT[stg][r] = st_; sShiftM[stg][r] = sm; sShiftG[stg][r] = sg; }
            }
        }
        cp_async_commit();
    };

    double ge_acc = 0.0;
    int stg = 0;
    if (LL_STAGES == 2) prefetch((int64_t)blockIdx.y * LR, 0);
    for (int64_t n0 = (int64_t)blockIdx.y * LR; n0 < N; n0 += stride, stg ^= (LL_STAGES - 1)) {
        if (LL_STAGES == 2) {
            prefetch(n0 + stride, stg ^ 1);
            cp_async_wait_group<1>();
        } else {
            prefetch(n0, 0);
            cp_async_wait_all();
        }
        __syncthreads();
        if (use_tma) {
            mbar_wait(&mbar[stg], phase[stg]);
            phase[stg] ^= 1;
        }
        R* sT = reinterpret_cast<R*>(ll_smem + stg * stage_bytes);
        R* sG = sT + LR * capt;
        TD* sX = reinterpret_cast<TD*>(sG + LR * LL_CAPM);
        TM* sK = reinterpret_cast<TM*>(sX + LR * capx);
        const int nr = (int)min((int64_t)LR, N - n0);
        if (active) {
            // one row loop per evaluator (the type of a thread's variable is fixed): see loglik_fwd_k
            auto rows = [&](auto&& eval) {
#pragma unroll 1
                for (int r = 0; r < nr; r++) {
                    const int sm_ = uni ? shift_m : sShiftM[stg][r];
                    const int sx_ = uni ? shift_x : sShift[stg][r];
                    const int st_ = uni ? shift_t : sShiftT[stg][r];
                    const R m_ = (R)sK[r * LL_CAPM + sm_ + lt];
                    const R g_ = gs + (g_lp ? sG[r * LL_CAPM + (uni ? shift_g : sShiftG[stg][r]) + lt] : R(0));
                    R ge = R(0);
                    eval(sX + r * capx + sx_ + v.xo, sT + r * capt + st_ + v.po, m_ != R(0), g_ * m_, ge);
                    ge_acc += (double)ge;
                }
            };
            if (!v.ok) {
                for (int r = 0; r < nr; r++)
                    for (int c = 0; c < v.C && v.po >= 0 && v.po + c < cap; c++)
                        sT[r * capt + sShiftT[stg][r] + v.po + c] = (R)NAN;
            } else if (v.kind == HLVAE_VAR_REAL) {
                rows([&](const TD* x, R* t, bool, R g, R& ge) {
                    const R rr = (R)x[0] * v.idiv - (v.snv * t[0] + v.nm);
                    t[0] = g * v.snv * rr * v.ivar;
                    ge = g * (R(0.5) * rr * rr * v.ivar - R(0.5)) * v.sg8;
                });
            } else if (v.kind == HLVAE_VAR_CAT && v.C == 5) {
                rows([&](const TD* x, R* t, bool, R g, R&) { cat_backward<R, TD, 5>(x, t, 5, g); });
            } else if (v.kind == HLVAE_VAR_POS) {
                rows([&](const TD* x, R* t, bool obs, R g, R& ge) { var_backward<R, TD, HLVAE_VAR_POS>(v, x, t, obs, g, ge); });
            } else if (v.kind == HLVAE_VAR_COUNT) {
                rows([&](const TD* x, R* t, bool obs, R g, R& ge) { var_backward<R, TD, HLVAE_VAR_COUNT>(v, x, t, obs, g, ge); });
            } else if (v.kind == HLVAE_VAR_CAT) {
                rows([&](const TD* x, R* t, bool obs, R g, R& ge) { var_backward<R, TD, HLVAE_VAR_CAT>(v, x, t, obs, g, ge); });
            } else if (v.C == 5) {
                rows([&](const TD* x, R* t, bool obs, R g, R& ge) { var_backward<R, TD, HLVAE_VAR_ORDINAL, 5>(v, x, t, obs, g, ge); });
            } else {
                rows([&](const TD* x, R* t, bool obs, R g, R& ge) { var_backward<R, TD, HLVAE_VAR_ORDINAL>(v, x, t, obs, g, ge); });
            }
        }
        if (use_tma) fence_async_smem();
        __syncthreads();
        if (uni_t) {
            tma_unstage_batch<TS, LR>(g_theta + n0 * ld_theta + ps0, ld_theta, sT, capt, shift_t, span_p, nr, tid);
            if (tid % (LL_THREADS / LR) == 0) {
                bulk_commit();
                bulk_wait_read();
            }
        } else if (use_tma) {
            for (int r = 0; r < nr; r++)
                tma_unstage_row<TS>(g_theta + (n0 + r) * ld_theta + ps0, sT + r * capt, sShiftT[stg][r], span_p, tid, r);
            if (tid < nr) {
                bulk_commit();
                bulk_wait_read();
            }
        } else {
            for (int r = 0; r < nr; r++)
                unstage_row<TS>(g_theta + (n0 + r) * ld_theta + ps0, sT + r * capt, sShiftT[stg][r], span_p, tid, t_ok);
        }
        __syncthreads();
    }
    cp_async_wait_all();
    if (g_lvy && active && (v.kind == HLVAE_VAR_REAL || v.kind == HLVAE_VAR_POS) && ge_acc != 0.0)
        atomicAdd(g_lvy + d, ge_acc);
}

// ------------------------------------------------------------------------------------
// Stand-alone monitoring transforms for callers that hold `params` / `data` only
// (training.py:84-91 calls them separately from the likelihood).
template <typename TS>
__global__ void __launch_bounds__(256)
statistics_k(int64_t N, int D, int64_t ld_theta, const int32_t* __restrict__ var_kind,
             const int32_t* __restrict__ var_nclass, const int32_t* __restrict__ var_pcol,
             const double* __restrict__ vparam, const TS* __restrict__ params, TS* __restrict__ mean,
             TS* __restrict__ mode) {
    const int d = blockIdx.x * 256 + threadIdx.x;
    if (d >= D) return;
    const int kind = var_kind[d], C = var_nclass[d], pcol = var_pcol[d];
    const double v = exp(vparam[2 * D + d]);                                 // read_functions.py:284
    for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
        const TS* p = params + n * ld_theta + pcol;
        double a, b;
        if (kind == HLVAE_VAR_REAL) {
            a = b = (double)p[0];                                            // :275-278
        } else if (kind == HLVAE_VAR_POS) {
            a = exp((double)p[0] + 0.5 * v) - 1.0;                           // :287
            b = exp((double)p[0] - v) - 1.0;                                 // :289
        } else if (kind == HLVAE_VAR_COUNT) {
            a = (double)p[0];
            b = floor(a);                                                    // :293-295
        } else {
            int am = 0;
            TS best = p[0];
            for (int c = 1; c < C; c++)
                if (p[c] > best) { best = p[c]; am = c; }                    // :296-302, first index on ties
            a = b = am;
        }
        mean[n * D + d] = (TS)a;
        mode[n * D + d] = (TS)b;
    }
}

template <typename TS>
__global__ void __launch_bounds__(256)
discrete_transform_k(int64_t N, int D, int64_t ld_data, const int32_t* __restrict__ var_kind,
                     const int32_t* __restrict__ var_nclass, const int32_t* __restrict__ var_dcol,
                     const TS* __restrict__ data, TS* __restrict__ out) {
    const int d = blockIdx.x * 256 + threadIdx.x;
    if (d >= D) return;
    const int kind = var_kind[d], C = var_nclass[d], dcol = var_dcol[d];
    for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
        const TS* x = data + n * ld_data + dcol;
        double r;
        if (kind == HLVAE_VAR_CAT) {                                         // read_functions.py:224-227
            int am = 0;
            TS best = x[0];
            for (int c = 1; c < C; c++)
                if (x[c] > best) { best = x[c]; am = c; }
            r = am;
        } else if (kind == HLVAE_VAR_ORDINAL) {                              // :228-230
            double s = 0.0;
            for (int c = 0; c < C; c++) s += (double)x[c];
            r = s - 1.0;
        } else {
            r = (double)x[0];                                                // :233
        }
        out[n * D + d] = (TS)r;
    }
}

bool args_ok(int64_t N, int D, const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* d,
             const double* vp, const void* data, const void* theta, const void* mask, int dtype, int data_dtype,
             int mask_dtype) {
    if (!(N >= 0 && D > 0 && a && b && c && d && vp && data && theta && mask)) return false;
    if (dtype != HLVAE_F32 && dtype != HLVAE_F64) return false;
    if (data_dtype != dtype && data_dtype != HLVAE_U8) return false;
    if (mask_dtype != dtype && mask_dtype != HLVAE_U8) return false;
    return true;
}

struct Tiling {
    int tile_vars, n_tiles;
    unsigned rows;
};

// Variable tiles of equal size (<= LL_THREADS) and LL_WAVES times as many row stripes as stay resident at once.
// Measured (waves 1 / 2 / 4 / 6 / 8 / 12 / 25 at 16 000 rows, D4: forward 0.376 / 0.368 / 0.362 / 0.357 / 0.354 /
// 0.354 / 0.367 ms; tabular D = 256 at 64 000 rows: 0.46 -> 0.39 ms forward, 0.51 -> 0.40 ms backward at 6): shorter
// CTAs let the SMs that finish early pick up more work instead of idling through the tail of a single wave.
// (re-measured at the end of r02, 4 / 6 / 8 / 10: D4 forward 0.283 / 0.277 / 0.271 / 0.267 ms, backward 0.181 / 0.179 / 0.179 /
// 0.183 ms; tabular forward 0.244 / 0.235 / 0.236 / 0.238 ms, backward 0.194 / 0.192 / 0.195 / 0.199 ms: 6 stays)
#ifndef HLVAE_LL_WAVES
#define HLVAE_LL_WAVES 6
#endif
constexpr int LL_WAVES = HLVAE_LL_WAVES;
Tiling make_tiling(int64_t N, int D, int ctas_per_sm, int rows_per_batch = LL_ROWS, int max_tile_vars = LL_THREADS) {
    N = (N + rows_per_batch - 1) / rows_per_batch;            // row batches
    Tiling t;
    t.n_tiles = (D + max_tile_vars - 1) / max_tile_vars;
    t.tile_vars = (D + t.n_tiles - 1) / t.n_tiles;
    t.n_tiles = (D + t.tile_vars - 1) / t.tile_vars;
    int64_t want = ((int64_t)148 * ctas_per_sm * LL_WAVES) / t.n_tiles;
    if (want > N) want = N;
    if (want > 65535) want = 65535;
    if (want < 1) want = 1;
    t.rows = (unsigned)want;
    return t;
}

// Variables per tile such that the stage of one CTA (rows x (theta + data [+ upstream gradient]) of the tile's widest
// possible span, tile_vars * max_class columns) stays within LL_SMEM_BUDGET: float64 storage with 16-class variables
// would otherwise ask for 272 KB (more than an SM has) - the tile shrinks instead.  Every layout of the reference's
// data sets (<= 5 classes) keeps the full 128-variable tile.
constexpr size_t LL_SMEM_BUDGET = 96 * 1024;                  // two CTAs per SM at least
int max_tile_vars(int rows, int max_class, size_t col_bytes, size_t var_bytes) {
    int tv = LL_THREADS;
    while (tv > 8 && (size_t)LL_STAGES * rows * (((size_t)tv * max_class + 16) * col_bytes + LL_CAPM * var_bytes) > LL_SMEM_BUDGET)
        tv -= 8;
    return tv;
}

template <typename K>
int set_smem(K kern, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

#define HLVAE_LL_DISPATCH(CALL)                                                                      \
    if (dtype == HLVAE_F64) {                                                                        \
        if (data_dtype == HLVAE_U8) {                                                                \
            if (mask_dtype == HLVAE_U8) { CALL(double, uint8_t, uint8_t); } else { CALL(double, uint8_t, double); } \
        } else {                                                                                     \
            if (mask_dtype == HLVAE_U8) { CALL(double, double, uint8_t); } else { CALL(double, double, double); }   \
        }                                                                                            \
    } else {                                                                                         \
        if (data_dtype == HLVAE_U8) {                                                                \
            if (mask_dtype == HLVAE_U8) { CALL(float, uint8_t, uint8_t); } else { CALL(float, uint8_t, float); }    \
        } else {                                                                                     \
            if (mask_dtype == HLVAE_U8) { CALL(float, float, uint8_t); } else { CALL(float, float, float); }        \
        }                                                                                            \
    }

extern "C" int hlvae_loglik_fwd(int64_t N, int D, int64_t ld_data, int64_t ld_theta, const int32_t* var_kind,
                                const int32_t* var_nclass, const int32_t* var_dcol, const int32_t* var_pcol,
                                const double* vparam, const void* data, const void* theta, const void* mask,
                                int dtype, int data_dtype, int mask_dtype, int max_class, void* log_p_x,
                                void* log_p_x_missing, void* params, void* recon_mean, void* recon_mode,
                                void* data_tr, double* ll_total, void* stream) {
    if (!args_ok(N, D, var_kind, var_nclass, var_dcol, var_pcol, vparam, data, theta, mask, dtype, data_dtype,
                 mask_dtype))
        return HLVAE_E_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (max_class < 1 || max_class > HLVAE_MAX_CLASS) return HLVAE_E_ARG;
    const size_t esz = dtype == HLVAE_F64 ? 8 : 4;
    const size_t xsz = data_dtype == HLVAE_U8 ? 1 : esz;
    const size_t msz = mask_dtype == HLVAE_U8 ? 1 : esz;
    const int tv_max = max_tile_vars(LL_ROWS, max_class, esz + xsz, msz);
    Tiling tl = make_tiling(N, D, 1, LL_ROWS, tv_max);
    const int cap = (tl.tile_vars * max_class + 15) & ~15;
    const size_t smem = (size_t)LL_STAGES * LL_ROWS * ((cap + 8) * esz + (cap + 16) * xsz + LL_CAPM * msz);
#define HLVAE_LL_FWD(TS, TD, TM)                                                                                     \
    {                                                                                                                \
        auto kern = loglik_fwd_k<TS, TD, TM>;                                                                        \
        int rc = set_smem(kern, smem);                                                                               \
        if (rc) return rc;                                                                                           \
        int nb = 1;                                                                                                  \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, LL_THREADS, smem);                                  \
        tl = make_tiling(N, D, nb < 1 ? 1 : nb, LL_ROWS, tv_max);                                                                   \
        dim3 grid(tl.n_tiles, tl.rows);                                                                              \
        kern<<<grid, LL_THREADS, smem, st>>>(N, D, tl.tile_vars, cap, ld_data, ld_theta, var_kind, var_nclass,       \
                                             var_dcol, var_pcol, vparam, (const TD*)data, (const TS*)theta,          \
                                             (const TM*)mask, (TS*)log_p_x, (TS*)log_p_x_missing, (TS*)params, (TS*)recon_mean,       \
                                             (TS*)recon_mode, (TS*)data_tr, ll_total);                               \
    }
    HLVAE_LL_DISPATCH(HLVAE_LL_FWD)
#undef HLVAE_LL_FWD
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_loglik_bwd(int64_t N, int D, int64_t ld_data, int64_t ld_theta, const int32_t* var_kind,
                                const int32_t* var_nclass, const int32_t* var_dcol, const int32_t* var_pcol,
                                const double* vparam, const void* data, const void* theta, const void* mask,
                                int dtype, int data_dtype, int mask_dtype, int max_class, const void* g_lp,
                                const double* g_scalar, void* g_theta, double* g_lvy, void* stream) {
    if (!args_ok(N, D, var_kind, var_nclass, var_dcol, var_pcol, vparam, data, theta, mask, dtype, data_dtype,
                 mask_dtype) ||
        !g_theta || (!g_lp && !g_scalar))
        return HLVAE_E_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (max_class < 1 || max_class > HLVAE_MAX_CLASS) return HLVAE_E_ARG;
    const size_t esz = dtype == HLVAE_F64 ? 8 : 4;
    const size_t xsz = data_dtype == HLVAE_U8 ? 1 : esz;
    const size_t msz = mask_dtype == HLVAE_U8 ? 1 : esz;
    // (the tile is sized for the 8-row stage; instantiations that stage 4 rows just use half the budget)
    const int tv_max = max_tile_vars(LL_ROWS, max_class, esz + xsz, msz + esz);
    Tiling tl = make_tiling(N, D, 1, LL_ROWS, tv_max);
    const int cap = (tl.tile_vars * max_class + 15) & ~15;
#define HLVAE_LL_BWD(TS, TD, TM)                                                                                     \
    {                                                                                                                \
        auto kern = loglik_bwd_k<TS, TD, TM>;                                                                        \
        constexpr int LR = BwdRows<TS, TD>::value;                                                                   \
        const size_t smem = (size_t)LL_STAGES * LR * ((cap + 8 + LL_CAPM) * esz + (cap + 16) * xsz + LL_CAPM * msz); \
        int rc = set_smem(kern, smem);                                                                               \
        if (rc) return rc;                                                                                           \
        int nb = 1;                                                                                                  \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, LL_THREADS, smem);                                  \
        tl = make_tiling(N, D, nb < 1 ? 1 : nb, LR, tv_max);                                                                   \
        dim3 grid(tl.n_tiles, tl.rows);                                                                              \
        kern<<<grid, LL_THREADS, smem, st>>>(N, D, tl.tile_vars, cap, ld_data, ld_theta, var_kind, var_nclass,       \
                                             var_dcol, var_pcol, vparam, (const TD*)data, (const TS*)theta,          \
                                             (const TM*)mask, (const TS*)g_lp, g_scalar, (TS*)g_theta, g_lvy);                        \
    }
    HLVAE_LL_DISPATCH(HLVAE_LL_BWD)
#undef HLVAE_LL_BWD
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_statistics(int64_t N, int D, int64_t ld_theta, const int32_t* var_kind, const int32_t* var_nclass,
                                const int32_t* var_pcol, const double* vparam, const void* params, int dtype,
                                void* mean, void* mode, void* stream) {
    if (N < 0 || D <= 0 || !var_kind || !var_nclass || !var_pcol || !vparam || !params || !mean || !mode ||
        (dtype != HLVAE_F32 && dtype != HLVAE_F64))
        return HLVAE_E_ARG;
    if (N == 0) return 0;
    const unsigned gx = (unsigned)((D + 255) / 256);
    unsigned gy = (unsigned)((148 * 8 + gx - 1) / gx);
    if ((int64_t)gy > N) gy = (unsigned)N;
    dim3 grid(gx, gy);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HLVAE_F64)
        statistics_k<double><<<grid, 256, 0, st>>>(N, D, ld_theta, var_kind, var_nclass, var_pcol, vparam,
                                                   (const double*)params, (double*)mean, (double*)mode);
    else
        statistics_k<float><<<grid, 256, 0, st>>>(N, D, ld_theta, var_kind, var_nclass, var_pcol, vparam,
                                                  (const float*)params, (float*)mean, (float*)mode);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_discrete_transform(int64_t N, int D, int64_t ld_data, const int32_t* var_kind,
                                        const int32_t* var_nclass, const int32_t* var_dcol, const void* data,
                                        int dtype, void* out, void* stream) {
    if (N < 0 || D <= 0 || !var_kind || !var_nclass || !var_dcol || !data || !out ||
        (dtype != HLVAE_F32 && dtype != HLVAE_F64))
        return HLVAE_E_ARG;
    if (N == 0) return 0;
    const unsigned gx = (unsigned)((D + 255) / 256);
    unsigned gy = (unsigned)((148 * 8 + gx - 1) / gx);
    if ((int64_t)gy > N) gy = (unsigned)N;
    dim3 grid(gx, gy);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HLVAE_F64)
        discrete_transform_k<double><<<grid, 256, 0, st>>>(N, D, ld_data, var_kind, var_nclass, var_dcol,
                                                           (const double*)data, (double*)out);
    else
        discrete_transform_k<float><<<grid, 256, 0, st>>>(N, D, ld_data, var_kind, var_nclass, var_dcol,
                                                          (const float*)data, (float*)out);
    HLVAE_CHECK_LAUNCH();
    return 0;
}
