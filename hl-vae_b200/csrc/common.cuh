// Shared device helpers for the hlvae_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hlvae_b200.h"

#define HLVAE_WARP 32

namespace hlvae {

// Per-latent-dimension constrained hyper-parameters of one additive kernel, expanded into
// the forms the inner loops need.
struct KParams {
    double os[HLVAE_MAX_COMPS];     // outputscale
    double hil2[HLVAE_MAX_COMPS];   // 0.5 / lengthscale^2
    double il2[HLVAE_MAX_COMPS];    // 1 / lengthscale^2
    double il3[HLVAE_MAX_COMPS];    // 1 / lengthscale^3
};

__device__ __forceinline__ void load_kparams(KParams& kp, const hlvae_kspec_t& sp, const double* __restrict__ os,
                                             const double* __restrict__ ls, int L, int l) {
#pragma unroll
    for (int r = 0; r < HLVAE_MAX_COMPS; r++) {
        if (r < sp.ncomp) {
            double s = os[(int64_t)r * L + l];
            double e = ls[(int64_t)r * L + l];
            kp.os[r] = s;
            double i2 = 1.0 / (e * e);
            kp.hil2[r] = 0.5 * i2;
            kp.il2[r] = i2;
            kp.il3[r] = i2 / e;
        } else {
            kp.os[r] = 0.0; kp.hil2[r] = 0.0; kp.il2[r] = 0.0; kp.il3[r] = 0.0;
        }
    }
}

// exp(x) for x <= 0 (every squared-exponential argument is -(d^2) / (2 l^2)).  Under half the
// instructions of the general library exp: no overflow / NaN handling, one rounding step
// (x = k ln2 + r, |r| <= ln2 / 2), a degree-13 Taylor polynomial and an exponent insert.
// Branch-free.  Error <= ~1.5 ulp; arguments below -708 are clamped (3e-308 instead of a denormal or 0).
__device__ __forceinline__ double exp_nonpos(double x) {
    x = fmax(x, -708.0);
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: adding it rounds to nearest integer
    const double t = fma(x, 1.4426950408889634074, MAGIC);
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -6.93147180369123816490e-01, x);      // ln2 high part
    r = fma(kf, -1.90821492927058770002e-10, r);             // ln2 low part
    // even and odd halves as two independent Horner chains in r^2: every DFMA takes its single new constant
    // straight from the constant bank (Estrin pairs need two 64-bit constants per DFMA and spend more
    // instructions moving them into registers than they save); dependency depth 8
    const double s = r * r;
    double pe = 1.0 / 479001600.0;                           // 1/12!
    double po = 1.0 / 6227020800.0;                          // 1/13!
    pe = fma(pe, s, 1.0 / 3628800.0);
    po = fma(po, s, 1.0 / 39916800.0);
    pe = fma(pe, s, 1.0 / 40320.0);
    po = fma(po, s, 1.0 / 362880.0);
    pe = fma(pe, s, 1.0 / 720.0);
    po = fma(po, s, 1.0 / 5040.0);
    pe = fma(pe, s, 1.0 / 24.0);
    po = fma(po, s, 1.0 / 120.0);
    pe = fma(pe, s, 0.5);
    po = fma(po, s, 1.0 / 6.0);
    pe = fma(pe, s, 1.0);
    po = fma(po, s, 1.0);
    const double p = fma(r, po, pe);
    return p * __hiloint2double((k + 1023) << 20, 0);        // k >= -1022 here
}

// Table-driven variant for the kernels that can afford 512 bytes of shared memory: x = (64 k + j) ln2 / 64 + r with
// |r| <= ln2 / 128, exp(x) = 2^k * T[j] * (1 + r + r^2/2 + ... + r^5/120), T[j] = 2^(j/64) from `tab` (filled by
// exp2_table_fill).  The series needs 5 dependent DFMAs instead of 2 x 7 and the table entry is only needed by the
// last operation, so its shared-memory latency hides behind the polynomial.  Truncation r^6/720 < 4e-17 relative;
// error <= ~1.5 ulp like exp_nonpos.
constexpr int HLVAE_EXP_TAB = 64;

__device__ __forceinline__ void exp2_table_fill(double* __restrict__ tab, int tid, int nthreads) {
    for (int j = tid; j < HLVAE_EXP_TAB; j += nthreads) tab[j] = exp2((double)j / HLVAE_EXP_TAB);
}

__device__ __forceinline__ double exp_nonpos_tab(double x, const double* __restrict__ tab) {
    x = fmax(x, -708.0);
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(x, 64.0 * 1.4426950408889634074, MAGIC);
    const int n = __double2loint(t);                         // round(x * 64 / ln2) = 64 k + j
    const double nf = t - MAGIC;
    double r = fma(nf, -6.93147180369123816490e-01 / 64.0, x);       // (ln2 high part) / 64: exact scaling
    r = fma(nf, -1.90821492927058770002e-10 / 64.0, r);
    const double s = tab[n & (HLVAE_EXP_TAB - 1)];
    const int k = n >> 6;                                    // arithmetic shift = floor(n / 64)
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p *= r;                                                  // exp(r) - 1
    return fma(s, p, s) * __hiloint2double((k + 1023) << 20, 0);     // k >= -1022 here
}

// Discrete factors of one component: CatKernel (kernel_spec.py:26-32) and BinKernel
// (kernel_spec.py:9-23); true when every factor equals 1.
__device__ __forceinline__ bool disc_match(const hlvae_comp_t& c, const double* __restrict__ xa,
                                           const double* __restrict__ xb, int sa, int sb) {
    bool ok = true;
#pragma unroll
    for (int f = 0; f < HLVAE_MAX_DISC; f++) {
        if (f < c.ndisc) {
            double a = xa[c.disc_col[f] * sa], b = xb[c.disc_col[f] * sb];
            ok = ok && ((c.disc_kind[f] == HLVAE_KIND_CAT) ? (a - b == 0.0) : (a + b == 2.0));
        }
    }
    return ok;
}

// Unscaled value of component r at (xa, xb): prod(disc) * exp(-(xa-xb)^2 / (2 l^2))
// (GP_model.py:43-69 form of the gpytorch RBFKernel the reference builds at kernel_spec.py:58-69).
__device__ __forceinline__ double comp_value(const hlvae_comp_t& c, double hil2, const double* __restrict__ xa,
                                             const double* __restrict__ xb, int sa, int sb, double& d_out) {
    d_out = 0.0;
    if (!disc_match(c, xa, xb, sa, sb)) return 0.0;
    if (c.se_col < 0) return 1.0;
    double d = xa[c.se_col * sa] - xb[c.se_col * sb];
    d_out = d;
    return exp_nonpos(-(d * d) * hil2);
}

// K(xa, xb) = sum_r os_r * comp_r   (AdditiveKernel of ScaleKernels, kernel_gen.py:219-310)
// xa[c * sa], xb[c * sb] address covariate column c (sa = sb = 1 for row-major rows).
__device__ __forceinline__ double eval_additive(const hlvae_kspec_t& sp, const KParams& kp,
                                                const double* __restrict__ xa, const double* __restrict__ xb,
                                                int sa = 1, int sb = 1) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < HLVAE_MAX_COMPS; r++) {
        if (r < sp.ncomp) {
            double d;
            double k = comp_value(sp.comp[r], kp.hil2[r], xa, xb, sa, sb, d);
            acc = fma(kp.os[r], k, acc);
        }
    }
    return acc;
}

// Accumulate g * dK/d{os_r, ls_r} and (optionally) g * dK/dxb[se_col_r] into per-component
// register accumulators.  dK/dls_r = os_r k_r d^2 / l^3 ; dK/dxb = os_r k_r d / l^2 (d = xa - xb).
template <bool WITH_GX>
__device__ __forceinline__ void accum_grads(const hlvae_kspec_t& sp, const KParams& kp,
                                            const double* __restrict__ xa, const double* __restrict__ xb, double g,
                                            double (&gos)[HLVAE_MAX_COMPS], double (&gls)[HLVAE_MAX_COMPS],
                                            double (&gxb)[HLVAE_MAX_COMPS], int sa = 1, int sb = 1) {
#pragma unroll
    for (int r = 0; r < HLVAE_MAX_COMPS; r++) {
        if (r < sp.ncomp) {
            double d;
            double k = comp_value(sp.comp[r], kp.hil2[r], xa, xb, sa, sb, d);
            double gk = g * k;
            gos[r] += gk;
            double t = gk * kp.os[r] * d;
            gls[r] = fma(t * d, kp.il3[r], gls[r]);
            if (WITH_GX) gxb[r] = fma(t, kp.il2[r], gxb[r]);
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor pipe.
// A: lane holds a[row = lane/4][col = lane%4]; B: b[row = lane%4][col = lane/4];
// C: c[row = lane/4][col = 2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

template <typename T>
__device__ __forceinline__ double ld_as_double(const void* p, int64_t i) {
    return (double)reinterpret_cast<const T*>(p)[i];
}

__device__ __forceinline__ void report_status(int32_t* status, int code, int l, int s) {
    if (status && atomicCAS(status, 0, code) == 0) {
        status[1] = l;
        status[2] = s;
    }
}

}  // namespace hlvae

#define HLVAE_CHECK_LAUNCH()                              \
    do {                                                  \
        cudaError_t e__ = cudaGetLastError();             \
        if (e__ != cudaSuccess) return (int)e__;          \
    } while (0)
