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

// The same with the discrete factor of a product component folded in: the result is exp(x) when `ok` and 0
// otherwise.  Two instructions of the stand-alone form are replaced by integer ones: the clamp x >= -708 is an
// unsigned minimum on the high word (non-positive doubles order like their bit patterns; -708.0 = 0xC0862000_00000000;
// the low word of a clamped argument is left as it is: at most 2^-43 below -708), and the select acts on the one
// non-zero word of the power-of-two scale instead of on the 64-bit result.
template <bool SELECT>
__device__ __forceinline__ double exp_nonpos_tab_sel(double x, const double* __restrict__ tab, bool ok) {
    const unsigned hi = min((unsigned)__double2hiint(x), 0xC0862000u);
    x = __hiloint2double((int)hi, __double2loint(x));
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(x, 64.0 * 1.4426950408889634074, MAGIC);
    const int n = __double2loint(t);
    const double nf = t - MAGIC;
    double r = fma(nf, -6.93147180369123816490e-01 / 64.0, x);
    r = fma(nf, -1.90821492927058770002e-10 / 64.0, r);
    const double s = tab[n & (HLVAE_EXP_TAB - 1)];
    int sc = ((n >> 6) + 1023) << 20;
    if (SELECT) sc = ok ? sc : 0;
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p *= r;
    return fma(s, p, s) * __hiloint2double(sc, 0);
}

// ---- Tensor memory (TMEM, 512 columns x 128 lanes x 32 bit per SM) as thread-private scratch: a thread of warp w
// reaches lane 32 (w % 4) + its lane id of every column through tcgen05.st / tcgen05.ld (shape 32x32b: one 32-bit
// register per column), so a block of columns is storage that costs neither registers nor shared memory.
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {                // one whole warp; COLS = power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_u32(slot)), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {               // one whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st_d2(uint32_t a, const double* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(__double2loint(v[0])),
                 "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_d4(uint32_t a, const double* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a),
                 "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])),
                 "r"(__double2hiint(v[1])), "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])),
                 "r"(__double2loint(v[3])), "r"(__double2hiint(v[3]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_d8(uint32_t a, const double* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15, %16};" ::"r"(a),
        "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
        "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])),
        "r"(__double2loint(v[4])), "r"(__double2hiint(v[4])), "r"(__double2loint(v[5])), "r"(__double2hiint(v[5])),
        "r"(__double2loint(v[6])), "r"(__double2hiint(v[6])), "r"(__double2loint(v[7])), "r"(__double2hiint(v[7]))
        : "memory");
}
// loads are issued without waiting; tmem_load_doubles waits once for all of them
__device__ __forceinline__ void tmem_ld_r4(uint32_t a, int* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_r8(uint32_t a, int* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(a)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_r16(uint32_t a, int* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a)
        : "memory");
}
template <int NR>
__device__ __forceinline__ void tmem_issue_loads(uint32_t a, int* r) {
    if constexpr (NR >= 16) {
        tmem_ld_r16(a, r);
        tmem_issue_loads<NR - 16>(a + 16, r + 16);
    } else if constexpr (NR >= 8) {
        tmem_ld_r8(a, r);
        tmem_issue_loads<NR - 8>(a + 8, r + 8);
    } else if constexpr (NR >= 4) {
        tmem_ld_r4(a, r);
    }
}
// N doubles of one thread <-> 2 N consecutive columns (N even), split into the power-of-two vector widths
template <int N>
__device__ __forceinline__ void tmem_store_doubles(uint32_t a, const double* v) {
    static_assert(N % 2 == 0 && N >= 0, "even counts");
    if constexpr (N >= 8) {
        tmem_st_d8(a, v);
        tmem_store_doubles<N - 8>(a + 16, v + 8);
    } else if constexpr (N >= 4) {
        tmem_st_d4(a, v);
        tmem_store_doubles<N - 4>(a + 8, v + 4);
    } else if constexpr (N >= 2) {
        tmem_st_d2(a, v);
    }
}
template <int N>
__device__ __forceinline__ void tmem_load_doubles(uint32_t a, double* v) {
    static_assert(N % 2 == 0 && N >= 2, "even counts");
    int r[2 * N];
    tmem_issue_loads<2 * N>(a, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 2 * N; i++) asm volatile("" : "+r"(r[i]));      // the registers are read after the wait
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
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

// Asynchronous global -> shared copies of one 4- / 8- / 16-byte element each (LDGSTS): gathers and scatters whose
// latency is taken off the issuing thread; completed by cp_async_wait_all of the same thread, then a barrier.
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async element size");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr_u32(smem_dst)), "l"(gsrc), "n"(BYTES)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- TMA bulk copy global -> shared (cp.async.bulk), completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_addr_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr_u32(bar)) : "memory");
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
