// Per-variable observation heads: HLVAE.theta_estimation (HLVAE.py:416-453) over the Observation_*
// modules (HLVAE.py:11-89) as ONE streaming pass.  The reference evaluates every head twice (observed and,
// under no_grad, missing entries), multiplies by masks and scatters through boolean indices; for 0/1 masks that
// is  theta[n, p] = act_p(bias[p] + sum_k weight[p, k] y[n, var(p), k])  in the forward direction, with the
// gradient flowing only through observed (mask = 1) entries.
//
// CTA = tile of whole variables covering <= 256 theta columns x a stripe of rows, thread = theta column with
// its weights in registers.  y is read straight from global memory: the lanes of a warp own the columns of
// adjacent variables, so for either layout of y (conv: [N, Y, D] viewed as [N, D, Y]; dense: [N, D, Y]) a warp's
// loads fall into one or two sectors per k and the y_dim loads of RB rows are all in flight together - no
// staging, no barrier; theta rows are written coalesced (thread = column).  Backward: thread = column
// accumulates d/d{weight, bias} in float64 registers over the whole stripe (one atomic per column and CTA at the
// end) and leaves its masked upstream gradient in (double-buffered) shared memory; after ONE barrier per row
// batch, thread = (variable, k) forms d/dy from the <= 16 columns of the variable, decoded so that consecutive
// threads write consecutive addresses of the caller's layout.  HBM-bound: y and theta cross once.
// Layouts with <= 5 theta columns per variable and y_dim <= 8 take a thread-per-VARIABLE backward kernel instead (below).
#include "common.cuh"

#ifndef HLVAE_TH_BWD_BYTES
#define HLVAE_TH_BWD_BYTES 80      // bytes of y values in flight per thread in the backward kernel
#endif
#ifndef HLVAE_TH_BWD_CTAS
#define HLVAE_TH_BWD_CTAS 3
#endif
#ifndef HLVAE_TV_BWD_CTAS
// CTAs per SM worth of row stripes, thread-per-variable backward kernel.  Measured with 80 bytes of y per thread in
// flight, 8 / 12 / 16 / 32 / 64: 0.406 / 0.381 / 0.389 / 0.399 / 0.453 ms at the configs[1] batch, 0.269 / 0.257 / 0.270 /
// 0.283 / 0.382 ms on the tabular layout; 16 with 40 bytes (the default): 0.383 / 0.253 ms
#define HLVAE_TV_BWD_CTAS 16
#endif

namespace {

constexpr int TH_THREADS = 256;   // = max theta columns and max variables per tile
// rows in flight per thread: about 40 registers of y values (y_dim = 5: 8 rows in float32, 4 in float64)
template <typename TS, int YP, bool BWD = false>
struct ThRows {
    static constexpr int raw = (BWD ? HLVAE_TH_BWD_BYTES : 160) / (YP * (int)sizeof(TS));
    // backward: a power of two, the rows of a column travel through shared memory as one vector
    static constexpr int value = BWD ? (raw >= 8 ? 8 : raw >= 4 ? 4 : raw >= 2 ? 2 : 1) : (raw < 1 ? 1 : (raw > 8 ? 8 : raw));
};

template <typename TS, int RB>
struct alignas(sizeof(TS) * RB > 16 ? 16 : sizeof(TS) * RB) ThRowVec {
    TS v[RB];
};

template <typename TS>
__device__ __forceinline__ TS sigmoid_t(TS z);
template <>
__device__ __forceinline__ double sigmoid_t<double>(double z) { return 1.0 / (1.0 + exp(-z)); }
template <>
__device__ __forceinline__ float sigmoid_t<float>(float z) { return 1.0f / (1.0f + expf(-z)); }

struct ThTile {
    int d0, nv, p0, ncols;
    int64_t r_begin, r_end;
};

__device__ __forceinline__ ThTile th_tile(const int32_t* __restrict__ tile_var, const int32_t* __restrict__ var_pcol,
                                          int64_t N, int TH_RB) {
    ThTile t;
    t.d0 = tile_var[blockIdx.x];
    const int d1 = tile_var[blockIdx.x + 1];
    t.nv = d1 - t.d0;
    t.p0 = var_pcol[t.d0];
    t.ncols = var_pcol[d1] - t.p0;
    const int64_t batches = (N + TH_RB - 1) / TH_RB;
    const int64_t per = (batches + gridDim.y - 1) / gridDim.y;
    t.r_begin = (int64_t)blockIdx.y * per * TH_RB;
    t.r_end = t.r_begin + per * TH_RB;
    if (t.r_end > N) t.r_end = N;
    return t;
}

// R rows of one theta column.  yb / ob point at the first row; the row and k offsets are 32-bit and uniform over
// the CTA (r and k are compile-time), so each access costs one wide multiply-add on top of the memory instruction.
template <typename TS, int YP, int R>
__device__ __forceinline__ void th_fwd_rows(const TS* __restrict__ yb, TS* __restrict__ ob, int sn, int sk, int ld, int Y,
                                            const TS (&wk)[YP], TS b, int mode, bool ydep) {
    // keep the strides opaque per call: otherwise the compiler hoists all R x y_dim 64-bit offsets out of the
    // row loop as loop invariants (80 registers of addresses, spills, ~9 integer instructions per load)
    asm volatile("" : "+r"(sn), "+r"(sk), "+r"(ld));
    TS z[R];
#pragma unroll
    for (int r = 0; r < R; r++) z[r] = b;
    if (ydep) {
        TS yk[R][YP];
        const TS* pr = yb;                       // one pointer per row, k offsets shared by all rows
#pragma unroll
        for (int r = 0; r < R; r++) {
#pragma unroll
            for (int k = 0; k < YP; k++) yk[r][k] = (k < Y) ? pr[k * sk] : (TS)0;
            pr += sn;
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
#pragma unroll
            for (int k = 0; k < YP; k++) z[r] = fma(wk[k], yk[r][k], z[r]);
            if (mode == HLVAE_HEAD_SIGMOID) z[r] = sigmoid_t<TS>(z[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) ob[r * ld] = z[r];
}

template <typename TS, int YP>
__global__ void __launch_bounds__(TH_THREADS, 3)
theta_fwd_k(int64_t N, int Y, const int32_t* __restrict__ col_var, const int32_t* __restrict__ col_mode,
            const int32_t* __restrict__ var_pcol, const int32_t* __restrict__ tile_var,
            const double* __restrict__ weight, const double* __restrict__ bias, const TS* __restrict__ y, int64_t sn,
            int64_t sd, int64_t sk, TS* __restrict__ theta, int64_t ld_theta) {
    constexpr int TH_RB = ThRows<TS, YP>::value;
    Y = (YP == 16) ? Y : YP;      // exact instantiations (y_dim <= 8): y_dim is a compile-time constant from here on
    const int tid = threadIdx.x;
    const ThTile t = th_tile(tile_var, var_pcol, N, TH_RB);
    if (tid >= t.ncols || t.r_begin >= t.r_end) return;
    const int p = t.p0 + tid;
    const int mode = col_mode[p];
    const bool ydep = (mode == HLVAE_HEAD_AFFINE || mode == HLVAE_HEAD_SIGMOID);
    const TS b = (mode == HLVAE_HEAD_ZERO) ? (TS)0 : (TS)bias[p];
    TS wk[YP];
#pragma unroll
    for (int k = 0; k < YP; k++) wk[k] = (ydep && k < Y) ? (TS)weight[(int64_t)p * Y + k] : (TS)0;
    const TS* yv = y + (int64_t)col_var[p] * sd;
    TS* out = theta + p;
    const int sn32 = (int)sn, sk32 = (int)sk, ld32 = (int)ld_theta;
    int64_t n0 = t.r_begin;
    for (; n0 + TH_RB <= t.r_end; n0 += TH_RB)
        th_fwd_rows<TS, YP, TH_RB>(yv + n0 * sn, out + n0 * ld_theta, sn32, sk32, ld32, Y, wk, b, mode, ydep);
    for (; n0 < t.r_end; n0++)
        th_fwd_rows<TS, YP, 1>(yv + n0 * sn, out + n0 * ld_theta, sn32, sk32, ld32, Y, wk, b, mode, ydep);
}

// One batch of R rows of the backward pass (see the kernel comment); called with R = RB for full batches and
// R = 1 for the tail rows of a stripe.  All threads of the CTA call it (it contains the barrier).
template <typename TS, typename TM, int YP, int R>
__device__ __forceinline__ void th_bwd_rows(bool live, bool ydep, int mode, TS b, const TS (&wk)[YP], double (&gw)[YP],
                                            double& gb, const TS* __restrict__ yb, const TM* __restrict__ mb,
                                            const TS* __restrict__ gb_in, TS* __restrict__ gyb, int sn, int sd, int sk,
                                            int D, int ld, int Y, TS* __restrict__ gsb, const TS* __restrict__ Ws,
                                            const int* __restrict__ pairs, int npairs) {
    const int tid = threadIdx.x;
    asm volatile("" : "+r"(sn), "+r"(sk), "+r"(ld), "+r"(D));      // see th_fwd_rows
    // ---- thread = column: masked upstream gradient, d/d{weight, bias}
    if (live) {
        ThRowVec<TS, R> gout;
        if (mode == HLVAE_HEAD_ZERO) {
#pragma unroll
            for (int r = 0; r < R; r++) gout.v[r] = (TS)0;
        } else {
            TS g[R];
#pragma unroll
            for (int r = 0; r < R; r++) g[r] = gb_in[r * ld];
            TM m[R];
#pragma unroll
            for (int r = 0; r < R; r++) m[r] = mb[r * D];
            if (ydep) {
                TS yk[R][YP];
                const TS* pr = yb;
#pragma unroll
                for (int r = 0; r < R; r++) {
#pragma unroll
                    for (int k = 0; k < YP; k++) yk[r][k] = (k < Y) ? pr[k * sk] : (TS)0;
                    pr += sn;
                }
                // sums over the batch's rows in the storage type, one float64 accumulation per batch (conversions
                // to float64 run at a fraction of the FMA rate)
                TS gwb[YP], gbb = (TS)0;
#pragma unroll
                for (int k = 0; k < YP; k++) gwb[k] = (TS)0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    TS gr = (m[r] != (TM)0) ? g[r] : (TS)0;
                    if (mode == HLVAE_HEAD_SIGMOID) {
                        TS z = b;
#pragma unroll
                        for (int k = 0; k < YP; k++) z = fma(wk[k], yk[r][k], z);
                        const TS sg = sigmoid_t<TS>(z);
                        gr *= sg * ((TS)1 - sg);
                    }
#pragma unroll
                    for (int k = 0; k < YP; k++) gwb[k] = fma(gr, yk[r][k], gwb[k]);
                    gbb += gr;
                    gout.v[r] = gr;
                }
#pragma unroll
                for (int k = 0; k < YP; k++) gw[k] += (double)gwb[k];
                gb += (double)gbb;
            } else {   // HLVAE_HEAD_BIAS: the column is its bias
                TS gbb = (TS)0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    gbb += (m[r] != (TM)0) ? g[r] : (TS)0;
                    gout.v[r] = (TS)0;
                }
                gb += (double)gbb;
            }
        }
        *reinterpret_cast<ThRowVec<TS, R>*>(gsb + tid * R) = gout;
    }
    __syncthreads();
    // ---- thread = (variable, k): d/dy[n, d, k] = sum over the variable's columns of g * weight[col, k]
    for (int q = tid; q < npairs; q += TH_THREADS) {
        const int pq = pairs[q];                 // (variable, k, first column, column count), tabulated per tile
        const int v = pq & 255, k = (pq >> 8) & 31, lp = (pq >> 13) & 511, nc = pq >> 22;
        TS a[R];
#pragma unroll
        for (int r = 0; r < R; r++) a[r] = (TS)0;
        for (int c = 0; c < nc; c++) {
            const TS wv = Ws[(lp + c) * Y + k];
            const ThRowVec<TS, R> gv = *reinterpret_cast<const ThRowVec<TS, R>*>(gsb + (lp + c) * R);
#pragma unroll
            for (int r = 0; r < R; r++) a[r] = fma(gv.v[r], wv, a[r]);
        }
        TS* dst = gyb + v * sd + k * sk;
#pragma unroll
        for (int r = 0; r < R; r++) dst[r * sn] = a[r];
    }
}

template <typename TS, typename TM, int YP>
__global__ void __launch_bounds__(TH_THREADS, HLVAE_TH_BWD_CTAS)
theta_bwd_k(int64_t N, int D, int Y, const int32_t* __restrict__ col_var, const int32_t* __restrict__ col_mode,
            const int32_t* __restrict__ var_pcol, const int32_t* __restrict__ tile_var,
            const double* __restrict__ weight, const double* __restrict__ bias, const TS* __restrict__ y, int64_t sn,
            int64_t sd, int64_t sk, const TM* __restrict__ mask, const TS* __restrict__ g_theta, int64_t ld_theta,
            TS* __restrict__ g_y, double* __restrict__ g_weight, double* __restrict__ g_bias) {
    constexpr int TH_RB = ThRows<TS, YP, true>::value;
    // (y_dim stays a run-time value here: making it a compile-time constant as in the forward kernel was measured
    // 5 % slower - 0.547 vs 0.519 ms - the fully unrolled d/dy phase schedules worse)
    extern __shared__ __align__(16) unsigned char th_smem[];
    TS* gs = reinterpret_cast<TS*>(th_smem);                  // [2][256][RB] masked upstream gradient per column
    TS* Ws = gs + 2 * TH_RB * TH_THREADS;                     // [256][Y] weights of the tile's y-dependent columns
    int* pairs = reinterpret_cast<int*>(Ws + (size_t)TH_THREADS * Y);   // [nv * Y] packed (v, k, first column, #columns)
    const int tid = threadIdx.x;
    const ThTile t = th_tile(tile_var, var_pcol, N, TH_RB);
    if (t.r_begin >= t.r_end) return;

    const bool live = tid < t.ncols;
    const int p = t.p0 + (live ? tid : 0);
    const int mode = live ? col_mode[p] : HLVAE_HEAD_ZERO;
    const bool ydep = (mode == HLVAE_HEAD_AFFINE || mode == HLVAE_HEAD_SIGMOID);
    const int dvar = col_var[p];
    const TS b = (TS)bias[p];
    TS wk[YP];
    double gw[YP], gb = 0.0;
#pragma unroll
    for (int k = 0; k < YP; k++) {
        wk[k] = (ydep && k < Y) ? (TS)weight[(int64_t)p * Y + k] : (TS)0;
        gw[k] = 0.0;
    }
    for (int k = 0; k < Y; k++) Ws[tid * Y + k] = ydep ? (TS)weight[(int64_t)p * Y + k] : (TS)0;
    // (variable, k) pairs of the d/dy phase, ordered so that consecutive threads write consecutive addresses of the
    // caller's layout (variable-fastest when y is [N, Y, D] viewed as [N, D, Y])
    const int npairs = t.nv * Y;
    const bool v_fastest = sd < sk;
    for (int q = tid; q < npairs; q += TH_THREADS) {
        int v, k;
        if (v_fastest) { k = q / t.nv; v = q - k * t.nv; } else { v = q / Y; k = q - v * Y; }
        const int lp = var_pcol[t.d0 + v] - t.p0, nc = var_pcol[t.d0 + v + 1] - var_pcol[t.d0 + v];
        pairs[q] = v | (k << 8) | (lp << 13) | (nc << 22);
    }
    __syncthreads();
    const TS* yv = y + (int64_t)dvar * sd;
    const TM* mv = mask + dvar;
    const TS* gin = g_theta + p;
    TS* gyt = g_y + (int64_t)t.d0 * sd;
    const int sn32 = (int)sn, sd32 = (int)sd, sk32 = (int)sk, ld32 = (int)ld_theta;
    int buf = 0;
    int64_t n0 = t.r_begin;
    for (; n0 + TH_RB <= t.r_end; n0 += TH_RB, buf ^= 1)
        th_bwd_rows<TS, TM, YP, TH_RB>(live, ydep, mode, b, wk, gw, gb, yv + n0 * sn, mv + n0 * D, gin + n0 * ld_theta,
                                       gyt + n0 * sn, sn32, sd32, sk32, D, ld32, Y, gs + buf * TH_RB * TH_THREADS, Ws, pairs,
                                       npairs);
    for (; n0 < t.r_end; n0++, buf ^= 1)
        th_bwd_rows<TS, TM, YP, 1>(live, ydep, mode, b, wk, gw, gb, yv + n0 * sn, mv + n0 * D, gin + n0 * ld_theta,
                                   gyt + n0 * sn, sn32, sd32, sk32, D, ld32, Y, gs + buf * TH_RB * TH_THREADS, Ws, pairs, npairs);
    if (live && mode != HLVAE_HEAD_ZERO) {
        if (gb != 0.0) atomicAdd(g_bias + p, gb);
        if (ydep) {
#pragma unroll
            for (int k = 0; k < YP; k++)
                if (k < Y && gw[k] != 0.0) atomicAdd(g_weight + (int64_t)p * Y + k, gw[k]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Thread-per-VARIABLE backward kernel: layouts whose variables own at most TV_CV theta columns (up to 5 classes:
// every layout the reference ships) with y_dim <= 8.  The thread-per-column kernel above re-loads a variable's y
// values once per column, crosses shared memory for d/dy and spends 435 instructions per (row, variable); here a
// thread owns a variable: its heads and its weight / bias gradient sums sit in registers, the g_theta rows of the tile
// (coalesced row segments) and the thread's own y values of the NEXT row batch arrive by cp.async while the current
// batch is evaluated, d/dy is formed and written by the same thread (coalesced in the convolutional layout):
// 270 instructions per (row, variable), configs[1] batch 0.478 -> 0.374 ms, tabular 64 000 rows 0.317 -> 0.253 ms.
// (The forward direction was measured the same way and dropped: 0.250 against 0.205 ms for thread-per-column, whose
// stores are already coalesced and which needs no staging at all.)
constexpr int TV_THREADS = 128;   // variables per tile
constexpr int TV_CV = 5;          // theta columns per variable held in registers
#ifndef HLVAE_TV_BWD_MIN_CTAS
#define HLVAE_TV_BWD_MIN_CTAS 1
#endif
#ifndef HLVAE_TV_BWD_BYTES
#define HLVAE_TV_BWD_BYTES 40      // y values of one row batch per thread (y_dim 5, float32: 2 rows); 80: see above, 160: 0.60 ms
#endif
template <typename TS, int YP>
struct TvRows {
    static constexpr int raw = HLVAE_TV_BWD_BYTES / (YP * (int)sizeof(TS));
    static constexpr int value = raw < 1 ? 1 : (raw > 8 ? 8 : raw);
};

template <typename TS, int YP>
struct TvVar {            // one variable's heads
    TS w[TV_CV][YP], b[TV_CV];
    int pc, lp, nc;       // first theta column (global / inside the tile), number of columns
    unsigned dep, sig, bo;   // bit c: column c depends on y / has the Sigmoid / is its bias (ordinal thresholds)
};

template <typename TS, int YP>
__device__ __forceinline__ void tv_load(TvVar<TS, YP>& v, bool live, int d, int p0, const int32_t* __restrict__ var_pcol,
                                        const int32_t* __restrict__ col_mode, const double* __restrict__ weight,
                                        const double* __restrict__ bias) {
    v.pc = live ? var_pcol[d] : p0;
    v.lp = v.pc - p0;
    v.nc = live ? var_pcol[d + 1] - v.pc : 0;
    v.dep = v.sig = v.bo = 0;
#pragma unroll
    for (int c = 0; c < TV_CV; c++) {
        const bool on = c < v.nc;
        const int mode = on ? col_mode[v.pc + c] : HLVAE_HEAD_ZERO;
        const bool ydep = (mode == HLVAE_HEAD_AFFINE || mode == HLVAE_HEAD_SIGMOID);
        if (ydep) v.dep |= 1u << c;
        if (mode == HLVAE_HEAD_SIGMOID) v.sig |= 1u << c;
        if (mode == HLVAE_HEAD_BIAS) v.bo |= 1u << c;
        v.b[c] = (mode != HLVAE_HEAD_ZERO) ? (TS)bias[v.pc + c] : (TS)0;
#pragma unroll
        for (int k = 0; k < YP; k++) v.w[c][k] = ydep ? (TS)weight[(int64_t)(v.pc + c) * YP + k] : (TS)0;
    }
}

struct TvTile {
    int d0, d1, p0, ncols;
    int64_t r_begin, r_end;
};
__device__ __forceinline__ TvTile tv_tile(const int32_t* __restrict__ var_pcol, int D, int64_t N, int RB) {
    TvTile t;
    t.d0 = blockIdx.x * TV_THREADS;
    t.d1 = min(D, t.d0 + TV_THREADS);
    t.p0 = var_pcol[t.d0];
    t.ncols = var_pcol[t.d1] - t.p0;
    const int64_t batches = (N + RB - 1) / RB;
    const int64_t per = (batches + gridDim.y - 1) / gridDim.y;
    t.r_begin = (int64_t)blockIdx.y * per * RB;
    t.r_end = t.r_begin + per * RB;
    if (t.r_end > N) t.r_end = N;
    return t;
}

// Rows of the tile's column span, global (row stride ld) -> shared ([RB][cap]) by cp.async: thread i of the CTA moves
// columns i, i + 128, ... (every warp instruction covers 128 contiguous bytes; no registers, no wait on the issuing thread)
template <typename TS, int RB>
__device__ __forceinline__ void tv_rows_in(TS* __restrict__ dst, int cap, const TS* __restrict__ src, int ld, int ncols,
                                           int nr, int tid) {
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (r < nr) {
#pragma unroll
            for (int j = 0; j < TV_CV; j++) {
                const int i = tid + j * TV_THREADS;
                if (i < ncols) hlvae::cp_async<(int)sizeof(TS)>(dst + r * cap + i, src + r * ld + i);
            }
        }
    }
    hlvae::cp_async_commit();
}

// A thread's own y values of one row batch, global -> shared by cp.async: sY[r][k][thread].  Thread-private (the
// thread that copies an element is the only one that reads it), so cp.async.wait_all alone orders copy and use.
template <typename TS, int YP, int RB>
__device__ __forceinline__ void tv_y_in(TS* __restrict__ sy, const TS* __restrict__ src, int sn, int sk, int nr, int tid) {
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (r < nr) {
#pragma unroll
            for (int k = 0; k < YP; k++)
                hlvae::cp_async<(int)sizeof(TS)>(sy + (r * YP + k) * TV_THREADS + tid, src + r * sn + k * sk);
        }
    }
}

template <typename TS, typename TM, int YP>
__global__ void __launch_bounds__(TV_THREADS, sizeof(TS) == 4 ? HLVAE_TV_BWD_MIN_CTAS : 1)
theta_bwd_var_k(int64_t N, int D, int cap, const int32_t* __restrict__ col_mode, const int32_t* __restrict__ var_pcol,
                const double* __restrict__ weight, const double* __restrict__ bias, const TS* __restrict__ y, int64_t sn,
                int64_t sd, int64_t sk, const TM* __restrict__ mask, const TS* __restrict__ g_theta, int64_t ld_theta,
                TS* __restrict__ g_y, double* __restrict__ g_weight, double* __restrict__ g_bias) {
    constexpr int RB = TvRows<TS, YP>::value;
    constexpr int FLUSH = (256 + RB - 1) / RB;   // batches between two flushes of the storage-type parameter-gradient sums
    extern __shared__ __align__(16) unsigned char tv_smem[];
    TS* sG = reinterpret_cast<TS*>(tv_smem);                  // [2][RB][cap] upstream gradient rows of the tile
    TS* sY = sG + 2 * RB * cap;                               // [2][RB][YP][128] the threads' y values
    const int tid = threadIdx.x;
    const TvTile t = tv_tile(var_pcol, D, N, RB);
    if (t.r_begin >= t.r_end) return;
    const int d = t.d0 + tid;
    const bool live = d < t.d1;
    const int dd = live ? d : t.d0;
    const TS* yv = y + (int64_t)dd * sd;
    TS* gyv = g_y + (int64_t)dd * sd;
    const TM* mv = mask + dd;
    const TS* gin = g_theta + t.p0;
    int sn32 = (int)sn, sk32 = (int)sk, D32 = D;
    // software pipeline: g_theta rows (shared by the CTA) and the thread's own y values of batch j + 1 travel to the
    // other buffer (cp.async) and its mask entries to registers while batch j is evaluated
    auto issue = [&](int64_t n0, int b) {
        const int nr = (int)min((int64_t)RB, t.r_end - n0);
        tv_y_in<TS, YP, RB>(sY + b * RB * YP * TV_THREADS, yv + n0 * sn, sn32, sk32, nr, tid);
        tv_rows_in<TS, RB>(sG + b * RB * cap, cap, gin + n0 * ld_theta, (int)ld_theta, t.ncols, nr, tid);   // commits
    };
    auto load_mask = [&](TM (&m)[RB], int64_t n0) {
        const TM* pm = mv + n0 * D;
#pragma unroll
        for (int r = 0; r < RB; r++) m[r] = (n0 + r < t.r_end) ? pm[r * D32] : (TM)0;
    };
    if (t.ncols > cap) {      // more columns than max_cols promised for 128 variables (a caller error): the tile does not
        if (live)             // fit its staging buffer - poison its d/dy instead of overrunning shared memory
            for (int64_t n = t.r_begin; n < t.r_end; n++)
                for (int k = 0; k < YP; k++) gyv[n * sn + k * sk] = (TS)NAN;
        return;
    }
    issue(t.r_begin, 0);
    TM m[RB], m_next[RB];
    load_mask(m, t.r_begin);
    TvVar<TS, YP> v;
    tv_load<TS, YP>(v, live, d, t.p0, var_pcol, col_mode, weight, bias);
    const unsigned grad_cols = v.dep | v.bo;                  // columns that receive a gradient at all
    // a variable with more columns than the caller's max_cols promised (a caller error): its d/dy is poisoned with
    // NaN instead of silently missing the columns beyond TV_CV
    const bool wide = v.nc > TV_CV;
    TS gw[TV_CV][YP], gb[TV_CV];
#pragma unroll
    for (int c = 0; c < TV_CV; c++) {
        gb[c] = (TS)0;
#pragma unroll
        for (int k = 0; k < YP; k++) gw[c][k] = (TS)0;
    }
    auto flush = [&]() {
#pragma unroll
        for (int c = 0; c < TV_CV; c++) {
            if ((grad_cols >> c) & 1u) {
                if (gb[c] != (TS)0) atomicAdd(g_bias + v.pc + c, (double)gb[c]);
                if ((v.dep >> c) & 1u) {
#pragma unroll
                    for (int k = 0; k < YP; k++)
                        if (gw[c][k] != (TS)0) atomicAdd(g_weight + (int64_t)(v.pc + c) * YP + k, (double)gw[c][k]);
                }
            }
            gb[c] = (TS)0;
#pragma unroll
            for (int k = 0; k < YP; k++) gw[c][k] = (TS)0;
        }
    };
    int buf = 0, since = 0;
    for (int64_t n0 = t.r_begin; n0 < t.r_end; n0 += RB, buf ^= 1) {
        const int nr = (int)min((int64_t)RB, t.r_end - n0);
        const TS* sg = sG + buf * RB * cap;
        const TS* sy = sY + buf * RB * YP * TV_THREADS + tid;
        asm volatile("" : "+r"(sn32), "+r"(sk32), "+r"(D32));
        hlvae::cp_async_wait_all();
        __syncthreads();                                      // batch j's g rows are visible; batch j - 1 is done with buf ^ 1
        const bool more = n0 + RB < t.r_end;
        if (more) {
            issue(n0 + RB, buf ^ 1);
            load_mask(m_next, n0 + RB);
        }
        if (live) {
            TS* pg = gyv + n0 * sn;
#pragma unroll
            for (int r = 0; r < RB; r++) {
                if (r < nr) {
                    TS yk[YP], dy[YP];
#pragma unroll
                    for (int k = 0; k < YP; k++) {
                        yk[k] = sy[(r * YP + k) * TV_THREADS];
                        dy[k] = (TS)0;
                    }
#pragma unroll
                    for (int c = 0; c < TV_CV; c++) {
                        if (c < v.nc) {
                            TS gr = (m[r] != (TM)0 && ((grad_cols >> c) & 1u)) ? sg[r * cap + v.lp + c] : (TS)0;
                            if ((v.sig >> c) & 1u) {
                                TS z = v.b[c];
#pragma unroll
                                for (int k = 0; k < YP; k++) z = fma(v.w[c][k], yk[k], z);
                                const TS s_ = sigmoid_t<TS>(z);
                                gr *= s_ * ((TS)1 - s_);
                            }
                            gb[c] += gr;
#pragma unroll
                            for (int k = 0; k < YP; k++) {
                                dy[k] = fma(gr, v.w[c][k], dy[k]);
                                gw[c][k] = fma(gr, yk[k], gw[c][k]);
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < YP; k++) pg[r * sn32 + k * sk32] = wide ? (TS)NAN : dy[k];
                }
            }
            if (++since == FLUSH) {
                flush();
                since = 0;
            }
        }
        if (more) {
#pragma unroll
            for (int r = 0; r < RB; r++) m[r] = m_next[r];
        }
    }
    if (live) flush();
}

// 32-bit in-batch offsets: (rows in flight) * row stride + y_dim * k stride must fit in an int
bool th_offsets_fit(int64_t sn, int64_t sd, int64_t sk, int64_t ld, int Y) {
    const int64_t lim = (int64_t)1 << 31;
    return sn >= 0 && sd >= 0 && sk >= 0 && 8 * sn + Y * sk + 256 * sd < lim && 8 * ld < lim;
}

int grid_stripes(int64_t N, int n_tiles, int RB, int ctas_per_sm) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t batches = (N + RB - 1) / RB;
    int64_t want = ((int64_t)sms * ctas_per_sm + n_tiles - 1) / n_tiles;
    if (want > batches) want = batches;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    return (int)want;
}

template <typename TS, int YP>
int launch_fwd(int64_t N, int Y, int n_tiles, const int32_t* col_var, const int32_t* col_mode, const int32_t* var_pcol,
               const int32_t* tile_var, const double* weight, const double* bias, const void* y, int64_t sn, int64_t sd,
               int64_t sk, void* theta, int64_t ld_theta, cudaStream_t st) {
    auto kern = theta_fwd_k<TS, YP>;
    // CTAs per SM worth of row stripes: 8 / 16 / 32 / 64 / 128 measured 0.228 / 0.212 / 0.205 / 0.205 / 0.216 ms
    dim3 grid(n_tiles, grid_stripes(N, n_tiles, ThRows<TS, YP>::value, 32));
    kern<<<grid, TH_THREADS, 0, st>>>(N, Y, col_var, col_mode, var_pcol, tile_var, weight, bias, (const TS*)y, sn, sd,
                                         sk, (TS*)theta, ld_theta);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

template <typename TS, typename TM, int YP>
int launch_bwd(int64_t N, int D, int Y, int n_tiles, const int32_t* col_var, const int32_t* col_mode,
               const int32_t* var_pcol, const int32_t* tile_var, const double* weight, const double* bias, const void* y,
               int64_t sn, int64_t sd, int64_t sk, const void* mask, const void* g_theta, int64_t ld_theta, void* g_y,
               double* g_weight, double* g_bias, cudaStream_t st) {
    auto kern = theta_bwd_k<TS, TM, YP>;
    constexpr int RB = ThRows<TS, YP, true>::value;
    const size_t smem = ((size_t)2 * RB * TH_THREADS + (size_t)TH_THREADS * Y) * sizeof(TS) + (size_t)TH_THREADS * Y * sizeof(int);
    if (smem > 48 * 1024) {      // float64 storage with y_dim 15 / 16: above the default dynamic shared-memory limit
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    // CTAs per SM worth of row stripes: 9 / 18 / 36 / 72 measured 0.517 / 0.484 / 0.478 / 0.497 ms at the configs[1] batch
    dim3 grid(n_tiles, grid_stripes(N, n_tiles, RB, 36));
    kern<<<grid, TH_THREADS, smem, st>>>(N, D, Y, col_var, col_mode, var_pcol, tile_var, weight, bias, (const TS*)y, sn,
                                         sd, sk, (const TM*)mask, (const TS*)g_theta, ld_theta, (TS*)g_y, g_weight,
                                         g_bias);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

template <typename TS, typename TM, int YP>
int launch_bwd_var(int64_t N, int D, int max_cols, const int32_t* col_mode, const int32_t* var_pcol, const double* weight,
                   const double* bias, const void* y, int64_t sn, int64_t sd, int64_t sk, const void* mask,
                   const void* g_theta, int64_t ld_theta, void* g_y, double* g_weight, double* g_bias, cudaStream_t st) {
    auto kern = theta_bwd_var_k<TS, TM, YP>;
    constexpr int RB = TvRows<TS, YP>::value;
    const int cap = TV_THREADS * max_cols, n_tiles = (D + TV_THREADS - 1) / TV_THREADS;
    const size_t smem = (size_t)2 * RB * (cap + YP * TV_THREADS) * sizeof(TS);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(n_tiles, grid_stripes(N, n_tiles, RB, HLVAE_TV_BWD_CTAS));
    kern<<<grid, TV_THREADS, smem, st>>>(N, D, cap, col_mode, var_pcol, weight, bias, (const TS*)y, sn, sd, sk,
                                         (const TM*)mask, (const TS*)g_theta, ld_theta, (TS*)g_y, g_weight, g_bias);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

// the thread-per-variable kernels cover this call
bool tv_covers(int max_cols, int Y) { return max_cols >= 1 && max_cols <= TV_CV && Y <= 8; }

}  // namespace

extern "C" int hlvae_theta_fwd(int64_t N, int D, int P, int Y, int n_tiles, const int32_t* col_var,
                               const int32_t* col_mode, const int32_t* var_pcol, const int32_t* tile_var,
                               const double* weight, const double* bias, const void* y, int64_t sn, int64_t sd,
                               int64_t sk, int dtype, void* theta, int64_t ld_theta, void* stream) {
    if (N < 0 || D <= 0 || P <= 0 || Y <= 0 || n_tiles <= 0 || !col_var || !col_mode || !var_pcol || !tile_var ||
        !weight || !bias || !y || !theta || ld_theta < P)
        return HLVAE_E_ARG;
    if (Y > HLVAE_MAX_Y || !th_offsets_fit(sn, sd, sk, ld_theta, Y)) return HLVAE_E_UNSUPPORTED;
    if (dtype != HLVAE_F32 && dtype != HLVAE_F64) return HLVAE_E_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define HLVAE_TH_FWD(TS, YP) \
    return launch_fwd<TS, YP>(N, Y, n_tiles, col_var, col_mode, var_pcol, tile_var, weight, bias, y, sn, sd, sk, theta, ld_theta, st)
#define HLVAE_TH_FWD_Y(TS)                                                                                             \
    switch (Y) {                                                                                                       \
        case 1: HLVAE_TH_FWD(TS, 1); case 2: HLVAE_TH_FWD(TS, 2); case 3: HLVAE_TH_FWD(TS, 3); case 4: HLVAE_TH_FWD(TS, 4); \
        case 5: HLVAE_TH_FWD(TS, 5); case 6: HLVAE_TH_FWD(TS, 6); case 7: HLVAE_TH_FWD(TS, 7); case 8: HLVAE_TH_FWD(TS, 8); \
        default: HLVAE_TH_FWD(TS, 16);                                                                                 \
    }
    if (dtype == HLVAE_F32) { HLVAE_TH_FWD_Y(float) }
    HLVAE_TH_FWD_Y(double)
#undef HLVAE_TH_FWD_Y
#undef HLVAE_TH_FWD
}

extern "C" int hlvae_theta_bwd(int64_t N, int D, int P, int Y, int n_tiles, int max_cols, const int32_t* col_var,
                               const int32_t* col_mode, const int32_t* var_pcol, const int32_t* tile_var,
                               const double* weight, const double* bias, const void* y, int64_t sn, int64_t sd,
                               int64_t sk, int dtype, const void* mask, int mask_dtype, const void* g_theta,
                               int64_t ld_theta, void* g_y, double* g_weight, double* g_bias, void* stream) {
    if (N < 0 || D <= 0 || P <= 0 || Y <= 0 || n_tiles <= 0 || !col_var || !col_mode || !var_pcol || !tile_var ||
        !weight || !bias || !y || !mask || !g_theta || !g_y || !g_weight || !g_bias || ld_theta < P)
        return HLVAE_E_ARG;
    if (Y > HLVAE_MAX_Y || !th_offsets_fit(sn, sd, sk, ld_theta, Y)) return HLVAE_E_UNSUPPORTED;
    if (dtype != HLVAE_F32 && dtype != HLVAE_F64) return HLVAE_E_ARG;
    if (mask_dtype != dtype && mask_dtype != HLVAE_U8) return HLVAE_E_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (tv_covers(max_cols, Y)) {
#define HLVAE_TV_BWD(TS, TM, YP)                                                                                       \
    return launch_bwd_var<TS, TM, YP>(N, D, max_cols, col_mode, var_pcol, weight, bias, y, sn, sd, sk, mask, g_theta,  \
                                      ld_theta, g_y, g_weight, g_bias, st)
#define HLVAE_TV_BWD_Y(TS, TM)                                                                                         \
    switch (Y) {                                                                                                       \
        case 1: HLVAE_TV_BWD(TS, TM, 1); case 2: HLVAE_TV_BWD(TS, TM, 2); case 3: HLVAE_TV_BWD(TS, TM, 3);             \
        case 4: HLVAE_TV_BWD(TS, TM, 4); case 5: HLVAE_TV_BWD(TS, TM, 5); case 6: HLVAE_TV_BWD(TS, TM, 6);             \
        case 7: HLVAE_TV_BWD(TS, TM, 7); default: HLVAE_TV_BWD(TS, TM, 8);                                             \
    }
        if (dtype == HLVAE_F32) {
            if (mask_dtype == HLVAE_U8) { HLVAE_TV_BWD_Y(float, unsigned char) }
            HLVAE_TV_BWD_Y(float, float)
        }
        if (mask_dtype == HLVAE_U8) { HLVAE_TV_BWD_Y(double, unsigned char) }
        HLVAE_TV_BWD_Y(double, double)
#undef HLVAE_TV_BWD_Y
#undef HLVAE_TV_BWD
    }
#define HLVAE_TH_BWD(TS, TM, YP)                                                                                       \
    return launch_bwd<TS, TM, YP>(N, D, Y, n_tiles, col_var, col_mode, var_pcol, tile_var, weight, bias, y, sn, sd, sk, \
                                  mask, g_theta, ld_theta, g_y, g_weight, g_bias, st)
#define HLVAE_TH_BWD_Y(TS, TM)                                                                                         \
    switch (Y) {                                                                                                       \
        case 1: HLVAE_TH_BWD(TS, TM, 1); case 2: HLVAE_TH_BWD(TS, TM, 2); case 3: HLVAE_TH_BWD(TS, TM, 3);             \
        case 4: HLVAE_TH_BWD(TS, TM, 4); case 5: HLVAE_TH_BWD(TS, TM, 5); case 6: HLVAE_TH_BWD(TS, TM, 6);             \
        case 7: HLVAE_TH_BWD(TS, TM, 7); case 8: HLVAE_TH_BWD(TS, TM, 8);                                              \
        default: HLVAE_TH_BWD(TS, TM, 16);                                                                             \
    }
    if (dtype == HLVAE_F32) {
        if (mask_dtype == HLVAE_U8) { HLVAE_TH_BWD_Y(float, unsigned char) }
        HLVAE_TH_BWD_Y(float, float)
    }
    if (mask_dtype == HLVAE_U8) { HLVAE_TH_BWD_Y(double, unsigned char) }
    HLVAE_TH_BWD_Y(double, double)
#undef HLVAE_TH_BWD_Y
#undef HLVAE_TH_BWD
}
