// Replicated M x M stage of the KL upper bound, float64, one CTA per latent dimension:
//   hlvae_mxm_pre   : K0zz + eps I, Cholesky, explicit inverses iK, iH, w = iK m, G = iK H iK - iK
//                     (elbo_functions.py:148,153-157,162-163,171 / :223-231)
//   hlvae_mxm_post  : D, E, kld_qu_pu, kld_total (:170-181 / :259-277), the natural-gradient pieces
//                     (:186-191 / :279-283) and d kld / d{K0zz, m, H} in closed form
//   hlvae_natgrad_update : training.py:130-137
// Small dense algebra is done CTA-wide on matrices held in shared memory (M <= 64) or in a
// caller-provided global workspace (64 < M <= 128; L2 resident).
#include "common.cuh"

using namespace hlvae;

namespace hlvae {
bool spec_valid(const hlvae_kspec_t* sp, int Q);
}

namespace {

constexpr int MX_THREADS = 256;
constexpr int MX_NBUF = 5;

struct Mat {
    double* p;
    int ld;
    __device__ __forceinline__ double& operator()(int i, int j) const { return p[i * ld + j]; }
};

// C = X Y (TRANSX = false) or X^T Y (TRANSX = true); all M x M.  C must not alias X or Y.
// On the FP64 tensor pipe (mma.sync.m8n8k4): a warp owns 8 x 8 output tiles (tile t = warp, warp + 8, ...), per k-step
// one A and one B fragment element per lane and one DMMA.  r02: 4 x 4 register tiles with scalar FMAs spent 75
// instructions per k on 16 FMAs (per-element address arithmetic and bounds predicates inside the k loop) - 80 % of
// hlvae_mxm_post's instructions.  Sizes that are not a multiple of 8 take the predicated (zero-filled) loads.
// `K_FROM_DIAG`: X^T Y with X, Y lower triangular - the sum may start at the tile's first row / column because rows
// above the diagonal hold zeros.
template <bool TRANSX, bool K_FROM_DIAG, int NB>              // NB column tiles in flight per warp: independent DMMA chains
__device__ void mm_tiles(Mat C, Mat X, Mat Y, int M, double* __restrict__ gout) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ar = lane >> 2, ac = lane & 3;
    const int nt = (M + 7) >> 3;
    const int ng = (nt + NB - 1) / NB;
    const int M4 = (M + 3) & ~3;
    const bool full = (M & 7) == 0 && (nt % NB) == 0;
    for (int t = warp; t < nt * ng; t += MX_THREADS / 32) {
        const int it = t / ng, jg = (t - it * ng) * NB;
        const int i = it * 8 + ar;
        const int kbeg = K_FROM_DIAG ? (max(it, jg) * 8) : 0;
        double c[NB][2];
#pragma unroll
        for (int u = 0; u < NB; u++) c[u][0] = c[u][1] = 0.0;
        if (full) {
            const double* ap = TRANSX ? &X(kbeg + ac, i) : &X(i, kbeg + ac);
            const double* bp = &Y(kbeg + ac, jg * 8 + ar);
            const int astep = TRANSX ? 4 * X.ld : 4, bstep = 4 * Y.ld;
#pragma unroll 2
            for (int k0 = kbeg; k0 < M; k0 += 4) {
                const double a = *ap;
#pragma unroll
                for (int u = 0; u < NB; u++) dmma884(c[u][0], c[u][1], a, bp[u * 8]);
                ap += astep;
                bp += bstep;
            }
        } else {
            for (int k0 = kbeg; k0 < M4; k0 += 4) {
                const int kk = k0 + ac;
                const double a = (i < M && kk < M) ? (TRANSX ? X(kk, i) : X(i, kk)) : 0.0;
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    const int j = (jg + u) * 8 + ar;
                    const double b = (j < M && kk < M) ? Y(kk, j) : 0.0;
                    dmma884(c[u][0], c[u][1], a, b);
                }
            }
        }
        if (i < M) {
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int oj = (jg + u) * 8 + 2 * ac;
                if (oj < M) {
                    C(i, oj) = c[u][0];
                    if (gout) gout[i * M + oj] = c[u][0];
                }
                if (oj + 1 < M) {
                    C(i, oj + 1) = c[u][1];
                    if (gout) gout[i * M + oj + 1] = c[u][1];
                }
            }
        }
    }
    __syncthreads();
}

// Shared-memory operands (M <= 64): four independent accumulation chains per warp (a single chain is bound by the
// DMMA latency: hlvae_mxm_post 0.074 ms against 0.054 ms); operands in the L2-resident workspace (M > 64): one tile at
// a time (four in flight: 0.55 ms against 0.40 ms at M = 120 - the predicated path over a 15 x 15 tile grid).
template <bool TRANSX, bool K_FROM_DIAG = false>
__device__ void mm(Mat C, Mat X, Mat Y, int M, double* __restrict__ gout = nullptr) {
    if (M <= 64) mm_tiles<TRANSX, K_FROM_DIAG, 4>(C, X, Y, M, gout);
    else mm_tiles<TRANSX, K_FROM_DIAG, 1>(C, X, Y, M, gout);
}

// In-place lower Cholesky of A, column by column.  Returns false when a pivot is not positive.  logdet = 2 sum
// log L_jj.  `tmp` holds M doubles in shared memory.  Four threads share a row (each sums a quarter of the dot
// product, two shuffles combine them), the column is scaled by one reciprocal square root instead of a square root
// and a division per row, and the logarithms are taken once after the loop: a column costs ~450 cycles instead of
// ~1100 (the factorisations are the latency chain of hlvae_natgrad_update and hlvae_mxm_pre).  Measured and dropped: the
// register-resident scheme of kl_subject_k on two warps (thread = row, 64 x 64 fully unrolled): straight-line code that
// runs once per launch is bound by instruction fetch (hlvae_natgrad_update 0.089 -> 0.098 ms at M = 64).
__device__ bool chol(Mat A, int M, double* tmp, double& logdet) {
    const int i = threadIdx.x >> 2, q = threadIdx.x & 3;       // MX_THREADS = 4 x 64 rows; M <= 64 rows per pass
    const int passes = (M + MX_THREADS / 4 - 1) / (MX_THREADS / 4);
    for (int j = 0; j < M; j++) {
        for (int ps = 0; ps < passes; ps++) {
            const int row = i + ps * (MX_THREADS / 4);
            double s = 0.0;
            if (row >= j && row < M) {
                for (int k = q; k < j; k += 4) s = fma(-A(row, k), A(j, k), s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == 0 && row >= j && row < M) tmp[row] = A(row, j) + s;
        }
        __syncthreads();
        const double djj = tmp[j];
        if (!(djj > 0.0)) return false;       // uniform: every thread reads the same value
        const double rs = rsqrt(djj);
        for (int row = threadIdx.x; row < M; row += MX_THREADS)
            if (row >= j) A(row, j) = (row == j) ? djj * rs : tmp[row] * rs;
        __syncthreads();
    }
    // every thread needs the sum, in a fixed order: one logarithm per diagonal entry through tmp (free again)
    for (int j = threadIdx.x; j < M; j += MX_THREADS) tmp[j] = 2.0 * log(A(j, j));
    __syncthreads();
    double tot = 0.0;
    for (int j = 0; j < M; j++) tot += tmp[j];
    __syncthreads();
    logdet = tot;
    return true;
}

// B = L^-1 for lower-triangular L (column c: forward substitution L y = e_c).  B must not alias L.  Four threads
// share a column (each sums a quarter of every dot product, two shuffles combine them - the columns of the first rows
// are the long ones and only M of the 256 threads had work before); `dinv` (M doubles of shared scratch) receives the
// reciprocal diagonal first: one division per row instead of one per entry.
__device__ void tri_inv(Mat B, Mat Lm, int M, double* dinv) {
    for (int c = threadIdx.x; c < M; c += MX_THREADS) dinv[c] = 1.0 / Lm(c, c);
    __syncthreads();
    const int q = threadIdx.x & 3;
    const int passes = (M + MX_THREADS / 4 - 1) / (MX_THREADS / 4);
    for (int ps = 0; ps < passes; ps++) {
        const int c = (threadIdx.x >> 2) + ps * (MX_THREADS / 4);
        const bool live = c < M;
        if (live && q == 0) {
            for (int i = 0; i < c; i++) B(i, c) = 0.0;
            B(c, c) = dinv[c];
        }
        __syncwarp();
        // the four threads of a column walk the rows together; rows beyond a column's range cost nothing but the
        // shuffles (all 32 lanes take part in them)
        int cmin = ((threadIdx.x & ~31) >> 2) + ps * (MX_THREADS / 4);      // first column of this warp
        for (int i = cmin + 1; i < M; i++) {
            double s = 0.0;
            if (live && i > c)
                for (int k = c + q; k < i; k += 4) s = fma(Lm(i, k), B(k, c), s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (live && i > c && q == 0) B(i, c) = -s * dinv[i];
            __syncwarp();
        }
    }
    __syncthreads();
}

// C = B^T B for lower-triangular B (rows above the diagonal are zero); optional copy to global.
__device__ void ata_lower(Mat C, Mat B, int M, double* __restrict__ gout) { mm<true, true>(C, B, B, M, gout); }

__device__ void load(Mat A, const double* __restrict__ g, int M) {
    for (int e = threadIdx.x; e < M * M; e += MX_THREADS) A(e / M, e % M) = g[e];
    __syncthreads();
}

__device__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < MX_THREADS / 32; w++) s += red[w];
    __syncthreads();
    return s;
}

// Buffers: M <= 64 -> all MX_NBUF in shared memory; otherwise buffer 0 in shared memory and the
// rest in the global workspace ws[l][MX_NBUF-1][M][M].
struct Bufs {
    Mat b[MX_NBUF];
    double* vec;   // 4 * M + 16 doubles of shared scratch
};

__device__ Bufs carve(double* smem, int M, double* ws, int l) {
    Bufs B;
    const int ld = M + 1;
    const bool all_smem = (M <= 64);
    B.b[0] = Mat{smem, ld};
    double* p = smem + (size_t)M * ld;
    for (int u = 1; u < MX_NBUF; u++) {
        if (all_smem) {
            B.b[u] = Mat{p, ld};
            p += (size_t)M * ld;
        } else {
            B.b[u] = Mat{ws + ((size_t)l * (MX_NBUF - 1) + (u - 1)) * M * M, M};
        }
    }
    B.vec = p;
    return B;
}

size_t mx_smem_bytes(int M) {
    size_t mats = (M <= 64) ? MX_NBUF : 1;
    return (mats * (size_t)M * (M + 1) + 4 * (size_t)M + 16) * sizeof(double);
}

// =====================================================================================
__global__ void __launch_bounds__(MX_THREADS)
mxm_pre_k(const __grid_constant__ hlvae_kspec_t sp0, const double* __restrict__ os0, const double* __restrict__ ls0,
          int L, int Q, int M, const double* __restrict__ z, double eps, const double* __restrict__ m,
          const double* __restrict__ H, double* __restrict__ iK, double* __restrict__ iH, double* __restrict__ w,
          double* __restrict__ G, double* __restrict__ pre, double* __restrict__ ws, int32_t* __restrict__ status) {
    extern __shared__ double smem[];
    const int l = blockIdx.x, tid = threadIdx.x;
    // grid (L, 2): the K0zz factorisation and the H factorisation are independent, one CTA each
    const bool h_role = blockIdx.y == 1;
    Bufs B = carve(smem, M, ws, (int)(blockIdx.y * gridDim.x) + l);
    double* tmp = B.vec;            // [M]
    double* mv = B.vec + M;         // [M] m
    double* wv = B.vec + 2 * M;     // [M] w
    double* red = B.vec + 4 * M;    // [8]
    const size_t mm_off = (size_t)l * M * M;
    if (h_role) {                                                      // iH, log det H (:162-163 / :227-228)
        Mat A = B.b[0], P1 = B.b[1];
        double logdetH = 0.0;
        load(A, H + mm_off, M);
        if (!chol(A, M, tmp, logdetH)) {
            if (tid == 0) report_status(status, HLVAE_STATUS_NOT_PD, l, -2);
            return;
        }
        tri_inv(P1, A, M, tmp);
        ata_lower(A, P1, M, iH + mm_off);
        if (tid == 0) pre[l * 4 + 1] = logdetH;
        return;
    }
    KParams kp;
    load_kparams(kp, sp0, os0, ls0, L, l);
    const double* zl = z + (size_t)l * M * Q;
    Mat A = B.b[0], P1 = B.b[1], P2 = B.b[2];
    // K0zz + eps I (elbo_functions.py:148,153 / :223-224)
    for (int e = tid; e < M * M; e += MX_THREADS) {
        int i = e / M, j = e % M;
        A(i, j) = eval_additive(sp0, kp, zl + i * Q, zl + j * Q) + (i == j ? eps : 0.0);
    }
    for (int i = tid; i < M; i += MX_THREADS) mv[i] = m[(size_t)l * M + i];
    __syncthreads();
    double logdetK = 0.0;
    if (!chol(A, M, tmp, logdetK)) {                                   // :154 / :225
        if (tid == 0) report_status(status, HLVAE_STATUS_NOT_PD, l, -1);
        return;
    }
    tri_inv(P1, A, M, tmp);
    ata_lower(A, P1, M, iK + mm_off);                                  // iK (:155 / :226), kept in A
    // w = iK m, qf1 = m^T iK m (:177 / :272)
    for (int i = tid; i < M; i += MX_THREADS) {
        double a = 0.0;
        for (int k = 0; k < M; k++) a = fma(A(i, k), mv[k], a);
        wv[i] = a;
        w[(size_t)l * M + i] = a;
    }
    __syncthreads();
    double qf = 0.0;
    for (int i = tid; i < M; i += MX_THREADS) qf += mv[i] * wv[i];
    qf = block_sum(qf, red);
    // P2 = H iK ; tr1 = trace(H iK) (:176 / :271)
    load(P1, H + mm_off, M);
    mm<false>(P2, P1, A, M);
    double tr1 = 0.0;
    for (int i = tid; i < M; i += MX_THREADS) tr1 += P2(i, i);
    tr1 = block_sum(tr1, red);
    // G = sym(iK H iK) - iK   (:171 / :231 minus the iK of :170)
    mm<false>(P1, A, P2, M);
    for (int e = tid; e < M * M; e += MX_THREADS) {
        int i = e / M, j = e % M;
        G[mm_off + e] = 0.5 * (P1(i, j) + P1(j, i)) - A(i, j);
    }
    if (tid == 0) {
        pre[l * 4 + 0] = logdetK;
        pre[l * 4 + 2] = tr1;
        pre[l * 4 + 3] = qf;
    }
}

// =====================================================================================
// f = c0 (1/2 tr(G S) + w^T gw) + kld_qu_pu with G = iK H iK - iK, w = iK m:
//   Gamma = df/d(iK) = c0/2 (Ss iK H + H iK Ss - Ss) + c0/2 (gw m^T + m gw^T) + 1/2 (H + m m^T)
//   df/dK = -iK Gamma iK + 1/2 iK          df/dH = c0/2 iK Ss iK + 1/2 iK - 1/2 iH
//   df/dm = c0 iK gw + iK m
__global__ void __launch_bounds__(MX_THREADS)
mxm_post_k(int L, int M, double c0, double constant, const double* __restrict__ iK, const double* __restrict__ iH,
           const double* __restrict__ H, const double* __restrict__ m, const double* __restrict__ w,
           const double* __restrict__ G, const double* __restrict__ pre, const double* __restrict__ S,
           const double* __restrict__ p, const double* __restrict__ gw, const double* __restrict__ scal,
           double* __restrict__ kld, double* __restrict__ gK_over_c0, double* __restrict__ gH, double* __restrict__ gm,
           double* __restrict__ ng_m, double* __restrict__ ng_H, double* __restrict__ ws) {
    extern __shared__ double smem[];
    const int l = blockIdx.x, tid = threadIdx.x;
    Bufs B = carve(smem, M, ws, l);
    double* mv = B.vec;             // m
    double* gv = B.vec + M;         // gw
    double* pv = B.vec + 2 * M;     // p
    double* wv = B.vec + 3 * M;     // w
    double* red = B.vec + 4 * M;
    const size_t mm_off = (size_t)l * M * M;
    Mat Q0 = B.b[0], Q1 = B.b[1], Q2 = B.b[2], Q3 = B.b[3], Q4 = B.b[4];
    double js = 0.0;
    for (int e = tid; e < M * M; e += MX_THREADS) {
        int i = e / M, j = e % M;
        double ss = 0.5 * (S[mm_off + e] + S[mm_off + (size_t)j * M + i]);
        Q0(i, j) = ss;                                                 // Ss
        js = fma(G[mm_off + e], ss, js);
        Q1(i, j) = iK[mm_off + e];
        Q2(i, j) = H[mm_off + e];
    }
    for (int i = tid; i < M; i += MX_THREADS) {
        mv[i] = m[(size_t)l * M + i];
        gv[i] = gw[(size_t)l * M + i];
        pv[i] = p[(size_t)l * M + i];
        wv[i] = w[(size_t)l * M + i];
    }
    __syncthreads();
    js = 0.5 * block_sum(js, red);                                     // S parts of D (:170) and E (:172)
    mm<false>(Q3, Q1, Q0, M);                                          // U = iK Ss
    mm<false>(Q4, Q3, Q1, M);                                          // iK Ss iK  (:189 / :281)
    for (int e = tid; e < M * M; e += MX_THREADS) {
        int i = e / M, j = e % M;
        double bm = Q4(i, j), ik = Q1(i, j), ih = iH[mm_off + e];
        if (ng_H) ng_H[mm_off + e] = 0.5 * (-ih + bm + ik);           // grad_H (:191 / :283)
        gH[mm_off + e] = 0.5 * c0 * bm + 0.5 * ik - 0.5 * ih;
    }
    for (int i = tid; i < M; i += MX_THREADS) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int k = 0; k < M; k++) {
            a = fma(Q1(i, k), pv[k], a);                               // iK p
            b = fma(Q4(i, k) + Q1(i, k), mv[k], b);                    // (iK S iK + iK) m
            c = fma(Q1(i, k), gv[k], c);                               // iK gw
        }
        if (ng_m) ng_m[(size_t)l * M + i] = -a + b;                    // grad_m (:190 / :282)
        gm[(size_t)l * M + i] = c0 * c + wv[i];
    }
    __syncthreads();
    mm<true>(Q4, Q3, Q2, M);                                           // V1 = U^T H = Ss iK H
    for (int e = tid; e < M * M; e += MX_THREADS) {
        int i = e / M, j = e % M;
        if (i <= j) {
            double g = 0.5 * c0 * (Q4(i, j) + Q4(j, i) - Q0(i, j)) + 0.5 * c0 * (gv[i] * mv[j] + mv[i] * gv[j]) +
                       0.5 * (Q2(i, j) + mv[i] * mv[j]);
            Q0(i, j) = g;                                              // Gamma (symmetric)
            Q0(j, i) = g;
        }
    }
    __syncthreads();
    mm<false>(Q3, Q1, Q0, M);                                          // iK Gamma
    mm<false>(Q4, Q3, Q1, M);                                          // iK Gamma iK
    const double ic0 = 1.0 / c0;
    for (int e = tid; e < M * M; e += MX_THREADS) {
        int i = e / M, j = e % M;
        gK_over_c0[mm_off + e] = (-Q4(i, j) + 0.5 * Q1(i, j)) * ic0;
    }
    if (tid == 0) {
        const double* sc = scal + (size_t)l * HLVAE_NSCAL;
        const double* pr = pre + (size_t)l * 4;
        double kq = 0.5 * (pr[2] + pr[3] - (double)M + pr[0] - pr[1]);                     // :180 / :275
        double v = c0 * (0.5 * (sc[0] + sc[1] + sc[2] - sc[3]) + js) + kq;               // :181 / :277
        if (l == 0) v -= constant;
        // deterministic total (the replicated loss must be bit-identical on every data-parallel rank and from launch
        // to launch): per-l values are parked in kld[2 + l]; the CTA that arrives last adds them up in index order
        kld[2 + l] = v;
        __threadfence();
        const double arrived = atomicAdd(kld + 1, 1.0);
        if (arrived == (double)(gridDim.x - 1)) {
            __threadfence();
            double tot = 0.0;
            for (int q = 0; q < (int)gridDim.x; q++) tot += __ldcg(kld + 2 + q);
            kld[0] += tot;
        }
    }
}

// =====================================================================================
// training.py:130-137
__global__ void __launch_bounds__(MX_THREADS)
natgrad_k(int L, int M, double lr, const double* __restrict__ m, const double* __restrict__ H,
          const double* __restrict__ iH_in, const double* __restrict__ grad_m, const double* __restrict__ grad_H,
          double* __restrict__ m_out, double* __restrict__ H_out, double* __restrict__ ws,
          int32_t* __restrict__ status) {
    extern __shared__ double smem[];
    const int l = blockIdx.x, tid = threadIdx.x;
    Bufs B = carve(smem, M, ws, l);
    double* tmp = B.vec;
    double* mv = B.vec + M;
    double* v = B.vec + 2 * M;
    const size_t mm_off = (size_t)l * M * M;
    Mat A = B.b[0], P1 = B.b[1], P2 = B.b[2], P3 = B.b[3];
    for (int i = tid; i < M; i += MX_THREADS) mv[i] = m[(size_t)l * M + i];
    double ldet;
    if (iH_in) {                                                       // H^-1 already known (hlvae_mxm_pre)
        load(P2, iH_in + mm_off, M);
    } else {
        load(A, H + mm_off, M);
        if (!chol(A, M, tmp, ldet)) {                                  // :131
            if (tid == 0) report_status(status, HLVAE_STATUS_NOT_PD, l, -3);
            return;
        }
        tri_inv(P1, A, M, tmp);
        ata_lower(P2, P1, M, nullptr);                                 // iH (:132)
    }
    // grad_H goes through P1 (free until the triangular inverse): coalesced loads once, instead of a strided global
    // read per thread and k in the product below and a transposed one in the update
    load(P1, grad_H + mm_off, M);
    // v = iH m - lr (grad_m - 2 grad_H m)   (:136-137): four threads per row
    {
        const int q = tid & 3;
        for (int i0 = 0; i0 < M; i0 += MX_THREADS / 4) {               // uniform trip count: every lane shuffles
            const int i = i0 + (tid >> 2);
            double a = 0.0, b = 0.0;
            for (int k = q; k < M && i < M; k += 4) {
                a = fma(P2(i, k), mv[k], a);
                b = fma(P1(i, k), mv[k], b);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            b += __shfl_xor_sync(0xffffffffu, b, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            b += __shfl_xor_sync(0xffffffffu, b, 2);
            if (q == 0 && i < M) v[i] = a - lr * (grad_m[(size_t)l * M + i] - 2.0 * b);
        }
    }
    // iH_new = iH + lr (grad_H + grad_H^T)   (:133)
    for (int e = tid; e < M * M; e += MX_THREADS) {
        int i = e / M, j = e % M;
        A(i, j) = P2(i, j) + lr * (P1(i, j) + P1(j, i));
    }
    __syncthreads();
    if (!chol(A, M, tmp, ldet)) {                                      // :134
        if (tid == 0) report_status(status, HLVAE_STATUS_NOT_PD, l, -4);
        return;
    }
    tri_inv(P1, A, M, tmp);
    ata_lower(P3, P1, M, H_out + mm_off);                              // H_new (:135)
    {
        const int q = tid & 3;
        for (int i0 = 0; i0 < M; i0 += MX_THREADS / 4) {
            const int i = i0 + (tid >> 2);
            double a = 0.0;
            for (int k = q; k < M && i < M; k += 4) a = fma(P3(i, k), v[k], a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (q == 0 && i < M) m_out[(size_t)l * M + i] = a;         // m_new (:136)
        }
    }
}

// =====================================================================================
// M x M epilogue shared by the evaluation-time users of the streaming statistics S = K0zx iB K0xz, p = K0zx iB y:
//   W = K0zz + eps I + sym(S)          (elbo_functions.py:43-45,91-93; validation.py:54-56; utils.py:136,158)
//   scal = [log det(K0zz + eps I), log det W, |L_W^-1 p|^2, sum(S * iK)]     (:47-53,95-105 / validation.py:57-66)
//   a = W^-1 p (utils.py:162 / :246),  c = (K0zz + eps I)^-1 (p - S a) = iK K0zx mu_tilde (utils.py:169 / :249),
//   iW = W^-1 (validation.py:71, elbo_functions.py:111)
// Cholesky factors + explicit triangular inverses in float64; every output is optional.
__global__ void __launch_bounds__(MX_THREADS)
mxm_aux_k(const __grid_constant__ hlvae_kspec_t sp0, const double* __restrict__ os0, const double* __restrict__ ls0,
          int L, int Q, int M, const double* __restrict__ z, double eps, const double* __restrict__ S,
          const double* __restrict__ p, double* __restrict__ scal, double* __restrict__ a_out,
          double* __restrict__ c_out, double* __restrict__ iW_out, double* __restrict__ ws,
          int32_t* __restrict__ status) {
    extern __shared__ double smem[];
    const int l = blockIdx.x, tid = threadIdx.x;
    Bufs B = carve(smem, M, ws, l);
    double* tmp = B.vec;            // [M]
    double* pv = B.vec + M;         // [M] p
    double* yv = B.vec + 2 * M;     // [M] L_W^-1 p, later p - S a
    double* av = B.vec + 3 * M;     // [M] a
    double* red = B.vec + 4 * M;
    const size_t mm_off = (size_t)l * M * M;
    KParams kp;
    load_kparams(kp, sp0, os0, ls0, L, l);
    const double* zl = z + (size_t)l * M * Q;
    Mat A = B.b[0], P1 = B.b[1], P2 = B.b[2], Wm = B.b[3];
    for (int e = tid; e < M * M; e += MX_THREADS) {
        const int i = e / M, j = e % M;
        const double k = eval_additive(sp0, kp, zl + i * Q, zl + j * Q) + (i == j ? eps : 0.0);
        A(i, j) = k;
        const double s = S ? 0.5 * (S[mm_off + e] + S[mm_off + (size_t)j * M + i]) : 0.0;
        Wm(i, j) = k + s;
    }
    for (int i = tid; i < M; i += MX_THREADS) pv[i] = p ? p[(size_t)l * M + i] : 0.0;
    __syncthreads();
    double logdetK = 0.0, logdetW = 0.0;
    if (!chol(A, M, tmp, logdetK)) {
        if (tid == 0) report_status(status, HLVAE_STATUS_NOT_PD, l, -1);
        return;
    }
    tri_inv(P1, A, M, tmp);
    ata_lower(P2, P1, M, nullptr);                                     // iK
    double trS = 0.0;
    if (S)
        for (int e = tid; e < M * M; e += MX_THREADS) trS = fma(S[mm_off + e], P2(e / M, e % M), trS);
    trS = block_sum(trS, red);
    if (!chol(Wm, M, tmp, logdetW)) {
        if (tid == 0) report_status(status, HLVAE_STATUS_NOT_PD, l, -5);
        return;
    }
    tri_inv(P1, Wm, M, tmp);                                                // L_W^-1 (lower)
    for (int i = tid; i < M; i += MX_THREADS) {
        double s = 0.0;
        for (int k = 0; k <= i; k++) s = fma(P1(i, k), pv[k], s);
        yv[i] = s;
    }
    __syncthreads();
    double qf2 = 0.0;
    for (int i = tid; i < M; i += MX_THREADS) qf2 = fma(yv[i], yv[i], qf2);
    qf2 = block_sum(qf2, red);
    for (int i = tid; i < M; i += MX_THREADS) {                        // a = L_W^-T y
        double s = 0.0;
        for (int k = i; k < M; k++) s = fma(P1(k, i), yv[k], s);
        av[i] = s;
        if (a_out) a_out[(size_t)l * M + i] = s;
    }
    __syncthreads();
    if (c_out) {
        for (int i = tid; i < M; i += MX_THREADS) {                    // p - S a  (S as accumulated, not symmetrised)
            double s = 0.0;
            if (S)
                for (int k = 0; k < M; k++) s = fma(S[mm_off + (size_t)i * M + k], av[k], s);
            yv[i] = pv[i] - s;
        }
        __syncthreads();
        for (int i = tid; i < M; i += MX_THREADS) {
            double s = 0.0;
            for (int k = 0; k < M; k++) s = fma(P2(i, k), yv[k], s);
            c_out[(size_t)l * M + i] = s;
        }
    }
    if (iW_out) ata_lower(A, P1, M, iW_out + mm_off);
    if (tid == 0 && scal) {
        scal[l * 4 + 0] = logdetK;
        scal[l * 4 + 1] = logdetW;
        scal[l * 4 + 2] = qf2;
        scal[l * 4 + 3] = trS;
    }
}

template <typename K>
int set_smem(K kern, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

extern "C" int64_t hlvae_mxm_workspace_doubles(int L, int M) {
    return (M <= 64) ? 0 : (int64_t)2 * L * (MX_NBUF - 1) * M * M;      // hlvae_mxm_pre runs 2 CTAs per latent dim
}

extern "C" int hlvae_mxm_pre(const hlvae_kspec_t* spec0, const double* os0, const double* ls0, int L, int Q, int M,
                             const double* z, double eps, const double* m, const double* H, double* iK, double* iH,
                             double* w, double* G, double* pre, double* ws, int32_t* status, void* stream) {
    if (!hlvae::spec_valid(spec0, Q) || L <= 0 || Q <= 0 || Q > HLVAE_MAX_Q || M <= 0 || !z || !m || !H || !iK ||
        !iH || !w || !G || !pre)
        return HLVAE_E_ARG;
    if (M > 128) return HLVAE_E_UNSUPPORTED;
    if (M > 64 && !ws) return HLVAE_E_ARG;
    size_t smem = mx_smem_bytes(M);
    int rc = set_smem(mxm_pre_k, smem);
    if (rc) return rc;
    mxm_pre_k<<<dim3(L, 2), MX_THREADS, smem, (cudaStream_t)stream>>>(*spec0, os0, ls0, L, Q, M, z, eps, m, H, iK, iH, w, G,
                                                             pre, ws, status);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_mxm_post(int L, int M, double c0, double constant, const double* iK, const double* iH,
                              const double* H, const double* m, const double* w, const double* G, const double* pre,
                              const double* S, const double* p, const double* gw, const double* scal, double* kld,
                              double* gK_over_c0, double* gH, double* gm, double* ng_m, double* ng_H, double* ws,
                              void* stream) {
    if (L <= 0 || M <= 0 || !(c0 > 0.0) || !iK || !iH || !H || !m || !w || !G || !pre || !S || !p || !gw || !scal ||
        !kld || !gK_over_c0 || !gH || !gm)
        return HLVAE_E_ARG;
    if (M > 128) return HLVAE_E_UNSUPPORTED;
    if (M > 64 && !ws) return HLVAE_E_ARG;
    size_t smem = mx_smem_bytes(M);
    int rc = set_smem(mxm_post_k, smem);
    if (rc) return rc;
    mxm_post_k<<<L, MX_THREADS, smem, (cudaStream_t)stream>>>(L, M, c0, constant, iK, iH, H, m, w, G, pre, S, p, gw,
                                                              scal, kld, gK_over_c0, gH, gm, ng_m, ng_H, ws);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_natgrad_update(int L, int M, double lr, const double* m, const double* H, const double* iH,
                                    const double* grad_m, const double* grad_H, double* m_out, double* H_out,
                                    double* ws, int32_t* status, void* stream) {
    if (L <= 0 || M <= 0 || !m || !H || !grad_m || !grad_H || !m_out || !H_out) return HLVAE_E_ARG;
    if (M > 128) return HLVAE_E_UNSUPPORTED;
    if (M > 64 && !ws) return HLVAE_E_ARG;
    size_t smem = mx_smem_bytes(M);
    int rc = set_smem(natgrad_k, smem);
    if (rc) return rc;
    natgrad_k<<<L, MX_THREADS, smem, (cudaStream_t)stream>>>(L, M, lr, m, H, iH, grad_m, grad_H, m_out, H_out, ws,
                                                             status);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_mxm_aux(const hlvae_kspec_t* spec0, const double* os0, const double* ls0, int L, int Q, int M,
                             const double* z, double eps, const double* S, const double* p, double* scal, double* a,
                             double* c, double* iW, double* ws, int32_t* status, void* stream) {
    if (!hlvae::spec_valid(spec0, Q) || L <= 0 || Q <= 0 || Q > HLVAE_MAX_Q || M <= 0 || !z) return HLVAE_E_ARG;
    if (M > 128) return HLVAE_E_UNSUPPORTED;
    if (M > 64 && !ws) return HLVAE_E_ARG;
    size_t smem = mx_smem_bytes(M);
    int rc = set_smem(mxm_aux_k, smem);
    if (rc) return rc;
    mxm_aux_k<<<L, MX_THREADS, smem, (cudaStream_t)stream>>>(*spec0, os0, ls0, L, Q, M, z, eps, S, p, scal, a, c, iW, ws,
                                                             status);
    HLVAE_CHECK_LAUNCH();
    return 0;
}
