// Streaming part of the KL upper bound (elbo_functions.py:118-285): everything that scales
// with the minibatch.  Kernels:
//   kl_subject_k     : one warp per (subject, latent dim)  - T x T work (B_s, Cholesky, inverse), T <= 32
//   kl_subject_big_k : one CTA per (subject, latent dim) for subjects of 33 .. HLVAE_TMAX rows
//   kl_panel_k       : one CTA per (latent dim, subject chunk) - K0xz rows on the fly (values parked in tensor
//                      memory for the gradient pass), B^-1 K0xz, sufficient statistics on the FP64 tensor pipe,
//                      gradients; panel inputs software-pipelined with cp.async, Z by a TMA bulk copy.
// Forward and backward are produced in one pass: with w = iK m and G = iK H iK - iK held
// fixed, J = 1/2 (A + B + C + D + E - F) is linear in S, so dJ/d(inputs) needs nothing
// from a later stage ("single-pass gradient").
#include "common.cuh"

using namespace hlvae;

namespace hlvae {
bool spec_valid(const hlvae_kspec_t* sp, int Q);
}

namespace {

struct AccOff {
    int64_t o[10];
};

// =====================================================================================
// kl_subject_k
// =====================================================================================
constexpr int SJ_WARPS = 2;
constexpr int SJ_TW = 32;     // longest subject of the warp-per-pair kernel (a subject's rows are the lanes of a warp)

// One component's descriptor pulled into registers field by field (a struct copy indexed by a
// run-time component number would be placed in local memory).
struct CompRegs {
    int se_col, ndisc;
    int disc_kind[HLVAE_MAX_DISC], disc_col[HLVAE_MAX_DISC];
    __device__ __forceinline__ void load(const hlvae_kspec_t& sp, int r) {
        se_col = sp.comp[r].se_col;
        ndisc = sp.comp[r].ndisc;
#pragma unroll
        for (int f = 0; f < HLVAE_MAX_DISC; f++) {
            disc_kind[f] = sp.comp[r].disc_kind[f];
            disc_col[f] = sp.comp[r].disc_col[f];
        }
    }
    // unscaled value of the component at rows (xa, xb) of one covariate matrix; d = xa - xb on its SE column.
    // ndisc is the same for every lane, so the factor count selects a branch without divergence: components with
    // no or one discrete factor (all the generators of kernel_gen.py produce) skip the predicated general loop.
    __device__ __forceinline__ bool match1(const double* __restrict__ xa, const double* __restrict__ xb, int f) const {
        const double a = xa[disc_col[f]], b = xb[disc_col[f]];
        return (disc_kind[f] == HLVAE_KIND_CAT) ? (a - b == 0.0) : (a + b == 2.0);
    }
    __device__ __forceinline__ double value(const double* __restrict__ xa, const double* __restrict__ xb, double hil2,
                                            double& d, const double* __restrict__ etab) const {
        d = 0.0;
        bool ok = true;
        if (ndisc == 1) {
            ok = match1(xa, xb, 0);
        } else if (ndisc > 1) {
#pragma unroll
            for (int f = 0; f < HLVAE_MAX_DISC; f++)
                if (f < ndisc) ok = ok && match1(xa, xb, f);
        }
        if (!ok) return 0.0;
        if (se_col < 0) return 1.0;
        d = xa[se_col] - xb[se_col];
        return exp_nonpos_tab(-(d * d) * hil2, etab);
    }
    // The same value through straight-line bodies for the three shapes kernel_gen.py:199-310 builds (SE, SE x
    // categorical, categorical): no data-dependent branch, the discrete factor folded into the exponent insert of the
    // exponential.  The shape is the same for every lane (and loop invariant for the caller).  For a mismatch the
    // value is 0 and d is left at the covariate difference (every consumer multiplies d by the value).
    __device__ __forceinline__ double value_fast(const double* __restrict__ xa, const double* __restrict__ xb, double hil2,
                                                 double& d, const double* __restrict__ etab) const {
        const bool cat0 = disc_kind[0] == HLVAE_KIND_CAT;
        if (se_col >= 0 && (ndisc == 0 || (ndisc == 1 && cat0))) {
            d = xa[se_col] - xb[se_col];
            const double arg = (d * -hil2) * d;
            if (ndisc == 0) return exp_nonpos_tab_sel<false>(arg, etab, true);
            return exp_nonpos_tab_sel<true>(arg, etab, xa[disc_col[0]] == xb[disc_col[0]]);
        }
        if (se_col < 0 && ndisc == 1 && cat0) {
            d = 0.0;
            return (xa[disc_col[0]] == xb[disc_col[0]]) ? 1.0 : 0.0;
        }
        return value(xa, xb, hil2, d, etab);
    }
};

// Per-warp shared memory: covariate rows, three T x T matrices, and the hyper-parameters of this
// latent dimension by component ({outputscale, 1/(2 l^2), 1/l^3} x {K0, K1}).
constexpr int SJ_KP = 6 * HLVAE_MAX_COMPS;

// =====================================================================================
// kl_subject_k: one warp per (subject, latent dim).  The T x T algebra is laid out around what bounded the first
// version of this kernel (ncu r01k: shared-memory bandwidth - two conflicting LDS.64 per DFMA in every product;
// 0.63 ms at configs[1], this form 0.38 ms):
//   * Cholesky and the triangular inverse keep "their" row / column in REGISTERS (loops fully unrolled over the
//     padded size TP, so every register index is static); the only shared-memory operand left is a row of L that all
//     lanes read at the same address (broadcast, conflict-free, two entries per LDS.128);
//   * the three T^3 products - B^-1 = L^-T L^-1, X = Ktil B^-1 and B^-1 X - run on the FP64 tensor pipe
//     (mma.sync.m8n8k4: one LDS.64 per operand element per 8 x 8 x 4 block instead of two per multiply-add), with the
//     leading dimension TP + 4 (= 4 or 12 mod 16), for which both fragment shapes are bank-conflict free;
//   * dJ/dB is contracted with dB/d(theta_1) straight from the accumulator fragments (never stored).
// Rows T..TP-1 are padded with an identity block (B) / zeros (Ktil), which leaves every result untouched.
// =====================================================================================
template <int TP>
struct SubjShape {
    static constexpr int LDA = TP + 4;
    static constexpr int NT = TP / 8;
    static constexpr int NTRI = TP * (TP + 1) / 2;
    static constexpr int per_warp = SJ_TW * HLVAE_MAX_Q + 2 * TP * LDA + TP + SJ_KP;
    // per CTA: exp table, then the (i, j) pairs of the lower-triangle walk (16 bits each)
    static constexpr int shared_doubles = (HLVAE_EXP_TAB + (NTRI + 3) / 4 + 1) & ~1;   // even: 16-byte aligned rows
    static constexpr int min_blocks = TP <= 24 ? 8 : 4;
};

template <int TP, typename TS>
__global__ void __launch_bounds__(SJ_WARPS * 32, SubjShape<TP>::min_blocks)
kl_subject_k(const __grid_constant__ hlvae_kspec_t sp0, const double* __restrict__ os0, const double* __restrict__ ls0,
              const __grid_constant__ hlvae_kspec_t sp1, const double* __restrict__ os1, const double* __restrict__ ls1,
              const double* __restrict__ noise, int L, int Q, const double* __restrict__ x, int64_t ldx,
              const int32_t* __restrict__ row_idx, const int32_t* __restrict__ subj_ptr,
              const int32_t* __restrict__ tt_ptr, int n_subj, const TS* __restrict__ log_v, int64_t ld_lv,
              double* __restrict__ binv, int64_t tt_total, double* __restrict__ acc, const AccOff off,
              TS* __restrict__ g_logv, double gscale, int32_t* __restrict__ status) {
    using S2 = SubjShape<TP>;
    constexpr int LDA = S2::LDA, NT = S2::NT;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* etab = smem;
    unsigned short* tri = reinterpret_cast<unsigned short*>(smem + HLVAE_EXP_TAB);   // element t -> (i << 8) | j
    double* xs = smem + S2::shared_doubles + (size_t)warp * S2::per_warp;
    double* Am = xs + SJ_TW * HLVAE_MAX_Q;             // B -> L -> B^-1
    double* Bm = Am + TP * LDA;                         // L^-1 -> K0ss / Ktil -> X
    double* dinv = Bm + TP * LDA;                       // 1 / L_jj
    double* kp = dinv + TP;
    const int ar = lane >> 2, ac = lane & 3;            // fragment coordinates

    exp2_table_fill(etab, threadIdx.x, SJ_WARPS * 32);
    for (int i = threadIdx.x; i < TP; i += SJ_WARPS * 32)
        for (int j = 0; j <= i; j++) tri[i * (i + 1) / 2 + j] = (unsigned short)((i << 8) | j);
    __syncthreads();
    const int64_t n_pairs = (int64_t)n_subj * L;
    for (int64_t pair = (int64_t)blockIdx.x * SJ_WARPS + warp; pair < n_pairs; pair += (int64_t)gridDim.x * SJ_WARPS) {
    int s, l;
    if (n_pairs <= 0x7fffffff) {                          // 32-bit division (the 64-bit one is ~50 instructions)
        s = (int)((unsigned)pair / (unsigned)L);
        l = (int)((unsigned)pair - (unsigned)s * (unsigned)L);
    } else {
        s = (int)(pair / L);
        l = (int)(pair % L);
    }
    const int r0 = subj_ptr[s];
    const int T = subj_ptr[s + 1] - r0;
    if (T <= 0) continue;
    if (T > SJ_TW && T <= HLVAE_TMAX) continue;         // kl_subject_big_k's
    if (T > TP) {
        if (lane == 0) report_status(status, HLVAE_STATUS_T_TOO_LARGE, l, s);
        continue;
    }
    const int TL = T * (T + 1) / 2;
    int g = -1;
    double ev = 0.0, lv = 0.0;
    if (lane < T) {
        g = row_idx[r0 + lane];
        for (int q = 0; q < Q; q++) xs[lane * Q + q] = x[(int64_t)g * ldx + q];
        lv = (double)log_v[(int64_t)g * ld_lv + l];
        ev = exp(lv);
    }
    if (lane < 2 * HLVAE_MAX_COMPS) {
        const int which = lane >> 3, r = lane & 7;
        const int nc = which ? sp1.ncomp : sp0.ncomp;
        double o = 0.0, h = 0.0, i3 = 0.0;
        if (r < nc) {
            o = (which ? os1 : os0)[(int64_t)r * L + l];
            const double e_ = (which ? ls1 : ls0)[(int64_t)r * L + l];
            const double i2 = 1.0 / (e_ * e_);
            h = 0.5 * i2;
            i3 = i2 / e_;
        }
        kp[which * 24 + r] = o;
        kp[which * 24 + 8 + r] = h;
        kp[which * 24 + 16 + r] = i3;
    }
    // zero both matrices (contiguous, 16-byte aligned: 128-bit stores), then the identity padding of B
    for (int e = lane; e < TP * LDA; e += 32) reinterpret_cast<double2*>(Am)[e] = make_double2(0.0, 0.0);
    __syncwarp();
    if (lane >= T && lane < TP) Am[lane * LDA + lane] = 1.0;
    __syncwarp();
    const double nz = noise[l];
    // B_s = K1(x_s, x_s) + noise I (elbo_functions.py:249-250): lower triangle, mirrored
    {   // components outermost (descriptor and hyper-parameters in registers), lower triangle, then mirrored
        for (int t = lane; t < TL; t += 32) {
            const int ij = tri[t], i = ij >> 8, j = ij & 255;
            if (i == j) Am[i * LDA + j] = nz;                // (off-diagonal entries were zero-filled above)
        }
        for (int r = 0; r < sp1.ncomp; r++) {
            CompRegs c;
            c.load(sp1, r);
            const double osr = kp[24 + r], hil2 = kp[32 + r];
            for (int t = lane; t < TL; t += 32) {
                const int ij = tri[t], i = ij >> 8, j = ij & 255;
                double d;
                Am[i * LDA + j] = fma(osr, c.value_fast(xs + i * Q, xs + j * Q, hil2, d, etab), Am[i * LDA + j]);
            }
        }
        for (int t = lane; t < TL; t += 32) {
            const int ij = tri[t], i = ij >> 8, j = ij & 255;
            Am[j * LDA + i] = Am[i * LDA + j];
        }
    }
    __syncwarp();

    // ---- Cholesky (:251), left-looking; lane i keeps row i in registers, finished columns are published to Am so
    // that row j can be read back by every lane at one address
    double rr[TP];
    {
        const int li = lane < TP ? lane : TP - 1;
#pragma unroll
        for (int k = 0; k < TP; k++) rr[k] = Am[li * LDA + k];
    }
    __syncwarp();
    bool bad = false;
    double ljj = 1.0;
#pragma unroll
    for (int j = 0; j < TP; j++) {
        double s0 = rr[j], s1 = 0.0;
        const double* rj = Am + j * LDA;                 // L[j][0..j) are final (written at steps < j)
#pragma unroll
        for (int k = 0; k + 1 < j; k += 2) {
            const double2 v = *reinterpret_cast<const double2*>(rj + k);
            s0 = fma(-rr[k], v.x, s0);
            s1 = fma(-rr[k + 1], v.y, s1);
        }
        if (j & 1) s0 = fma(-rr[j - 1], rj[j - 1], s0);
        const double sum = s0 + s1;
        const double djj = __shfl_sync(0xffffffffu, sum, j);
        if (!(djj > 0.0)) bad = true;
        const double rd = rsqrt(djj);
        const double val = (lane == j) ? djj * rd : sum * rd;
        rr[j] = val;
        if (lane == j) {
            ljj = val;
            dinv[j] = rd;                                 // 1 / L_jj for the triangular solve
        }
        if (lane >= j && lane < TP) Am[lane * LDA + j] = val;
        __syncwarp();
    }
    if (bad) {                                           // uniform: djj is a broadcast value
        if (lane == 0) report_status(status, HLVAE_STATUS_NOT_PD, l, s);
        continue;
    }
    // C term (:258): log det B = 2 sum log L_ii
    const double logdet = warp_sum(lane < T ? 2.0 * log(ljj) : 0.0);
    // ---- L^-1: lane c solves L y = e_c with y in registers (zero above the diagonal); L[i][k] is a broadcast read
    {
#pragma unroll
        for (int i = 0; i < TP; i++) {
            double a0 = 0.0, a1 = 0.0;
            const double* ri = Am + i * LDA;
#pragma unroll
            for (int k = 0; k + 1 < i; k += 2) {
                const double2 v = *reinterpret_cast<const double2*>(ri + k);
                a0 = fma(v.x, rr[k], a0);                 // rr[k] now holds y[k] for k < i
                a1 = fma(v.y, rr[k + 1], a1);
            }
            if (i & 1) a0 = fma(ri[i - 1], rr[i - 1], a0);
            const double rdi = dinv[i];
            const double y = (lane == i) ? rdi : (lane < i ? -(a0 + a1) * rdi : 0.0);
            rr[i] = y;                                   // overwrites L[lane][i] (dead: row entries k < i only feed the solve)
        }
    }
    // NOTE the solve above reuses rr: entry rr[k] (k < i) must be y[k], and it is - y[k] was stored at iteration k;
    // entries k >= i still hold L[lane][k], which the loop never reads (k < i only).
    // publish L^-1 (lane c owns column c) for the tensor-pipe products
#pragma unroll
    for (int i = 0; i < TP; i++)
        if (lane < TP) Bm[i * LDA + lane] = rr[i];
    __syncwarp();
    // ---- B^-1 = L^-T L^-1 (explicit inverse, as :252): lower tiles, k from the tile's first row on
    double* bout = binv + (int64_t)l * tt_total + tt_ptr[s];
    {
        double cfr[NT * (NT + 1) / 2][2];
        int q = 0;
#pragma unroll
        for (int ti = 0; ti < NT; ti++)
#pragma unroll
            for (int tj = 0; tj <= ti; tj++, q++) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int k0 = ti * 8; k0 < TP; k0 += 4) {
                    const double a = Bm[(k0 + ac) * LDA + ti * 8 + ar];     // (L^-1)^T[i][k] = L^-1[k][i]
                    const double b = Bm[(k0 + ac) * LDA + tj * 8 + ar];
                    dmma884(c0, c1, a, b);
                }
                cfr[q][0] = c0;
                cfr[q][1] = c1;
            }
        __syncwarp();
        q = 0;
#pragma unroll
        for (int ti = 0; ti < NT; ti++)
#pragma unroll
            for (int tj = 0; tj <= ti; tj++, q++) {
                const int i = ti * 8 + ar, j = tj * 8 + 2 * ac;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const double v = cfr[q][u];
                    Am[i * LDA + j + u] = v;
                    Am[(j + u) * LDA + i] = v;
                    if (i < T && j + u < T) {
                        bout[i * T + j + u] = v;
                        bout[(j + u) * T + i] = v;
                    }
                }
            }
    }
    __syncwarp();

    // K0(x_s, x_s) (:248), one evaluation per component and entry.  With wgt = B^-1_ij (x2 off the diagonal):
    //   B + D1 terms (:257,259) = sum_r os_r sum wgt v_r + sum_i B^-1_ii e^logv_i,   dJ/dK0ss = B^-1 / 2.
    // (L^-1 is dead: its buffer is zeroed with 128-bit stores and collects K0ss)
    for (int e = lane; e < TP * LDA / 2; e += 32) reinterpret_cast<double2*>(Bm)[e] = make_double2(0.0, 0.0);
    __syncwarp();
    double bd = 0.0;
    for (int r = 0; r < sp0.ncomp; r++) {
        CompRegs c;
        c.load(sp0, r);
        const double osr = kp[r], hil2 = kp[8 + r], il3 = kp[16 + r];
        double gos = 0.0, gls = 0.0;
        for (int t = lane; t < TL; t += 32) {
            const int ij = tri[t], i = ij >> 8, j = ij & 255;
            double d;
            const double v = c.value_fast(xs + i * Q, xs + j * Q, hil2, d, etab);
            const double wv = ((i == j) ? 1.0 : 2.0) * Am[i * LDA + j] * v;
            gos += wv;
            gls = fma(wv * d, d, gls);
            Bm[i * LDA + j] = fma(osr, v, Bm[i * LDA + j]);
        }
        gos = warp_sum(gos);
        gls = warp_sum(gls);
        bd = fma(osr, gos, bd);
        if (lane == 0) {
            atomicAdd(acc + off.o[HLVAE_ACC_GOS0] + (int64_t)r * L + l, 0.5 * gos);
            if (c.se_col >= 0) atomicAdd(acc + off.o[HLVAE_ACC_GLS0] + (int64_t)r * L + l, 0.5 * gls * osr * il3);
        }
    }
    __syncwarp();
    {   // lower triangle of K0ss is in place (zeros elsewhere): mirror
        for (int t = lane; t < TL; t += 32) {
            const int ij = tri[t], i = ij >> 8, j = ij & 255;
            Bm[j * LDA + i] = Bm[i * LDA + j];
        }
    }
    __syncwarp();
    double bdiag = 0.0;
    if (lane < T) {   // Ktil = K0ss + diag(e^logv); its B term and the log-variance gradient
        const double bii = Am[lane * LDA + lane];
        Bm[lane * LDA + lane] += ev;
        bdiag = bii * ev;
        g_logv[(int64_t)g * L + l] = (TS)(gscale * 0.5 * (bii * ev - 1.0));
    }
    bd += warp_sum(bdiag);
    __syncwarp();
    // ---- X = Ktil B^-1 (all tiles, accumulators in registers), then stored over Ktil
    {
        double xf[NT * NT][2];
#pragma unroll
        for (int ti = 0; ti < NT; ti++)
#pragma unroll
            for (int tj = 0; tj < NT; tj++) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int k0 = 0; k0 < TP; k0 += 4) {
                    const double a = Bm[(ti * 8 + ar) * LDA + k0 + ac];
                    const double b = Am[(k0 + ac) * LDA + tj * 8 + ar];
                    dmma884(c0, c1, a, b);
                }
                xf[ti * NT + tj][0] = c0;
                xf[ti * NT + tj][1] = c1;
            }
        __syncwarp();
#pragma unroll
        for (int ti = 0; ti < NT; ti++)
#pragma unroll
            for (int tj = 0; tj < NT; tj++)
                *reinterpret_cast<double2*>(&Bm[(ti * 8 + ar) * LDA + tj * 8 + 2 * ac]) =
                    make_double2(xf[ti * NT + tj][0], xf[ti * NT + tj][1]);
    }
    __syncwarp();
    // ---- dJ/dB_s (part without K0xz) = 1/2 (B^-1 - B^-1 Ktil B^-1): lower tiles on the tensor pipe, contracted with
    // dB/d(theta1) from the accumulator fragments (symmetric: off-diagonal entries count twice)
    {
        double gq[NT * (NT + 1) / 2][2];
        int q = 0;
#pragma unroll
        for (int ti = 0; ti < NT; ti++)
#pragma unroll
            for (int tj = 0; tj <= ti; tj++, q++) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int k0 = 0; k0 < TP; k0 += 4) {
                    const double a = Am[(ti * 8 + ar) * LDA + k0 + ac];
                    const double b = Bm[(k0 + ac) * LDA + tj * 8 + ar];
                    dmma884(c0, c1, a, b);
                }
                const int i = ti * 8 + ar, j = tj * 8 + 2 * ac;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const double cu = u ? c1 : c0;
                    const bool use = i < T && j + u <= i;
                    gq[q][u] = use ? ((i == j + u) ? 0.5 : 1.0) * (Am[i * LDA + j + u] - cu) : 0.0;
                }
            }
        for (int r = 0; r < sp1.ncomp; r++) {
            CompRegs c;
            c.load(sp1, r);
            const double osr = kp[24 + r], hil2 = kp[32 + r], il3 = kp[40 + r];
            double gos = 0.0, gls = 0.0;
            q = 0;
#pragma unroll
            for (int ti = 0; ti < NT; ti++)
#pragma unroll
                for (int tj = 0; tj <= ti; tj++, q++) {
                    const int i = ti * 8 + ar, j = tj * 8 + 2 * ac;
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        if (gq[q][u] != 0.0) {
                            double d;
                            const double gv = gq[q][u] * c.value_fast(xs + i * Q, xs + (j + u) * Q, hil2, d, etab);
                            gos += gv;
                            gls = fma(gv * d, d, gls);
                        }
                    }
                }
            gos = warp_sum(gos);
            gls = warp_sum(gls);
            if (lane == 0) {
                atomicAdd(acc + off.o[HLVAE_ACC_GOS1] + (int64_t)r * L + l, gos);
                if (c.se_col >= 0) atomicAdd(acc + off.o[HLVAE_ACC_GLS1] + (int64_t)r * L + l, gls * osr * il3);
            }
        }
    }
    const double fsum = warp_sum(lv);
    if (lane == 0) {
        double* scal = acc + off.o[HLVAE_ACC_SCAL] + (int64_t)l * HLVAE_NSCAL;
        atomicAdd(scal + 1, bd);
        atomicAdd(scal + 2, logdet);
        atomicAdd(scal + 3, fsum);
    }
    __syncwarp();
    }   // pairs
}

// =====================================================================================
// kl_subject_big_k: the same per-(subject, latent dim) work for subjects of 33 .. HLVAE_TMAX rows, which do not fit
// the lanes of one warp: one CTA of SB_THREADS threads per pair, the three T x T matrices in shared memory, thread =
// row in the factorisation (two barriers per column), thread = entries (i, j) in the products.  Plain FP64 FMAs -
// long subjects are the exception (the reference's Python loop, elbo_functions.py:243-266, takes any length), so this
// path is built for correctness first; the shipped data sets (T = 20) never reach it.
// =====================================================================================
constexpr int SB_THREADS = 128;
constexpr int SB_LD = HLVAE_TMAX + 1;      // odd leading dimension: conflict-free rows and columns
constexpr size_t SB_SMEM_DOUBLES = (size_t)HLVAE_TMAX * HLVAE_MAX_Q + 3 * (size_t)HLVAE_TMAX * SB_LD + 2 * HLVAE_TMAX +
                                   SJ_KP + HLVAE_EXP_TAB + 2 * (SB_THREADS / 32) + 2;

__device__ __forceinline__ double block_sum(double v, double* red, int tid) {   // every thread gets the sum
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < SB_THREADS / 32; w++) s += red[w];
    return s;
}

template <typename TS>
__global__ void __launch_bounds__(SB_THREADS)
kl_subject_big_k(const __grid_constant__ hlvae_kspec_t sp0, const double* __restrict__ os0, const double* __restrict__ ls0,
                 const __grid_constant__ hlvae_kspec_t sp1, const double* __restrict__ os1, const double* __restrict__ ls1,
                 const double* __restrict__ noise, int L, int Q, const double* __restrict__ x, int64_t ldx,
                 const int32_t* __restrict__ row_idx, const int32_t* __restrict__ subj_ptr,
                 const int32_t* __restrict__ tt_ptr, int n_subj, const TS* __restrict__ log_v, int64_t ld_lv,
                 double* __restrict__ binv, int64_t tt_total, double* __restrict__ acc, const AccOff off,
                 TS* __restrict__ g_logv, double gscale, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) double smem[];
    constexpr int LD = SB_LD;
    double* xs = smem;                                   // [T][Q]
    double* Am = xs + HLVAE_TMAX * HLVAE_MAX_Q;          // B -> L -> B^-1
    double* Bm = Am + HLVAE_TMAX * LD;                   // L^-1 -> K0ss / Ktil
    double* Cm = Bm + HLVAE_TMAX * LD;                   // products
    double* evs = Cm + HLVAE_TMAX * LD;                  // e^logv
    double* dcol = evs + HLVAE_TMAX;                     // diagonal of L
    double* kp = dcol + HLVAE_TMAX;                      // hyper-parameters, as in kl_subject_k
    double* etab = kp + SJ_KP;
    double* red = etab + HLVAE_EXP_TAB;
    int* flag = reinterpret_cast<int*>(red + SB_THREADS / 32);
    const int tid = threadIdx.x;
    exp2_table_fill(etab, tid, SB_THREADS);
    const int64_t n_pairs = (int64_t)n_subj * L;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const int s = (int)(pair / L), l = (int)(pair % L);
        const int r0 = subj_ptr[s];
        const int T = subj_ptr[s + 1] - r0;
        if (T <= SJ_TW) continue;                        // kl_subject_k's
        if (T > HLVAE_TMAX) {
            if (tid == 0) report_status(status, HLVAE_STATUS_T_TOO_LARGE, l, s);
            continue;
        }
        __syncthreads();                                 // the previous pair is done with shared memory
        int g = -1;
        double lv = 0.0;
        if (tid < T) {
            g = row_idx[r0 + tid];
            for (int q = 0; q < Q; q++) xs[tid * Q + q] = x[(int64_t)g * ldx + q];
            lv = (double)log_v[(int64_t)g * ld_lv + l];
            evs[tid] = exp(lv);
        }
        if (tid < 2 * HLVAE_MAX_COMPS) {
            const int which = tid >> 3, r = tid & 7;
            const int nc = which ? sp1.ncomp : sp0.ncomp;
            double o = 0.0, h = 0.0, i3 = 0.0;
            if (r < nc) {
                o = (which ? os1 : os0)[(int64_t)r * L + l];
                const double e_ = (which ? ls1 : ls0)[(int64_t)r * L + l];
                const double i2 = 1.0 / (e_ * e_);
                h = 0.5 * i2;
                i3 = i2 / e_;
            }
            kp[which * 24 + r] = o;
            kp[which * 24 + 8 + r] = h;
            kp[which * 24 + 16 + r] = i3;
        }
        if (tid == 0) *flag = 0;
        __syncthreads();
        const double nz = noise[l];
        // ---- B_s = K1(x_s, x_s) + noise I (elbo_functions.py:249-250)
        for (int e = tid; e < T * T; e += SB_THREADS) {
            const int i = e / T, j = e - i * T;
            if (j > i) continue;
            double k1 = (i == j) ? nz : 0.0;
            for (int r = 0; r < sp1.ncomp; r++) {
                CompRegs c;
                c.load(sp1, r);
                double d;
                k1 = fma(kp[24 + r], c.value_fast(xs + i * Q, xs + j * Q, kp[32 + r], d, etab), k1);
            }
            Am[i * LD + j] = k1;
            Am[j * LD + i] = k1;
        }
        __syncthreads();
        // ---- Cholesky (:251), left-looking: thread i owns row i
        for (int j = 0; j < T; j++) {
            double sum = 0.0;
            if (tid >= j && tid < T) {
                sum = Am[tid * LD + j];
                for (int k = 0; k < j; k++) sum = fma(-Am[tid * LD + k], Am[j * LD + k], sum);
                if (tid == j) {
                    if (!(sum > 0.0)) *flag = 1;
                    dcol[j] = sqrt(sum);
                }
            }
            __syncthreads();
            if (tid >= j && tid < T) Am[tid * LD + j] = (tid == j) ? dcol[j] : sum / dcol[j];
            __syncthreads();
        }
        if (*flag) {
            if (tid == 0) report_status(status, HLVAE_STATUS_NOT_PD, l, s);
            continue;
        }
        // C term (:258)
        const double logdet = block_sum(tid < T ? 2.0 * log(dcol[tid]) : 0.0, red, tid);
        // ---- L^-1: thread c solves L y = e_c down its own column of Bm
        if (tid < T) {
            const int c = tid;
            for (int i = 0; i < T; i++) {
                double y = 0.0;
                if (i >= c) {
                    double a = (i == c) ? 1.0 : 0.0;
                    for (int k = c; k < i; k++) a = fma(-Am[i * LD + k], Bm[k * LD + c], a);
                    y = a / dcol[i];
                }
                Bm[i * LD + c] = y;
            }
        }
        __syncthreads();
        // ---- B^-1 = L^-T L^-1 (:252) -> Cm, then over Am and out to global memory
        double* bout = binv + (int64_t)l * tt_total + tt_ptr[s];
        for (int e = tid; e < T * T; e += SB_THREADS) {
            const int i = e / T, j = e - i * T;
            if (j > i) continue;
            double a = 0.0;
            for (int k = i; k < T; k++) a = fma(Bm[k * LD + i], Bm[k * LD + j], a);
            Cm[i * LD + j] = a;
            Cm[j * LD + i] = a;
        }
        __syncthreads();
        for (int e = tid; e < T * T; e += SB_THREADS) {
            const int i = e / T, j = e - i * T;
            const double v = Cm[i * LD + j];
            Am[i * LD + j] = v;
            bout[e] = v;
        }
        __syncthreads();
        // ---- K0(x_s, x_s) (:248) into Bm; B + D1 terms (:257,259) and the K0 hyper-gradients (dJ/dK0ss = B^-1 / 2)
        for (int e = tid; e < T * T; e += SB_THREADS) Bm[(e / T) * LD + e % T] = 0.0;
        __syncthreads();
        double bd = 0.0;
        for (int r = 0; r < sp0.ncomp; r++) {
            CompRegs c;
            c.load(sp0, r);
            const double osr = kp[r], hil2 = kp[8 + r], il3 = kp[16 + r];
            double gos = 0.0, gls = 0.0;
            for (int e = tid; e < T * T; e += SB_THREADS) {
                const int i = e / T, j = e - i * T;
                double d;
                const double v = c.value_fast(xs + i * Q, xs + j * Q, hil2, d, etab);
                const double wv = Am[i * LD + j] * v;
                gos += wv;
                gls = fma(wv * d, d, gls);
                Bm[i * LD + j] = fma(osr, v, Bm[i * LD + j]);
            }
            gos = block_sum(gos, red, tid);
            gls = block_sum(gls, red, tid);
            bd = fma(osr, gos, bd);
            if (tid == 0) {
                atomicAdd(acc + off.o[HLVAE_ACC_GOS0] + (int64_t)r * L + l, 0.5 * gos);
                if (c.se_col >= 0) atomicAdd(acc + off.o[HLVAE_ACC_GLS0] + (int64_t)r * L + l, 0.5 * gls * osr * il3);
            }
        }
        __syncthreads();
        double bdiag = 0.0;
        if (tid < T) {   // Ktil = K0ss + diag(e^logv); its B term and the log-variance gradient
            const double bii = Am[tid * LD + tid], ev = evs[tid];
            Bm[tid * LD + tid] += ev;
            bdiag = bii * ev;
            g_logv[(int64_t)g * L + l] = (TS)(gscale * 0.5 * (bii * ev - 1.0));
        }
        bd += block_sum(bdiag, red, tid);
        // ---- X = Ktil B^-1 -> Cm
        for (int e = tid; e < T * T; e += SB_THREADS) {
            const int i = e / T, j = e - i * T;
            double a = 0.0;
            for (int k = 0; k < T; k++) a = fma(Bm[i * LD + k], Am[k * LD + j], a);
            Cm[i * LD + j] = a;
        }
        __syncthreads();
        // ---- dJ/dB_s (part without K0xz) = 1/2 (B^-1 - B^-1 Ktil B^-1) -> Bm, contracted with dB/d(theta1)
        for (int e = tid; e < T * T; e += SB_THREADS) {
            const int i = e / T, j = e - i * T;
            double a = 0.0;
            for (int k = 0; k < T; k++) a = fma(Am[i * LD + k], Cm[k * LD + j], a);
            Bm[i * LD + j] = 0.5 * (Am[i * LD + j] - a);
        }
        __syncthreads();
        for (int r = 0; r < sp1.ncomp; r++) {
            CompRegs c;
            c.load(sp1, r);
            const double osr = kp[24 + r], hil2 = kp[32 + r], il3 = kp[40 + r];
            double gos = 0.0, gls = 0.0;
            for (int e = tid; e < T * T; e += SB_THREADS) {
                const int i = e / T, j = e - i * T;
                double d;
                const double gv = Bm[i * LD + j] * c.value_fast(xs + i * Q, xs + j * Q, hil2, d, etab);
                gos += gv;
                gls = fma(gv * d, d, gls);
            }
            gos = block_sum(gos, red, tid);
            gls = block_sum(gls, red, tid);
            if (tid == 0) {
                atomicAdd(acc + off.o[HLVAE_ACC_GOS1] + (int64_t)r * L + l, gos);
                if (c.se_col >= 0) atomicAdd(acc + off.o[HLVAE_ACC_GLS1] + (int64_t)r * L + l, gls * osr * il3);
            }
        }
        const double fsum = block_sum(lv, red, tid);
        if (tid == 0) {
            double* scal = acc + off.o[HLVAE_ACC_SCAL] + (int64_t)l * HLVAE_NSCAL;
            atomicAdd(scal + 1, bd);
            atomicAdd(scal + 2, logdet);
            atomicAdd(scal + 3, fsum);
        }
    }
}

template <typename TS>
int launch_subject_big(const hlvae_kspec_t* spec0, const double* os0, const double* ls0, const hlvae_kspec_t* spec1,
                       const double* os1, const double* ls1, const double* noise, int L, int Q, const double* x,
                       int64_t ldx, const int32_t* row_idx, const int32_t* subj_ptr, const int32_t* tt_ptr, int n_subj,
                       const void* log_v, int64_t ld_lv, double* binv, int64_t tt_total, double* acc, const AccOff& off,
                       void* g_logv, double gscale, int32_t* status, cudaStream_t st) {
    const size_t smem = SB_SMEM_DOUBLES * sizeof(double);
    auto kern = kl_subject_big_k<TS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t pairs = (int64_t)n_subj * L;
    const unsigned grid = (unsigned)(pairs < 148 * 8 ? pairs : 148 * 8);
    kern<<<grid, SB_THREADS, smem, st>>>(*spec0, os0, ls0, *spec1, os1, ls1, noise, L, Q, x, ldx, row_idx, subj_ptr,
                                         tt_ptr, n_subj, (const TS*)log_v, ld_lv, binv, tt_total, acc, off, (TS*)g_logv,
                                         gscale, status);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

template <int TP, typename TS>
int launch_subject(const hlvae_kspec_t* spec0, const double* os0, const double* ls0, const hlvae_kspec_t* spec1,
                    const double* os1, const double* ls1, const double* noise, int L, int Q, const double* x,
                    int64_t ldx, const int32_t* row_idx, const int32_t* subj_ptr, const int32_t* tt_ptr, int n_subj,
                    const void* log_v, int64_t ld_lv, double* binv, int64_t tt_total, double* acc, const AccOff& off,
                    void* g_logv, double gscale, int32_t* status, cudaStream_t st) {
    const size_t smem = ((size_t)SJ_WARPS * SubjShape<TP>::per_warp + SubjShape<TP>::shared_doubles) * sizeof(double);
    auto kern = kl_subject_k<TP, TS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t pairs = (int64_t)n_subj * L;
    const unsigned grid = (unsigned)((pairs + SJ_WARPS - 1) / SJ_WARPS);
    kern<<<grid, SJ_WARPS * 32, smem, st>>>(*spec0, os0, ls0, *spec1, os1, ls1, noise, L, Q, x, ldx, row_idx, subj_ptr,
                                            tt_ptr, n_subj, (const TS*)log_v, ld_lv, binv, tt_total, acc, off,
                                            (TS*)g_logv, gscale, status);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

// =====================================================================================
// kl_panel_k
// =====================================================================================
constexpr int PN_SMAX = 16;   // subjects per panel
// The unscaled values of the K0 components are kept from the K0xz pass to the gradient pass in TENSOR MEMORY: a thread
// only ever re-reads the values it produced itself (same inducing point, same rows), and a thread of warp w owns lane
// 32 (w % 4) + lane id of every TMEM column, so 128 columns per group of four warps are 64 private doubles per thread
// that cost neither registers nor shared memory (r02 kept them in 41 - 98 KB of shared memory, two components at most
// in the two-CTAs-per-SM shape; the third was evaluated twice).
// NC = components whose gradient sums stay in registers over the CTA's whole chunk (the others are reduced per panel).
constexpr int PN_TMEM_COLS_PER_GROUP = 128;
constexpr int PN_CHUNK_CSR = 256;   // subjects of a chunk whose CSR offsets are copied to shared memory
template <int MP, int RP, bool G_SMEM, int NC>
struct PanelSmem {
    static constexpr int LD = MP + 4;          // leading dim of row panels: conflict-free DMMA fragment loads
    static constexpr int LDB = RP + 4;         // leading dim of the dense block-diagonal B^-1 panel
    static constexpr size_t doubles = (size_t)HLVAE_MAX_Q * MP /*Zs*/ + MP /*ws*/ + (G_SMEM ? (size_t)MP * LD : 0) +
                                      2 * (size_t)RP * LD /*Kb,Vb*/ + (size_t)RP * LDB /*Bp*/ +
                                      2 * (size_t)RP * HLVAE_MAX_Q /*xs*/ + 5 * RP /*mus, mraw, rv, rho*/ +
                                      (size_t)MP * HLVAE_MAX_COMPS /*zacc*/ + 4 * HLVAE_MAX_COMPS + 8 /*hyper acc + A*/ +
                                      8 * HLVAE_MAX_COMPS /*kps, kps1*/ + HLVAE_EXP_TAB /*etab*/;
    static constexpr int n_lower_tiles = (RP / 8) * (RP / 8 + 1) / 2;
    static constexpr int n_s_tiles = (MP / 8) * (MP / 8 + 1) / 2;
    static constexpr size_t ints = 4 * RP + 4 * (PN_SMAX + 1) + 8 + n_lower_tiles + n_s_tiles + 2 * (PN_CHUNK_CSR + 1);
    static constexpr size_t bytes = doubles * 8 + ints * 4;
};

// NT threads per CTA: 512 (one CTA per SM) or 256 (two CTAs per SM whose barrier and latency stalls cover each other;
// register file: 2 x 256 x 128).  Warps form a (NT / 128) x 4 grid over the tiles of the [RP x MP] products.
template <int MP, int RP, bool G_SMEM, int NT, int NC, typename TS>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1)
kl_panel_k(const __grid_constant__ hlvae_kspec_t sp0, const double* __restrict__ os0, const double* __restrict__ ls0,
           const __grid_constant__ hlvae_kspec_t sp1, const double* __restrict__ os1, const double* __restrict__ ls1,
           int L, int Q, int M, const double* __restrict__ x, int64_t ldx, const double* __restrict__ z,
           const int32_t* __restrict__ row_idx, const int32_t* __restrict__ subj_ptr,
           const int32_t* __restrict__ tt_ptr, int n_subj, int subj_per_chunk, const TS* __restrict__ mu,
           int64_t ld_mu, const double* __restrict__ w, const double* __restrict__ G,
           const double* __restrict__ binv, int64_t tt_total, double* __restrict__ acc, const AccOff off,
           TS* __restrict__ g_mu, TS* __restrict__ qdiag, double gscale, int32_t* __restrict__ status) {
    using SM = PanelSmem<MP, RP, G_SMEM, NC>;
    constexpr int PN_NCACHE = NC;
    constexpr int LD = SM::LD;
    constexpr int LDB = SM::LDB;
    constexpr int PN_THREADS = NT;
    constexpr int WGI = NT / 128;               // warp grid: WGI x 4
    // S = sum K0xz^T B^-1 K0xz is symmetric: only the 8x8 tiles on or below the diagonal are accumulated (36 of 64
    // at M = 64, 136 of 256 at M = 128), handed out to the warps in runs of the row-major order so that consecutive
    // tiles of a warp mostly share their row; the flush mirrors the off-diagonal tiles.
    constexpr int SNT = (MP / 8) * (MP / 8 + 1) / 2;            // lower tiles of S
    constexpr int SBASE = SNT / (NT / 32), SREM = SNT % (NT / 32);   // every warp owns SBASE tiles, the first SREM one more
    constexpr int SPW = SBASE + (SREM ? 1 : 0);                 // accumulator slots per warp
    constexpr int WR = (RP / 8 + WGI - 1) / WGI;   // row tiles per warp for [RP x MP] outputs (the last warp row may own fewer)
    constexpr int WC = (MP / 8) / 4;            // col tiles per warp for [RP x MP] outputs
    constexpr int NGRP = PN_THREADS / MP;       // row groups in the element-wise phases
    constexpr int RPT = RP / NGRP;              // rows per thread in the element-wise phases
    constexpr int TM_COLS = (NT / 128) * PN_TMEM_COLS_PER_GROUP;                       // TMEM columns of this CTA
    // PARK (two-CTA shapes): the S accumulators and the per-component gradient sums live in this thread's TMEM
    // columns between the phases that use them, which frees ~40 registers for the exponential pipelines of P1 / P5b
    constexpr bool PARK = NT == 256;
    constexpr int HGN = (3 * NC + 1) & ~1;                                              // gradient sums (padded: even)
    constexpr int PARK_S = PARK ? 4 * SPW : 0;                                          // columns (2 per double)
    constexpr int PARK_H = PARK ? 2 * HGN : 0;
    constexpr int NCT_RAW = (PN_TMEM_COLS_PER_GROUP - PARK_S - PARK_H) / (2 * RPT);
    constexpr int NCT = NCT_RAW < HLVAE_MAX_COMPS ? NCT_RAW : HLVAE_MAX_COMPS;          // components cached in TMEM
    static_assert(NCT >= 1, "TMEM column plan");
    static_assert(WR >= 1 && WC >= 1 && SPW >= 1 && RPT >= 2 && RPT % 2 == 0 && RP % NGRP == 0, "tile shape");
    static_assert(NT / 32 >= RP / 8, "one warp per row tile in the r = K0xz w - mu product");
    static_assert(TM_COLS == 256 || TM_COLS == 512, "TMEM allocation: a power of two");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Zs = reinterpret_cast<double*>(smem_raw);          // [Q][MP] transposed
    double* ws = Zs + HLVAE_MAX_Q * MP;
    double* Gs = ws + MP;                                      // [MP][LD] if G_SMEM
    double* Kb = Gs + (G_SMEM ? MP * LD : 0);                  // [RP][LD]  K0xz, later W = V G
    double* Vb = Kb + RP * LD;                                 // [RP][LD]  B^-1 K0xz
    double* Bp = Vb + RP * LD;                                 // [RP][LDB] dense block-diagonal B^-1, later dJ/dB
    double* xs = Bp + RP * LDB;                                // [2][RP][8] covariate rows of this / the next panel
    double* mus = xs + 2 * RP * HLVAE_MAX_Q;
    TS* mraw = reinterpret_cast<TS*>(mus + RP);                // [2][RP] mu as stored, this / the next panel
    double* rv = mus + 3 * RP;
    double* rho = rv + RP;
    double* zacc = rho + RP;                                   // [MP][MAX_COMPS]
    double* hyp = zacc + MP * HLVAE_MAX_COMPS;                 // gos0, gls0, gos1, gls1, A
    double* kps = hyp + 4 * HLVAE_MAX_COMPS + 8;               // K0 hyper-parameters by component: os, hil2, il2, il3
    double* kps1 = kps + 4 * HLVAE_MAX_COMPS;                  // same for K1
    double* etab = kps1 + 4 * HLVAE_MAX_COMPS;                 // 2^(j/64) for exp_nonpos_tab
    // panel descriptors, two sets: the panel being computed and the one being loaded
    int* grow_b = reinterpret_cast<int*>(etab + HLVAE_EXP_TAB);    // [2][RP] minibatch row of a panel row
    int* sor_b = grow_b + 2 * RP;                              // [2][RP] subject (within the panel) of a panel row, -1 = padding
    int* sub_r0_b = sor_b + 2 * RP;                            // [2][PN_SMAX+1] first panel row of a subject
    int* sub_b0_b = sub_r0_b + 2 * (PN_SMAX + 1);              // [2][PN_SMAX+1] offsets of the T x T blocks
    int* meta = sub_b0_b + 2 * (PN_SMAX + 1);                  // [0,1] subjects, [2,3] first subject, [4] cursor, [5] TMEM base
    int* tl_tab = meta + 8;                                    // (row tile << 8 | column tile) of the lower triangle
    int* s_tab = tl_tab + SM::n_lower_tiles;                   // the same for the M x M tiles of S
    int* csr_r = s_tab + SM::n_s_tiles;                        // subj_ptr[s_begin ..] of this chunk
    int* csr_t = csr_r + PN_CHUNK_CSR + 1;                     // tt_ptr[s_begin ..]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wi = warp >> 2, wj = warp & 3;
    const int l = blockIdx.y;
    const int s_begin = blockIdx.x * subj_per_chunk;
    const int s_end = min(n_subj, s_begin + subj_per_chunk);
    const int em = tid % MP;                                   // this thread's inducing point in element-wise phases
    const int eg = tid / MP;                                   // and its row group

    // ---- per-CTA setup: inducing points of this latent dim (transposed), w, G.  The [M, Q] block of Z is one
    // contiguous 16-byte aligned span (whenever M Q is even): a single TMA bulk copy (cp.async.bulk + mbarrier) brings
    // it into the not yet used K0xz buffer while the rest of the set-up runs; it is transposed from there.
    const bool z_bulk = ((M * Q) & 1) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0;
    unsigned long long* zbar = reinterpret_cast<unsigned long long*>(rho);       // (rho is first written in P3a)
    if (z_bulk) {
        if (tid == 0) {
            tma_mbar_init(zbar, 1);
            tma_mbar_expect_tx(zbar, (unsigned)(M * Q * 8));
            tma_bulk_g2s(Kb, z + (int64_t)l * M * Q, (unsigned)(M * Q * 8), zbar);
        }
    } else {
        for (int e = tid; e < HLVAE_MAX_Q * MP; e += PN_THREADS) {
            int q = e / MP, m = e % MP;
            Zs[e] = (q < Q && m < M) ? z[((int64_t)l * M + m) * Q + q] : 0.0;
        }
    }
    for (int m = tid; m < MP; m += PN_THREADS) ws[m] = (m < M) ? w[(int64_t)l * M + m] : 0.0;
    exp2_table_fill(etab, tid, PN_THREADS);
    if (warp == 0) tmem_alloc<TM_COLS>(reinterpret_cast<uint32_t*>(meta + 5));
    for (int i = tid; i < RP / 8; i += PN_THREADS)
        for (int j = 0; j <= i; j++) tl_tab[i * (i + 1) / 2 + j] = (i << 8) | j;
    for (int i = tid; i < MP / 8; i += PN_THREADS)
        for (int j = 0; j <= i; j++) s_tab[i * (i + 1) / 2 + j] = (i << 8) | j;
    for (int i = tid; i <= PN_CHUNK_CSR && s_begin + i <= s_end; i += PN_THREADS) {
        csr_r[i] = subj_ptr[s_begin + i];
        csr_t[i] = tt_ptr[s_begin + i];
    }
    if (G_SMEM) {
        for (int e = tid; e < MP * LD; e += PN_THREADS) {
            int i = e / LD, j = e % LD;
            Gs[e] = (i < M && j < M) ? G[((int64_t)l * M + i) * M + j] : 0.0;
        }
    }
    for (int e = tid; e < MP * HLVAE_MAX_COMPS + 4 * HLVAE_MAX_COMPS + 8; e += PN_THREADS) zacc[e] = 0.0;
    if (tid < HLVAE_MAX_COMPS) {
        double o = 0.0, h = 0.0, i2 = 0.0, i3 = 0.0;
        if (tid < sp0.ncomp) {
            o = os0[(int64_t)tid * L + l];
            const double e_ = ls0[(int64_t)tid * L + l];
            i2 = 1.0 / (e_ * e_);
            h = 0.5 * i2;
            i3 = i2 / e_;
        }
        kps[tid] = o;
        kps[HLVAE_MAX_COMPS + tid] = h;
        kps[2 * HLVAE_MAX_COMPS + tid] = i2;
        kps[3 * HLVAE_MAX_COMPS + tid] = i3;
    } else if (tid < 2 * HLVAE_MAX_COMPS) {
        const int r = tid - HLVAE_MAX_COMPS;
        double o = 0.0, h = 0.0, i2 = 0.0, i3 = 0.0;
        if (r < sp1.ncomp) {
            o = os1[(int64_t)r * L + l];
            const double e_ = ls1[(int64_t)r * L + l];
            i2 = 1.0 / (e_ * e_);
            h = 0.5 * i2;
            i3 = i2 / e_;
        }
        kps1[r] = o;
        kps1[HLVAE_MAX_COMPS + r] = h;
        kps1[2 * HLVAE_MAX_COMPS + r] = i2;
        kps1[3 * HLVAE_MAX_COMPS + r] = i3;
    }
    const double* Gl = G + (int64_t)l * M * M;

    double sacc[SPW][2];
#pragma unroll
    for (int t = 0; t < SPW; t++) sacc[t][0] = sacc[t][1] = 0.0;
    double p_acc = 0.0, gw_acc = 0.0, a_acc = 0.0;
    // gradient sums of the cached K0 components, kept per thread over the CTA's whole chunk (reduced once at the end)
    double hgs[HGN];                                           // [3 r + {0, 1, 2}]: sums of g v, g v d, g v d^2
#pragma unroll
    for (int r = 0; r < HGN; r++) hgs[r] = 0.0;

    if (tid == 0) meta[4] = s_begin;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (z_bulk) {                                              // Z has landed in Kb as [M][Q]: transpose, zero padded
        tma_mbar_wait(zbar, 0);
        for (int e = tid; e < HLVAE_MAX_Q * MP; e += PN_THREADS) {
            int q = e / MP, m = e % MP;
            Zs[e] = (q < Q && m < M) ? Kb[m * Q + q] : 0.0;
        }
        __syncthreads();                                       // Kb is free again; Zs is complete
    }
    const uint32_t tmem_base = (uint32_t)meta[5];
    // this thread's private columns: lane quadrant of its warp, column block of its group of four warps
    const uint32_t tm_mine = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * PN_TMEM_COLS_PER_GROUP);
    const uint32_t tm_comp = tm_mine + PARK_S + PARK_H;       // the cached component values start here
    if constexpr (PARK) {
        tmem_store_doubles<2 * SPW>(tm_mine, &sacc[0][0]);
        tmem_store_doubles<HGN>(tm_mine + PARK_S, hgs);
        tmem_wait_st();
    }
    const double wm = ws[em];
    // CSR offsets of subject s (s_begin <= s <= s_end): from the shared-memory copy where it reaches
    auto rows_at = [&](int s) { return s - s_begin <= PN_CHUNK_CSR ? csr_r[s - s_begin] : subj_ptr[s]; };
    auto tt_at = [&](int s) { return s - s_begin <= PN_CHUNK_CSR ? csr_t[s - s_begin] : tt_ptr[s]; };

    // ---- P0, software pipelined: while panel n is computed, thread 0 packs whole subjects into panel n + 1 (at most
    // RP rows) and every thread issues the asynchronous copies (cp.async) that gather its covariate rows, its mu
    // column and the subjects' B^-1 blocks - scattered onto the block diagonal of the dense panel, the rest of each
    // row zero-filled by the same thread - so that a panel starts with its inputs in shared memory instead of
    // waiting for L2 / HBM behind two barriers (r02: 14 % of the kernel's stall samples).
    auto pack = [&](int pz) {                                  // the 32 lanes of one warp: lane i looks at subject cursor + i
        int* r0 = sub_r0_b + pz * (PN_SMAX + 1);
        int* b0 = sub_b0_b + pz * (PN_SMAX + 1);
        int* so = sor_b + pz * RP;
        int s0 = meta[4], ns, T, incl, incl2;
        while (true) {
            const int s = s0 + lane;
            const bool valid = lane < PN_SMAX && s < s_end;
            T = valid ? rows_at(s + 1) - rows_at(s) : 0;
            const bool big = T > HLVAE_TMAX || T > RP;         // cannot be packed: skipped and reported
            incl = T;
            incl2 = T * T;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {                 // inclusive prefix sums of T and T^2
                const int a = __shfl_up_sync(0xffffffffu, incl, o), b2 = __shfl_up_sync(0xffffffffu, incl2, o);
                if (lane >= o) { incl += a; incl2 += b2; }
            }
            const unsigned okm = __ballot_sync(0xffffffffu, valid && !big && incl <= RP);
            ns = __ffs(~okm) - 1;                              // the leading run of subjects that fit (lanes >= PN_SMAX never do)
            if (ns == 0 && (__ballot_sync(0xffffffffu, big) & 1u)) {
                if (lane == 0) report_status(status, HLVAE_STATUS_T_TOO_LARGE, l, s0);
                s0++;
                continue;
            }
            break;
        }
        if (lane < ns) {
            r0[lane + 1] = incl;
            b0[lane + 1] = incl2;
            for (int t = 0; t < T; t++) so[incl - T + t] = lane;
        }
        const int rows = ns > 0 ? __shfl_sync(0xffffffffu, incl, ns - 1) : 0;
        for (int row = rows + lane; row < RP; row += 32) so[row] = -1;
        if (lane == 0) {
            r0[0] = 0;
            b0[0] = 0;
            meta[pz] = ns;
            meta[2 + pz] = s0;
            meta[4] = s0 + ns;
        }
    };
    // minibatch row of this thread's panel row in descriptor set pz (-1: none)
    auto next_row = [&](int pz) {
        const int ns = meta[pz];
        if (ns == 0 || tid >= sub_r0_b[pz * (PN_SMAX + 1) + ns]) return -1;
        return (int)row_idx[rows_at(meta[2 + pz]) + tid];
    };
    auto issue_loads = [&](int pz, int g) {
        const int ns = meta[pz];
        if (ns > 0) {
            const int* r0 = sub_r0_b + pz * (PN_SMAX + 1);
            const int* b0 = sub_b0_b + pz * (PN_SMAX + 1);
            const int* so = sor_b + pz * RP;
            if (tid < RP) {
                double* xn = xs + (pz * RP + tid) * HLVAE_MAX_Q;
                if (g >= 0) {
                    grow_b[pz * RP + tid] = g;
                    const double* src = x + (int64_t)g * ldx;
                    for (int q = 0; q < Q; q++) cp_async<8>(xn + q, src + q);
                    cp_async<(int)sizeof(TS)>(mraw + pz * RP + tid, mu + (int64_t)g * ld_mu + l);
                } else {
#pragma unroll
                    for (int q = 0; q < HLVAE_MAX_Q; q++) xn[q] = 0.0;
                }
            }
            const double* bsrc = binv + (int64_t)l * tt_total + tt_at(meta[2 + pz]);
            for (int row = warp; row < RP; row += PN_THREADS / 32) {
                const int k = so[row];
                int rs = 0, T = 0, bo = 0;
                if (k >= 0) {
                    rs = r0[k];
                    T = r0[k + 1] - rs;
                    bo = b0[k] + (row - rs) * T - rs;          // block element (row - rs, col - rs)
                }
                for (int col = lane; col < LDB; col += 32) {
                    if (col >= rs && col < rs + T) cp_async<8>(&Bp[row * LDB + col], bsrc + bo + col);
                    else Bp[row * LDB + col] = 0.0;
                }
            }
        }
        cp_async_commit();
    };
    if (warp == PN_THREADS / 32 - 1) pack(0);
    __syncthreads();
    issue_loads(0, next_row(0));
    int par = 0;

    while (true) {
        cp_async_wait_all();
        __syncthreads();      // the inputs of this panel have landed; every thread is done with the previous one
        const int nsub = meta[par];
        if (nsub == 0) break;
        const int* sub_r0 = sub_r0_b + par * (PN_SMAX + 1);
        const int* sub_of_row = sor_b + par * RP;
        const int* grow = grow_b + par * RP;
        const double* xsc = xs + par * RP * HLVAE_MAX_Q;       // [RP][8]
        const int R = sub_r0[nsub];
        const int R8 = (R + 7) & ~7;
        if (tid < R) mus[tid] = (double)mraw[par * RP + tid];

        // ---- P1: K0xz rows (elbo_functions.py:147 / :222), zero padded to [R8][MP].
        // Thread = (inducing point em, row group eg); components outermost so that the spec and the inducing point's
        // covariates stay in registers and row covariates are warp broadcasts.  Straight-line bodies: the RPT rows of
        // a thread are independent, so their exponentials interleave and hide each other's latency; rows >= R read
        // zero-filled covariates and inducing points >= M zeros (their values are finite and never used).  The shape
        // of a component is the same for every thread and selects one of four bodies without divergence: the three
        // forms kernel_gen.py:199-310 builds - SE, SE x categorical, categorical - read one packed {SE covariate,
        // discrete covariate} pair per row (one LDS.128) and fold the discrete factor into the exponent insert of the
        // exponential; everything else takes the general body.  The unscaled values go to this thread's TMEM columns.
        {
            double kacc[RPT];
#pragma unroll
            for (int k = 0; k < RPT; k++) kacc[k] = 0.0;
            for (int r = 0; r < sp0.ncomp; r++) {
                CompRegs c;
                c.load(sp0, r);
                const int nd = c.ndisc;
                const bool has_se = c.se_col >= 0;
                const bool cat0 = c.disc_kind[0] == HLVAE_KIND_CAT;
                const double zse = has_se ? Zs[c.se_col * MP + em] : 0.0;
                const double z0 = nd > 0 ? Zs[c.disc_col[0] * MP + em] : 0.0;
                const double nh = -kps[HLVAE_MAX_COMPS + r], osr = kps[r];
                // this thread's rows are eg + k NGRP: fixed offsets from one base per covariate column
                const double* xa = xsc + eg * HLVAE_MAX_Q + (has_se ? c.se_col : 0);
                const double* xb = xsc + eg * HLVAE_MAX_Q + (nd > 0 ? c.disc_col[0] : 0);
                constexpr int XS = NGRP * HLVAE_MAX_Q;
                double vv[RPT];
                if (nd == 0 && has_se) {
#pragma unroll
                    for (int k = 0; k < RPT; k++) {
                        const double d = xa[k * XS] - zse;
                        vv[k] = exp_nonpos_tab_sel<false>((d * nh) * d, etab, true);
                    }
                } else if (nd == 1 && cat0 && has_se) {
#pragma unroll
                    for (int k = 0; k < RPT; k++) {
                        const double d = xa[k * XS] - zse;
                        vv[k] = exp_nonpos_tab_sel<true>((d * nh) * d, etab, xb[k * XS] == z0);
                    }
                } else if (nd == 1 && cat0) {
#pragma unroll
                    for (int k = 0; k < RPT; k++) vv[k] = (xb[k * XS] == z0) ? 1.0 : 0.0;
                } else {
                    const bool cat1 = c.disc_kind[1] == HLVAE_KIND_CAT, cat2 = c.disc_kind[2] == HLVAE_KIND_CAT;
                    const int dc0 = nd > 0 ? c.disc_col[0] : 0, dc1 = nd > 1 ? c.disc_col[1] : 0,
                              dc2 = nd > 2 ? c.disc_col[2] : 0;
                    const double z1 = nd > 1 ? Zs[dc1 * MP + em] : 0.0, z2 = nd > 2 ? Zs[dc2 * MP + em] : 0.0;
#pragma unroll
                    for (int k = 0; k < RPT; k++) {
                        const double* xr = xsc + (eg + k * NGRP) * HLVAE_MAX_Q;
                        const double a0 = xr[dc0], a1 = xr[dc1], a2 = xr[dc2];
                        bool ok = true;
                        ok = ok && (nd < 1 || (cat0 ? (a0 == z0) : (a0 + z0 == 2.0)));
                        ok = ok && (nd < 2 || (cat1 ? (a1 == z1) : (a1 + z1 == 2.0)));
                        ok = ok && (nd < 3 || (cat2 ? (a2 == z2) : (a2 + z2 == 2.0)));
                        const double d = xa[k * XS] - zse;
                        const double e_ = has_se ? exp_nonpos_tab_sel<false>((d * nh) * d, etab, true) : 1.0;
                        vv[k] = ok ? e_ : 0.0;
                    }
                }
#pragma unroll
                for (int k = 0; k < RPT; k++) kacc[k] = fma(osr, vv[k], kacc[k]);
                if (r < NCT) tmem_store_doubles<RPT>(tm_comp + r * 2 * RPT, vv);
            }
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                const int row = eg + k * NGRP;
                if (row < R8) Kb[row * LD + em] = (row < R && em < M) ? kacc[k] : 0.0;
            }
            tmem_wait_st();
        }
        __syncthreads();

        // ---- P2: V = B^-1 K0xz (block diagonal, :160 / :254) on the FP64 tensor pipe, k-range
        // limited to the subjects a row tile touches; r = K0xz w - mu (:166 / :230)
        {
            const int ar = lane >> 2, ac = lane & 3;
#pragma unroll
            for (int tr = 0; tr < WR; tr++) {
                const int rt = wi * WR + tr;
                if (rt * 8 < R8) {
                    double c[WC][2];
#pragma unroll
                    for (int t = 0; t < WC; t++) c[t][0] = c[t][1] = 0.0;
                    if (rt * 8 < R) {
                        const int rlast = min(rt * 8 + 7, R - 1);
                        const int klo = sub_r0[sub_of_row[rt * 8]] & ~3;
                        const int khi = (sub_r0[sub_of_row[rlast] + 1] + 3) & ~3;
                        for (int k0 = klo; k0 < khi; k0 += 4) {
                            const double a = Bp[(rt * 8 + ar) * LDB + k0 + ac];
#pragma unroll
                            for (int t = 0; t < WC; t++) {
                                const double b = Kb[(k0 + ac) * LD + (wj * WC + t) * 8 + ar];
                                dmma884(c[t][0], c[t][1], a, b);
                            }
                        }
                    }
#pragma unroll
                    for (int t = 0; t < WC; t++)
                        *reinterpret_cast<double2*>(&Vb[(rt * 8 + ar) * LD + (wj * WC + t) * 8 + 2 * ac]) =
                            make_double2(c[t][0], c[t][1]);
                }
            }
        }
        // r = K0xz w - mu (:166 / :230) as one more column tile of the tensor pipe: B fragment = w in column 0
        if (warp * 8 < R) {
            const int ar = lane >> 2, ac = lane & 3;
            double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
            for (int k0 = 0; k0 < MP; k0 += 4) {
                const double a = Kb[(warp * 8 + ar) * LD + k0 + ac];
                const double b = (ar == 0) ? ws[k0 + ac] : 0.0;
                dmma884(c0, c1, a, b);
            }
            const int row = warp * 8 + ar;
            if (ac == 0) rv[row] = row < R ? c0 - mus[row] : 0.0;       // rows R..R8-1: zeros (operand of rho below)
        }
        // the next panel is packed by the warp with the least to do in this phase (last row tiles, no part in r)
        if (warp == PN_THREADS / 32 - 1) pack(par ^ 1);
        __syncthreads();
        const int g_next = next_row(par ^ 1);

        // ---- P3a: rho = B^-1 r (one more column tile of the tensor pipe: B fragment = r in column 0, k-range = the
        // subjects a row tile touches); A += r . rho (:167 / :256); dJ/dmu = -rho
        if (warp * 8 < R) {
            const int ar = lane >> 2, ac = lane & 3;
            const int rlast = min(warp * 8 + 7, R - 1);
            const int klo = sub_r0[sub_of_row[warp * 8]] & ~3;
            const int khi = (sub_r0[sub_of_row[rlast] + 1] + 3) & ~3;
            double c0 = 0.0, c1 = 0.0;
            for (int k0 = klo; k0 < khi; k0 += 4) {
                const double a = Bp[(warp * 8 + ar) * LDB + k0 + ac];
                const double b = (ar == 0) ? rv[k0 + ac] : 0.0;
                dmma884(c0, c1, a, b);
            }
            const int row = warp * 8 + ar;
            if (ac == 0 && row < R) {
                rho[row] = c0;
                a_acc = fma(rv[row], c0, a_acc);
                g_mu[(int64_t)grow[row] * L + l] = (TS)(-c0 * gscale);
            }
        }
        // p = sum_r V[r] mu_r (:188 / :265) and dJ/dw = K0xz^T rho = V^T r (B^-1 is symmetric): per-thread partial
        // sums over this thread's rows, kept over the CTA's whole chunk; they need only V, mu and r, so they share this
        // barrier interval with rho and the S update instead of a serial two-warp phase of their own
        if (em < M) {
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                const int row = eg + k * NGRP;
                if (row < R) {
                    const double v = Vb[row * LD + em];
                    p_acc = fma(v, mus[row], p_acc);
                    gw_acc = fma(v, rv[row], gw_acc);
                }
            }
        }
        // ---- P3b: S += K0xz^T V on the FP64 tensor pipe (:161 / :254,266), lower tiles only
        {
            const int R4 = (R + 3) & ~3;
            const int kr = lane & 3, kc = lane >> 2;
            // slots 0 .. SBASE-1 are live in every warp (straight-line code); only the last slot depends on the warp
            const int t_first = warp * SBASE + min(warp, SREM);
            const bool last_live = SREM != 0 && warp < SREM;
            int ao[SPW], bo[SPW];                                       // column offsets of this warp's tiles in Kb / Vb
#pragma unroll
            for (int t = 0; t < SPW; t++) {
                const int rc = s_tab[(t < SBASE || last_live) ? t_first + t : 0];
                ao[t] = (rc >> 8) * 8 + kc;
                bo[t] = (rc & 255) * 8 + kc;
            }
            if constexpr (PARK) tmem_load_doubles<2 * SPW>(tm_mine, &sacc[0][0]);
            for (int k0 = 0; k0 < R4; k0 += 4) {
                const double* kr_ = Kb + (k0 + kr) * LD;
                const double* vr_ = Vb + (k0 + kr) * LD;
#pragma unroll
                for (int t = 0; t < SBASE; t++) dmma884(sacc[t][0], sacc[t][1], kr_[ao[t]], vr_[bo[t]]);
                if (SREM != 0 && last_live) dmma884(sacc[SPW - 1][0], sacc[SPW - 1][1], kr_[ao[SPW - 1]], vr_[bo[SPW - 1]]);
            }
            if constexpr (PARK) {
                tmem_store_doubles<2 * SPW>(tm_mine, &sacc[0][0]);
                tmem_wait_st();
            }
        }
        __syncthreads();

        issue_loads(par ^ 1, g_next);             // Bp, mraw and the other xs / descriptor set are free from here on

        // ---- P4: W = V G  (dJ/dS = G / 2 applied on both sides -> dJ/dK0xz = W + rho w^T), into Kb.
        // One G fragment per k-step serves all of the warp's row tiles.
        {
            const int ar = lane >> 2, ac = lane & 3;
            double c[WR][WC][2];
            bool live[WR];
#pragma unroll
            for (int tr = 0; tr < WR; tr++) {
                live[tr] = (wi * WR + tr) * 8 < R;
#pragma unroll
                for (int t = 0; t < WC; t++) c[tr][t][0] = c[tr][t][1] = 0.0;
            }
            if (live[0]) {
#pragma unroll 4
                for (int k0 = 0; k0 < MP; k0 += 4) {
                    double bfr[WC];
#pragma unroll
                    for (int t = 0; t < WC; t++) {
                        const int col = (wj * WC + t) * 8 + ar;       // B frag: row = k0 + lane%4, col = lane/4
                        if (G_SMEM) {
                            bfr[t] = Gs[(k0 + ac) * LD + col];
                        } else if (M == MP) {                          // full tile: no bounds tests (uniform branch)
                            bfr[t] = __ldg(Gl + (k0 + ac) * MP + col);
                        } else {
                            bfr[t] = (k0 + ac < M && col < M) ? __ldg(Gl + (int64_t)(k0 + ac) * M + col) : 0.0;
                        }
                    }
#pragma unroll
                    for (int tr = 0; tr < WR; tr++) {
                        if (live[tr]) {
                            const double a = Vb[((wi * WR + tr) * 8 + ar) * LD + k0 + ac];
#pragma unroll
                            for (int t = 0; t < WC; t++) dmma884(c[tr][t][0], c[tr][t][1], a, bfr[t]);
                        }
                    }
                }
            }
#pragma unroll
            for (int tr = 0; tr < WR; tr++) {
                if (live[tr]) {
#pragma unroll
                    for (int t = 0; t < WC; t++)
                        *reinterpret_cast<double2*>(&Kb[((wi * WR + tr) * 8 + ar) * LD + (wj * WC + t) * 8 + 2 * ac]) =
                            make_double2(c[tr][t][0], c[tr][t][1]);
                }
            }
        }
        __syncthreads();

        // optional output: diag(W V^T)_r = V_r G V_r^T per row (validation.py:69-72 with G = W^-1)
        if (qdiag) {
            for (int r = warp; r < R; r += PN_THREADS / 32) {
                double a = 0.0;
                for (int m = lane; m < MP; m += 32) a = fma(Kb[r * LD + m], Vb[r * LD + m], a);
                a = warp_sum(a);
                if (lane == 0) qdiag[(int64_t)grow[r] * L + l] = (TS)a;
            }
        }
        // ---- P5a: dJ/dB_s (K0xz part) = -1/2 (rho rho^T + W V^T), symmetric: only the 8x8 tiles on or below the
        // diagonal that meet the block diagonal, on the FP64 tensor pipe, and contracted with dB_s/d(theta1) straight
        // from the accumulator fragments (off-diagonal entries count twice) - the tiles are never stored
        {
            constexpr int NWARP = PN_THREADS / 32;
            constexpr int NTL = (SM::n_lower_tiles + NWARP - 1) / NWARP;       // tiles per warp
            const int nrt = R8 / 8;
            const int nlt = nrt * (nrt + 1) / 2;
            const int ar = lane >> 2, ac = lane & 3;
            double gq[NTL][2];
            int gi[NTL], gj[NTL];
            bool live[NTL];
#pragma unroll
            for (int t = 0; t < NTL; t++) {
                gq[t][0] = gq[t][1] = 0.0;
                gi[t] = gj[t] = 0;
                live[t] = false;
                const int tp = warp + t * NWARP;
                if (tp < nlt) {
                    const int rc = tl_tab[tp], rt = rc >> 8, ct = rc & 255;
                    if (sub_of_row[min(ct * 8 + 7, R - 1)] >= sub_of_row[rt * 8]) {
                        double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
                        for (int m0 = 0; m0 < MP; m0 += 4) {
                            const double a = Kb[(rt * 8 + ar) * LD + m0 + ac];
                            const double b = Vb[(ct * 8 + ar) * LD + m0 + ac];
                            dmma884(c0, c1, a, b);
                        }
                        const int i = rt * 8 + ar, j = ct * 8 + 2 * ac;
                        if (i < R) {
                            const int si = sub_of_row[i];
                            const double ri = rho[i];
                            if (j <= i && sub_of_row[j] == si) gq[t][0] = (i == j ? -0.5 : -1.0) * (c0 + ri * rho[j]);
                            if (j + 1 <= i && sub_of_row[j + 1] == si)
                                gq[t][1] = (i == j + 1 ? -0.5 : -1.0) * (c1 + ri * rho[j + 1]);
                        }
                        gi[t] = min(i, RP - 1);
                        gj[t] = min(j, RP - 2);
                        live[t] = true;
                    }
                }
            }
            for (int r = 0; r < sp1.ncomp; r++) {
                CompRegs c;
                c.load(sp1, r);
                const double osr = kps1[r], hil2 = kps1[HLVAE_MAX_COMPS + r], il3 = kps1[3 * HLVAE_MAX_COMPS + r];
                const double nh = -hil2;
                const bool has_se = c.se_col >= 0, cat0 = c.disc_kind[0] == HLVAE_KIND_CAT;
                const double* xa = xsc + (has_se ? c.se_col : 0);
                const double* xb = xsc + (c.ndisc > 0 ? c.disc_col[0] : 0);
                double gos = 0.0, gls = 0.0;
#pragma unroll
                for (int t = 0; t < NTL; t++) {
                    if (live[t]) {
                        if (c.ndisc == 1 && cat0) {                 // categorical (x SE): the forms kernel_gen.py builds
                            const double xia = xa[gi[t] * HLVAE_MAX_Q], xib = xb[gi[t] * HLVAE_MAX_Q];
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                const double xja = xa[(gj[t] + u) * HLVAE_MAX_Q], xjb = xb[(gj[t] + u) * HLVAE_MAX_Q];
                                const double d = xia - xja;
                                const double v = has_se ? exp_nonpos_tab_sel<true>((d * nh) * d, etab, xib == xjb)
                                                        : ((xib == xjb) ? 1.0 : 0.0);
                                const double gv = gq[t][u] * v;
                                gos += gv;
                                gls = fma(gv * d, d, gls);
                            }
                        } else {
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                double d;
                                const double gv = gq[t][u] * c.value(xsc + gi[t] * HLVAE_MAX_Q, xsc + (gj[t] + u) * HLVAE_MAX_Q, hil2, d, etab);
                                gos += gv;
                                gls = fma(gv * d, d, gls);
                            }
                        }
                    }
                }
                gos = warp_sum(gos);
                gls = warp_sum(gls) * osr * il3;
                if (lane == 0) {
                    atomicAdd(&hyp[2 * HLVAE_MAX_COMPS + r], gos);
                    atomicAdd(&hyp[3 * HLVAE_MAX_COMPS + r], gls);
                }
            }
        }
        // ---- P5b: dJ/dK0xz = W + rho w^T, contracted with dK0xz/d{outputscale, lengthscale, Z}
        {
            if constexpr (PARK) tmem_load_doubles<HGN>(tm_mine + PARK_S, hgs);
            double gk[RPT];
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                const int row = eg + k * NGRP;
                gk[k] = (row < R && em < M) ? Kb[row * LD + em] + rho[row] * wm : 0.0;
            }
            for (int r = 0; r < sp0.ncomp; r++) {
                CompRegs c;
                c.load(sp0, r);
                const double zse = (c.se_col >= 0) ? Zs[c.se_col * MP + em] : 0.0;
                const double osr = kps[r], hil2 = kps[HLVAE_MAX_COMPS + r], il2 = kps[2 * HLVAE_MAX_COMPS + r],
                             il3 = kps[3 * HLVAE_MAX_COMPS + r];
                // sums of g v, g v d, g v d^2 over this thread's rows (d = x - z); outputscale and lengthscale
                // factors are applied once at the end
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                if (r < NCT) {
                    // values of this component back from the thread's TMEM columns; rows >= R and inducing points
                    // >= M carry gk = 0
                    double vv[RPT];
                    tmem_load_doubles<RPT>(tm_comp + r * 2 * RPT, vv);
                    if (c.se_col >= 0) {
                        const double* xa = xsc + eg * HLVAE_MAX_Q + c.se_col;
#pragma unroll
                        for (int k = 0; k < RPT; k++) {
                            const double gkv = gk[k] * vv[k];
                            const double d = xa[k * NGRP * HLVAE_MAX_Q] - zse;
                            const double t = gkv * d;
                            s0 += gkv;
                            s1 += t;
                            s2 = fma(t, d, s2);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < RPT; k++) s0 = fma(gk[k], vv[k], s0);
                    }
                } else {
                    if (em < M) {
                        double zd[HLVAE_MAX_DISC];
#pragma unroll
                        for (int f = 0; f < HLVAE_MAX_DISC; f++) zd[f] = (f < c.ndisc) ? Zs[c.disc_col[f] * MP + em] : 0.0;
#pragma unroll
                        for (int k = 0; k < RPT; k++) {
                            const int row = eg + k * NGRP;
                            if (row < R) {
                                const double* xr = xsc + row * HLVAE_MAX_Q;
                                bool ok = true;
#pragma unroll
                                for (int f = 0; f < HLVAE_MAX_DISC; f++)
                                    if (f < c.ndisc) {
                                        const double a = xr[c.disc_col[f]];
                                        ok = ok && ((c.disc_kind[f] == HLVAE_KIND_CAT) ? (a == zd[f]) : (a + zd[f] == 2.0));
                                    }
                                if (ok) {
                                    double v = 1.0, d = 0.0;
                                    if (c.se_col >= 0) {
                                        d = xr[c.se_col] - zse;
                                        v = exp_nonpos_tab(-(d * d) * hil2, etab);
                                    }
                                    const double gkv = gk[k] * v;
                                    const double t = gkv * d;
                                    s0 += gkv;
                                    s1 += t;
                                    s2 = fma(t, d, s2);
                                }
                            }
                        }
                    }
                }
                if (r < PN_NCACHE) {
#pragma unroll
                    for (int q = 0; q < PN_NCACHE; q++)
                        if (q == r) { hgs[3 * q] += s0; hgs[3 * q + 1] += s1; hgs[3 * q + 2] += s2; }
                } else {
                    const double gos = warp_sum(s0);
                    const double gls = warp_sum(s2) * osr * il3;
                    if (lane == 0) {
                        atomicAdd(&hyp[r], gos);
                        atomicAdd(&hyp[HLVAE_MAX_COMPS + r], gls);
                    }
                    if (c.se_col >= 0 && s1 != 0.0) atomicAdd(&zacc[em * HLVAE_MAX_COMPS + r], s1 * osr * il2);
                }
            }
            if constexpr (PARK) {
                tmem_store_doubles<HGN>(tm_mine + PARK_S, hgs);
                tmem_wait_st();
            }
        }
        par ^= 1;
    }
    if constexpr (PARK) {
        tmem_load_doubles<2 * SPW>(tm_mine, &sacc[0][0]);
        tmem_load_doubles<HGN>(tm_mine + PARK_S, hgs);
    }

    // ---- cached components: reduce the per-thread gradient sums into hyp / zacc
#pragma unroll
    for (int r = 0; r < PN_NCACHE; r++) {
        if (r < sp0.ncomp) {
            const double osr = kps[r], il2 = kps[2 * HLVAE_MAX_COMPS + r], il3 = kps[3 * HLVAE_MAX_COMPS + r];
            const double gos = warp_sum(hgs[3 * r]);
            const double gls = warp_sum(hgs[3 * r + 2]) * osr * il3;
            if (lane == 0) {
                atomicAdd(&hyp[r], gos);
                atomicAdd(&hyp[HLVAE_MAX_COMPS + r], gls);
            }
            if (sp0.comp[r].se_col >= 0 && hgs[3 * r + 1] != 0.0)
                atomicAdd(&zacc[em * HLVAE_MAX_COMPS + r], hgs[3 * r + 1] * osr * il2);
        }
    }
    __syncthreads();
    // ---- flush CTA accumulators
    {
        double* Sg = acc + off.o[HLVAE_ACC_S] + (int64_t)l * M * M;
        const int cr = lane >> 2, cc = 2 * (lane & 3);
#pragma unroll
        for (int t = 0; t < SPW; t++) {
            const int tp = warp * SBASE + min(warp, SREM) + t;
            if (t < SBASE || (SREM != 0 && warp < SREM)) {
                const int rc = s_tab[tp], ta = rc >> 8, tb = rc & 255;
                const int i = ta * 8 + cr, j = tb * 8 + cc;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (i < M && j + u < M) {
                        atomicAdd(Sg + (int64_t)i * M + j + u, sacc[t][u]);
                        if (ta != tb) atomicAdd(Sg + (int64_t)(j + u) * M + i, sacc[t][u]);     // the mirrored tile
                    }
                }
            }
        }
        // every row group holds a partial sum of its inducing point: summed in a fixed order through shared memory
        // (the row panels are free by now), then one atomic per inducing point and CTA
        Kb[eg * MP + em] = p_acc;
        Vb[eg * MP + em] = gw_acc;
        __syncthreads();
        if (tid < M) {
            double pa = 0.0, ga = 0.0;
#pragma unroll
            for (int g2 = 0; g2 < NGRP; g2++) {
                pa += Kb[g2 * MP + tid];
                ga += Vb[g2 * MP + tid];
            }
            atomicAdd(acc + off.o[HLVAE_ACC_P] + (int64_t)l * M + tid, pa);
            atomicAdd(acc + off.o[HLVAE_ACC_GW] + (int64_t)l * M + tid, ga);
        }
        a_acc = warp_sum(a_acc);
        if (lane == 0 && a_acc != 0.0) atomicAdd(&hyp[4 * HLVAE_MAX_COMPS], a_acc);
        __syncthreads();
        if (tid == 0) atomicAdd(acc + off.o[HLVAE_ACC_SCAL] + (int64_t)l * HLVAE_NSCAL + 0, hyp[4 * HLVAE_MAX_COMPS]);
        if (tid < HLVAE_MAX_COMPS) {
            const int r = tid;
            if (r < sp0.ncomp) {
                atomicAdd(acc + off.o[HLVAE_ACC_GOS0] + (int64_t)r * L + l, hyp[r]);
                if (sp0.comp[r].se_col >= 0)
                    atomicAdd(acc + off.o[HLVAE_ACC_GLS0] + (int64_t)r * L + l, hyp[HLVAE_MAX_COMPS + r]);
            }
            if (r < sp1.ncomp) {
                atomicAdd(acc + off.o[HLVAE_ACC_GOS1] + (int64_t)r * L + l, hyp[2 * HLVAE_MAX_COMPS + r]);
                if (sp1.comp[r].se_col >= 0)
                    atomicAdd(acc + off.o[HLVAE_ACC_GLS1] + (int64_t)r * L + l, hyp[3 * HLVAE_MAX_COMPS + r]);
            }
        }
        double* gz = acc + off.o[HLVAE_ACC_GZ] + (int64_t)l * M * Q;
        for (int e = tid; e < M * HLVAE_MAX_COMPS; e += PN_THREADS) {
            int m = e / HLVAE_MAX_COMPS, r = e % HLVAE_MAX_COMPS;
            if (r < sp0.ncomp && sp0.comp[r].se_col >= 0) {
                double v = zacc[e];
                if (v != 0.0) atomicAdd(gz + (int64_t)m * Q + sp0.comp[r].se_col, v);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc<TM_COLS>(tmem_base);
}

template <int MP, int RP, bool G_SMEM, int NT, int NC, typename TS>
int launch_panel(const hlvae_kspec_t* spec0, const double* os0, const double* ls0, const hlvae_kspec_t* spec1,
                 const double* os1, const double* ls1, int L, int Q, int M, const double* x, int64_t ldx,
                 const double* z, const int32_t* row_idx, const int32_t* subj_ptr, const int32_t* tt_ptr, int n_subj,
                 int subj_per_chunk, const void* mu, int64_t ld_mu, const double* w, const double* G,
                 const double* binv, int64_t tt_total, double* acc, const AccOff& off, void* g_mu, void* qdiag,
                 double gscale, int32_t* status, cudaStream_t st) {
    using SM = PanelSmem<MP, RP, G_SMEM, NC>;
    auto kern = kl_panel_k<MP, RP, G_SMEM, NT, NC, TS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::bytes);
    if (e != cudaSuccess) return (int)e;
    int n_chunks = (n_subj + subj_per_chunk - 1) / subj_per_chunk;
    dim3 grid(n_chunks, L);
    kern<<<grid, NT, SM::bytes, st>>>(*spec0, os0, ls0, *spec1, os1, ls1, L, Q, M, x, ldx, z, row_idx,
                                              subj_ptr, tt_ptr, n_subj, subj_per_chunk, (const TS*)mu, ld_mu, w, G,
                                              binv, tt_total, acc, off, (TS*)g_mu, (TS*)qdiag, gscale, status);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

template <typename TS>
int dispatch_panel(int M, const hlvae_kspec_t* spec0, const double* os0, const double* ls0,
                   const hlvae_kspec_t* spec1, const double* os1, const double* ls1, int L, int Q, const double* x,
                   int64_t ldx, const double* z, const int32_t* row_idx, const int32_t* subj_ptr,
                   const int32_t* tt_ptr, int n_subj, int subj_per_chunk, const void* mu, int64_t ld_mu,
                   const double* w, const double* G, const double* binv, int64_t tt_total, double* acc,
                   const AccOff& off, void* g_mu, void* qdiag, double gscale, int32_t* status, int row_panel,
                   cudaStream_t st) {
#define HLVAE_PANEL(MP, RP, GS, NT, NC)                                                                               \
    return launch_panel<MP, RP, GS, NT, NC, TS>(spec0, os0, ls0, spec1, os1, ls1, L, Q, M, x, ldx, z, row_idx, subj_ptr,     \
                                        tt_ptr, n_subj, subj_per_chunk, mu, ld_mu, w, G, binv, tt_total, acc, off,   \
                                        g_mu, qdiag, gscale, status, st)
    // Shapes (row_panel = rows per panel, 0 = the default of the M class; measured at 16 000 rows, T = 20):
    //   M <= 32 : two 256-thread CTAs per SM, 64-row panels (0.52 ms against 0.67 ms for one 512-thread CTA);
    //   M <= 64 : two 256-thread CTAs per SM walking 40-row panels (they cover each other's barrier and latency
    //             stalls; 0.94 ms) when whole subjects fill 40 rows about as well as 64 (T = 20: 2 of 2 against 3 of
    //             3.2), else one 512-thread CTA with 64-row panels (1.05 ms);
    //   M <= 128: one 512-thread CTA, 64- / 48- / 32-row panels (M = 120: 2.56 / 3.10 / 4.16 ms - the shared memory
    //             the component values left when they moved to tensor memory pays for the larger panels).
    // Measured and dropped: 256 threads with 32-row panels at M = 64 (+9 %), 48-row panels at M = 64 (+10 %), 1024
    // threads at 64 registers (+18 %: spills, per-thread set-up doubled), G in shared memory at M = 64 (one CTA per SM).
    if (M <= 32) { HLVAE_PANEL(32, 64, true, 256, 3); }
    if (M <= 64) {
        if (row_panel == 40) { HLVAE_PANEL(64, 40, false, 256, 3); }
        HLVAE_PANEL(64, 64, false, 512, 3);
    }
    if (M <= 128) {
        if (row_panel == 32) { HLVAE_PANEL(128, 32, false, 512, 3); }
        if (row_panel == 48) { HLVAE_PANEL(128, 48, false, 512, 3); }
        HLVAE_PANEL(128, 64, false, 512, 3);
    }
#undef HLVAE_PANEL
    return HLVAE_E_UNSUPPORTED;
}

void fill_offsets(int L, int M, int Q, int64_t* o) {
    int64_t p = 0;
    o[HLVAE_ACC_S] = p;    p += (int64_t)L * M * M;
    o[HLVAE_ACC_P] = p;    p += (int64_t)L * M;
    o[HLVAE_ACC_GW] = p;   p += (int64_t)L * M;
    o[HLVAE_ACC_SCAL] = p; p += (int64_t)L * HLVAE_NSCAL;
    o[HLVAE_ACC_GZ] = p;   p += (int64_t)L * M * Q;
    o[HLVAE_ACC_GOS0] = p; p += (int64_t)HLVAE_MAX_COMPS * L;
    o[HLVAE_ACC_GLS0] = p; p += (int64_t)HLVAE_MAX_COMPS * L;
    o[HLVAE_ACC_GOS1] = p; p += (int64_t)HLVAE_MAX_COMPS * L;
    o[HLVAE_ACC_GLS1] = p; p += (int64_t)HLVAE_MAX_COMPS * L;
    o[HLVAE_ACC_TOTAL] = p;
}

}  // namespace

extern "C" int hlvae_kl_acc_layout(int L, int M, int Q, int64_t* offsets) {
    if (L <= 0 || M <= 0 || Q <= 0 || !offsets) return HLVAE_E_ARG;
    fill_offsets(L, M, Q, offsets);
    return 0;
}

extern "C" int hlvae_kl_subject(const hlvae_kspec_t* spec0, const double* os0, const double* ls0,
                                const hlvae_kspec_t* spec1, const double* os1, const double* ls1, const double* noise,
                                int L, int Q, const double* x, int64_t ldx, const int32_t* row_idx,
                                const int32_t* subj_ptr, const int32_t* tt_ptr, int n_subj, int t_cap,
                                const void* log_v, int64_t ld_lv, int dtype, double* binv, int64_t tt_total,
                                double* acc, int M, void* g_logv, double gscale, int32_t* status, void* stream) {
    if (!hlvae::spec_valid(spec0, Q) || !hlvae::spec_valid(spec1, Q) || L <= 0 || Q <= 0 || Q > HLVAE_MAX_Q ||
        M <= 0 || n_subj < 0 || t_cap <= 0 || t_cap > HLVAE_TMAX || !x || !row_idx || !subj_ptr || !tt_ptr || !log_v ||
        !binv || !acc || !g_logv || !noise)
        return HLVAE_E_ARG;
    if (dtype != HLVAE_F32 && dtype != HLVAE_F64) return HLVAE_E_ARG;
    if (n_subj == 0) return 0;
    AccOff off;
    fill_offsets(L, M, Q, off.o);
    cudaStream_t st = (cudaStream_t)stream;
#define HLVAE_SUBJ2(TP)                                                                                              \
    return dtype == HLVAE_F64                                                                                        \
               ? launch_subject<TP, double>(spec0, os0, ls0, spec1, os1, ls1, noise, L, Q, x, ldx, row_idx, subj_ptr, \
                                             tt_ptr, n_subj, log_v, ld_lv, binv, tt_total, acc, off, g_logv, gscale,  \
                                             status, st)                                                              \
               : launch_subject<TP, float>(spec0, os0, ls0, spec1, os1, ls1, noise, L, Q, x, ldx, row_idx, subj_ptr,  \
                                            tt_ptr, n_subj, log_v, ld_lv, binv, tt_total, acc, off, g_logv, gscale,   \
                                            status, st)
    if (t_cap > SJ_TW) {     // subjects of 33 .. HLVAE_TMAX rows: CTA per pair (each kernel skips the other's subjects)
        const int rc = dtype == HLVAE_F64
                           ? launch_subject_big<double>(spec0, os0, ls0, spec1, os1, ls1, noise, L, Q, x, ldx, row_idx,
                                                        subj_ptr, tt_ptr, n_subj, log_v, ld_lv, binv, tt_total, acc, off,
                                                        g_logv, gscale, status, st)
                           : launch_subject_big<float>(spec0, os0, ls0, spec1, os1, ls1, noise, L, Q, x, ldx, row_idx,
                                                       subj_ptr, tt_ptr, n_subj, log_v, ld_lv, binv, tt_total, acc, off,
                                                       g_logv, gscale, status, st);
        if (rc != 0) return rc;
    }
    if (t_cap <= 8) { HLVAE_SUBJ2(8); }
    if (t_cap <= 16) { HLVAE_SUBJ2(16); }
    if (t_cap <= 24) { HLVAE_SUBJ2(24); }
    HLVAE_SUBJ2(32);
#undef HLVAE_SUBJ2
}

extern "C" int hlvae_kl_panel(const hlvae_kspec_t* spec0, const double* os0, const double* ls0,
                              const hlvae_kspec_t* spec1, const double* os1, const double* ls1, int L, int Q, int M,
                              const double* x, int64_t ldx, const double* z, const int32_t* row_idx,
                              const int32_t* subj_ptr, const int32_t* tt_ptr, int n_subj, int subj_per_chunk,
                              const void* mu, int64_t ld_mu, int dtype, const double* w, const double* G,
                              const double* binv, int64_t tt_total, double* acc, void* g_mu, void* qdiag,
                              double gscale, int32_t* status, int row_panel, void* stream) {
    if (!hlvae::spec_valid(spec0, Q) || !hlvae::spec_valid(spec1, Q) || L <= 0 || Q <= 0 || Q > HLVAE_MAX_Q ||
        M <= 0 || n_subj < 0 || subj_per_chunk <= 0 || !x || !z || !row_idx || !subj_ptr || !tt_ptr || !mu || !w ||
        !G || !binv || !acc || !g_mu)
        return HLVAE_E_ARG;
    if (M > 128) return HLVAE_E_UNSUPPORTED;
    if (n_subj == 0) return 0;
    AccOff off;
    fill_offsets(L, M, Q, off.o);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HLVAE_F64)
        return dispatch_panel<double>(M, spec0, os0, ls0, spec1, os1, ls1, L, Q, x, ldx, z, row_idx, subj_ptr, tt_ptr,
                                      n_subj, subj_per_chunk, mu, ld_mu, w, G, binv, tt_total, acc, off, g_mu, qdiag,
                                      gscale, status, row_panel, st);
    if (dtype == HLVAE_F32)
        return dispatch_panel<float>(M, spec0, os0, ls0, spec1, os1, ls1, L, Q, x, ldx, z, row_idx, subj_ptr, tt_ptr,
                                     n_subj, subj_per_chunk, mu, ld_mu, w, G, binv, tt_total, acc, off, g_mu, qdiag,
                                     gscale, status, row_panel, st);
    return HLVAE_E_ARG;
}
