// Dense additive-kernel evaluation K[l, i, j] and its backward.
// Stands behind `covar_module(x1, x2).evaluate()` (elbo_functions.py:147-148, 222-223).
#include "common.cuh"

using namespace hlvae;

namespace {

constexpr int EV_THREADS = 256;

// grid: (ceil(n1*n2 / EV_THREADS), L)
__global__ void __launch_bounds__(EV_THREADS)
kernel_eval_fwd_k(const __grid_constant__ hlvae_kspec_t sp, const double* __restrict__ os,
                  const double* __restrict__ ls, int L, int Q, const double* __restrict__ x1, int n1, int64_t ld1,
                  int64_t bs1, const double* __restrict__ x2, int n2, int64_t ld2, int64_t bs2,
                  double* __restrict__ out) {
    const int l = blockIdx.y;
    KParams kp;
    load_kparams(kp, sp, os, ls, L, l);
    int64_t e = (int64_t)blockIdx.x * EV_THREADS + threadIdx.x;
    if (e >= (int64_t)n1 * n2) return;
    int i = (int)(e / n2), j = (int)(e % n2);
    const double* xa = x1 + l * bs1 + (int64_t)i * ld1;
    const double* xb = x2 + l * bs2 + (int64_t)j * ld2;
    out[((int64_t)l * n1 + i) * n2 + j] = eval_additive(sp, kp, xa, xb);
}

// grid: (ceil(n1*n2 / EV_THREADS), L).  Hyper-parameter gradients: block reduce + one atomic
// per block; covariate gradients: one atomic per (element, SE component) - this op serves the
// small M x M matrix K0zz and API-compatibility calls, not the streaming path.
__global__ void __launch_bounds__(EV_THREADS)
kernel_eval_bwd_k(const __grid_constant__ hlvae_kspec_t sp, const double* __restrict__ os,
                  const double* __restrict__ ls, int L, int Q, const double* __restrict__ x1, int n1, int64_t ld1,
                  int64_t bs1, const double* __restrict__ x2, int n2, int64_t ld2, int64_t bs2,
                  const double* __restrict__ g_out, double* __restrict__ g_os, double* __restrict__ g_ls,
                  double* __restrict__ g_x1, double* __restrict__ g_x2) {
    __shared__ double red[2 * HLVAE_MAX_COMPS][EV_THREADS / 32];
    const int l = blockIdx.y;
    KParams kp;
    load_kparams(kp, sp, os, ls, L, l);
    double gos[HLVAE_MAX_COMPS], gls[HLVAE_MAX_COMPS], gxb[HLVAE_MAX_COMPS];
#pragma unroll
    for (int r = 0; r < HLVAE_MAX_COMPS; r++) { gos[r] = 0; gls[r] = 0; gxb[r] = 0; }
    int64_t e = (int64_t)blockIdx.x * EV_THREADS + threadIdx.x;
    if (e < (int64_t)n1 * n2) {
        int i = (int)(e / n2), j = (int)(e % n2);
        const double* xa = x1 + l * bs1 + (int64_t)i * ld1;
        const double* xb = x2 + l * bs2 + (int64_t)j * ld2;
        double g = g_out[((int64_t)l * n1 + i) * n2 + j];
        accum_grads<true>(sp, kp, xa, xb, g, gos, gls, gxb);
#pragma unroll
        for (int r = 0; r < HLVAE_MAX_COMPS; r++) {
            if (r < sp.ncomp && sp.comp[r].se_col >= 0 && gxb[r] != 0.0) {
                int c = sp.comp[r].se_col;
                if (g_x2) atomicAdd(g_x2 + ((int64_t)l * n2 + j) * Q + c, gxb[r]);
                if (g_x1) atomicAdd(g_x1 + ((int64_t)l * n1 + i) * Q + c, -gxb[r]);
            }
        }
    }
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
    for (int r = 0; r < HLVAE_MAX_COMPS; r++) {
        if (r < sp.ncomp) {
            double a = warp_sum(gos[r]), b = warp_sum(gls[r]);
            if (lane == 0) { red[r][warp] = a; red[HLVAE_MAX_COMPS + r][warp] = b; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * HLVAE_MAX_COMPS) {
        int r = threadIdx.x % HLVAE_MAX_COMPS;
        if (r < sp.ncomp) {
            double s = 0;
            for (int w = 0; w < EV_THREADS / 32; w++) s += red[threadIdx.x][w];
            double* dst = (threadIdx.x < HLVAE_MAX_COMPS) ? g_os : g_ls;
            if (s != 0.0) atomicAdd(dst + (int64_t)r * L + l, s);
        }
    }
}

bool spec_ok(const hlvae_kspec_t* sp, int Q) {
    if (!sp || sp->ncomp < 0 || sp->ncomp > HLVAE_MAX_COMPS) return false;
    for (int r = 0; r < sp->ncomp; r++) {
        const hlvae_comp_t& c = sp->comp[r];
        if (c.se_col >= Q || c.ndisc < 0 || c.ndisc > HLVAE_MAX_DISC) return false;
        for (int f = 0; f < c.ndisc; f++)
            if (c.disc_col[f] < 0 || c.disc_col[f] >= Q ||
                (c.disc_kind[f] != HLVAE_KIND_CAT && c.disc_kind[f] != HLVAE_KIND_BIN))
                return false;
    }
    return true;
}


// out[i, l] = sum_{j in subject sid[i]} K_l(xt[i], x[row_idx[j]]) * v[row_idx[j], l]
// (utils.py:175-186: K1(X*, x) mu_tilde; every K1 term carries the id kernel, so only the rows of the test
// row's own subject contribute).  grid: (ceil(Nt / 256), L); sid[i] < 0: subject absent -> 0.
__global__ void __launch_bounds__(EV_THREADS)
subject_matvec_k(const __grid_constant__ hlvae_kspec_t sp, const double* __restrict__ os,
                 const double* __restrict__ ls, int L, int Q, const double* __restrict__ xt, int nt,
                 const double* __restrict__ x, const int32_t* __restrict__ row_idx,
                 const int32_t* __restrict__ subj_ptr, const int32_t* __restrict__ sid,
                 const double* __restrict__ v, double* __restrict__ out) {
    const int l = blockIdx.y;
    KParams kp;
    load_kparams(kp, sp, os, ls, L, l);
    const int i = blockIdx.x * EV_THREADS + threadIdx.x;
    if (i >= nt) return;
    const int s = sid[i];
    double acc = 0.0;
    if (s >= 0) {
        const double* xa = xt + (int64_t)i * Q;
        for (int j = subj_ptr[s]; j < subj_ptr[s + 1]; j++) {
            const int g = row_idx[j];
            acc = fma(eval_additive(sp, kp, xa, x + (int64_t)g * Q), v[(int64_t)g * L + l], acc);
        }
    }
    out[(int64_t)i * L + l] = acc;
}

// out[i, l] = sum_j K_l(x1[i], x2[l, j]) v[l, j]: a dense block of the additive kernel applied to one vector per
// latent dimension without materialising it (utils.py:169 / :249: K0Xz (iK K0zx mu_tilde)).  One warp per (row, l).
__global__ void __launch_bounds__(EV_THREADS)
kernel_matvec_k(const __grid_constant__ hlvae_kspec_t sp, const double* __restrict__ os,
                const double* __restrict__ ls, int L, int Q, const double* __restrict__ x1, int n1,
                const double* __restrict__ x2, int n2, const double* __restrict__ v, double* __restrict__ out) {
    const int l = blockIdx.y;
    KParams kp;
    load_kparams(kp, sp, os, ls, L, l);
    const int i = blockIdx.x * (EV_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n1) return;
    const double* xa = x1 + (int64_t)i * Q;
    double acc = 0.0;
    for (int j = lane; j < n2; j += 32)
        acc = fma(eval_additive(sp, kp, xa, x2 + ((int64_t)l * n2 + j) * Q), v[(int64_t)l * n2 + j], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[(int64_t)i * L + l] = acc;
}

}  // namespace

namespace hlvae {
bool spec_valid(const hlvae_kspec_t* sp, int Q) { return spec_ok(sp, Q); }
}

extern "C" int hlvae_version(void) { return HLVAE_ABI_VERSION; }
extern "C" int hlvae_sizeof_kspec(void) { return (int)sizeof(hlvae_kspec_t); }

extern "C" int hlvae_kernel_eval_fwd(const hlvae_kspec_t* spec, const double* outputscale, const double* lengthscale,
                                     int L, int Q, const double* x1, int n1, int64_t ld1, int64_t bs1,
                                     const double* x2, int n2, int64_t ld2, int64_t bs2, double* out, void* stream) {
    if (!spec_ok(spec, Q) || L <= 0 || Q <= 0 || Q > HLVAE_MAX_Q || !x1 || !x2 || !out || n1 < 0 || n2 < 0)
        return HLVAE_E_ARG;
    if (n1 == 0 || n2 == 0) return 0;
    int64_t ne = (int64_t)n1 * n2;
    dim3 grid((unsigned)((ne + EV_THREADS - 1) / EV_THREADS), L);
    kernel_eval_fwd_k<<<grid, EV_THREADS, 0, (cudaStream_t)stream>>>(*spec, outputscale, lengthscale, L, Q, x1, n1,
                                                                     ld1, bs1, x2, n2, ld2, bs2, out);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_kernel_eval_bwd(const hlvae_kspec_t* spec, const double* outputscale, const double* lengthscale,
                                     int L, int Q, const double* x1, int n1, int64_t ld1, int64_t bs1,
                                     const double* x2, int n2, int64_t ld2, int64_t bs2, const double* g_out,
                                     double* g_os, double* g_ls, double* g_x1, double* g_x2, void* stream) {
    if (!spec_ok(spec, Q) || L <= 0 || Q <= 0 || Q > HLVAE_MAX_Q || !x1 || !x2 || !g_out || !g_os || !g_ls ||
        n1 < 0 || n2 < 0)
        return HLVAE_E_ARG;
    if (n1 == 0 || n2 == 0) return 0;
    int64_t ne = (int64_t)n1 * n2;
    dim3 grid((unsigned)((ne + EV_THREADS - 1) / EV_THREADS), L);
    kernel_eval_bwd_k<<<grid, EV_THREADS, 0, (cudaStream_t)stream>>>(*spec, outputscale, lengthscale, L, Q, x1, n1,
                                                                     ld1, bs1, x2, n2, ld2, bs2, g_out, g_os, g_ls,
                                                                     g_x1, g_x2);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_subject_matvec(const hlvae_kspec_t* spec, const double* outputscale, const double* lengthscale,
                                    int L, int Q, const double* xt, int nt, const double* x, const int32_t* row_idx,
                                    const int32_t* subj_ptr, const int32_t* sid, const double* v, double* out,
                                    void* stream) {
    if (!hlvae::spec_valid(spec, Q) || L <= 0 || Q <= 0 || Q > HLVAE_MAX_Q || nt < 0 || !xt || !x || !row_idx ||
        !subj_ptr || !sid || !v || !out)
        return HLVAE_E_ARG;
    if (nt == 0) return 0;
    dim3 grid((unsigned)((nt + EV_THREADS - 1) / EV_THREADS), (unsigned)L);
    subject_matvec_k<<<grid, EV_THREADS, 0, (cudaStream_t)stream>>>(*spec, outputscale, lengthscale, L, Q, xt, nt, x,
                                                                    row_idx, subj_ptr, sid, v, out);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_kernel_matvec(const hlvae_kspec_t* spec, const double* outputscale, const double* lengthscale,
                                   int L, int Q, const double* x1, int n1, const double* x2, int n2, const double* v,
                                   double* out, void* stream) {
    if (!spec_ok(spec, Q) || L <= 0 || Q <= 0 || Q > HLVAE_MAX_Q || n1 < 0 || n2 < 0 || !x1 || !x2 || !v || !out)
        return HLVAE_E_ARG;
    if (n1 == 0) return 0;
    dim3 grid((unsigned)((n1 + EV_THREADS / 32 - 1) / (EV_THREADS / 32)), (unsigned)L);
    kernel_matvec_k<<<grid, EV_THREADS, 0, (cudaStream_t)stream>>>(*spec, outputscale, lengthscale, L, Q, x1, n1, x2,
                                                                   n2, v, out);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Constrained hyper-parameters of both additive kernels in one launch: out[row, l] = softplus(raw_row[l]) + lb_row
// (gpytorch's Positive / GreaterThan constraints as the reference's kernels use them, kernel_spec.py:58-69 and
// kernel_gen.py:219-310; torch.nn.functional.softplus semantics: beta 1, threshold 20), 1 for rows without a
// parameter.  The backward direction multiplies the upstream gradient by softplus' (and sums over l for a scalar
// parameter).  Replaces ~20 stock element-wise / concatenation kernels per step in front of the KL kernels.
namespace {
struct HyperRows {
    const double* raw[HLVAE_MAX_HYPER_ROWS];
    double lb[HLVAE_MAX_HYPER_ROWS];
    int bcast[HLVAE_MAX_HYPER_ROWS];
};

__global__ void __launch_bounds__(256)
hyper_constrain_k(const HyperRows rows, int n_rows, int L, double* __restrict__ out, const double* __restrict__ g_out,
                  double* __restrict__ g_raw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * L) return;
    const int row = e / L, l = e - row * L;
    const double* raw = rows.raw[row];
    if (!g_out) {
        double v = 1.0;
        if (raw) {
            const double x = raw[rows.bcast[row] ? 0 : l];
            v = (x > 20.0 ? x : log1p(exp(x))) + rows.lb[row];
        }
        out[e] = v;
    } else {
        double g = 0.0;
        if (raw) {
            if (!rows.bcast[row]) {
                const double x = raw[l];
                const double z = exp(x);
                g = x > 20.0 ? g_out[e] : g_out[e] * z / (z + 1.0);
            } else if (l == 0) {
                const double x = raw[0];
                const double z = exp(x);
                const double d = x > 20.0 ? 1.0 : z / (z + 1.0);
                for (int k = 0; k < L; k++) g += g_out[row * L + k];
                g *= d;
            }
        }
        g_raw[e] = g;
    }
}
}  // namespace

extern "C" int hlvae_hyper_constrain(int n_rows, int L, const double* const* raw, const int32_t* bcast, const double* lb,
                                     double* out, const double* g_out, double* g_raw, void* stream) {
    if (n_rows <= 0 || n_rows > HLVAE_MAX_HYPER_ROWS || L <= 0 || !raw || !bcast || !lb || (!g_out && !out) ||
        (g_out && !g_raw))
        return HLVAE_E_ARG;
    HyperRows rows;
    for (int i = 0; i < HLVAE_MAX_HYPER_ROWS; i++) {
        rows.raw[i] = i < n_rows ? raw[i] : nullptr;
        rows.lb[i] = i < n_rows ? lb[i] : 0.0;
        rows.bcast[i] = i < n_rows ? bcast[i] : 0;
    }
    const int n = n_rows * L;
    hyper_constrain_k<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rows, n_rows, L, out, g_out, g_raw);
    HLVAE_CHECK_LAUNCH();
    return 0;
}
