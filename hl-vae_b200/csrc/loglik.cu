// Fused masked heterogeneous log-likelihood: HLVAE.loglik_and_reconstruction
// (HLVAE.py:381-414) over the per-type functions of HL_VAE/loglik.py:27-213, plus
// read_functions.statistics (:268-302) and discrete_variables_transformation (:221-235).
// One thread per (row, variable); consecutive threads take consecutive variables of a row, so
// a warp touches one contiguous span of data / theta / params.  Arithmetic is float64 whatever
// the storage type, which also makes the categorical / ordinal argmax decisions reproduce the
// float64 reference (first index wins ties, as torch.argmax).
#include "common.cuh"

using namespace hlvae;

namespace {

constexpr int LL_THREADS = 256;
constexpr double LOG_2PI = 1.8378770664093454835606594728112;

__device__ __forceinline__ double softplus_d(double x) {
    // torch.nn.functional.softplus (beta=1, threshold=20)
    return x > 20.0 ? x : log1p(exp(x));
}
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }
// derivative of softplus_d (1 above the linear threshold, as torch's softplus backward)
__device__ __forceinline__ double dsoftplus_d(double x) { return x > 20.0 ? 1.0 : sigmoid_d(x); }

template <typename TM>
__device__ __forceinline__ double load_mask(const void* mask, int64_t i) {
    return (double)reinterpret_cast<const TM*>(mask)[i];
}

template <typename TS>
__device__ __forceinline__ void st(void* p, int64_t i, double v) {
    if (p) reinterpret_cast<TS*>(p)[i] = (TS)v;
}

// ------------------------------------------------------------------------------------
template <typename TS, typename TM>
__global__ void __launch_bounds__(LL_THREADS)
loglik_fwd_k(int64_t N, int D, int64_t ld_data, int64_t ld_theta, const int32_t* __restrict__ var_kind,
             const int32_t* __restrict__ var_nclass, const int32_t* __restrict__ var_dcol,
             const int32_t* __restrict__ var_pcol, const double* __restrict__ vparam, const TS* __restrict__ data,
             const TS* __restrict__ theta, const void* __restrict__ mask, void* __restrict__ log_p_x,
             void* __restrict__ log_p_x_missing, void* __restrict__ params, void* __restrict__ recon_mean,
             void* __restrict__ recon_mode, void* __restrict__ data_tr, double* __restrict__ ll_total) {
    __shared__ double red[LL_THREADS / 32];
    const int64_t idx = (int64_t)blockIdx.x * LL_THREADS + threadIdx.x;
    double lp_obs = 0.0;
    if (idx < N * D) {
        const int64_t n = idx / D;
        const int d = (int)(idx % D);
        const int kind = var_kind[d];
        const int C = var_nclass[d];
        const TS* x = data + n * ld_data + var_dcol[d];
        const TS* th = theta + n * ld_theta + var_pcol[d];
        const int64_t pbase = n * ld_theta + var_pcol[d];
        const double m = load_mask<TM>(mask, n * D + d);
        double lp = 0.0, rmean = 0.0, rmode = 0.0, dtr = 0.0;
        if (kind == HLVAE_VAR_REAL) {
            // loglik.py:27-70 (extra_params path)
            const double nm = vparam[d], nv = vparam[D + d], e = vparam[2 * D + d], div = vparam[3 * D + d];
            const double xv = (double)x[0] / div;                       // HLVAE.py:393-394
            const double lvy = -8.0 + softplus_d(e + 8.0);              // :51
            const double var = nv * exp(lvy);                            // :52,56
            const double mean = sqrt(nv) * (double)th[0] + nm;           // :55
            const double r = xv - mean;
            lp = -0.5 * r * r / var - 0.5 * LOG_2PI - 0.5 * log(var);    // :58
            st<TS>(params, pbase, mean);
            rmean = mean; rmode = mean;                                  // read_functions.py:275-278
            dtr = (double)x[0];                                          // read_functions.py:233
        } else if (kind == HLVAE_VAR_POS) {
            // loglik.py:73-121
            const double nm = vparam[d], nv = vparam[D + d], e = vparam[2 * D + d];
            const double ld = log(1.0 + (double)x[0]);                   // :84
            const double mean = sqrt(nv) * (double)th[0] + nm;           // :96
            const double var = nv * exp(e);                              // :100
            const double r = ld - mean;
            lp = -0.5 * r * r / var - 0.5 * log(2.0 * 3.14159265358979323846 * var) - ld;   // :102
            st<TS>(params, pbase, mean);
            const double v = exp(e);                                     // read_functions.py:284
            rmean = exp(mean + 0.5 * v) - 1.0;                           // :287
            rmode = exp(mean - v) - 1.0;                                 // :289
            dtr = (double)x[0];
        } else if (kind == HLVAE_VAR_COUNT) {
            // loglik.py:191-213
            double lam = softplus_d((double)th[0]);
            lam = fmin(fmax(lam, 1e-6), 1e20);                           // :203
            const double xv = (double)x[0];
            lp = xv * log(lam) - lam - lgamma(xv + 1.0);                 // Poisson.log_prob
            st<TS>(params, pbase, lam);
            rmean = lam; rmode = floor(lam);                             // read_functions.py:293-295
            dtr = xv;
        } else if (kind == HLVAE_VAR_CAT) {
            // loglik.py:124-146
            double t[HLVAE_MAX_CLASS];
            double mx = -INFINITY;
#pragma unroll 4
            for (int c = 0; c < C; c++) { t[c] = (double)th[c]; mx = fmax(mx, t[c]); }
            double se = 0.0;
            for (int c = 0; c < C; c++) se += exp(t[c] - mx);
            const double lse = mx + log(se);                             // torch.logsumexp
            // params = theta - lse (:134); log_p_x uses log_softmax of that again (:135)
            double se2 = 0.0, mx2 = -INFINITY;
            for (int c = 0; c < C; c++) { t[c] = t[c] - lse; mx2 = fmax(mx2, t[c]); }
            for (int c = 0; c < C; c++) se2 += exp(t[c] - mx2);
            const double lse2 = mx2 + log(se2);
            int am = 0, dam = 0;
            double best = t[0], dbest = (double)x[0];
            for (int c = 0; c < C; c++) {
                const double xv = (double)x[c];
                lp += xv * (t[c] - lse2);
                st<TS>(params, pbase + c, t[c]);
                if (t[c] > best) { best = t[c]; am = c; }                // argmax, first index on ties
                if (xv > dbest) { dbest = xv; dam = c; }
            }
            rmean = am; rmode = am;                                      // read_functions.py:296-302
            dtr = dam;                                                   // read_functions.py:226-227
        } else {
            // ordinal, loglik.py:149-188
            const double eps = 1e-6;
            const double loc = softplus_d((double)th[C - 1]);            // :163
            double p[HLVAE_MAX_CLASS];
            double cum = 0.0, prev = 0.0, tot = 0.0;
            int vals = 0;
            for (int c = 0; c < C; c++) {
                double sg = 1.0;
                if (c < C - 1) {
                    cum += fmin(fmax(softplus_d((double)th[c]), eps), 1e20);   // :164
                    sg = sigmoid_d(cum - loc);                                    // :165
                }
                p[c] = fmin(fmax(sg - prev, eps), 1.0);                          // :166-169
                prev = sg;
                tot += p[c];
                vals += (int)(double)x[c];                                       // :172
            }
            if (m == 0.0) vals = 1;                                              // :173
            int am = 0;
            double best = -1.0, lsum = 0.0, py = 0.0;
            for (int c = 0; c < C; c++) {
                const double ph = p[c] / tot;                                    // :178
                st<TS>(params, pbase + c, ph);
                if (ph > best) { best = ph; am = c; }
                lsum += ph;
                if (c == vals - 1) py = ph;
            }
            lp = log(py) - log(lsum);                                            // :179 log_softmax(log p)
            rmean = am; rmode = am;
            double sx = 0.0;
            for (int c = 0; c < C; c++) sx += (double)x[c];
            dtr = sx - 1.0;                                                      // read_functions.py:229-230
        }
        lp_obs = lp * m;
        st<TS>(log_p_x, n * D + d, lp_obs);
        st<TS>(log_p_x_missing, n * D + d, lp * (1.0 - m));
        st<TS>(recon_mean, n * D + d, rmean);
        st<TS>(recon_mode, n * D + d, rmode);
        st<TS>(data_tr, n * D + d, dtr);
    }
    if (ll_total) {
        lp_obs = warp_sum(lp_obs);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lp_obs;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < LL_THREADS / 32; w++) s += red[w];
            atomicAdd(ll_total, s);
        }
    }
}

// ------------------------------------------------------------------------------------
template <typename TS, typename TM>
__global__ void __launch_bounds__(LL_THREADS)
loglik_bwd_k(int64_t N, int D, int64_t ld_data, int64_t ld_theta, const int32_t* __restrict__ var_kind,
             const int32_t* __restrict__ var_nclass, const int32_t* __restrict__ var_dcol,
             const int32_t* __restrict__ var_pcol, const double* __restrict__ vparam, const TS* __restrict__ data,
             const TS* __restrict__ theta, const void* __restrict__ mask, const TS* __restrict__ g_lp, double g_scalar,
             TS* __restrict__ g_theta, double* __restrict__ g_lvy) {
    // grid: (ceil(D / LL_THREADS), row blocks); thread owns one variable, loops over a stripe of rows,
    // so the per-variable log-variance gradient reduces in a register.
    const int d = blockIdx.x * LL_THREADS + threadIdx.x;
    if (d >= D) return;
    const int kind = var_kind[d];
    const int C = var_nclass[d];
    const int dcol = var_dcol[d], pcol = var_pcol[d];
    const double nm = vparam[d], nv = vparam[D + d], e = vparam[2 * D + d], div = vparam[3 * D + d];
    const double snv = sqrt(nv);
    double ge = 0.0;
    double var = 1.0, sg8 = 0.0;
    if (kind == HLVAE_VAR_REAL) {
        var = nv * exp(-8.0 + softplus_d(e + 8.0));
        sg8 = dsoftplus_d(e + 8.0);
    } else if (kind == HLVAE_VAR_POS) {
        var = nv * exp(e);
    }
    for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
        const double m = load_mask<TM>(mask, n * D + d);
        const double g = (g_lp ? (double)g_lp[n * D + d] : g_scalar) * m;
        const TS* x = data + n * ld_data + dcol;
        const TS* th = theta + n * ld_theta + pcol;
        TS* gt = g_theta + n * ld_theta + pcol;
        if (kind == HLVAE_VAR_REAL) {
            const double r = (double)x[0] / div - (snv * (double)th[0] + nm);
            gt[0] = (TS)(g * snv * r / var);
            ge += g * (0.5 * r * r / var - 0.5) * sg8;
        } else if (kind == HLVAE_VAR_POS) {
            const double r = log(1.0 + (double)x[0]) - (snv * (double)th[0] + nm);
            gt[0] = (TS)(g * snv * r / var);
            ge += g * (0.5 * r * r / var - 0.5);
        } else if (kind == HLVAE_VAR_COUNT) {
            const double t0 = (double)th[0];
            const double sp = softplus_d(t0);
            double gl = 0.0;
            if (sp >= 1e-6 && sp <= 1e20) gl = ((double)x[0] / sp - 1.0) * dsoftplus_d(t0);
            gt[0] = (TS)(g * gl);
        } else if (kind == HLVAE_VAR_CAT) {
            double t[HLVAE_MAX_CLASS];
            double mx = -INFINITY, sx = 0.0;
            for (int c = 0; c < C; c++) { t[c] = (double)th[c]; mx = fmax(mx, t[c]); sx += (double)x[c]; }
            double se = 0.0;
            for (int c = 0; c < C; c++) { t[c] = exp(t[c] - mx); se += t[c]; }
            for (int c = 0; c < C; c++) gt[c] = (TS)(g * ((double)x[c] - t[c] / se * sx));
        } else {
            const double eps = 1e-6;
            const double t_loc = (double)th[C - 1];
            const double loc = softplus_d(t_loc);
            double sgm[HLVAE_MAX_CLASS], p[HLVAE_MAX_CLASS], q[HLVAE_MAX_CLASS];
            double cum = 0.0, prev = 0.0, tot = 0.0;
            int vals = 0;
            for (int c = 0; c < C; c++) {
                double sg = 1.0;
                if (c < C - 1) {
                    cum += fmin(fmax(softplus_d((double)th[c]), eps), 1e20);
                    sg = sigmoid_d(cum - loc);
                }
                sgm[c] = sg;
                q[c] = sg - prev;
                p[c] = fmin(fmax(q[c], eps), 1.0);
                prev = sg;
                tot += p[c];
                vals += (int)(double)x[c];
            }
            if (m == 0.0) vals = 1;
            const int y = vals - 1;
            // lp = log p_y - log tot ; clamp passes gradient inside [eps, 1]
            double gq[HLVAE_MAX_CLASS];
            for (int c = 0; c < C; c++) {
                double gp = ((c == y) ? 1.0 / p[c] : 0.0) - 1.0 / tot;
                gq[c] = (q[c] >= eps && q[c] <= 1.0) ? gp : 0.0;
            }
            // q_c = sg_c - sg_{c-1}: d/dsg_c = gq_c - gq_{c+1}, c < C-1; u_c = cum_c - loc
            double g_loc = 0.0, run = 0.0;
            for (int c = C - 2; c >= 0; c--) {
                const double gu = (gq[c] - gq[c + 1]) * sgm[c] * (1.0 - sgm[c]);
                g_loc -= gu;
                run += gu;                                   // reverse cumulative sum -> d/da_c
                const double tc = (double)th[c];
                const double sp = softplus_d(tc);
                const double ga = (sp >= eps && sp <= 1e20) ? run * dsoftplus_d(tc) : 0.0;
                gt[c] = (TS)(g * ga);
            }
            gt[C - 1] = (TS)(g * g_loc * dsoftplus_d(t_loc));
        }
    }
    if (g_lvy && (kind == HLVAE_VAR_REAL || kind == HLVAE_VAR_POS) && ge != 0.0) atomicAdd(g_lvy + d, ge);
}

// ------------------------------------------------------------------------------------
// Stand-alone monitoring transforms for callers that hold `params` / `data` only
// (training.py:84-91 calls them separately from the likelihood).
template <typename TS>
__global__ void __launch_bounds__(LL_THREADS)
statistics_k(int64_t N, int D, int64_t ld_theta, const int32_t* __restrict__ var_kind,
             const int32_t* __restrict__ var_nclass, const int32_t* __restrict__ var_pcol,
             const double* __restrict__ vparam, const TS* __restrict__ params, TS* __restrict__ mean,
             TS* __restrict__ mode) {
    const int64_t idx = (int64_t)blockIdx.x * LL_THREADS + threadIdx.x;
    if (idx >= N * D) return;
    const int64_t n = idx / D;
    const int d = (int)(idx % D);
    const int kind = var_kind[d], C = var_nclass[d];
    const TS* p = params + n * ld_theta + var_pcol[d];
    double a, b;
    if (kind == HLVAE_VAR_REAL) {
        a = b = (double)p[0];                                            // read_functions.py:275-278
    } else if (kind == HLVAE_VAR_POS) {
        const double v = exp(vparam[2 * D + d]);                         // :284
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
    mean[idx] = (TS)a;
    mode[idx] = (TS)b;
}

template <typename TS>
__global__ void __launch_bounds__(LL_THREADS)
discrete_transform_k(int64_t N, int D, int64_t ld_data, const int32_t* __restrict__ var_kind,
                     const int32_t* __restrict__ var_nclass, const int32_t* __restrict__ var_dcol,
                     const TS* __restrict__ data, TS* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * LL_THREADS + threadIdx.x;
    if (idx >= N * D) return;
    const int64_t n = idx / D;
    const int d = (int)(idx % D);
    const int kind = var_kind[d], C = var_nclass[d];
    const TS* x = data + n * ld_data + var_dcol[d];
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
    out[idx] = (TS)r;
}

bool args_ok(int64_t N, int D, const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* d,
             const double* vp, const void* data, const void* theta, const void* mask, int dtype) {
    return N >= 0 && D > 0 && a && b && c && d && vp && data && theta && mask && (dtype == HLVAE_F32 || dtype == HLVAE_F64);
}

}  // namespace

extern "C" int hlvae_loglik_fwd(int64_t N, int D, int64_t ld_data, int64_t ld_theta, const int32_t* var_kind,
                                const int32_t* var_nclass, const int32_t* var_dcol, const int32_t* var_pcol,
                                const double* vparam, const void* data, const void* theta, const void* mask,
                                int dtype, int mask_u8, void* log_p_x, void* log_p_x_missing, void* params,
                                void* recon_mean, void* recon_mode, void* data_tr, double* ll_total, void* stream) {
    if (!args_ok(N, D, var_kind, var_nclass, var_dcol, var_pcol, vparam, data, theta, mask, dtype)) return HLVAE_E_ARG;
    if (N == 0) return 0;
    const int64_t total = N * D;
    const unsigned grid = (unsigned)((total + LL_THREADS - 1) / LL_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
#define HLVAE_LL_FWD(TS, TM)                                                                                         \
    loglik_fwd_k<TS, TM><<<grid, LL_THREADS, 0, st>>>(N, D, ld_data, ld_theta, var_kind, var_nclass, var_dcol,       \
                                                      var_pcol, vparam, (const TS*)data, (const TS*)theta, mask,     \
                                                      log_p_x, log_p_x_missing, params, recon_mean, recon_mode,      \
                                                      data_tr, ll_total)
    if (dtype == HLVAE_F64) {
        if (mask_u8) HLVAE_LL_FWD(double, uint8_t); else HLVAE_LL_FWD(double, double);
    } else {
        if (mask_u8) HLVAE_LL_FWD(float, uint8_t); else HLVAE_LL_FWD(float, float);
    }
#undef HLVAE_LL_FWD
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_loglik_bwd(int64_t N, int D, int64_t ld_data, int64_t ld_theta, const int32_t* var_kind,
                                const int32_t* var_nclass, const int32_t* var_dcol, const int32_t* var_pcol,
                                const double* vparam, const void* data, const void* theta, const void* mask,
                                int dtype, int mask_u8, const void* g_lp, double g_scalar, void* g_theta,
                                double* g_lvy, void* stream) {
    if (!args_ok(N, D, var_kind, var_nclass, var_dcol, var_pcol, vparam, data, theta, mask, dtype) || !g_theta)
        return HLVAE_E_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // enough row stripes to fill the machine: 148 SMs x 8 resident CTAs, capped by N
    const unsigned gx = (unsigned)((D + LL_THREADS - 1) / LL_THREADS);
    unsigned gy = (unsigned)((148 * 8 + gx - 1) / gx);
    if ((int64_t)gy > N) gy = (unsigned)N;
    dim3 grid(gx, gy);
#define HLVAE_LL_BWD(TS, TM)                                                                                         \
    loglik_bwd_k<TS, TM><<<grid, LL_THREADS, 0, st>>>(N, D, ld_data, ld_theta, var_kind, var_nclass, var_dcol,       \
                                                      var_pcol, vparam, (const TS*)data, (const TS*)theta, mask,     \
                                                      (const TS*)g_lp, g_scalar, (TS*)g_theta, g_lvy)
    if (dtype == HLVAE_F64) {
        if (mask_u8) HLVAE_LL_BWD(double, uint8_t); else HLVAE_LL_BWD(double, double);
    } else {
        if (mask_u8) HLVAE_LL_BWD(float, uint8_t); else HLVAE_LL_BWD(float, float);
    }
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
    const unsigned grid = (unsigned)((N * D + LL_THREADS - 1) / LL_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HLVAE_F64)
        statistics_k<double><<<grid, LL_THREADS, 0, st>>>(N, D, ld_theta, var_kind, var_nclass, var_pcol, vparam,
                                                          (const double*)params, (double*)mean, (double*)mode);
    else
        statistics_k<float><<<grid, LL_THREADS, 0, st>>>(N, D, ld_theta, var_kind, var_nclass, var_pcol, vparam,
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
    const unsigned grid = (unsigned)((N * D + LL_THREADS - 1) / LL_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HLVAE_F64)
        discrete_transform_k<double><<<grid, LL_THREADS, 0, st>>>(N, D, ld_data, var_kind, var_nclass, var_dcol,
                                                                  (const double*)data, (double*)out);
    else
        discrete_transform_k<float><<<grid, LL_THREADS, 0, st>>>(N, D, ld_data, var_kind, var_nclass, var_dcol,
                                                                 (const float*)data, (float*)out);
    HLVAE_CHECK_LAUNCH();
    return 0;
}
