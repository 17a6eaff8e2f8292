// Likelihood branches outside the fused kernel's five-type fast path (csrc/loglik.cu):
//   REAL_ROWVAR : loglik_real with a variance network  (HL_VAE/loglik.py:45-48: the log-variance comes per ROW from theta)
//   POS_ROWVAR  : loglik_pos  with a variance network  (HL_VAE/loglik.py:89,104-108)
//   BETA        : loglik_beta                          (HL_VAE/loglik.py:216-256)
// One type group per launch, thread = (row, variable) element, float64 arithmetic, float32 / float64 storage.
// These are non-default configurations of the reference (logvar_network = False in config/hlvae_config_file.txt:50;
// no beta variable in the shipped data sets); the element-wise streaming form is HBM-bound as it stands.
#include "common.cuh"

using namespace hlvae;

namespace {

constexpr int AX_THREADS = 256;
constexpr double LOG_2PI = 1.8378770664093454835606594728112;

template <typename T> __device__ __forceinline__ double ldv(const void* p, int64_t i) {
    return (double)reinterpret_cast<const T*>(p)[i];
}
__device__ __forceinline__ double ld_any(const void* p, int64_t i, int dtype) {
    return dtype == HLVAE_F64 ? ldv<double>(p, i) : dtype == HLVAE_F32 ? ldv<float>(p, i) : ldv<uint8_t>(p, i);
}
__device__ __forceinline__ void st_any(void* p, int64_t i, double v, int dtype) {
    if (!p) return;
    if (dtype == HLVAE_F64) reinterpret_cast<double*>(p)[i] = v;
    else reinterpret_cast<float*>(p)[i] = (float)v;
}
__device__ __forceinline__ double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }      // torch threshold 20
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

// digamma(x), x > 0: recurrence up to x >= 6, then the asymptotic series (error < 1e-15 there)
__device__ __forceinline__ double digamma_d(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 +
                     f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}

struct AuxArgs {
    int mode;
    int64_t N;
    int Dg;
    const void* data; int64_t ld_data; int data_dtype;
    const void* mask; int64_t ld_mask; int mask_dtype;
    const void* th_a; int64_t ld_a; int64_t cs_a;
    const void* th_b; int64_t ld_b;
    int dtype;
    const double* vparam;      // [4, Dg]
    const double* disp;        // BETA: raw dispersion (1 element)
};

// value (and, when WITH_GRAD, the derivatives of log p w.r.t. th_a, th_b / the dispersion) of one element
template <bool WITH_GRAD>
__device__ __forceinline__ void eval_elem(const AuxArgs& a, int64_t n, int d, double& lp, double& pa, double& pb,
                                          double& ga, double& gb) {
    const double x = ld_any(a.data, n * a.ld_data + d, a.data_dtype);
    const double tha = ld_any(a.th_a, n * a.ld_a + d * a.cs_a, a.dtype);
    ga = gb = 0.0;
    if (a.mode == HLVAE_AUX_BETA) {
        const double dmin = a.vparam[d], dmax = a.vparam[a.Dg + d];
        const double xc = (x - dmin) / (dmax - dmin) + 1e-6;                       // loglik.py:224
        const double raw = a.disp[0];
        const double sp = softplus_d(raw);
        const double phi = fmin(fmax(sp, 1e-6), 1e20);                             // :240
        const double cdf = 0.5 * (1.0 + erf(tha * 0.70710678118654752440));       // td.Normal(0, 1).cdf, :242
        const double al = phi * cdf, be = phi * (1.0 - cdf);                       // :244-245
        const double lx = log(xc), l1x = log(1.0 - xc);
        lp = (al - 1.0) * lx + (be - 1.0) * l1x - lgamma(al) - lgamma(be) + lgamma(al + be);   // :247-248
        pa = al;
        pb = be;
        if (WITH_GRAD) {
            const double psi_s = digamma_d(al + be);
            const double dla = lx - digamma_d(al) + psi_s, dlb = l1x - digamma_d(be) + psi_s;
            const double pdf = exp(-0.5 * tha * tha) * 0.39894228040143267794;
            ga = phi * pdf * (dla - dlb);
            const double dphi = (sp > 1e-6 && sp < 1e20) ? sigmoid_d(raw) : 0.0;
            gb = (cdf * dla + (1.0 - cdf) * dlb) * dphi;                           // d log p / d raw dispersion
        }
        return;
    }
    const double nm = a.vparam[d], nv = a.vparam[a.Dg + d], dv = a.vparam[3 * a.Dg + d];
    const double e = ld_any(a.th_b, n * a.ld_b + d, a.dtype);
    const double mean = sqrt(nv) * tha + nm;                                       // :55 / :97
    if (a.mode == HLVAE_AUX_REAL) {
        const double lvy = -8.0 + softplus_d(e + 8.0);                             // :47
        const double var = nv * exp(lvy);                                          // :48,56
        const double r = x / dv - mean;
        lp = -0.5 * r * r / var - 0.5 * LOG_2PI - 0.5 * log(var);                  // :58
        pa = mean;
        pb = var;
        if (WITH_GRAD) {
            ga = sqrt(nv) * r / var;
            gb = (0.5 * r * r / var - 0.5) * (e + 8.0 > 20.0 ? 1.0 : sigmoid_d(e + 8.0));
        }
    } else {                                                                        // HLVAE_AUX_POS
        const double y = log(1.0 + x);                                             // :85
        const double var = nv * exp(e);                                            // :108
        const double r = y - mean;
        lp = -0.5 * r * r / var - 0.5 * log(2.0 * 3.14159265358979323846 * var) - y;   // :110
        pa = mean;
        pb = var;
        if (WITH_GRAD) {
            ga = sqrt(nv) * r / var;
            gb = 0.5 * r * r / var - 0.5;
        }
    }
}

__global__ void __launch_bounds__(AX_THREADS)
loglik_aux_fwd_k(const AuxArgs a, void* __restrict__ lpx, void* __restrict__ lpm, void* __restrict__ prm_a,
                 void* __restrict__ prm_b) {
    const int64_t total = a.N * a.Dg;
    for (int64_t e = (int64_t)blockIdx.x * AX_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * AX_THREADS) {
        const int64_t n = e / a.Dg;
        const int d = (int)(e - n * a.Dg);
        double lp, pa, pb, ga, gb;
        eval_elem<false>(a, n, d, lp, pa, pb, ga, gb);
        const double m = ld_any(a.mask, n * a.ld_mask + d, a.mask_dtype);
        st_any(lpx, e, lp * m, a.dtype);                                           // :62 / :113 / :251
        st_any(lpm, e, lp * (1.0 - m), a.dtype);
        st_any(prm_a, e, pa, a.dtype);
        st_any(prm_b, e, pb, a.dtype);
    }
}

// g_a, g_b: [N, Dg] per-element gradients of sum(g_lpx * log_p_x) (+ g_scalar * sum(log_p_x)); for BETA g_b is not
// written and the dispersion gradient is accumulated into g_disp (one atomic per warp).
__global__ void __launch_bounds__(AX_THREADS)
loglik_aux_bwd_k(const AuxArgs a, const void* __restrict__ g_lpx, const double* __restrict__ g_scalar,
                 void* __restrict__ g_a, void* __restrict__ g_b, double* __restrict__ g_disp) {
    const int64_t total = a.N * a.Dg;
    const double gs = g_scalar ? g_scalar[0] : 0.0;
    double disp_acc = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * AX_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * AX_THREADS) {
        const int64_t n = e / a.Dg;
        const int d = (int)(e - n * a.Dg);
        double lp, pa, pb, ga, gb;
        eval_elem<true>(a, n, d, lp, pa, pb, ga, gb);
        const double m = ld_any(a.mask, n * a.ld_mask + d, a.mask_dtype);
        const double g = ((g_lpx ? ld_any(g_lpx, e, a.dtype) : 0.0) + gs) * m;
        st_any(g_a, e, g * ga, a.dtype);
        if (a.mode == HLVAE_AUX_BETA) disp_acc = fma(g, gb, disp_acc);
        else st_any(g_b, e, g * gb, a.dtype);
    }
    if (a.mode == HLVAE_AUX_BETA && g_disp) {
        disp_acc = warp_sum(disp_acc);
        if ((threadIdx.x & 31) == 0 && disp_acc != 0.0) atomicAdd(g_disp, disp_acc);
    }
}

bool args_ok(int mode, int64_t N, int Dg, const void* data, const void* mask, const void* th_a, const void* th_b,
             int dtype, int data_dtype, int mask_dtype, const double* vparam, const double* disp) {
    if (mode != HLVAE_AUX_REAL && mode != HLVAE_AUX_POS && mode != HLVAE_AUX_BETA) return false;
    if (N < 0 || Dg <= 0 || !data || !mask || !th_a || !vparam) return false;
    if (mode == HLVAE_AUX_BETA ? !disp : !th_b) return false;
    if (dtype != HLVAE_F32 && dtype != HLVAE_F64) return false;
    auto okd = [](int c) { return c == HLVAE_F32 || c == HLVAE_F64 || c == HLVAE_U8; };
    return okd(data_dtype) && okd(mask_dtype);
}

unsigned grid_for(int64_t total) {
    int64_t b = (total + AX_THREADS - 1) / AX_THREADS;
    return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int hlvae_loglik_aux_fwd(int mode, int64_t N, int Dg, const void* data, int64_t ld_data, int data_dtype,
                                    const void* mask, int64_t ld_mask, int mask_dtype, const void* th_a, int64_t ld_a,
                                    int64_t cs_a, const void* th_b, int64_t ld_b, int dtype, const double* vparam,
                                    const double* disp, void* lpx, void* lpm, void* prm_a, void* prm_b, void* stream) {
    if (!args_ok(mode, N, Dg, data, mask, th_a, th_b, dtype, data_dtype, mask_dtype, vparam, disp) || !lpx || !lpm)
        return HLVAE_E_ARG;
    if (N == 0) return 0;
    AuxArgs a{mode, N, Dg, data, ld_data, data_dtype, mask, ld_mask, mask_dtype, th_a, ld_a, cs_a, th_b, ld_b, dtype,
              vparam, disp};
    loglik_aux_fwd_k<<<grid_for(N * Dg), AX_THREADS, 0, (cudaStream_t)stream>>>(a, lpx, lpm, prm_a, prm_b);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

extern "C" int hlvae_loglik_aux_bwd(int mode, int64_t N, int Dg, const void* data, int64_t ld_data, int data_dtype,
                                    const void* mask, int64_t ld_mask, int mask_dtype, const void* th_a, int64_t ld_a,
                                    int64_t cs_a, const void* th_b, int64_t ld_b, int dtype, const double* vparam,
                                    const double* disp, const void* g_lpx, const double* g_scalar, void* g_a, void* g_b,
                                    double* g_disp, void* stream) {
    if (!args_ok(mode, N, Dg, data, mask, th_a, th_b, dtype, data_dtype, mask_dtype, vparam, disp) || !g_a ||
        (mode != HLVAE_AUX_BETA && !g_b) || (!g_lpx && !g_scalar))
        return HLVAE_E_ARG;
    if (N == 0) return 0;
    AuxArgs a{mode, N, Dg, data, ld_data, data_dtype, mask, ld_mask, mask_dtype, th_a, ld_a, cs_a, th_b, ld_b, dtype,
              vparam, disp};
    loglik_aux_bwd_k<<<grid_for(N * Dg), AX_THREADS, 0, (cudaStream_t)stream>>>(a, g_lpx, g_scalar, g_a, g_b, g_disp);
    HLVAE_CHECK_LAUNCH();
    return 0;
}
