// Batch normalisation of the heterogeneous data batch: HL_VAE/utils.py:88-143 (`batch_normalization`), the
// encoder's input X_list and the normalisation parameters of the real / positive likelihoods, on the packed
// [N, E_x] layout.  Three streaming launches: per-variable masked count and sum, then the sum of squared masked
// deviations from the mean (the reference's two-pass variance, utils.py:106-107 / :125-126), then the
// element-wise transform.  Only real (non-convolutional) and positive variables take part in the first two.
#include "common.cuh"

namespace {

constexpr int BN_THREADS = 256;
constexpr int BN_COLS = 32;                    // statistic variables per CTA
constexpr int BN_LANES = BN_THREADS / BN_COLS; // row lanes per column

template <typename TS>
__device__ __forceinline__ TS log_t(TS v);
template <>
__device__ __forceinline__ double log_t<double>(double v) { return log(v); }
template <>
__device__ __forceinline__ float log_t<float>(float v) { return logf(v); }

// value the statistics are taken over: real -> d * m, pos -> log(1 + d * m)   (utils.py:98,104 / :122-124)
template <typename TS>
__device__ __forceinline__ TS bn_value(int kind, TS d, TS m) {
    const TS obs = d * m;
    return kind == HLVAE_VAR_POS ? log_t<TS>((TS)1 + obs) : obs;
}

template <typename TS, typename TD, typename TM>
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_k(int64_t N, int D, int64_t ld_data, const int32_t* __restrict__ var_kind, const int32_t* __restrict__ var_dcol,
           const int32_t* __restrict__ stat_vars, int n_stat, const TD* __restrict__ data, const TM* __restrict__ mask,
           int pass, double* __restrict__ stats) {
    __shared__ double red[2][BN_LANES][BN_COLS];
    const int c = threadIdx.x % BN_COLS, rl = threadIdx.x / BN_COLS;
    const int si = blockIdx.x * BN_COLS + c;
    const bool live = si < n_stat;
    const int d = live ? stat_vars[si] : 0;
    const int kind = var_kind[d];
    const TD* dp = data + var_dcol[d];
    const TM* mp = mask + d;
    const int64_t per = (N + gridDim.y - 1) / gridDim.y;
    const int64_t n_begin = (int64_t)blockIdx.y * per;
    const int64_t n_end = n_begin + per < N ? n_begin + per : N;
    double a0 = 0.0, a1 = 0.0;
    if (live) {
        if (pass == 0) {
            for (int64_t n = n_begin + rl; n < n_end; n += BN_LANES) {
                const TS m = (TS)mp[n * D];
                a0 += (double)m;
                a1 += (double)(bn_value<TS>(kind, (TS)dp[n * ld_data], m) * m);
            }
        } else {
            const TS mean = (TS)(stats[D + d] / stats[d]);
            for (int64_t n = n_begin + rl; n < n_end; n += BN_LANES) {
                const TS m = (TS)mp[n * D];
                const TS dev = (bn_value<TS>(kind, (TS)dp[n * ld_data], m) - mean) * m;
                a0 += (double)(dev * dev);
            }
        }
    }
    red[0][rl][c] = a0;
    red[1][rl][c] = a1;
    __syncthreads();
    if (rl == 0 && live) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < BN_LANES; k++) { s0 += red[0][k][c]; s1 += red[1][k][c]; }
        if (pass == 0) {
            atomicAdd(stats + d, s0);
            atomicAdd(stats + D + d, s1);
        } else {
            atomicAdd(stats + 2 * D + d, s0);
        }
    }
}

constexpr int BN_RB = 8;       // rows in flight per thread in the element-wise kernel

template <typename TS>
__device__ __forceinline__ TS bn_transform(int kind, bool standardise, TS x, TS m, TS mean, TS sd) {
    if (standardise) return (bn_value<TS>(kind, x, m) - mean) / sd * m;           // utils.py:108 / :128
    if (kind == HLVAE_VAR_REAL) return x * m / (TS)255;                            // :99-103
    if (kind == HLVAE_VAR_COUNT) return (m == (TS)0) ? (TS)0 : log_t<TS>(x * m);   // :113-119
    return x * m;                                                                  // :132-140
}

// thread = data column, rows of a stripe in batches of BN_RB with all loads of a batch issued before the first use
// (a 4-columns-per-thread variant with vector accesses was measured 2x slower: a quarter of the threads, same
// number of mask loads)
template <typename TS, typename TD, typename TM>
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_k(int64_t N, int D, int64_t ld_data, const int32_t* __restrict__ var_kind, const int32_t* __restrict__ dcol_var,
           const TD* __restrict__ data, const TM* __restrict__ mask, int conv, const double* __restrict__ meanvar,
           TS* __restrict__ out) {
    const int col = blockIdx.x * BN_THREADS + threadIdx.x;
    if (col >= ld_data) return;
    const int d = dcol_var[col];
    const int kind = var_kind[d];
    const bool standardise = (kind == HLVAE_VAR_POS) || (kind == HLVAE_VAR_REAL && !conv);
    const TS mean = standardise ? (TS)meanvar[d] : (TS)0;
    const TS sd = standardise ? (TS)sqrt(meanvar[D + d] + 1e-5) : (TS)1;
    const int64_t per = (N + gridDim.y - 1) / gridDim.y;
    const int64_t n_begin = (int64_t)blockIdx.y * per;
    const int64_t n_end = n_begin + per < N ? n_begin + per : N;
    const TD* dp = data + col;
    const TM* mp = mask + d;
    TS* op = out + col;
    int64_t n = n_begin;
    for (; n + BN_RB <= n_end; n += BN_RB) {
        TD x[BN_RB];
        TM m[BN_RB];
#pragma unroll
        for (int r = 0; r < BN_RB; r++) {
            x[r] = dp[(n + r) * ld_data];
            m[r] = mp[(n + r) * D];
        }
#pragma unroll
        for (int r = 0; r < BN_RB; r++)
            op[(n + r) * ld_data] = bn_transform<TS>(kind, standardise, (TS)x[r], (TS)m[r], mean, sd);
    }
    for (; n < n_end; n++)
        op[n * ld_data] = bn_transform<TS>(kind, standardise, (TS)dp[n * ld_data], (TS)mp[n * D], mean, sd);
}

int stripes_for(int64_t N, int tiles, int min_rows, int ctas_per_sm = 8) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = ((int64_t)sms * ctas_per_sm + tiles - 1) / tiles;
    const int64_t cap = (N + min_rows - 1) / min_rows;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    return (int)want;
}

template <typename TS, typename TD, typename TM>
int launch_stats(int64_t N, int D, int64_t ld, const int32_t* var_kind, const int32_t* var_dcol, const int32_t* stat_vars,
                 int n_stat, const void* data, const void* mask, int pass, double* stats, cudaStream_t st) {
    const int tiles = (n_stat + BN_COLS - 1) / BN_COLS;
    dim3 grid(tiles, stripes_for(N, tiles, 4 * BN_LANES));
    bn_stats_k<TS, TD, TM><<<grid, BN_THREADS, 0, st>>>(N, D, ld, var_kind, var_dcol, stat_vars, n_stat, (const TD*)data,
                                                        (const TM*)mask, pass, stats);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

template <typename TS, typename TD, typename TM>
int launch_apply(int64_t N, int D, int64_t ld, const int32_t* var_kind, const int32_t* dcol_var, const void* data,
                 const void* mask, int conv, const double* meanvar, void* out, cudaStream_t st) {
    const int tiles = (int)((ld + BN_THREADS - 1) / BN_THREADS);
    dim3 grid(tiles, stripes_for(N, tiles, 8 * BN_RB, 64));
    bn_apply_k<TS, TD, TM><<<grid, BN_THREADS, 0, st>>>(N, D, ld, var_kind, dcol_var, (const TD*)data, (const TM*)mask,
                                                        conv, meanvar, (TS*)out);
    HLVAE_CHECK_LAUNCH();
    return 0;
}

bool codes_ok(int dtype, int data_dtype, int mask_dtype) {
    return (dtype == HLVAE_F32 || dtype == HLVAE_F64) && (data_dtype == dtype || data_dtype == HLVAE_U8) &&
           (mask_dtype == dtype || mask_dtype == HLVAE_U8);
}

}  // namespace

#define HLVAE_BN_DISPATCH(CALL)                                                               \
    if (dtype == HLVAE_F32) {                                                                 \
        if (data_dtype == HLVAE_U8) {                                                         \
            if (mask_dtype == HLVAE_U8) return CALL(float, unsigned char, unsigned char);     \
            return CALL(float, unsigned char, float);                                         \
        }                                                                                     \
        if (mask_dtype == HLVAE_U8) return CALL(float, float, unsigned char);                 \
        return CALL(float, float, float);                                                     \
    }                                                                                         \
    if (data_dtype == HLVAE_U8) {                                                             \
        if (mask_dtype == HLVAE_U8) return CALL(double, unsigned char, unsigned char);        \
        return CALL(double, unsigned char, double);                                           \
    }                                                                                         \
    if (mask_dtype == HLVAE_U8) return CALL(double, double, unsigned char);                   \
    return CALL(double, double, double);

extern "C" int hlvae_batch_norm_stats(int64_t N, int D, int64_t ld_data, const int32_t* var_kind, const int32_t* var_dcol,
                                      const int32_t* stat_vars, int n_stat, const void* data, const void* mask, int dtype,
                                      int data_dtype, int mask_dtype, int pass, double* stats, void* stream) {
    if (N < 0 || D <= 0 || ld_data <= 0 || n_stat < 0 || !var_kind || !var_dcol || !data || !mask || !stats ||
        (n_stat > 0 && !stat_vars) || (pass != 0 && pass != 1) || !codes_ok(dtype, data_dtype, mask_dtype))
        return HLVAE_E_ARG;
    if (N == 0 || n_stat == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define HLVAE_BN_STATS(TS, TD, TM) \
    launch_stats<TS, TD, TM>(N, D, ld_data, var_kind, var_dcol, stat_vars, n_stat, data, mask, pass, stats, st)
    HLVAE_BN_DISPATCH(HLVAE_BN_STATS)
#undef HLVAE_BN_STATS
}

extern "C" int hlvae_batch_norm_apply(int64_t N, int D, int64_t ld_data, const int32_t* var_kind,
                                      const int32_t* dcol_var, const void* data, const void* mask, int dtype,
                                      int data_dtype, int mask_dtype, int conv, const double* meanvar, void* out,
                                      void* stream) {
    if (N < 0 || D <= 0 || ld_data <= 0 || !var_kind || !dcol_var || !data || !mask || !meanvar || !out ||
        !codes_ok(dtype, data_dtype, mask_dtype))
        return HLVAE_E_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define HLVAE_BN_APPLY(TS, TD, TM) \
    launch_apply<TS, TD, TM>(N, D, ld_data, var_kind, dcol_var, data, mask, conv, meanvar, out, st)
    HLVAE_BN_DISPATCH(HLVAE_BN_APPLY)
#undef HLVAE_BN_APPLY
}
