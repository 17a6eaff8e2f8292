// A/B probe for the sufficient-statistics contraction S = K^T V (elbo_functions.py:161 / :254,266) on the two
// tensor pipes of a B200, at the accuracy the path needs:
//
//   mode 0  FP64 tensor pipe: mma.sync.m8n8k4.f64 from shared-memory row panels - the form hlvae_kl_panel uses;
//   mode 1  5th-generation tensor cores: tcgen05.mma kind::i8 with int32 accumulators in TMEM, fed with an
//           ERROR-FREE SPLITTING of the float64 operands (Ozaki scheme): K and V are scaled to fixed point and cut
//           into `nslice` 8-bit slices written to shared memory in the UMMA K-major layout; every slice pair with
//           a + b < nslice is one exact integer GEMM; TMEM is drained with tcgen05.ld and recombined in float64.
//
// Why this and not a TF32 / BF16 split with fp32 accumulation: oracle/emulate_tensor_contraction.py shows that every
// fp32-ACCUMULATED scheme - direct or whitened, 3 or 6 products - misses the 1e-4 gate on grad_H / grad_m (and more)
// in all three numerical regimes, because cond(K0zz + eps I) ~ 5e7 amplifies the 2^-24 accumulation error; exact
// integer accumulation is the only tcgen05 arithmetic that can pass, and it needs 7 slices (56 bits).
// The probe measures what that costs next to the FP64 pipe on the same data (bench.py --workload contraction,
// tests/test_gpu_contraction_probe.py); DESIGN.md section 4 has the numbers and the decision.
//
// Not on the product path: hlvae_kl_panel keeps the FP64 pipe.  M = 64 only.
#include "common.cuh"

using namespace hlvae;

namespace {

constexpr int PM = 64;               // inducing points (UMMA N; two slices are stacked to fill UMMA M = 128)
constexpr int PK = 32;               // rows per tcgen05.mma (K of kind::i8)
constexpr int P_MAXSLICE = 8;
constexpr int A_BYTES = 128 * PK;    // one stacked A operand (two K slices): 4 KB
constexpr int B_BYTES = PM * PK;     // one V slice: 2 KB
constexpr uint32_t SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle ("interleave"), K-major: core matrix = 8 rows x 16 bytes stored as 128
// contiguous bytes; LBO = distance between the two 16-byte K chunks, SBO = distance between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version of sm_100
    return d;                        // base offset 0, layout type 0 = no swizzle
}

// instruction descriptor of kind::i8: int32 accumulate, a / b format 0 = unsigned, 1 = signed 8 bit, both K-major
__device__ __forceinline__ uint32_t umma_idesc_i8(int a_signed, int b_signed, int M, int N) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// bounded wait: false when the barrier did not complete (the caller then skips to the tear-down instead of hanging)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < SPIN_LIMIT; spin++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}

// 32 lanes x 32 columns of int32 accumulators -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// first stacked operand p whose pair (p, b = g - 2 p) exists for accumulator g: the MMA that initialises it
__device__ __forceinline__ int first_pair(int g, int nslice) {
    for (int p = 0;; p++)
        if (g - 2 * p < nslice) return p;
}

// byte `B` (0 = least significant, compile time) of four 64-bit fixed-point values packed into one word (value 0 in
// byte 0): three byte-permute instructions
template <int B>
__device__ __forceinline__ uint32_t pack_byte(const long long (&q)[16], int first) {
    uint32_t x[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
        x[u] = B < 4 ? (uint32_t)(unsigned long long)q[first + u] : (uint32_t)((unsigned long long)q[first + u] >> 32);
    constexpr uint32_t sel = (uint32_t)(B & 3) | ((uint32_t)(4 + (B & 3)) << 4);     // byte B of x[0], byte B of x[1]
    const uint32_t t0 = __byte_perm(x[0], x[1], sel), t1 = __byte_perm(x[2], x[3], sel);
    return __byte_perm(t0, t1, 0x5410);
}

template <int NS, int A>
__device__ __forceinline__ void store_slices(const long long (&qk)[16], const long long (&qv)[16], uint8_t* a_st,
                                             uint8_t* b_st, int ks, int kc, int m) {
    if constexpr (A < NS) {
        constexpr int B = NS - 1 - A;                      // byte of the fixed-point value that is slice A
        const int mr = (A & 1) * 64 + m;
        uint4 w, y;
        w.x = pack_byte<B>(qk, 0); w.y = pack_byte<B>(qk, 4); w.z = pack_byte<B>(qk, 8); w.w = pack_byte<B>(qk, 12);
        y.x = pack_byte<B>(qv, 0); y.y = pack_byte<B>(qv, 4); y.z = pack_byte<B>(qv, 8); y.w = pack_byte<B>(qv, 12);
        // element (row m', k) of a K-major operand: (m' % 8) * 16 + (m' / 8) * 256 + (k / 16) * 128 + k % 16
        *reinterpret_cast<uint4*>(a_st + (ks * (P_MAXSLICE / 2) + (A >> 1)) * A_BYTES + (mr & 7) * 16 + (mr >> 3) * 256 + kc * 128) = w;
        *reinterpret_cast<uint4*>(b_st + (ks * P_MAXSLICE + A) * B_BYTES + (m & 7) * 16 + (m >> 3) * 256 + kc * 128) = y;
        store_slices<NS, A + 1>(qk, qv, a_st, b_st, ks, kc, m);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// mode 1.  CTA = 256 threads, one CTA per SM (448 of the 512 TMEM columns).  A stage is 64 rows = two K = 32 steps;
// thread (m = tid % 64, h = tid / 64) converts rows 16 h .. 16 h + 15 of column m.  Two stages of operand buffers:
// while the tensor core reads stage s, the threads slice stage s + 1 (the completion barrier of a stage is only
// waited for before that stage is overwritten).  Thread 0 issues the MMAs; warps 0..3 drain the 128 TMEM lanes.
// Slice a (0 = most significant) of K sits in stacked operand a / 2, rows (a % 2) * 64 + m; accumulator g collects
// the pairs (operand p, V slice b) with 2 p + b = g: lanes 0..63 carry weight 2^-8g, lanes 64..127 weight 2^-8(g+1).
constexpr int PI_THREADS = 256;
constexpr int PI_STAGE_ROWS = 2 * PK;
constexpr int PI_A_STAGE = 2 * (P_MAXSLICE / 2) * A_BYTES;      // [kstep][pair][A_BYTES]
constexpr int PI_B_STAGE = 2 * P_MAXSLICE * B_BYTES;            // [kstep][slice][B_BYTES]
constexpr int PI_SMEM = 2 * (PI_A_STAGE + PI_B_STAGE);

template <int NS>
__global__ void __launch_bounds__(PI_THREADS, 1)
probe_i8_k(int64_t N, const double* __restrict__ K, const double* __restrict__ V, double k_scale,
           double v_scale, int rows_per_cta, double* __restrict__ S, int32_t* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t pi_smem[];
    __shared__ __align__(8) uint64_t bar_free[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int l = blockIdx.y;
    const int64_t row_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t row_end = min(N, row_begin + rows_per_cta);
    const double* Kl = K + (int64_t)l * N * PM;
    const double* Vl = V + (int64_t)l * N * PM;
    constexpr int nslice = NS;
    constexpr int nacc = NS;                           // accumulators g = 0 .. nslice - 1 (64 columns each)
    constexpr int np = (NS + 1) / 2;                   // stacked A operands

    if (tid == 0) {
        mbar_init(&bar_free[0], 1);
        mbar_init(&bar_free[1], 1);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // an odd slice count leaves the lower half of the last stacked operand unused: it must read as zeros
    for (int e = tid; e < PI_SMEM / 16; e += PI_THREADS) reinterpret_cast<uint4*>(pi_smem)[e] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    const int m = tid & 63, h = tid >> 6;              // h: 16-row group of the stage; K step h / 2, K chunk h % 2
    const double kmul = ldexp(1.0, 8 * nslice) / k_scale;          // K / k_scale in [0, 1) -> unsigned 8 nslice bits
    const double vmul = ldexp(1.0, 8 * nslice - 1) / v_scale;      // V / v_scale in [-1, 1) -> signed
    uint32_t phase[2] = {0, 0};
    bool ok = true;
    int step = 0;
    for (int64_t r0 = row_begin; r0 < row_end && ok; r0 += PI_STAGE_ROWS, step++) {
        const int st = step & 1;
        uint8_t* a_st = pi_smem + st * (PI_A_STAGE + PI_B_STAGE);
        uint8_t* b_st = a_st + PI_A_STAGE;
        // ---- load and convert this thread's 16 rows (rows beyond the end contribute zeros)
        long long qk[16], qv[16];
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int64_t r = r0 + h * 16 + u;
            const bool in = r < row_end;
            const double kv = in ? Kl[r * PM + m] : 0.0;
            const double vv = in ? Vl[r * PM + m] : 0.0;
            qk[u] = __double2ll_rd(kv * kmul);
            qv[u] = __double2ll_rd(vv * vmul);
        }
        if (step >= 2) {                                                   // the MMAs that read this stage are done?
            ok = mbar_wait(&bar_free[st], phase[st]);
            phase[st] ^= 1;
        }
        store_slices<NS, 0>(qk, qv, a_st, b_st, h >> 1, h & 1, m);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> tensor-core reads
        __syncthreads();
        if (tid == 0 && ok) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int kk = 0; kk < 2; kk++)
                for (int g = 0; g < nacc; g++)
                    for (int p = 0; p < np; p++) {
                        const int b = g - 2 * p;
                        if (b < 0 || b >= nslice) continue;
                        const uint64_t ad = umma_desc(smem_u32(a_st + (kk * (P_MAXSLICE / 2) + p) * A_BYTES), 128, 256);
                        const uint64_t bd = umma_desc(smem_u32(b_st + (kk * P_MAXSLICE + b) * B_BYTES), 128, 256);
                        // K slices are unsigned; the most significant V slice carries the sign.  The very first
                        // MMA into an accumulator overwrites it.
                        umma_i8(tmem + 64 * g, ad, bd, umma_idesc_i8(0, b == 0 ? 1 : 0, 128, PM),
                                (step == 0 && kk == 0 && p == first_pair(g, nslice)) ? 0u : 1u);
                    }
            umma_commit(&bar_free[st]);                                    // arrives when the MMAs above have read smem
        }
    }
    // all MMAs complete <=> the last commit of each stage has arrived
    for (int st = 0; st < 2 && ok; st++) {
        const int uses = (step + 1 - st) / 2;                              // steps that used stage st
        if (uses > 0) ok = mbar_wait(&bar_free[st], phase[st]);
    }
    if (!ok && tid == 0) report_status(status, 3, l, (int)blockIdx.x);

    // ---- drain: warp w < 4 owns TMEM lanes 32 w .. 32 w + 31 (row m' = 32 w + lane of every accumulator)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok && row_end > row_begin && warp < 4) {
        const int mrow = warp * 32 + lane;                                 // 0..127: upper half = next slice weight
        const int mi = mrow & 63;
        double out[PM];
#pragma unroll
        for (int n = 0; n < PM; n++) out[n] = 0.0;
        for (int g = 0; g < nacc; g++) {
            // value = sum over (a, b) of slice_a(K) slice_b(V) 2^(8 (nslice-1-a) + 8 (nslice-1-b)) / (kmul vmul)
            const int a_weight = g + (mrow >> 6);                          // a + b of this half
            const double wgt = ldexp(1.0, 8 * (2 * nslice - 2 - a_weight)) / (kmul * vmul);
#pragma unroll
            for (int c0 = 0; c0 < PM; c0 += 32) {
                int32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 64 * g + c0, v);
#pragma unroll
                for (int n = 0; n < 32; n++) out[c0 + n] = fma((double)v[n], wgt, out[c0 + n]);
            }
        }
        double* Sl = S + (int64_t)l * PM * PM + (int64_t)mi * PM;
#pragma unroll
        for (int n = 0; n < PM; n++) atomicAdd(Sl + n, out[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// mode 0: the FP64 tensor pipe, 512 threads, 64-row panels staged in shared memory, accumulators in registers over
// the CTA's whole chunk - the S update of hlvae_kl_panel (kl_stream.cu, "P3b") on its own.
__global__ void __launch_bounds__(512, 1)
probe_dmma_k(int64_t N, const double* __restrict__ K, const double* __restrict__ V, int rows_per_cta,
             double* __restrict__ S) {
    constexpr int LD = PM + 4, RP = 64;
    extern __shared__ __align__(16) double probe_smem[];
    double* Kb = probe_smem;
    double* Vb = probe_smem + RP * LD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wi = warp >> 2, wj = warp & 3;
    const int l = blockIdx.y;
    const int64_t row_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t row_end = min(N, row_begin + rows_per_cta);
    const double* Kl = K + (int64_t)l * N * PM;
    const double* Vl = V + (int64_t)l * N * PM;
    double sacc[2][2][2] = {};
    const int kr = lane & 3, kc = lane >> 2;
    for (int64_t r0 = row_begin; r0 < row_end; r0 += RP) {
        for (int e = tid; e < RP * PM; e += 512) {
            const int r = e / PM, c = e % PM;
            const bool in = r0 + r < row_end;
            Kb[r * LD + c] = in ? Kl[(r0 + r) * PM + c] : 0.0;
            Vb[r * LD + c] = in ? Vl[(r0 + r) * PM + c] : 0.0;
        }
        __syncthreads();
        for (int k0 = 0; k0 < RP; k0 += 4) {
            double af[2], bf[2];
#pragma unroll
            for (int t = 0; t < 2; t++) af[t] = Kb[(k0 + kr) * LD + (wi * 2 + t) * 8 + kc];
#pragma unroll
            for (int t = 0; t < 2; t++) bf[t] = Vb[(k0 + kr) * LD + (wj * 2 + t) * 8 + kc];
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) dmma884(sacc[a][b][0], sacc[a][b][1], af[a], bf[b]);
        }
        __syncthreads();
    }
    double* Sl = S + (int64_t)l * PM * PM;
    const int cr = lane >> 2, cc = 2 * (lane & 3);
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int i = (wi * 2 + a) * 8 + cr, j = (wj * 2 + b) * 8 + cc;
            atomicAdd(Sl + i * PM + j, sacc[a][b][0]);
            atomicAdd(Sl + i * PM + j + 1, sacc[a][b][1]);
        }
}

}  // namespace

extern "C" int hlvae_contraction_probe(int mode, int nslice, int L, int64_t N, int M, const double* K, const double* V,
                                       double k_scale, double v_scale, int rows_per_cta, double* S, int32_t* status,
                                       void* stream) {
    if (L <= 0 || N < 0 || !K || !V || !S || rows_per_cta <= 0) return HLVAE_E_ARG;
    if (M != PM) return HLVAE_E_UNSUPPORTED;
    if (N == 0) return 0;
    dim3 grid((unsigned)((N + rows_per_cta - 1) / rows_per_cta), (unsigned)L);
    if (mode == 0) {
        const int smem = 2 * 64 * (PM + 4) * (int)sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(probe_dmma_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        probe_dmma_k<<<grid, 512, smem, (cudaStream_t)stream>>>(N, K, V, rows_per_cta, S);
    } else if (mode == 1) {
        // int32 accumulators: a row contributes < 2^16 per slice pair, an accumulator sums <= 4 pairs
        // (7 slices = 56 bits cover the 53-bit significand; 8 would overflow the 64-bit fixed-point value)
        if (nslice < 2 || nslice > 7 || !(k_scale > 0.0) || !(v_scale > 0.0) || rows_per_cta > 8192)
            return HLVAE_E_ARG;
#define HLVAE_PROBE_I8(NS)                                                                                          \
    case NS: {                                                                                                      \
        cudaError_t e = cudaFuncSetAttribute(probe_i8_k<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, PI_SMEM); \
        if (e != cudaSuccess) return (int)e;                                                                        \
        probe_i8_k<NS><<<grid, PI_THREADS, PI_SMEM, (cudaStream_t)stream>>>(N, K, V, k_scale, v_scale, rows_per_cta, \
                                                                            S, status);                             \
        break;                                                                                                      \
    }
        switch (nslice) {
            HLVAE_PROBE_I8(2) HLVAE_PROBE_I8(3) HLVAE_PROBE_I8(4) HLVAE_PROBE_I8(5) HLVAE_PROBE_I8(6) HLVAE_PROBE_I8(7)
        }
#undef HLVAE_PROBE_I8
    } else {
        return HLVAE_E_ARG;
    }
    HLVAE_CHECK_LAUNCH();
    return 0;
}
