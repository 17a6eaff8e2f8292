/*
 * hlvae_b200 - C ABI of the B200-native HL-VAE ELBO hot path.
 *
 * The reference (MineOgre/HL-VAE) is pure Python and has no FFI: its boundary for this
 * path is a set of Python call sites.  Each entry point below names the reference
 * interface it sits behind (paths relative to the reference root); the Python host layer
 * in hl-vae_b200/ mirrors those interfaces and calls these functions through ctypes.
 *
 * Conventions
 *  - extern "C", plain pointers and integer sizes only; no torch types.
 *  - Every pointer is a DEVICE pointer owned by the caller for the duration of the call
 *    (the library allocates nothing and keeps no pointer after return), except
 *    `hlvae_kspec_t*`, which is a HOST pointer copied into kernel parameters.
 *  - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work; no host sync.
 *  - Return value: 0 on success, negative on argument error (HLVAE_E_*), positive =
 *    cudaError_t from the launch.  Numerical failures (non positive-definite block,
 *    subject longer than HLVAE_TMAX) are reported through the device status words
 *    `status[0..3]` = {code, latent, subject, 0}; the host wrapper reads them.
 *  - Functions are re-entrant and hold no global mutable state.
 *  - Floating-point arrays are float64 unless a `dtype` argument says otherwise
 *    (HLVAE_F32 / HLVAE_F64 select the storage type of streamed arrays; arithmetic is
 *    float64 inside the kernels either way).
 */
#ifndef HLVAE_B200_H
#define HLVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HLVAE_ABI_VERSION 1

#define HLVAE_MAX_COMPS 8   /* additive components per kernel                      */
#define HLVAE_MAX_DISC 3    /* categorical / binary factors per component          */
#define HLVAE_MAX_Q 8       /* covariate columns                                   */
#define HLVAE_TMAX 64       /* rows per subject (T <= 32: warp per (subject, latent dim); 33..64: CTA per pair) */

#define HLVAE_F32 0
#define HLVAE_F64 1
#define HLVAE_U8 2          /* exact small integers: masks, one-hot / thermometer codes, pixel values */

#define HLVAE_KIND_CAT 1    /* kernel_spec.py:26-32  CatKernel: x1 == x2           */
#define HLVAE_KIND_BIN 2    /* kernel_spec.py:9-23   BinKernel: x1 + x2 == 2       */

#define HLVAE_E_ARG (-1)
#define HLVAE_E_UNSUPPORTED (-2)

#define HLVAE_STATUS_NOT_PD 1
#define HLVAE_STATUS_T_TOO_LARGE 2

/* One ScaleKernel term of an additive kernel (kernel_gen.py:219-310): outputscale *
 * [SE(x[se_col]; lengthscale)] * prod_i disc_i.  se_col < 0: no squared-exponential
 * factor.  Component r uses outputscale[r][l] and lengthscale[r][l]. */
typedef struct {
    int32_t se_col;
    int32_t ndisc;
    int32_t disc_kind[HLVAE_MAX_DISC];
    int32_t disc_col[HLVAE_MAX_DISC];
} hlvae_comp_t;

typedef struct {
    int32_t ncomp;
    int32_t reserved;
    hlvae_comp_t comp[HLVAE_MAX_COMPS];
} hlvae_kspec_t;

int hlvae_version(void);
/* sizeof(hlvae_kspec_t) as compiled, so a binding can verify its struct layout. */
int hlvae_sizeof_kspec(void);

/* ------------------------------------------------------------------------------------
 * Dense additive-kernel evaluation.  Replaces `covar_module(x1, x2).evaluate()`
 * (elbo_functions.py:147-148,222-223; utils.py:128-130) for kernel objects built by
 * kernel_gen.generate_kernel_batched (kernel_gen.py:199-310).
 *   x1: [n1,Q] (bs1 = 0) or [L,n1,Q] (bs1 = n1*ld1) row-major with leading dim ld1
 *   out: [L,n1,n2];  outputscale, lengthscale: [ncomp, L] (constrained values)
 * bwd: g_out [L,n1,n2] -> g_os, g_ls [ncomp,L] (accumulated, caller zero-fills),
 *   g_x1 / g_x2 [L,n,Q] (nullable, accumulated; squared-exponential columns only).
 * ---------------------------------------------------------------------------------- */
int hlvae_kernel_eval_fwd(const hlvae_kspec_t* spec, const double* outputscale, const double* lengthscale,
                          int L, int Q, const double* x1, int n1, int64_t ld1, int64_t bs1,
                          const double* x2, int n2, int64_t ld2, int64_t bs2, double* out, void* stream);
int hlvae_kernel_eval_bwd(const hlvae_kspec_t* spec, const double* outputscale, const double* lengthscale,
                          int L, int Q, const double* x1, int n1, int64_t ld1, int64_t bs1,
                          const double* x2, int n2, int64_t ld2, int64_t bs2, const double* g_out,
                          double* g_os, double* g_ls, double* g_x1, double* g_x2, void* stream);

/* K1(X*, x) v restricted to same-subject pairs, the last term of the GP posterior-mean predictors
 * (utils.py:175-186 / :255-267): out[i, l] = sum over the rows j of subject sid[i] of
 * K_l(xt[i], x[row_idx[j]]) v[row_idx[j], l];  sid[i] < 0 (subject not among the conditioning rows) -> 0.
 * xt [nt,Q], x [N,Q], v [N,L], out [nt,L] float64 row-major; (row_idx, subj_ptr) as in hlvae_kl_subject. */
int hlvae_subject_matvec(const hlvae_kspec_t* spec, const double* outputscale, const double* lengthscale,
                         int L, int Q, const double* xt, int nt, const double* x, const int32_t* row_idx,
                         const int32_t* subj_ptr, const int32_t* sid, const double* v, double* out, void* stream);

/* K(x1, x2_l) v_l without materialising the block: out[i, l] = sum_j K_l(x1[i], x2[l, j]) v[l, j], the
 * K0Xz (iK K0zx mu_tilde) product of the GP posterior-mean predictors (utils.py:169 / :249).
 * x1 [n1,Q], x2 [L,n2,Q], v [L,n2], out [n1,L] float64 row-major. */
int hlvae_kernel_matvec(const hlvae_kspec_t* spec, const double* outputscale, const double* lengthscale,
                        int L, int Q, const double* x1, int n1, const double* x2, int n2, const double* v,
                        double* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Streaming part of the KL upper bound: everything in
 * elbo_functions.minibatch_KLD_upper_bound (elbo_functions.py:118-193) and
 * minibatch_KLD_upper_bound_iter (:196-285) that scales with the minibatch.
 *
 * Subjects are described in CSR form over a row permutation: subject s owns rows
 * row_idx[subj_ptr[s] .. subj_ptr[s+1]) of x / mu / log_v; tt_ptr[s] = sum_{s'<s} T_s'^2.
 *
 * With J = 1/2 (A + B + C + D + E - F) of elbo_functions.py:166-173 (w = iK m and
 * G = iK H iK - iK held constant), the two calls produce, per latent dimension l:
 *   S = sum_s K0xz_s^T B_s^-1 K0xz_s   (:161 / :254,266)      p = sum_s K0xz_s^T B_s^-1 mu_s (:188 / :265)
 *   gw = dJ/dw = sum_s K0xz_s^T B_s^-1 (K0xz_s w - mu_s)
 *   scal[0] = A (:166-167), scal[1] = B + sum(iB * K0_st) (:168,170), scal[2] = C (:169), scal[3] = F (:173)
 * and dJ/d{mu, log_v, Z, outputscale0, lengthscale0, outputscale1, lengthscale1}.
 *
 * hlvae_kl_subject : per-(subject, l) warp: B_s = K1(x_s,x_s) + noise I (:150-151 / :249-250),
 *                    Cholesky, explicit inverse (:156-157 / :251-252), writes binv
 *                    [L, tt_total] and the terms that do not involve Z.
 * hlvae_kl_panel   : per-(l, subject chunk) CTA: K0xz rows on the fly, B^-1 K0xz, the
 *                    sufficient statistics (FP64 tensor-core mma), and the remaining gradients.
 * `acc` is one float64 buffer the caller zero-fills; hlvae_kl_acc_layout gives offsets
 * (in doubles) of {S, p, gw, scal, gZ, gos0, gls0, gos1, gls1, total}.  g_mu, g_logv:
 * [N, L] contiguous in the storage `dtype`, overwritten for the rows listed in row_idx with
 * gscale * dJ/dmu and gscale * dJ/dlog_v (gscale = P / P_batch of elbo_functions.py:181,277).
 * qdiag (nullable, [N, L], storage dtype): per row r the quadratic form (B^-1 K0xz)_r G (B^-1 K0xz)_r^T, the
 * diagonal that validation.validation_dubo needs with G = W^-1 (validation.py:69-72).
 * row_panel: rows per panel of hlvae_kl_panel (whole subjects are packed into panels, so every subject must have
 * <= row_panel rows; a subject that cannot be packed is skipped and reported as HLVAE_STATUS_T_TOO_LARGE); 0 = the
 * default (64 rows for every M); 40 = the two-CTAs-per-SM shape for 32 < M <= 64 (it pays when subjects fill 40-row
 * panels about as well as 64-row ones, e.g. T = 20); 32 / 48 = smaller panels for 64 < M <= 128.  Scheduling only:
 * results agree to float64 summation order.
 * ---------------------------------------------------------------------------------- */
#define HLVAE_ACC_S 0
#define HLVAE_ACC_P 1
#define HLVAE_ACC_GW 2
#define HLVAE_ACC_SCAL 3
#define HLVAE_ACC_GZ 4
#define HLVAE_ACC_GOS0 5
#define HLVAE_ACC_GLS0 6
#define HLVAE_ACC_GOS1 7
#define HLVAE_ACC_GLS1 8
#define HLVAE_ACC_TOTAL 9
#define HLVAE_NSCAL 4
int hlvae_kl_acc_layout(int L, int M, int Q, int64_t* offsets /* [10] */);

int hlvae_kl_subject(const hlvae_kspec_t* spec0, const double* os0, const double* ls0,
                     const hlvae_kspec_t* spec1, const double* os1, const double* ls1, const double* noise,
                     int L, int Q, const double* x, int64_t ldx,
                     const int32_t* row_idx, const int32_t* subj_ptr, const int32_t* tt_ptr, int n_subj, int t_cap,
                     const void* log_v, int64_t ld_lv, int dtype,
                     double* binv, int64_t tt_total, double* acc, int M, void* g_logv, double gscale,
                     int32_t* status, void* stream);

int hlvae_kl_panel(const hlvae_kspec_t* spec0, const double* os0, const double* ls0,
                   const hlvae_kspec_t* spec1, const double* os1, const double* ls1,
                   int L, int Q, int M, const double* x, int64_t ldx, const double* z,
                   const int32_t* row_idx, const int32_t* subj_ptr, const int32_t* tt_ptr, int n_subj,
                   int subj_per_chunk, const void* mu, int64_t ld_mu, int dtype,
                   const double* w, const double* G, const double* binv, int64_t tt_total,
                   double* acc, void* g_mu, void* qdiag, double gscale, int32_t* status, int row_panel, void* stream);

/* ------------------------------------------------------------------------------------
 * Constrained hyper-parameters (gpytorch Positive / GreaterThan: softplus(raw) + lower bound) of the outputscales
 * and lengthscales the reference's kernel objects own (kernel_spec.py:58-69, kernel_gen.py:219-310; the
 * `outputscale` / `lengthscale` arguments of every entry point above): one launch for up to HLVAE_MAX_HYPER_ROWS
 * parameters of L (or 1, broadcast) values each.  raw / bcast / lb are HOST arrays of n_rows entries (raw[i] = device
 * pointer or NULL for "no parameter": the row is 1).  Forward (g_out == NULL): out[n_rows, L].  Backward (g_out =
 * d loss / d out [n_rows, L]): g_raw[n_rows, L] = g_out * softplus'(raw); a broadcast row holds its sum in column 0.
 * ---------------------------------------------------------------------------------- */
#define HLVAE_MAX_HYPER_ROWS (4 * HLVAE_MAX_COMPS)
int hlvae_hyper_constrain(int n_rows, int L, const double* const* raw, const int32_t* bcast, const double* lb,
                          double* out, const double* g_out, double* g_raw, void* stream);

/* ------------------------------------------------------------------------------------
 * Replicated M x M stage (float64, one CTA per latent dimension, M <= 128).
 * hlvae_mxm_pre : K0zz = K0(Z,Z) + eps I (:148,153 / :223-224), Cholesky and explicit inverses
 *   iK, iH (:154-157,162-163 / :225-228), w = iK m, G = sym(iK H iK) - iK; pre[l] =
 *   {logdet K, logdet H, tr(iK H), m^T iK m} (:176-179 / :271-274).  All outputs [L,...] float64.
 * hlvae_mxm_post: from the (all-reduced) accumulators S, p, gw, scal: kld[0] += kld_total
 *   (:181 / :277, `constant` = L * N / 2 subtracted once; `kld` is a caller-zeroed float64 [2 + L]: kld[1] counts
 *   arrived CTAs and kld[2 + l] parks the per-latent terms, summed in index order by the last CTA so that the total
 *   is bit-identical from launch to launch and across data-parallel ranks), the natural-gradient pieces ng_m, ng_H
 *   (:186-191 / :279-283; nullable) and, with f = c0 (tr(G S)/2 + w^T gw) + kld_qu_pu,
 *   gK_over_c0 = (df/dK0zz) / c0, gH = df/dH, gm = df/dm.  c0 = P / P_batch.
 * hlvae_natgrad_update: training.py:130-137, (m, H, grad_m, grad_H, lr) -> (m_out, H_out); `iH` (nullable) is
 *   H^-1 when the caller already has it from hlvae_mxm_pre of the same step (skips :131-132).
 * hlvae_mxm_pre launches 2 CTAs per latent dimension (K0zz path and H path are independent).
 * `ws`: global workspace of hlvae_mxm_workspace_doubles(L, M) doubles (0 when M <= 64).
 * ---------------------------------------------------------------------------------- */
int64_t hlvae_mxm_workspace_doubles(int L, int M);
int hlvae_mxm_pre(const hlvae_kspec_t* spec0, const double* os0, const double* ls0, int L, int Q, int M,
                  const double* z, double eps, const double* m, const double* H, double* iK, double* iH,
                  double* w, double* G, double* pre, double* ws, int32_t* status, void* stream);
int hlvae_mxm_post(int L, int M, double c0, double constant, const double* iK, const double* iH,
                   const double* H, const double* m, const double* w, const double* G, const double* pre,
                   const double* S, const double* p, const double* gw, const double* scal, double* kld,
                   double* gK_over_c0, double* gH, double* gm, double* ng_m, double* ng_H, double* ws,
                   void* stream);
int hlvae_natgrad_update(int L, int M, double lr, const double* m, const double* H, const double* iH,
                         const double* grad_m, const double* grad_H, double* m_out, double* H_out, double* ws,
                         int32_t* status, void* stream);

/* M x M epilogue of the evaluation-time bounds and predictors (one CTA per latent dimension, float64):
 * from the streaming statistics S = K0zx iB K0xz [L,M,M] and p = K0zx iB y [L,M] (either nullable = 0), with
 * W = K0zz + eps I + sym(S):
 *   scal[l] = {log det(K0zz + eps I), log det W, |L_W^-1 p|^2, sum(S * iK)}
 *             (elbo_functions.py:47-53 / :95-105, validation.py:57-66)
 *   a = W^-1 p (utils.py:162 / :246), c = (K0zz + eps I)^-1 (p - S a) (utils.py:169 / :249), iW = W^-1
 *   (validation.py:71, elbo_functions.py:111) - each output nullable.
 * Replaces the removed torch.solve calls of the reference; `ws` as in hlvae_mxm_pre. */
int hlvae_mxm_aux(const hlvae_kspec_t* spec0, const double* os0, const double* ls0, int L, int Q, int M,
                  const double* z, double eps, const double* S, const double* p, double* scal, double* a,
                  double* c, double* iW, double* ws, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused masked heterogeneous log-likelihood.  Replaces HLVAE.loglik_and_reconstruction
 * (HLVAE.py:381-414) with the per-type functions of HL_VAE/loglik.py:27-213, plus the
 * per-step monitoring transforms read_functions.statistics (:268-302, incl. the
 * categorical / ordinal argmax imputation) and discrete_variables_transformation (:221-235).
 *
 * Variable d is described by var_kind[d] (HLVAE_VAR_*), var_nclass[d], var_dcol[d] (first
 * column in data), var_pcol[d] (first column in theta / params); vparam [4, D] float64 holds
 * per variable {normalisation mean, normalisation variance (already clamped), raw
 * log-variance parameter, data divisor (255 for conv real data, else 1)}.
 * Packed layout: variable d occupies columns [var_dcol[d], var_dcol[d] + nclass[d]) of data and
 * [var_pcol[d], ...) of theta, and var_dcol / var_pcol increase with d (read_functions.py:144-173
 * builds exactly this); a CTA stages the contiguous span of a tile of <= 128 variables.
 *   theta [N,P_theta] and every output: storage `dtype` (HLVAE_F32 / HLVAE_F64), which also selects
 *   the arithmetic: float64 storage -> float64 math; float32 storage -> float32 SFU math with exact
 *   float64 re-evaluation of categorical / ordinal argmax decisions that float32 cannot prove.
 *   data [N,E_x]: `data_dtype` = dtype or HLVAE_U8;  mask [N,D]: `mask_dtype` = dtype or HLVAE_U8.
 *   max_class = max_d nclass[d] (sizes the shared-memory staging of a row batch).
 * fwd outputs (nullable): log_p_x, log_p_x_missing [N,D], params [N,P_theta], recon_mean,
 *   recon_mode, data_tr [N,D]; ll_total[1] float64 += sum(log_p_x)  (HLVAE.py:377-379 summed).
 * bwd: upstream gradient of log_p_x = g_lp[n,d] (nullable) + *g_scalar (device scalar, nullable:
 *   the gradient of ll_total) -> g_theta [N,P_theta] (overwritten), g_lvy [D] float64 (accumulated).
 * ---------------------------------------------------------------------------------- */
#define HLVAE_VAR_REAL 0
#define HLVAE_VAR_POS 1
#define HLVAE_VAR_COUNT 2
#define HLVAE_VAR_CAT 3
#define HLVAE_VAR_ORDINAL 4
#define HLVAE_MAX_CLASS 16

int hlvae_loglik_fwd(int64_t N, int D, int64_t ld_data, int64_t ld_theta,
                     const int32_t* var_kind, const int32_t* var_nclass, const int32_t* var_dcol,
                     const int32_t* var_pcol, const double* vparam,
                     const void* data, const void* theta, const void* mask, int dtype, int data_dtype,
                     int mask_dtype, int max_class, void* log_p_x, void* log_p_x_missing, void* params,
                     void* recon_mean, void* recon_mode, void* data_tr, double* ll_total, void* stream);
int hlvae_loglik_bwd(int64_t N, int D, int64_t ld_data, int64_t ld_theta,
                     const int32_t* var_kind, const int32_t* var_nclass, const int32_t* var_dcol,
                     const int32_t* var_pcol, const double* vparam,
                     const void* data, const void* theta, const void* mask, int dtype, int data_dtype,
                     int mask_dtype, int max_class, const void* g_lp, const double* g_scalar, void* g_theta,
                     double* g_lvy, void* stream);

/* Stand-alone forms of the monitoring transforms for callers that hold only `params` or
 * `data` (training.py:84-91): read_functions.statistics (:268-302) -> mean, mode [N,D];
 * read_functions.discrete_variables_transformation (:221-235) -> out [N,D]. */
int hlvae_statistics(int64_t N, int D, int64_t ld_theta, const int32_t* var_kind, const int32_t* var_nclass,
                     const int32_t* var_pcol, const double* vparam, const void* params, int dtype,
                     void* mean, void* mode, void* stream);
int hlvae_discrete_transform(int64_t N, int D, int64_t ld_data, const int32_t* var_kind, const int32_t* var_nclass,
                             const int32_t* var_dcol, const void* data, int dtype, void* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Likelihood branches outside the fused five-type kernel, one type group [N, Dg] per launch:
 *   HLVAE_AUX_REAL : loglik_real with the variance network (HL_VAE/loglik.py:45-48) - th_a = means, th_b = per-row
 *                    raw log-variances;  HLVAE_AUX_POS : loglik_pos likewise (:89,104-108);
 *   HLVAE_AUX_BETA : loglik_beta (:216-256) - th_a = the probit-mean parameter, `disp` = raw dispersion (1 value).
 * th_a is addressed as th_a[n * ld_a + d * cs_a] (cs_a = 0: every variable of the group reads column 0, which is
 * what the reference's fallback indexing at loglik.py:232-235 does); th_b, data, mask as [n * ld + d].
 * vparam [4, Dg] float64: REAL / POS {normalisation mean, clamped normalisation variance, -, data divisor};
 * BETA {data_min, data_max, -, -}.  Outputs [N, Dg] contiguous in the storage `dtype`: log_p_x, log_p_x_missing,
 * prm_a / prm_b = (mean, variance) or (alpha, beta).
 * bwd: upstream g_lpx [N, Dg] (nullable) and / or a device scalar g_scalar (gradient of sum(log_p_x)) ->
 * g_a, g_b [N, Dg] element-wise gradients w.r.t. th_a, th_b (BETA: g_b unused, g_disp += d/d disp, caller zero-fills).
 * ---------------------------------------------------------------------------------- */
#define HLVAE_AUX_REAL 0
#define HLVAE_AUX_POS 1
#define HLVAE_AUX_BETA 2
int hlvae_loglik_aux_fwd(int mode, int64_t N, int Dg, const void* data, int64_t ld_data, int data_dtype,
                         const void* mask, int64_t ld_mask, int mask_dtype, const void* th_a, int64_t ld_a,
                         int64_t cs_a, const void* th_b, int64_t ld_b, int dtype, const double* vparam,
                         const double* disp, void* lpx, void* lpm, void* prm_a, void* prm_b, void* stream);
int hlvae_loglik_aux_bwd(int mode, int64_t N, int Dg, const void* data, int64_t ld_data, int data_dtype,
                         const void* mask, int64_t ld_mask, int mask_dtype, const void* th_a, int64_t ld_a,
                         int64_t cs_a, const void* th_b, int64_t ld_b, int dtype, const double* vparam,
                         const double* disp, const void* g_lpx, const double* g_scalar, void* g_a, void* g_b,
                         double* g_disp, void* stream);

/* ------------------------------------------------------------------------------------
 * Per-variable observation heads.  Replaces HLVAE.theta_estimation
 * (HLVAE.py:416-453) over the Observation_Count / _Real_Pos_Beta / _Cat / _Ordinal modules
 * (HLVAE.py:11-89) and the real-valued Sigmoid layer of the convolutional model (:292-295), logvar_network=False.
 * The reference evaluates each head on y * mask and, under no_grad, on y * (1 - mask), and merges the two
 * through boolean indexing; for 0/1 masks that is, per theta column p of variable d = col_var[p],
 *     theta[n, p] = act_p(bias[p] + sum_k weight[p, k] * y[n, d, k])
 * with gradients flowing through observed entries (mask[n, d] != 0) only.  col_mode[p]:
 *   HLVAE_HEAD_AFFINE   count / pos / real mean, categorical logits 1..C-1, ordinal region (:22,51,65,87)
 *   HLVAE_HEAD_SIGMOID  real mean of the convolutional model (affine, then Sigmoid)
 *   HLVAE_HEAD_ZERO     first categorical logit, fixed to 0 (:66-67)
 *   HLVAE_HEAD_BIAS     ordinal thresholds: bias[p] = weight_thresholds, independent of y (:85)
 * Packed layout as in hlvae_loglik_*: var_pcol[d] = first theta column of variable d, var_pcol[D] = P;
 * tile_var[0..n_tiles] = variable boundaries of tiles covering <= 256 theta columns and <= 256 variables each.
 *   y [N, D, Y] addressed with ELEMENT strides (sn, sd, sk): the convolutional decoder's y_grouped is a
 *   permuted view of [N, Y, D] (HLVAE.py:341-342, sd = 1, sk = D), the dense one is contiguous (:344).
 *   weight [P, Y], bias [P] float64 (rows of constant columns are ignored).  Y <= HLVAE_MAX_Y.
 *   y, theta, g_theta, g_y: storage `dtype` (HLVAE_F32 / HLVAE_F64), arithmetic in that type;
 *   mask [N, D] contiguous, `mask_dtype` = dtype or HLVAE_U8.
 * bwd: g_theta [N, ld_theta] -> g_y (same strides as y, overwritten), g_weight [P, Y], g_bias [P]
 *   float64 (accumulated, caller zero-fills).
 *   max_cols = the largest number of theta columns one variable owns (max_d var_pcol[d+1] - var_pcol[d]), or 0 when
 *   the caller does not know it: layouts with max_cols <= 5 and Y <= 8 take a thread-per-variable kernel (the
 *   variable's heads and gradient sums in registers, inputs of the next row batch in flight by cp.async), every other
 *   call the thread-per-column kernel.  max_cols below the true maximum is a caller error: d/dy of the wider variables
 *   comes back NaN.
 * ---------------------------------------------------------------------------------- */
#define HLVAE_MAX_Y 16
#define HLVAE_HEAD_AFFINE 0
#define HLVAE_HEAD_SIGMOID 1
#define HLVAE_HEAD_ZERO 2
#define HLVAE_HEAD_BIAS 3
int hlvae_theta_fwd(int64_t N, int D, int P, int Y, int n_tiles, const int32_t* col_var, const int32_t* col_mode,
                    const int32_t* var_pcol, const int32_t* tile_var, const double* weight, const double* bias,
                    const void* y, int64_t sn, int64_t sd, int64_t sk, int dtype, void* theta, int64_t ld_theta,
                    void* stream);
int hlvae_theta_bwd(int64_t N, int D, int P, int Y, int n_tiles, int max_cols, const int32_t* col_var,
                    const int32_t* col_mode, const int32_t* var_pcol, const int32_t* tile_var, const double* weight,
                    const double* bias, const void* y, int64_t sn, int64_t sd, int64_t sk, int dtype, const void* mask,
                    int mask_dtype, const void* g_theta, int64_t ld_theta, void* g_y, double* g_weight, double* g_bias,
                    void* stream);

/* ------------------------------------------------------------------------------------
 * Batch normalisation of the data batch.  Replaces HL_VAE/utils.py:88-143
 * `batch_normalization`: the encoder input X_list [N, E_x] and the normalisation parameters
 * [mean, var] of the real (non-convolutional) and positive variables that loglik_real / loglik_pos take.
 *   real, conv: observed / 255 (:99-103); real: (observed - mean) / sqrt(var + 1e-5) * mask with the masked
 *   two-pass mean / variance (:104-108); count: log(observed), 0 where missing (:113-119); pos: the same
 *   standardisation on log(1 + observed) (:120-131; the caller clamps the variance, :127); cat / ordinal /
 *   other: data * mask (:132-140).
 * hlvae_batch_norm_stats: pass 0 accumulates stats[0][d] += sum_n mask, stats[1][d] += sum_n value * mask;
 *   pass 1 (after pass 0 of ALL rows) stats[2][d] += sum_n ((value - mean) * mask)^2, for the variables listed
 *   in stat_vars[n_stat].  stats [3, D] float64, caller zero-fills.
 * hlvae_batch_norm_apply: out [N, E_x] (storage `dtype`) from data, mask and meanvar [2, D] float64
 *   (mean; variance as used at :108 / :128, i.e. already clamped for positive variables);
 *   dcol_var [E_x] = variable of each data column.
 * data [N, E_x]: `data_dtype` = dtype or HLVAE_U8; mask [N, D]: `mask_dtype` = dtype or HLVAE_U8.
 * ---------------------------------------------------------------------------------- */
int hlvae_batch_norm_stats(int64_t N, int D, int64_t ld_data, const int32_t* var_kind, const int32_t* var_dcol,
                           const int32_t* stat_vars, int n_stat, const void* data, const void* mask, int dtype,
                           int data_dtype, int mask_dtype, int pass, double* stats, void* stream);
int hlvae_batch_norm_apply(int64_t N, int D, int64_t ld_data, const int32_t* var_kind, const int32_t* dcol_var,
                           const void* data, const void* mask, int dtype, int data_dtype, int mask_dtype, int conv,
                           const double* meanvar, void* out, void* stream);

/* ------------------------------------------------------------------------------------
 * A/B probe (not on the product path): S[l] += K[l]^T V[l] for K, V [L, N, 64] float64 row-major, S [L, 64, 64],
 * the sufficient-statistics contraction of elbo_functions.py:161 / :254,266 on the two tensor pipes.
 *   mode 0: FP64 mma.sync (the form hlvae_kl_panel uses);
 *   mode 1: tcgen05.mma kind::i8 into TMEM with an error-free 8-bit splitting of the operands (`nslice` slices of
 *           K / k_scale in [0, 1) and V / v_scale in [-1, 1); k_scale, v_scale: powers of two bounding the data),
 *           drained with tcgen05.ld and recombined in float64.  rows_per_cta <= 8192 (int32 accumulators).
 * status: device words {code, l, cta}; code 3 = a tensor-core completion barrier timed out. */
int hlvae_contraction_probe(int mode, int nslice, int L, int64_t N, int M, const double* K, const double* V,
                            double k_scale, double v_scale, int rows_per_cta, double* S, int32_t* status,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HLVAE_B200_H */
