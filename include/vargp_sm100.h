/*
 * libvargp_sm100.so -- C ABI of the B200-native VAR-GP ELBO hot path.
 *
 * The reference (uber-research/vargp) has no FFI: its hot path is ATen calls issued from
 * var_gp/{kernels,gp_utils,likelihoods,vargp}.py.  Each entry point below replaces a group of those call
 * sites (cited as file:line relative to the reference root); INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add to call it.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless stated otherwise; sizes and strides are int64, in
 *     ELEMENTS; `stream` is a cudaStream_t passed as void*.
 *   - the library never allocates, frees or retains CALLER memory (its one own device allocation is an 8 KB pool of tile
 *     counters for the persistent GEMM, made by vargp_init); all work is asynchronous on `stream`.
 *   - return value: 0 ok, <0 invalid argument (see vargp_strerror), >0 a cudaError_t.
 *   - H = hyper samples, C = classes (output GPs), P = inducing points per class over all tasks,
 *     M = inducing points per task, S = P / M tasks, B = minibatch, D = input dims, F = likelihood samples.
 *   - theta is (H, D+1): log lengthscales then log scale factor, exactly `kern_samples` of
 *     var_gp/kernels.py:24.
 */
#ifndef VARGP_SM100_H_
#define VARGP_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VARGP_TRI_NONE 0
#define VARGP_TRI_LOWER 1
#define VARGP_TRI_UPPER 2

#define VARGP_EPI_NONE 0
#define VARGP_EPI_RBF 1      /* C = gamma2 * exp(acc - row[m]/2 - col[n]/2)                  */
#define VARGP_EPI_RBF_SYM 2  /* same, and C[m][m] = gamma2 exactly (var_gp/kernels.py:47-48,54) */

/* Batched strided GEMM  C = alpha * tri_a(A) * tri_b(B) [restricted to tri_c] + beta * C.
 * A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = B[k*b_rs + n*b_cs], C(m,n) = C[m*c_rs + n*c_cs].
 * Up to three batch dimensions (nb[0] slowest); a stride of 0 broadcasts an operand.
 * tri_a / tri_b declare structural zeros (the kernel never reads the other triangle);
 * tri_c writes only that triangle (the other is zero-filled when beta == 0, left untouched otherwise).
 * Replaces the einsum / bmm / triangular_solve call sites of var_gp/gp_utils.py:89-96,124-136,175-184
 * (every TRSM there becomes a GEMM with the explicit inverse factor W = chol(K)^-1). */
typedef struct {
  const float* A;
  const float* B;
  float* C;
  int64_t M, N, K;
  int64_t a_rs, a_cs, b_rs, b_cs, c_rs, c_cs;
  int64_t nb[3];
  int64_t a_bs[3], b_bs[3], c_bs[3];
  float alpha, beta;
  int32_t tri_a, tri_b, tri_c;
  /* optional fused epilogue (VARGP_EPI_*): row/col vectors are indexed like C's rows/cols */
  int32_t epi;
  const float* e_row;   /* (.., M) */
  const float* e_col;   /* (.., N) */
  int64_t e_row_bs[3], e_col_bs[3];
  const float* e_theta; /* gamma2 = exp(2 * e_theta[batch offset via e_theta_bs + e_D]) */
  int64_t e_theta_bs[3], e_D;
  /* scheduling hint: a product that runs BESIDE a critical chain (side stream) may be told to occupy at most this many
   * SMs (0 = all).  Honoured by the persistent 2-CTA kernel, whose CTAs would otherwise own every SM until it is done.
   * Negative: keep the 1-CTA kernel in its one-tile-per-CTA form (no persistent tile loop), so that launches of a
   * higher-priority stream can cut in between its waves. */
  int64_t sm_limit;
} vargp_gemm_t;

int vargp_init(int device);
const char* vargp_version(void);
const char* vargp_strerror(int code);
/* programmatic dependent launch between the library's kernels: 0 off, 1 on every launch, 2 (default) only on launches
 * with a small shared-memory footprint (a parked tensor-core GEMM CTA would take its SM away from the grid it waits
 * for).  Also settable through VARGP_PDL in the environment of vargp_init.  bench.py switches it off for its serialised
 * per-kernel timing pass.  Returns the previous setting. */
int vargp_set_pdl(int mode);
/* number of kernel launches issued through this library since load (for bench.py's gpu_launches) */
int64_t vargp_launch_count(void);

int vargp_gemm(const vargp_gemm_t* g, void* stream);

/* tcgen05 / TMA path (3xTF32): same contract as vargp_gemm restricted to K-contiguous operands
 * (a_cs == 1, b_rs == 1), 16-byte aligned rows; returns -2 if the problem does not qualify. */
int vargp_gemm_tc(const vargp_gemm_t* g, void* stream);
/* Launches of the 1-CTA kernel whose C goes out through the TMA engine may take its PERSISTENT form (one CTA per SM working
 * through the tiles handed out by a launch-wide counter, longest tiles first; gemm_tcp.cu).  mode: 0 off, 1 launches of
 * more than one wave of tiles (default), 2 every such launch, 3 from four waves; negative only queries.  Returns the
 * previous mode.  A descriptor with sm_limit < 0 always keeps the one-tile-per-CTA form. */
int64_t vargp_tc_persist_config(int64_t mode);
/* vargp_gemm_tc hands problems with at least `min_tiles` 256 x 256 output tiles to the persistent 2-CTA
 * (cta_group::2) kernel; < 0 disables it, INT64_MIN only queries.  Returns the previous setting.
 * vargp_tc2_launch_count: launches of that kernel since load. */
int64_t vargp_tc2_config(int64_t min_tiles);
int64_t vargp_tc2_launch_count(void);
/* profiling aid: CTA (0,0,0) of every following 1-CTA vargp_gemm_tc launch writes 8 clock64() stamps of its pipeline
 * (entry, setup, first slab landed, first slab issued, first partial sum, MMAs retired, stored, exit) to `buf`
 * (device memory, >= 8 int64); NULL switches it off. */
void vargp_tc_debug(long long* buf);

/* dst[h][r][:] = src[r][:] * exp(-theta[h][:D]);  norms[h][r] = |dst[h][r]|^2.
 * Replaces the `x / sigma` broadcasts and the Gram diagonals of var_gp/kernels.py:41-44,50-51,54. */
int vargp_scale_rows(const float* src, int64_t R, int64_t D, int64_t src_rs,
                     const float* theta, int64_t H, int64_t theta_rs,
                     float* dst, float* norms, void* stream);

/* L = chol(A + jitter*I) (lower; strict upper zeroed), batched; info[b] = 0 or 1 + index of the first
 * non-positive pivot.  A and L are (batch, n, n) with leading dimension ld and batch stride bs; may alias.
 * Replaces var_gp/gp_utils.py:5-11. */
int vargp_chol(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
               int64_t n, int64_t batch, float jitter, int32_t* info, void* stream);

/* Building block of the blocked (GEMM-driven) factorisation of large matrices: same as vargp_chol on one diagonal
 * block, but a failing pivot is reported as info_base + its 1-based index and, with accumulate != 0, only if
 * info[b] is still 0 (the first failure of the whole matrix wins). */
int vargp_chol_ex(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                  int64_t n, int64_t batch, float jitter, int32_t* info, int64_t info_base, int accumulate,
                  void* stream);

/* W = L^-1 (lower; strict upper zeroed), batched.  Stands in for every torch.triangular_solve of
 * var_gp/gp_utils.py (89,92,124,125,129,134,175,176,182). */
int vargp_trtri(const float* L, int64_t l_ld, int64_t l_bs, float* W, int64_t w_ld, int64_t w_bs,
                int64_t n, int64_t batch, void* stream);

/* L = chol(A + jitter*I) and W = L^-1 in one call (both lower, strict upper zeroed).  Matrices of at least
 * `min_n` rows (vargp_chol_config) are factored block-wise: nb x nb diagonal blocks by the vargp_chol / vargp_trtri
 * kernels, the trailing updates, panel solves and the block merges of the inverse as batched 3xTF32 tensor-core
 * GEMMs (potrf_blocked.cu); W and the strict upper triangle of L serve as scratch, so no workspace is needed.
 * A, L, W must not alias.  Replaces var_gp/gp_utils.py:5-11 plus the solves of :89-92,124-134,175-182. */
int vargp_chol_inv(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                   float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                   int32_t* info, void* stream);
/* Shared-memory-resident variant for n <= 128 (one CTA per matrix; potrf_small.cu): the whole factorisation of the
 * first tasks and the diagonal-block step of vargp_chol_inv.  A may alias W.  Pivot reporting as vargp_chol_ex. */
int vargp_chol_inv_small(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                         float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                         int32_t* info, int64_t info_base, int accumulate, void* stream);
/* Cluster-cooperative variant for 32 < n <= 320 (one thread-block cluster of 2 or 4 CTAs per matrix, block rows dealt
 * round-robin to the CTAs' shared memory, panels and rows of the inverse pushed through distributed shared memory, the
 * inverse formed during the factorisation sweep; potrf_cluster.cu): what vargp_chol_inv takes by default at the
 * Split-MNIST sizes (P = 60 ... 300).  A may alias L or W; L and W must differ.  Replaces var_gp/gp_utils.py:5-11 plus the
 * solves of :89-92,124-134,175-182 in ONE launch.
 * vargp_chol_cluster_config: routing window [min_n, max_n] of vargp_chol_inv (0, 0 disables; negative only queries),
 * returns the previous (min_n << 32) | max_n.  vargp_chol_cluster_wants: 1 if vargp_chol_inv would route n there. */
int vargp_chol_inv_cluster(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                           float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                           int32_t* info, void* stream);
/* same with the pivot reporting of vargp_chol_ex (info_base / accumulate): the diagonal-block step of the blocked
 * factorisation for block sizes 129 ... 320 (vargp_chol_config block = 256). */
int vargp_chol_inv_cluster_ex(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                              float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                              int32_t* info, int64_t info_base, int accumulate, void* stream);
int64_t vargp_chol_cluster_config(int64_t min_n, int64_t max_n);
int vargp_chol_cluster_wants(int64_t n);
/* profiling aid: clock64 stamps of the phases of matrix 0 (per CTA rank and block step) into a device buffer of
 * 4 * 16 * 16 int64 (scripts/chol_stamps.py); NULL turns it off. */
void vargp_chol_cluster_debug(void* buf);
/* block size (multiple of 32; 0 keeps it; 1 = automatic: 256 with the diagonal blocks on vargp_chol_inv_cluster_ex, 128 on
 * vargp_chol_inv_small when the cluster kernel is disabled) and minimum n (< 0 keeps it) of the blocked path;
 * returns (min_n << 32) | block after the update (block = 1 while automatic). */
int64_t vargp_chol_config(int64_t block, int64_t min_n);

/* packed row-major lower triangle -> (C, M, M) with softplus on the diagonal, and its adjoint.
 * Replaces var_gp/gp_utils.py:22-49. */
int vargp_tril_unpack(const float* vec, int64_t C, int64_t M, float* out, void* stream);
int vargp_tril_unpack_bwd(const float* Lbar, const float* vec, int64_t C, int64_t M, float* vec_bar, void* stream);

/* kl += (1/H) sum_hc [ -sum_{i in last block} log W_ii - sum_i log Lu_ii + (|T_last|_F^2 + |nu_last|^2 - M)/2 ]
 * W (H,C,P,P), T (H,C,S,M,M), nu (H,C,P), Lu (C,M,M).  Replaces var_gp/vargp.py:182-190. */
#define VARGP_KL_CHUNKS 16   /* partial sums per (h, c): work holds H*C*VARGP_KL_CHUNKS floats */
int vargp_kl_fwd(const float* W, const float* T, const float* nu, const float* Lu,
                 int64_t H, int64_t C, int64_t P, int64_t M, float* kl, float* work, void* stream);
int vargp_kl_bwd(const float* W, const float* T, const float* nu, const float* g_kl,
                 int64_t H, int64_t C, int64_t P, int64_t M, float* Wbar, float* Tbar, float* nubar, void* stream);
int vargp_kl_bwd_lu(const float* Lu, const float* g_kl, int64_t C, int64_t M, float* Lubar, void* stream);

/* Whitening of the variational parameters on the M x M diagonal task blocks, forward and adjoint, one shared-memory
 * kernel each (whiten.cu) for the (h, c) pairs of the rectangle [h0, h1) x [c0, c1):
 *   forward:  T_s = W_ss Lu_s, nu_s = W_ss m_s, N_ss += T_s T_s^T (N must already hold jitter W W^T) and, with kl != NULL,
 *             kl += (1/H) sum_hc KL_hc (same value as vargp_kl_fwd; deterministic).  `work`: vargp_whiten_fwd_work(H, C)
 *             floats, work[0] == 0 at launch and again at exit.
 *   adjoint:  k = g_kl / H on the last block (g_kl may be NULL), else 0;  Tbar_s = tril(2 G_ss T_s) + k T_s;
 *             nubar'_s = nubar_s + k nu_s;  Wbar_ss += tril(Tbar_s Lu_s^T + nubar'_s m_s^T) - k diag(1 / W_ii);  for the
 *             blocks s >= s_grad0:  Lubar[h][s - s_grad0][c] = tril(W_ss^T Tbar_s),  mbar[h][s - s_grad0][c] = W_ss^T nubar'_s.
 * W, N, G, Wbar (H, C, P, P); T (H, C, S, M, M); nu, nubar (H, C, P); Lu_all (S, C, M, M); m_all (S, C, M);
 * Lubar (H, S - s_grad0, C, M, M); mbar (H, S - s_grad0, C, M).  Returns -2 for M > vargp_whiten_max_m(adjoint).
 * Replaces the block-diagonal part of var_gp/vargp.py:35-88 (compute_q) and the KL of vargp.py:182-190. */
int64_t vargp_whiten_fwd_work(int64_t H, int64_t C);
int64_t vargp_whiten_max_m(int adjoint);
int vargp_whiten_fwd(const float* W, const float* Lu_all, const float* m_all, int64_t H, int64_t C, int64_t S, int64_t M,
                     int64_t P, int64_t h0, int64_t h1, int64_t c0, int64_t c1, float* T, float* nu, float* N, float* kl,
                     float* work, void* stream);
int vargp_whiten_bwd(const float* W, const float* T, const float* nu, const float* Lu_all, const float* m_all,
                     const float* G, const float* nubar, const float* g_kl, int64_t H, int64_t C, int64_t S, int64_t M,
                     int64_t P, int64_t h0, int64_t h1, int64_t c0, int64_t c1, int64_t s_grad0, float* Wbar, float* Lubar,
                     float* mbar, void* stream);

/* Predictive marginal (var_gp/gp_utils.py:178-186) from V = W Kzx and NV = N V, where
 * N = blockdiag(T_s T_s^T) + jitter W W^T (P x P, symmetric) collects everything quadratic in V:
 *   f_mean[g][b] = sum_p nu[g][p] V[g][p][b];
 *   f_var[g][b]  = gamma2 + sum_p V (NV - V)   ( = gamma2 - |V_b|^2 + sum_s |T_s^T V_sb|^2 + jitter |W^T V_b|^2 ). */
int vargp_marginal_reduce(const float* V, const float* NV, const float* nu,
                          const float* theta, int64_t theta_rs, int64_t D,
                          int64_t H, int64_t C, int64_t P, int64_t B,
                          float* f_mean, float* f_var, void* stream);
/* adjoint prologue: Vbar = nu gm^T + 2 gv (NV - V) (Vbar may alias NV);  Vg = gv V (so that Nbar = Vg V^T);
 * theta_bar[h][D] += 2 gamma2 sum_cb gv */
int vargp_marginal_bwd_prep(const float* V, const float* NV, const float* nu,
                            const float* g_mean, const float* g_var,
                            const float* theta, int64_t theta_rs, int64_t D,
                            int64_t H, int64_t C, int64_t P, int64_t B,
                            float* Vbar, float* Vg, float* theta_bar, void* stream);

/* batch of n x n, in place, only the lower triangle of X is read.
 * mirror == 0: X <- (Phi(X) + Phi(X)^T)/2 (Cholesky adjoint);  mirror != 0: X_ji <- X_ij (symmetric completion) */
int vargp_sym_phi(float* X, int64_t n, int64_t batch, int mirror, void* stream);

/* Kbar *= K (elementwise, in place); rsum[g][i] = row sums; csum[h][j] += sum over (c,i) (optional).
 * Kbar, K are (H, C, Pa, Pb).  With dsum (symmetric Gram, Pa == Pb) the diagonal products are moved to
 * dsum[g][i] and zeroed in Kbar / rsum: K_ii = gamma^2 does not depend on z or sigma, so keeping it out of
 * the z / sigma adjoint avoids an analytic cancellation in fp32. */
int vargp_rbf_bwd_prep(float* Kbar, const float* K, int64_t H, int64_t C, int64_t Pa, int64_t Pb,
                       float* rsum, float* csum, float* dsum, void* stream);
/* zs_bar = -(r1 + 2 r2) zs + Gz1 + 2 Gz2;  Zbar[c][i][d] = sum_h zs_bar * exp(-theta[h][d]);
 * theta_bar[h][d] += sum_ci (-zs zs_bar - zs Gz1);  theta_bar[h][D] += 2 sum_ci (r1 + r2 + dg).  (Gz1, r1), (Gz2, r2) may each be NULL pairs;
 * dg (the dsum of vargp_rbf_bwd_prep) may be NULL. */
int vargp_rbf_bwd_finish(const float* zs, const float* Gz1, const float* Gz2, const float* r1, const float* r2,
                         const float* dg, const float* theta, int64_t theta_rs, int64_t H, int64_t C, int64_t P, int64_t D,
                         float* Zbar, float* theta_bar, void* stream);
/* theta_bar[h][d] += sum_j csum[h][j] xs[h][j][d]^2;  optionally
 * xbar[j][d] = sum_h (-csum xs + sum_c Gx[h][c][j][d]) exp(-theta[h][d])   (Gx, xbar may be NULL) */
int vargp_rbf_bwd_xside(const float* xs, const float* csum, const float* Gx,
                        const float* theta, int64_t theta_rs, int64_t H, int64_t C, int64_t B, int64_t D,
                        float* theta_bar, float* xbar, void* stream);

/* Monte-Carlo softmax likelihood (var_gp/likelihoods.py:13-47), forward and adjoint in one pass:
 * nll += -(1/HF) sum_hfb log softmax_C(f_mean + sqrt(f_var) eps)[y_b];  g_mean, g_var = gscale * d nll / d(f_mean, f_var).
 * eps (H,F,C,B); y int64 (B).  The sum is deterministic (per-CTA partial sums combined in a fixed order, no float
 * atomics): `work` holds vargp_softmax_nll_work(H, B) floats; work[0] must be 0 at launch and is 0 again at exit. */
int64_t vargp_softmax_nll_work(int64_t H, int64_t B);
int vargp_softmax_nll(const float* f_mean, const float* f_var, const float* eps, const int64_t* y,
                      int64_t H, int64_t F, int64_t C, int64_t B,
                      float* nll, float* g_mean, float* g_var, float gscale, float* work, void* stream);
/* probs[b][c] = (1/HF) sum_hf softmax_C(f)[c]   (var_gp/likelihoods.py:49-63) */
int vargp_softmax_predict(const float* f_mean, const float* f_var, const float* eps,
                          int64_t H, int64_t F, int64_t C, int64_t B, float* probs, void* stream);

/* Kernel hyper-parameters (var_gp/kernels.py:62-77), H samples of D1 = D + 1 log-hyper-parameters:
 *   theta[h][d] = log_mean[d] + exp(log_logvar[d] / 2) * eps[h][d]                 (Normal.rsample)
 *   kl[0] = sum_d KL( N(log_mean, exp(log_logvar)) || N(prior_log_mean, exp(prior_log_logvar)) )   (kl may be NULL)
 * and the adjoint given theta_bar (H, D1) and the scalar g_kl on the device (either may be NULL). */
int vargp_hyper_fwd(const float* log_mean, const float* log_logvar, const float* prior_log_mean,
                    const float* prior_log_logvar, const float* eps, int64_t H, int64_t D1, float* theta, float* kl,
                    void* stream);
int vargp_hyper_bwd(const float* log_mean, const float* log_logvar, const float* prior_log_mean,
                    const float* prior_log_logvar, const float* eps, const float* theta_bar, const float* g_kl,
                    int64_t H, int64_t D1, float* log_mean_bar, float* log_logvar_bar, void* stream);

/* Prologue of the fused training step: the current task's parameters into the stacked operands of the step --
 * Zcat[c][P-M+i][:] = z[c][i][:] (Zcat is (C, P, D)), m_last (C, M) = u_mean, Lu_last (C, M, M) = vec2tril(u_tril_vec)
 * with a softplus diagonal.  Replaces the torch.cat / vec2tril calls of var_gp/vargp.py:52-59,151-152 (one launch). */
int vargp_step_assemble(const float* z, const float* u_mean, const float* u_tril_vec, int64_t C, int64_t M, int64_t D,
                        int64_t P, float* Zcat, float* m_last, float* Lu_last, void* stream);
/* Epilogue of the fused training step: the adjoints of the stacked operands back into parameter gradients, one launch --
 * z_grad (C, M, D) = Zbar[:, P-M:, :];  u_mean_grad (C, M) = sum_h mbar[h];  u_tril_vec_grad = adjoint of vec2tril applied
 * to sum_h Lubar[h] - diag(g_kl_u / Lu_ii) (g_kl_u may be NULL);  log_mean_grad, log_logvar_grad as vargp_hyper_bwd.
 * mbar (H, C, M) and Lubar (H, C, M, M) are addressed with the hyper-sample strides mbar_hs / Lubar_hs (elements). */
int vargp_step_grad_finish(const float* Zbar, const float* mbar, int64_t mbar_hs, const float* Lubar, int64_t Lubar_hs,
                           const float* Lu, const float* u_tril_vec, const float* g_kl_u,
                           const float* log_mean, const float* log_logvar, const float* prior_log_mean,
                           const float* prior_log_logvar, const float* eps, const float* theta_bar, const float* g_kl_h,
                           int64_t H, int64_t C, int64_t M, int64_t D, int64_t P,
                           float* z_grad, float* u_mean_grad, float* u_tril_vec_grad, float* log_mean_grad,
                           float* log_logvar_grad, void* stream);

/* Fused Yogi step over a flat parameter buffer (the optimizer the reference trains with,
 * experiments/vargp.py:23): m <- b1 m + (1-b1) g;  v <- v - (1-b2) sign(v - g^2) g^2;
 * p <- p - lr/(1-b1^t) * m / (sqrt(v/(1-b2^t)) + eps).  pows = {b1^t, b2^t} lives on the device and is
 * advanced in-stream (initialise to {1, 1}), so the step can be replayed from a CUDA graph. */
int vargp_yogi_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2,
                    float eps, float* pows, void* stream);

/* Data-parallel tail of the step as one compute + collective pair over NVLink peer memory (peer.cu): gradient
 * all-reduce (one-shot: every rank reads all ranks' staged gradients in rank order, so the sum is bit-identical
 * everywhere) fused with the Yogi update of vargp_yogi_step.  peer_bufs: HOST array of `world` DEVICE pointers, entry r =
 * rank r's buffer of vargp_peer_buffer_floats(n) floats in symmetric (peer-mapped) memory, zero-filled once before the
 * first step (the host side allocates and exchanges them, e.g. torch.distributed._symmetric_memory); ctr: 4 zero-initialised
 * device words owned by this rank.  n must be a multiple of 4 and flat_g / p / m / v 16-byte aligned.  On return (stream
 * order) flat_g holds the summed gradient and p, m, v, pows are advanced.  One cross-GPU synchronisation per call; every
 * rank must make the same sequence of calls. */
int64_t vargp_peer_buffer_floats(int64_t n);
int vargp_peer_allreduce_yogi(float* const* peer_bufs, int world, int rank, int64_t n, float* flat_g, float* p, float* m,
                              float* v, float lr, float b1, float b2, float eps, float* pows, uint32_t* ctr, void* stream);
/* Same with the in-switch reduction (NVLS): `multicast` is the multicast mapping of the ranks' symmetric buffers
 * (torch.distributed._symmetric_memory: handle.multicast_ptr); the reduce kernel then issues ONE
 * multimem.ld_reduce per 16 bytes instead of a load from every peer.  NULL = the peer-load form above. */
int vargp_peer_allreduce_yogi_nvls(float* const* peer_bufs, const float* multicast, int world, int rank, int64_t n,
                                   float* flat_g, float* p, float* m, float* v, float lr, float b1, float b2, float eps,
                                   float* pows, uint32_t* ctr, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* VARGP_SM100_H_ */
