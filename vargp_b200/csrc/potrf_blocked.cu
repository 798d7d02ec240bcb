// Blocked, GEMM-driven batched Cholesky + triangular inverse for large matrices (P > a few hundred):
//     L = chol(A + jitter I),   W = L^-1          (var_gp/gp_utils.py:5-11 and every triangular_solve behind it)
//
// The one-CTA-per-matrix kernels of chol.cu leave 118 of 148 SMs idle and run the O(n^3) update on the SIMT pipes:
// 43 + 10 ms at n = 2048, batch 30.  Here only the nb x nb diagonal blocks go through those kernels; everything
// else is a batched GEMM on the tcgen05 3xTF32 kernels (gemm_tc.cu / gemm_tc2.cu), i.e. the "trailing SYRK" and
// the panel TRSMs of a classical blocked factorisation run on the tensor cores.
//
// No workspace: W doubles as the scratch of the factorisation, the strict upper triangle of L as the scratch of
// the inversion.
//   init   W <- tril(A) + jitter I, strict upper of W <- 0
//   for each block column k (left-looking):
//     (i)   W[k:, k] -= L[k:, :k] L[k, :k]^T                  GEMM  (n - k0) x nb x k0     "SYRK + GEMM" update
//     (ii)  L[k, k]   = chol(W[k, k])                         chol_kernel on the nb x nb block
//     (iii) W[k, k]   = L[k, k]^-1                            trtri kernels on the block
//     (iv)  L[k+1:, k] = W[k+1:, k] W[k, k]^T                 GEMM  (n - k0 - nb) x nb x nb   panel "TRSM"
//   inversion, bottom-up over block sizes s = nb, 2 nb, 4 nb, ...: for every pair of adjacent inverted s-blocks
//     tmp = L21 W11  (into the L12 scratch),   W21 = -W22 tmp           2 batched GEMMs per level
//   final  strict upper of L <- 0
// A ragged last block (n not a multiple of s) pairs up with the last full block of its level in a separate call.
#include <cstdlib>

#include "common.cuh"

extern "C" int vargp_gemm_tc(const vargp_gemm_t* g, void* stream);
extern "C" int vargp_chol_inv_small(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                                    float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                                    int32_t* info, int64_t info_base, int accumulate, void* stream);

extern "C" int vargp_chol_inv_cluster(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                                      float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                                      int32_t* info, void* stream);
extern "C" int vargp_chol_cluster_wants(int64_t n);
extern "C" int vargp_chol_inv_cluster_ex(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                                         float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                                         int32_t* info, int64_t info_base, int accumulate, void* stream);

namespace vargp {

constexpr int kInitRows = 8;       // rows per CTA of the two passes below (one row per CTA was launch-bound)

__global__ void __launch_bounds__(256)
potrf_init_kernel(const float* __restrict__ A, int64_t a_ld, int64_t a_bs, float* __restrict__ W, int64_t w_ld,
                  int64_t w_bs, int n, float jitter) {
  pdl_enter();
  const int64_t b = blockIdx.z;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const int i0 = blockIdx.y * kInitRows;
  float v[kInitRows];
#pragma unroll
  for (int r = 0; r < kInitRows; ++r) {
    const int i = i0 + r;
    v[r] = (i < n && j <= i) ? A[b * a_bs + (int64_t)i * a_ld + j] : 0.f;
    if (j == i) v[r] += jitter;
  }
#pragma unroll
  for (int r = 0; r < kInitRows; ++r)
    if (i0 + r < n) W[b * w_bs + (int64_t)(i0 + r) * w_ld + j] = v[r];
}

// zero the strict upper triangle outside the nb x nb diagonal blocks (those are zero-filled by the block kernels)
__global__ void __launch_bounds__(256)
zero_upper_kernel(float* __restrict__ L, int64_t ld, int64_t bs, int n, int nb) {
  pdl_enter();
  const int64_t b = blockIdx.z;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const int i0 = blockIdx.y * kInitRows;
#pragma unroll
  for (int r = 0; r < kInitRows; ++r) {
    const int i = i0 + r;
    if (i < n && j >= (i / nb + 1) * nb) L[b * bs + (int64_t)i * ld + j] = 0.f;
  }
}

static int g_blk_nb = 0;          // diagonal block size of the blocked factorisation; 0 = auto: 256 when the cluster kernel
                                  // (potrf_cluster.cu) takes the diagonal blocks, else 128 (potrf_small.cu).  Measured (B200,
                                  // batch 30, us): n=500: 455 -> 320; 1000: 1178 -> 920; 2048: 3642 -> 2785
static bool g_no_small = false;   // VARGP_CHOL_NO_SMALL=1: diagonal blocks through the chol.cu kernels (A/B timing)
static int g_blk_min_n = 129;     // matrices at least this large take the blocked path (n <= 128: potrf_small.cu)

static int gemm_any(vargp_gemm_t& g, cudaStream_t s) {
  int rc = vargp_gemm_tc(&g, s);
  if (rc == VARGP_ERR_UNSUPPORTED) rc = vargp_gemm(&g, s);
  return rc;
}

static vargp_gemm_t gemm_desc(int64_t batch) {
  vargp_gemm_t g = {};
  g.nb[0] = 1; g.nb[1] = batch; g.nb[2] = 1;
  g.alpha = 1.f; g.beta = 0.f;
  return g;
}

}  // namespace vargp

using namespace vargp;

extern "C" int64_t vargp_chol_config(int64_t block, int64_t min_n) {
  if (block == 1) g_blk_nb = 0;                                   // 1 = auto (reported as 1, so that a query can be restored)
  if (block >= 32 && block <= 1024 && block % 32 == 0) g_blk_nb = (int)block;
  if (min_n >= 0) g_blk_min_n = (int)min_n;
  return ((int64_t)g_blk_min_n << 32) | (int64_t)(g_blk_nb ? g_blk_nb : 1);
}

extern "C" int vargp_chol_inv(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                              float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                              int32_t* info, void* stream) {
  if (!A || !L || !W || !info || n < 1 || batch < 1 || a_ld < n || l_ld < n || w_ld < n) return VARGP_ERR_ARG;
  if (L == W || A == W) return VARGP_ERR_ARG;
  if (batch > 65535 || n > 65535) return VARGP_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  static bool env_read = false;
  if (!env_read) {
    env_read = true;
    const char* e = getenv("VARGP_CHOL_BLOCK");
    if (e) vargp_chol_config(atoll(e), -1);
    e = getenv("VARGP_CHOL_MIN_N");
    if (e) vargp_chol_config(0, atoll(e));
    e = getenv("VARGP_CHOL_NO_SMALL");
    if (e) g_no_small = atoi(e) != 0;
  }
  if (vargp_chol_cluster_wants(n))                          // one cluster of 2 / 4 CTAs per matrix (potrf_cluster.cu)
    return vargp_chol_inv_cluster(A, a_ld, a_bs, L, l_ld, l_bs, W, w_ld, w_bs, n, batch, jitter, info, stream);
  const int nb = g_blk_nb ? g_blk_nb : (vargp_chol_cluster_wants(256) ? 256 : 128);
  if (n <= 128 && !g_no_small)       // whole matrix fits the shared-memory kernel
    return vargp_chol_inv_small(A, a_ld, a_bs, L, l_ld, l_bs, W, w_ld, w_bs, n, batch, jitter, info, 0, 0, stream);
  if (n < g_blk_min_n || n <= nb) {
    int rc = vargp_chol(A, a_ld, a_bs, L, l_ld, l_bs, n, batch, jitter, info, stream);
    if (rc) return rc;
    return vargp_trtri(L, l_ld, l_bs, W, w_ld, w_bs, n, batch, stream);
  }

  launch_k(potrf_init_kernel, dim3(dim3((unsigned)ceil_div(n, 256), (unsigned)ceil_div(n, kInitRows), (unsigned)batch)), dim3(256), 0, s, 
      A, a_ld, a_bs, W, w_ld, w_bs, (int)n, jitter);
  int rc = launch_status();
  if (rc) return rc;

  // ---- factorisation ----
  for (int64_t k0 = 0; k0 < n; k0 += nb) {
    const int64_t kb = (n - k0 < nb) ? n - k0 : nb;
    float* Wkk = W + k0 * w_ld + k0;
    float* Lkk = L + k0 * l_ld + k0;
    if (k0 > 0) {                                   // (i)  W[k0:, k0:k0+kb] -= L[k0:, :k0] L[k0:k0+kb, :k0]^T
      vargp_gemm_t g = gemm_desc(batch);
      g.A = L + k0 * l_ld; g.a_rs = l_ld; g.a_cs = 1; g.a_bs[1] = l_bs;
      g.B = L + k0 * l_ld; g.b_rs = 1; g.b_cs = l_ld; g.b_bs[1] = l_bs;
      g.C = Wkk; g.c_rs = w_ld; g.c_cs = 1; g.c_bs[1] = w_bs;
      g.M = n - k0; g.N = kb; g.K = k0;
      g.alpha = -1.f; g.beta = 1.f;
      rc = gemm_any(g, s);
      if (rc) return rc;
    }
    // (ii) diagonal block factor (first failing pivot of the whole matrix wins) and (iii) its inverse, in place of
    //      the consumed block of W: one shared-memory kernel for blocks up to 128, else the one-CTA kernels
    if (kb > 128 && kb <= 320 && vargp_chol_cluster_wants(kb) && !g_no_small) {
      // one cluster per diagonal block (potrf_cluster.cu), in place of the consumed block of W
      rc = vargp_chol_inv_cluster_ex(Wkk, w_ld, w_bs, Lkk, l_ld, l_bs, Wkk, w_ld, w_bs, kb, batch, 0.f, info, k0,
                                     k0 > 0 ? 1 : 0, stream);
      if (rc) return rc;
    } else if (kb <= 128 && !g_no_small) {
      rc = vargp_chol_inv_small(Wkk, w_ld, w_bs, Lkk, l_ld, l_bs, Wkk, w_ld, w_bs, kb, batch, 0.f, info, k0,
                                k0 > 0 ? 1 : 0, stream);
      if (rc) return rc;
    } else {
      rc = vargp_chol_ex(Wkk, w_ld, w_bs, Lkk, l_ld, l_bs, kb, batch, 0.f, info, k0, k0 > 0 ? 1 : 0, stream);
      if (rc) return rc;
      rc = vargp_trtri(Lkk, l_ld, l_bs, Wkk, w_ld, w_bs, kb, batch, stream);
      if (rc) return rc;
    }
    if (k0 + kb < n) {                              // (iv) L[k0+kb:, k0:k0+kb] = W[k0+kb:, k0:k0+kb] Wkk^T
      vargp_gemm_t g = gemm_desc(batch);
      g.A = W + (k0 + kb) * w_ld + k0; g.a_rs = w_ld; g.a_cs = 1; g.a_bs[1] = w_bs;
      g.B = Wkk; g.b_rs = 1; g.b_cs = w_ld; g.b_bs[1] = w_bs;
      g.C = L + (k0 + kb) * l_ld + k0; g.c_rs = l_ld; g.c_cs = 1; g.c_bs[1] = l_bs;
      g.M = n - k0 - kb; g.N = kb; g.K = kb;
      rc = gemm_any(g, s);
      if (rc) return rc;
    }
  }

  // ---- inversion: merge adjacent inverted blocks, bottom-up ----
  for (int64_t sz = nb; sz < n; sz *= 2) {
    const int64_t nfull = n / sz, rem = n % sz;
    const int64_t npair = nfull / 2;
    // full pairs (batched over pairs), then the ragged pair (last full block + remainder), if any
    for (int pass = 0; pass < 2; ++pass) {
      int64_t r0, rows2, pairs;
      if (pass == 0) {
        if (npair == 0) continue;
        r0 = 0; rows2 = sz; pairs = npair;
      } else {
        if (!(nfull % 2 == 1 && rem > 0)) continue;
        r0 = (nfull - 1) * sz; rows2 = rem; pairs = 1;
      }
      const int64_t pstep_l = 2 * sz * (l_ld + 1), pstep_w = 2 * sz * (w_ld + 1);
      float* L21 = L + (r0 + sz) * l_ld + r0;
      float* L12 = L + r0 * l_ld + (r0 + sz);       // scratch: sz x rows2 region ... used as rows2 x sz? see below
      float* W11 = W + r0 * w_ld + r0;
      float* W22 = W + (r0 + sz) * w_ld + (r0 + sz);
      float* W21 = W + (r0 + sz) * w_ld + r0;
      // tmp (rows2 x sz) = L21 (rows2 x sz) W11 (sz x sz, lower).  The L12 region is sz x rows2: store tmp transposed.
      {
        vargp_gemm_t g = gemm_desc(batch);
        g.nb[2] = pairs;
        g.A = L21; g.a_rs = l_ld; g.a_cs = 1; g.a_bs[1] = l_bs; g.a_bs[2] = pstep_l;
        g.B = W11; g.b_rs = w_ld; g.b_cs = 1; g.b_bs[1] = w_bs; g.b_bs[2] = pstep_w;
        g.C = L12; g.c_rs = 1; g.c_cs = l_ld; g.c_bs[1] = l_bs; g.c_bs[2] = pstep_l;
        g.M = rows2; g.N = sz; g.K = sz;
        g.tri_b = VARGP_TRI_LOWER;
        rc = gemm_any(g, s);
        if (rc) return rc;
      }
      // W21 = -W22 (rows2 x rows2, lower) tmp (rows2 x sz)
      {
        vargp_gemm_t g = gemm_desc(batch);
        g.nb[2] = pairs;
        g.A = W22; g.a_rs = w_ld; g.a_cs = 1; g.a_bs[1] = w_bs; g.a_bs[2] = pstep_w;
        g.B = L12; g.b_rs = 1; g.b_cs = l_ld; g.b_bs[1] = l_bs; g.b_bs[2] = pstep_l;
        g.C = W21; g.c_rs = w_ld; g.c_cs = 1; g.c_bs[1] = w_bs; g.c_bs[2] = pstep_w;
        g.M = rows2; g.N = sz; g.K = rows2;
        g.alpha = -1.f;
        g.tri_a = VARGP_TRI_LOWER;
        rc = gemm_any(g, s);
        if (rc) return rc;
      }
    }
  }
  if (n > nb) {
    launch_k(zero_upper_kernel, dim3(dim3((unsigned)ceil_div(n, 256), (unsigned)ceil_div(n, kInitRows), (unsigned)batch)), dim3(256), 0, s, L, l_ld, l_bs,
                                                                                                    (int)n, nb);
    rc = launch_status();
    if (rc) return rc;
  }
  return 0;
}
