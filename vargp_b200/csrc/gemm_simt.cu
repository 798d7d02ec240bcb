// Batched strided fp32 GEMM on the SIMT pipes, with triangular-operand skipping and the fused RBF epilogue.
//
// This is the general-shape workhorse of the hot path (every shape / stride / alignment); the tcgen05
// 3xTF32 kernel in gemm_tc.cu takes over for the K-contiguous, TMA-alignable problems.
// Roofline: fp32 FMA pipe (148 SMs x 128 FMA/clk); operands are staged through shared memory in
// BK-deep slabs, each thread owns a TM x TN register tile.
#include "common.cuh"

namespace vargp {

template <int BM, int BN, int BK, int TM, int TN>
struct GemmCfg {
  static constexpr int kThreads = (BM / TM) * (BN / TN);
  static constexpr int kRowChunks = TM / 4, kColChunks = TN / 4;
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(const vargp_gemm_t g) {
  pdl_enter();
  using Cfg = GemmCfg<BM, BN, BK, TM, TN>;
  constexpr int NT = Cfg::kThreads;
  constexpr int TX = BN / TN;       // threads along n
  constexpr int RCH = Cfg::kRowChunks, CCH = Cfg::kColChunks;
  constexpr int RSTEP = BM / RCH, CSTEP = BN / CCH;

  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;

  int64_t z = blockIdx.z;
  const int64_t i2 = z % g.nb[2]; z /= g.nb[2];
  const int64_t i1 = z % g.nb[1];
  const int64_t i0 = z / g.nb[1];
  const float* __restrict__ A = g.A + i0 * g.a_bs[0] + i1 * g.a_bs[1] + i2 * g.a_bs[2];
  const float* __restrict__ B = g.B + i0 * g.b_bs[0] + i1 * g.b_bs[1] + i2 * g.b_bs[2];
  float* __restrict__ C = g.C + i0 * g.c_bs[0] + i1 * g.c_bs[1] + i2 * g.c_bs[2];

  // output-triangle tile culling
  bool dead = false;
  if (g.tri_c == VARGP_TRI_LOWER && n0 > m0 + BM - 1) dead = true;
  if (g.tri_c == VARGP_TRI_UPPER && m0 > n0 + BN - 1) dead = true;

  // k-range implied by structural zeros
  int64_t k_lo = 0, k_hi = g.K;
  if (g.tri_a == VARGP_TRI_LOWER) k_hi = min(k_hi, m0 + BM);
  if (g.tri_a == VARGP_TRI_UPPER) k_lo = max(k_lo, m0);
  if (g.tri_b == VARGP_TRI_LOWER) k_lo = max(k_lo, n0);
  if (g.tri_b == VARGP_TRI_UPPER) k_hi = min(k_hi, n0 + BN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const bool a_kfast = g.a_cs <= g.a_rs;
  const bool b_nfast = g.b_cs <= g.b_rs;

  if (!dead) {
    for (int64_t k0 = k_lo; k0 < k_hi; k0 += BK) {
      // ---- stage A (BM x BK) and B (BK x BN) ----
#pragma unroll
      for (int e = tid; e < BM * BK; e += NT) {
        int mm, kk;
        if (a_kfast) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
        const int64_t m = m0 + mm, k = k0 + kk;
        float v = 0.f;
        bool ok = (m < g.M) && (k < k_hi);
        if (g.tri_a == VARGP_TRI_LOWER) ok = ok && (k <= m);
        if (g.tri_a == VARGP_TRI_UPPER) ok = ok && (k >= m);
        if (ok) v = __ldg(A + m * g.a_rs + k * g.a_cs);
        As[kk][mm] = v;
      }
#pragma unroll
      for (int e = tid; e < BN * BK; e += NT) {
        int nn, kk;
        if (b_nfast) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
        const int64_t n = n0 + nn, k = k0 + kk;
        float v = 0.f;
        bool ok = (n < g.N) && (k < k_hi);
        if (g.tri_b == VARGP_TRI_LOWER) ok = ok && (n <= k);
        if (g.tri_b == VARGP_TRI_UPPER) ok = ok && (n >= k);
        if (ok) v = __ldg(B + k * g.b_rs + n * g.b_cs);
        Bs[kk][nn] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int rc = 0; rc < RCH; ++rc) {
          const float4 t = *reinterpret_cast<const float4*>(&As[kk][rc * RSTEP + ty * 4]);
          a[rc * 4 + 0] = t.x; a[rc * 4 + 1] = t.y; a[rc * 4 + 2] = t.z; a[rc * 4 + 3] = t.w;
        }
#pragma unroll
        for (int cc = 0; cc < CCH; ++cc) {
          const float4 t = *reinterpret_cast<const float4*>(&Bs[kk][cc * CSTEP + tx * 4]);
          b[cc * 4 + 0] = t.x; b[cc * 4 + 1] = t.y; b[cc * 4 + 2] = t.z; b[cc * 4 + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue ----
  float gamma2 = 1.f;
  const float* e_row = nullptr;
  const float* e_col = nullptr;
  if (g.epi != VARGP_EPI_NONE) {
    gamma2 = expf(2.f * g.e_theta[i0 * g.e_theta_bs[0] + i1 * g.e_theta_bs[1] + i2 * g.e_theta_bs[2] + g.e_D]);
    e_row = g.e_row + i0 * g.e_row_bs[0] + i1 * g.e_row_bs[1] + i2 * g.e_row_bs[2];
    e_col = g.e_col + i0 * g.e_col_bs[0] + i1 * g.e_col_bs[1] + i2 * g.e_col_bs[2];
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + (i / 4) * RSTEP + ty * 4 + (i % 4);
    if (m >= g.M) continue;
    float rown = 0.f;
    if (g.epi != VARGP_EPI_NONE) rown = 0.5f * e_row[m];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + (j / 4) * CSTEP + tx * 4 + (j % 4);
      if (n >= g.N) continue;
      const bool masked = (g.tri_c == VARGP_TRI_LOWER && n > m) || (g.tri_c == VARGP_TRI_UPPER && n < m);
      float* cp = C + m * g.c_rs + n * g.c_cs;
      if (masked) {
        if (g.beta == 0.f) *cp = 0.f;
        continue;
      }
      float v = acc[i][j];
      if (g.epi != VARGP_EPI_NONE) {
        v = gamma2 * expf(v - rown - 0.5f * e_col[n]);
        if (g.epi == VARGP_EPI_RBF_SYM && m == n) v = gamma2;
      }
      v *= g.alpha;
      if (g.beta != 0.f) v = fmaf(g.beta, *cp, v);
      *cp = v;
    }
  }
}

// N == 1 with K-contiguous A: one warp per output row, lanes stride over k (coalesced), shuffle reduce.
// (nubar = V gm, nu = W_ss m: a 64x64 GEMM tile would waste 63/64 of its lanes.)
__global__ void __launch_bounds__(256)
gemv_kernel(const vargp_gemm_t g) {
  pdl_enter();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * 8 + wid;
  int64_t z = blockIdx.z;
  const int64_t i2 = z % g.nb[2]; z /= g.nb[2];
  const int64_t i1 = z % g.nb[1];
  const int64_t i0 = z / g.nb[1];
  if (m >= g.M) return;
  const float* __restrict__ A = g.A + i0 * g.a_bs[0] + i1 * g.a_bs[1] + i2 * g.a_bs[2] + m * g.a_rs;
  const float* __restrict__ B = g.B + i0 * g.b_bs[0] + i1 * g.b_bs[1] + i2 * g.b_bs[2];
  float* C = g.C + i0 * g.c_bs[0] + i1 * g.c_bs[1] + i2 * g.c_bs[2] + m * g.c_rs;
  int64_t k_lo = 0, k_hi = g.K;
  if (g.tri_a == VARGP_TRI_LOWER) k_hi = min(k_hi, m + 1);
  if (g.tri_a == VARGP_TRI_UPPER) k_lo = m;
  float acc = 0.f;
  for (int64_t k = k_lo + lane; k < k_hi; k += 32) acc = fmaf(__ldg(A + k * g.a_cs), __ldg(B + k * g.b_rs), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = g.alpha * acc;
    if (g.beta != 0.f) v = fmaf(g.beta, *C, v);
    *C = v;
  }
}

static int check_gemm(const vargp_gemm_t* g) {
  if (!g || !g->A || !g->B || !g->C) return VARGP_ERR_ARG;
  if (g->M < 0 || g->N < 0 || g->K < 0) return VARGP_ERR_ARG;
  for (int i = 0; i < 3; ++i)
    if (g->nb[i] < 1) return VARGP_ERR_ARG;
  if (g->epi != VARGP_EPI_NONE && (!g->e_row || !g->e_col || !g->e_theta)) return VARGP_ERR_ARG;
  return 0;
}

}  // namespace vargp

using namespace vargp;

extern "C" int vargp_gemm(const vargp_gemm_t* g, void* stream) {
  int rc = check_gemm(g);
  if (rc) return rc;
  if (g->M == 0 || g->N == 0) return 0;
  const int64_t nbatch = g->nb[0] * g->nb[1] * g->nb[2];
  if (nbatch > 65535) return VARGP_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (g->N == 1 && g->epi == VARGP_EPI_NONE && g->tri_b == VARGP_TRI_NONE && g->tri_c == VARGP_TRI_NONE &&
      g->a_cs <= g->a_rs && g->M >= 8) {
    dim3 grid((unsigned)ceil_div(g->M, 8), 1, (unsigned)nbatch);
    launch_k(gemv_kernel, dim3(grid), dim3(256), 0, s, *g);
    return launch_status();
  }
  // large tiles only when they still fill the machine (148 SMs)
  const int64_t big_tiles = ceil_div(g->M, 128) * ceil_div(g->N, 128) * nbatch;
  if (g->M >= 128 && g->N >= 128 && big_tiles >= 148) {
    dim3 grid((unsigned)ceil_div(g->N, 128), (unsigned)ceil_div(g->M, 128), (unsigned)nbatch);
    launch_k((gemm_simt_kernel<128, 128, 8, 8, 8>), dim3(grid), dim3(256), 0, s, *g);
  } else {
    dim3 grid((unsigned)ceil_div(g->N, 64), (unsigned)ceil_div(g->M, 64), (unsigned)nbatch);
    launch_k((gemm_simt_kernel<64, 64, 16, 4, 4>), dim3(grid), dim3(256), 0, s, *g);
  }
  return launch_status();
}
