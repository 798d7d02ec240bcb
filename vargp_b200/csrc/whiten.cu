// Whitening of the variational parameters against the Cholesky inverse, per (hyper sample h, class c, task s) on the
// M x M diagonal blocks -- forward and adjoint as ONE kernel each (M <= 128 forward, M <= 96 adjoint):
//
//   forward   T_s = W_ss Lu_s,  nu_s = W_ss m_s,  N_ss += T_s T_s^T,
//             KL_hc = -sum_i log W_ii - sum_i log Lu_t,ii + (|T_t|_F^2 + |nu_t|^2 - M) / 2        (last task block t)
//   adjoint   Tbar_s = tril(2 G_ss T_s) + k T_s,  nubar'_s = nubar_s + k nu_s            (k = g_kl / H on block t, else 0)
//             Wbar_ss += tril(Tbar_s Lu_s^T + nubar'_s m_s^T) - k diag(1 / W_ii)
//             Lubar_s = tril(W_ss^T Tbar_s),  mbar_s = W_ss^T nubar'_s                 (blocks whose gradients are wanted)
//
// These are the block-diagonal pieces of the autoregressive posterior of var_gp/vargp.py:35-88 / gp_utils.py:101-147 in
// whitened coordinates and of the KL of vargp.py:182-190.  As batched tensor-core GEMMs they were
// 150 problems of 60 x 60 x 60 each padded into a 128 x 128 tile: 4 launches forward (T, nu, KL, N += T T^T) and 6 in the
// backward pass (Tbar, KL adjoint, four whitening-adjoint products) of 20..45 us each on the critical chain of the
// Split-MNIST-shaped step -- for 65 MFLOP apiece.  Here a CTA keeps the blocks of one (h, c, s) in shared memory and does
// the three products on the fp32 FMA pipe (plain fp32: no TF32 splitting involved).
// Roofline: latency / shared-memory bandwidth of one SM per block (H*C*S CTAs); algorithmic work 3 * 2 M^3 flops per block.
#include "common.cuh"

namespace vargp {

constexpr int WT = 16;      // threads per tile dimension: thread (ti, tj) owns rows ti + 16 u, columns tj + 16 v

// acc[u][v] = sum_k A'(ti + 16 u, k) B'(k, tj + 16 v), operands in shared memory with leading dimension ld (odd)
template <int R, bool TA, bool TB>
__device__ __forceinline__ void block_mm(const float* __restrict__ A, const float* __restrict__ B, int M, int ld, int ti,
                                         int tj, float (&acc)[R][R]) {
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int v = 0; v < R; ++v) acc[u][v] = 0.f;
  // the buffers are zero-padded to 16 R rows / columns, so tiles that reach past M need no guards here
#pragma unroll 2
  for (int k = 0; k < M; ++k) {
    float a[R], b[R];
#pragma unroll
    for (int u = 0; u < R; ++u) a[u] = TA ? A[k * ld + ti + WT * u] : A[(ti + WT * u) * ld + k];
#pragma unroll
    for (int v = 0; v < R; ++v) b[v] = TB ? B[(tj + WT * v) * ld + k] : B[k * ld + tj + WT * v];
#pragma unroll
    for (int u = 0; u < R; ++u)
#pragma unroll
      for (int v = 0; v < R; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
  }
}

// M x M block at src (row stride lds) -> shared memory (row stride ld), rows / columns >= M zero-filled up to Mp
// (Mp = 16 R is a multiple of 16, so Mp * Mp is a multiple of the 256 threads: 8 independent loads per thread in flight)
__device__ __forceinline__ void load_block(float* __restrict__ dst, int ld, const float* __restrict__ src, int64_t lds, int M,
                                           int Mp) {
  constexpr int U = 8;
  for (int e0 = threadIdx.x; e0 < Mp * Mp; e0 += U * WT * WT) {
    float v[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int e = e0 + q * WT * WT;
      const int i = e / Mp, j = e - i * Mp;
      v[q] = (e < Mp * Mp && i < M && j < M) ? __ldg(src + (int64_t)i * lds + j) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int e = e0 + q * WT * WT;
      const int i = e / Mp, j = e - i * Mp;
      if (e < Mp * Mp) dst[i * ld + j] = v[q];
    }
  }
}

struct Rect { int h0, c0, Hs, Cs; };

// grid (S, Hs * Cs)
template <int R>
__global__ void __launch_bounds__(WT * WT)
whiten_fwd_kernel(const float* __restrict__ W, const float* __restrict__ Lu_all, const float* __restrict__ m_all,
                  int H, int C, int S, int M, int P, Rect rc, float* __restrict__ T, float* __restrict__ nu,
                  float* __restrict__ N, float* kl, float* work) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  constexpr int Mp = WT * R;
  const int ld = Mp + 1;
  float* sW = sm;
  float* sLu = sW + Mp * ld;
  float* sT = sLu + Mp * ld;
  float* sv = sT + Mp * ld;          // m (Mp) | nu (Mp)
  __shared__ float scratch[32];
  __shared__ bool s_last;
  const int s = blockIdx.x;
  const int h = rc.h0 + blockIdx.y / rc.Cs, c = rc.c0 + blockIdx.y % rc.Cs;
  const int64_t g = (int64_t)h * C + c;
  const int ti = threadIdx.x / WT, tj = threadIdx.x % WT;
  const float* Wg = W + g * P * P + ((int64_t)s * M) * P + (int64_t)s * M;
  load_block(sW, ld, Wg, P, M, Mp);
  load_block(sLu, ld, Lu_all + ((int64_t)s * C + c) * M * M, M, M, Mp);
  for (int i = threadIdx.x; i < Mp; i += blockDim.x) sv[i] = i < M ? m_all[((int64_t)s * C + c) * M + i] : 0.f;
  __syncthreads();

  float acc[R][R];
  block_mm<R, false, false>(sW, sLu, M, ld, ti, tj, acc);
  float* Tg = T + ((g * S + s) * M) * (int64_t)M;
  float klacc = 0.f;
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int v = 0; v < R; ++v) {
      const int i = ti + WT * u, j = tj + WT * v;
      const float t = (j <= i) ? acc[u][v] : 0.f;       // both factors are lower triangular: exact zeros above
      sT[i * ld + j] = (i < M && j < M) ? t : 0.f;
      if (i < M && j < M) {
        Tg[(int64_t)i * M + j] = t;
        if (s == S - 1) klacc = fmaf(0.5f * t, t, klacc);
      }
    }
  if (threadIdx.x < M) {        // nu_s = W_ss m_s  (row i; conflict-free: ld is odd)
    const int i = threadIdx.x;
    float a = 0.f;
    for (int k = 0; k <= i; ++k) a = fmaf(sW[i * ld + k], sv[k], a);
    nu[g * P + (int64_t)s * M + i] = a;
    if (s == S - 1) klacc += 0.5f * (a * a - 1.f) - logf(sW[i * ld + i]) - logf(sLu[i * ld + i]);
  }
  __syncthreads();
  block_mm<R, false, true>(sT, sT, M, ld, ti, tj, acc);
  float* Ng = N + g * P * P + ((int64_t)s * M) * P + (int64_t)s * M;
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int v = 0; v < R; ++v) {
      const int i = ti + WT * u, j = tj + WT * v;
      if (i < M && j < M) Ng[(int64_t)i * P + j] += acc[u][v];
    }
  if (kl && s == S - 1) {
    // deterministic: per-(h, c) partial sums, added up in a fixed order by the CTA that draws the last ticket
    const unsigned nparts = rc.Hs * rc.Cs;
    klacc = block_sum(klacc, scratch);
    unsigned* ticket = reinterpret_cast<unsigned*>(work);
    float* part = work + 1;
    if (threadIdx.x == 0) {
      part[blockIdx.y] = klacc;
      __threadfence();
      s_last = atomicAdd(ticket, 1u) == nparts - 1;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      float t = 0.f;
      for (unsigned i = threadIdx.x; i < nparts; i += blockDim.x) t += __ldcg(part + i);
      t = block_sum(t, scratch);
      if (threadIdx.x == 0) {
        kl[0] += t / (float)H;
        *ticket = 0u;
      }
    }
  }
}

// grid (S, Hs * Cs)
template <int R>
__global__ void __launch_bounds__(WT * WT)
whiten_bwd_kernel(const float* __restrict__ W, const float* __restrict__ T, const float* __restrict__ nu,
                  const float* __restrict__ Lu_all, const float* __restrict__ m_all, const float* __restrict__ Gm,
                  const float* __restrict__ nubar, const float* __restrict__ g_kl, int H, int C, int S, int M, int P,
                  Rect rc, int s_grad0, float* __restrict__ Wbar, float* __restrict__ Lubar, float* __restrict__ mbar) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  constexpr int Mp = WT * R;
  const int ld = Mp + 1;
  float* sA = sm;                    // G_ss, later W_ss
  float* sT = sA + Mp * ld;          // T_s
  float* sLu = sT + Mp * ld;         // Lu_s
  float* sTb = sLu + Mp * ld;        // Tbar_s
  float* sv = sTb + Mp * ld;         // m (Mp) | nubar' (Mp)
  const int s = blockIdx.x;
  const int h = rc.h0 + blockIdx.y / rc.Cs, c = rc.c0 + blockIdx.y % rc.Cs;
  const int64_t g = (int64_t)h * C + c;
  const int ti = threadIdx.x / WT, tj = threadIdx.x % WT;
  const int64_t blk = g * P * P + ((int64_t)s * M) * P + (int64_t)s * M;
  const float k = (g_kl && s == S - 1) ? g_kl[0] / (float)H : 0.f;
  load_block(sA, ld, Gm + blk, P, M, Mp);
  load_block(sT, ld, T + ((g * S + s) * M) * (int64_t)M, M, M, Mp);
  load_block(sLu, ld, Lu_all + ((int64_t)s * C + c) * M * M, M, M, Mp);
  for (int i = threadIdx.x; i < Mp; i += blockDim.x) {
    sv[i] = i < M ? m_all[((int64_t)s * C + c) * M + i] : 0.f;
    sv[Mp + i] = i < M ? fmaf(k, nu[g * P + (int64_t)s * M + i], nubar[g * P + (int64_t)s * M + i]) : 0.f;
  }
  __syncthreads();

  float acc[R][R];
  block_mm<R, false, false>(sA, sT, M, ld, ti, tj, acc);              // G_ss T_s
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int v = 0; v < R; ++v) {
      const int i = ti + WT * u, j = tj + WT * v;
      sTb[i * ld + j] = (i < M && j <= i) ? fmaf(2.f, acc[u][v], k * sT[i * ld + j]) : 0.f;
    }
  __syncthreads();
  const bool want = s >= s_grad0;
  if (want) load_block(sA, ld, W + blk, P, M, Mp);                    // G_ss is dead: its buffer takes W_ss
  block_mm<R, false, true>(sTb, sLu, M, ld, ti, tj, acc);             // Tbar_s Lu_s^T
  float* Wb = Wbar + blk;
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int v = 0; v < R; ++v) {
      const int i = ti + WT * u, j = tj + WT * v;
      if (i < M && j <= i) {
        float val = fmaf(sv[Mp + i], sv[j], acc[u][v]);
        if (i == j && k != 0.f) val -= k / W[blk + (int64_t)i * P + i];
        Wb[(int64_t)i * P + j] += val;
      }
    }
  if (!want) return;
  __syncthreads();
  block_mm<R, true, false>(sA, sTb, M, ld, ti, tj, acc);              // W_ss^T Tbar_s
  const int64_t Sg = S - s_grad0;
  float* Lb = Lubar + ((((int64_t)h * Sg + (s - s_grad0)) * C + c) * M) * M;
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int v = 0; v < R; ++v) {
      const int i = ti + WT * u, j = tj + WT * v;
      if (i < M && j < M) Lb[(int64_t)i * M + j] = (j <= i) ? acc[u][v] : 0.f;
    }
  if (threadIdx.x < M) {        // mbar_i = sum_k W[k][i] nubar'_k   (column i: consecutive threads, consecutive banks)
    const int i = threadIdx.x;
    float a = 0.f;
    for (int kk = i; kk < M; ++kk) a = fmaf(sA[kk * ld + i], sv[Mp + kk], a);
    mbar[(((int64_t)h * Sg + (s - s_grad0)) * C + c) * M + i] = a;
  }
}

template <int R> constexpr int whiten_smem(int nbuf) { return (nbuf * (WT * R) * (WT * R + 1) + 2 * WT * R) * (int)sizeof(float); }

}  // namespace vargp

using namespace vargp;

extern "C" int64_t vargp_whiten_fwd_work(int64_t H, int64_t C) { return 1 + H * C; }

// largest M the shared-memory kernels take (forward / adjoint); beyond that the caller uses the batched GEMMs
extern "C" int64_t vargp_whiten_max_m(int adjoint) { return adjoint ? 96 : 128; }

extern "C" int vargp_whiten_fwd(const float* W, const float* Lu_all, const float* m_all, int64_t H, int64_t C, int64_t S,
                                int64_t M, int64_t P, int64_t h0, int64_t h1, int64_t c0, int64_t c1, float* T, float* nu,
                                float* N, float* kl, float* work, void* stream) {
  if (!W || !Lu_all || !m_all || !T || !nu || !N || (kl && !work)) return VARGP_ERR_ARG;
  if (H < 1 || C < 1 || S < 1 || M < 1 || P != S * M || h0 < 0 || h1 > H || h0 >= h1 || c0 < 0 || c1 > C || c0 >= c1) return VARGP_ERR_ARG;
  if (M > 128 || (h1 - h0) * (c1 - c0) > 65535) return VARGP_ERR_UNSUPPORTED;
  const Rect rc = {(int)h0, (int)c0, (int)(h1 - h0), (int)(c1 - c0)};
  dim3 grid((unsigned)S, (unsigned)(rc.Hs * rc.Cs));
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(whiten_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, whiten_smem<8>(3));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(whiten_fwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, whiten_smem<6>(3));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(whiten_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, whiten_smem<4>(3));
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
#define VARGP_WF(R) launch_k((whiten_fwd_kernel<R>), dim3(grid), dim3(WT * WT), whiten_smem<R>(3), st, W, Lu_all, m_all, (int)H, \
                             (int)C, (int)S, (int)M, (int)P, rc, T, nu, N, kl, work)
  if (M <= 32) VARGP_WF(2);
  else if (M <= 64) VARGP_WF(4);
  else if (M <= 96) VARGP_WF(6);
  else VARGP_WF(8);
#undef VARGP_WF
  return launch_status();
}

extern "C" int vargp_whiten_bwd(const float* W, const float* T, const float* nu, const float* Lu_all, const float* m_all,
                                const float* G, const float* nubar, const float* g_kl, int64_t H, int64_t C, int64_t S,
                                int64_t M, int64_t P, int64_t h0, int64_t h1, int64_t c0, int64_t c1, int64_t s_grad0,
                                float* Wbar, float* Lubar, float* mbar, void* stream) {
  if (!W || !T || !nu || !Lu_all || !m_all || !G || !nubar || !Wbar || !Lubar || !mbar) return VARGP_ERR_ARG;
  if (H < 1 || C < 1 || S < 1 || M < 1 || P != S * M || h0 < 0 || h1 > H || h0 >= h1 || c0 < 0 || c1 > C || c0 >= c1 ||
      s_grad0 < 0 || s_grad0 >= S)
    return VARGP_ERR_ARG;
  if (M > 96 || (h1 - h0) * (c1 - c0) > 65535) return VARGP_ERR_UNSUPPORTED;
  const Rect rc = {(int)h0, (int)c0, (int)(h1 - h0), (int)(c1 - c0)};
  dim3 grid((unsigned)S, (unsigned)(rc.Hs * rc.Cs));
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(whiten_bwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, whiten_smem<6>(4));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(whiten_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, whiten_smem<4>(4));
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
#define VARGP_WB(R) launch_k((whiten_bwd_kernel<R>), dim3(grid), dim3(WT * WT), whiten_smem<R>(4), st, W, T, nu, Lu_all, m_all, G, \
                             nubar, g_kl, (int)H, (int)C, (int)S, (int)M, (int)P, rc, (int)s_grad0, Wbar, Lubar, mbar)
  if (M <= 32) VARGP_WB(2);
  else if (M <= 64) VARGP_WB(4);
  else VARGP_WB(6);
#undef VARGP_WB
  return launch_status();
}
