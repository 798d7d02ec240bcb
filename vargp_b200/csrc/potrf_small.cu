// Shared-memory-resident Cholesky + triangular inverse of small matrices (n <= 128), one CTA per matrix:
//     L = chol(A + jitter I),   W = L^-1
// This is the diagonal-block kernel of the blocked factorisation (potrf_blocked.cu) and the whole factorisation
// for the first tasks (P = M <= 128).  The one-CTA kernels of chol.cu keep their panel in registers but stream the
// factor through L2 for every 32-column update (measured 410 + 210 us at n = 300, ~150 us per 128-block); here the
// matrix is loaded once, factored and inverted out of shared memory, and written once.
//
//   load   S <- tril(A) + jitter I, padded with the identity up to a multiple of 32        (132-float rows: a lane
//          reading ITS row as float4 is conflict-free, the other operand is always a warp-wide broadcast)
//   for each 32-column block k (right-looking):
//     (a) warp 0 factors the 32 x 32 diagonal block: lane i keeps row i in registers, column j is broadcast
//         through shared memory (the sequential critical path: 32 steps of sqrt + scale + rank-1 update)
//     (b) panel: thread per row below the block solves x L_kk^T = a (forward substitution over registers) and
//         also leaves the panel transposed in PT (so that (c) can read it as broadcasts)
//     (c) trailing update: one warp per (row block, column block) pair of the lower triangle,
//         lane = row: acc[32] -= a[kk] * PT[kk][32 columns]
//   inverse: the four diagonal blocks (one warp each, lane = column, forward substitution), then block
//   sub-diagonal after block sub-diagonal  W_ij = -W_ii sum_{l=j..i-1} L_il W_lj ; every (block, 8-column slice)
//   is one warp's unit of work.
// Roofline: latency / FMA issue of ONE SM per matrix (30 matrices -> 30 SMs); n^3/3 + n^3/3 flops.
#include "common.cuh"

namespace vargp {

constexpr int SB = 32;
constexpr int SMAX = 128;
constexpr int SLD = 132;
constexpr int kSmallThreads = 256;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr int SCR_LD = 12;                                   // per-warp 32 x 8 scratch, padded rows
constexpr int kSmallSmemFloats = 2 * SMAX * SLD + SB * SLD + 32 + SMAX + 4;

__device__ __forceinline__ void ld_row32(const float* p, float (&a)[SB]) {
#pragma unroll
  for (int c4 = 0; c4 < SB; c4 += 4) {
    const float4 v = *reinterpret_cast<const float4*>(p + c4);
    a[c4] = v.x; a[c4 + 1] = v.y; a[c4 + 2] = v.z; a[c4 + 3] = v.w;
  }
}
__device__ __forceinline__ void st_row32(float* p, const float (&a)[SB]) {
#pragma unroll
  for (int c4 = 0; c4 < SB; c4 += 4) *reinterpret_cast<float4*>(p + c4) = make_float4(a[c4], a[c4 + 1], a[c4 + 2], a[c4 + 3]);
}

__global__ void __launch_bounds__(kSmallThreads, 1)
potrf_inv_small_kernel(const float* Ain, int64_t a_ld, int64_t a_bs, float* Lout, int64_t l_ld, int64_t l_bs,
                       float* Wout, int64_t w_ld, int64_t w_bs, int n, float jitter, int32_t* __restrict__ info,
                       int info_base, int accumulate) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  float* Ls = sm;
  float* Ws = Ls + SMAX * SLD;
  float* PT = Ws + SMAX * SLD;
  float* colj = PT + SB * SLD;
  float* dinv = colj + 32;
  int* s_info = reinterpret_cast<int*>(dinv + SMAX);     // dinv[i] = 1 / L_ii for the whole matrix

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nblk = (n + SB - 1) / SB, np = nblk * SB;
  const float* A = Ain + (int64_t)blockIdx.x * a_bs;
  if (tid == 0) *s_info = 0;

  // ---- load (coalesced along j; 16 independent loads in flight per thread: the matrix comes from L2 / HBM) ----
  {
    const int cpr = np / 32;                       // 32-column chunks per row
    const int nchunk = np * cpr;                   // (row, chunk) items, dealt to warps
    for (int c0 = wid * 16; c0 < nchunk; c0 += kSmallWarps * 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int c = c0 + u;
        const int i = c / cpr, j = (c % cpr) * 32 + lane;
        v[u] = (i == j) ? 1.f : 0.f;
        if (c < nchunk && i < n && j <= i) v[u] = A[(int64_t)i * a_ld + j] + ((i == j) ? jitter : 0.f);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int c = c0 + u;
        const int i = c / cpr, j = (c % cpr) * 32 + lane;
        if (c < nchunk) Ls[i * SLD + j] = v[u];
      }
    }
  }
  __syncthreads();

  // ---- factorisation ----
  for (int k = 0; k < nblk; ++k) {
    const int k0 = k * SB;
    if (wid == 0) {
      // (a) diagonal block: lane i owns row i; static register indices only (selects), see chol.cu
      float acc[SB];
      ld_row32(Ls + (k0 + lane) * SLD + k0, acc);
      // fully unrolled over the column j: every register index and every triangle predicate but `c <= lane` is
      // static (496 predicated FMAs in total instead of 32 x 32 selects per step)
#pragma unroll
      for (int j = 0; j < SB; ++j) {
        const float d = __shfl_sync(0xffffffffu, acc[j], j);
        if (!(d > 0.f) && lane == 0 && k0 + j < n && *s_info == 0) *s_info = k0 + j + 1;
        const float dj = sqrtf(d);
        const float inv = 1.f / dj;
        const float lij = (lane == j) ? dj : ((lane > j) ? acc[j] * inv : 0.f);
        acc[j] = lij;
        if (j + 1 < SB) {
          colj[lane] = lij;
          if (lane == j) dinv[k0 + j] = inv;
          __syncwarp();
#pragma unroll
          for (int c4 = ((j + 1) / 4) * 4; c4 < SB; c4 += 4) {
            const float4 l = *reinterpret_cast<const float4*>(&colj[c4]);
            const float lc[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int c = c4 + u;
              if (c > j && c <= lane) acc[c] = fmaf(-lij, lc[u], acc[c]);
            }
          }
          __syncwarp();
        } else if (lane == j) {
          dinv[k0 + j] = inv;
        }
      }
#pragma unroll
      for (int c = 0; c < SB; ++c)
        if (c > lane) acc[c] = 0.f;
      st_row32(Ls + (k0 + lane) * SLD + k0, acc);
    }
    __syncthreads();
    const int R = np - k0 - SB;                 // rows below the diagonal block
    if (R == 0) break;
    // (b) panel solve, thread per row; L_kk read as broadcasts
    if (tid < R) {
      float* rowp = Ls + (k0 + SB + tid) * SLD + k0;
      float acc[SB];
      ld_row32(rowp, acc);
#pragma unroll
      for (int j = 0; j < SB; ++j) {
        float s = acc[j];
#pragma unroll
        for (int l4 = 0; l4 < j; l4 += 4) {
          const float4 dd = *reinterpret_cast<const float4*>(Ls + (k0 + j) * SLD + k0 + l4);
          s = fmaf(-acc[l4 + 0], dd.x, s);
          if (l4 + 1 < j) s = fmaf(-acc[l4 + 1], dd.y, s);
          if (l4 + 2 < j) s = fmaf(-acc[l4 + 2], dd.z, s);
          if (l4 + 3 < j) s = fmaf(-acc[l4 + 3], dd.w, s);
        }
        acc[j] = s * dinv[k0 + j];
        PT[j * SLD + tid] = acc[j];
      }
      st_row32(rowp, acc);
    }
    __syncthreads();
    // (c) trailing update over the (row block, column block) pairs of the remaining lower triangle
    const int nrem = nblk - k - 1;
    const int npairs = nrem * (nrem + 1) / 2;
    for (int p = wid; p < npairs; p += kSmallWarps) {
      int rb = 0, rem = p;
      while (rem > rb) { rem -= rb + 1; ++rb; }          // p -> (rb, jb = rem), jb <= rb
      const int jb = rem;
      float* rowp = Ls + (k0 + SB + rb * SB + lane) * SLD;
      float a[SB], acc[SB];
      ld_row32(rowp + k0, a);
      ld_row32(rowp + k0 + SB + jb * SB, acc);
#pragma unroll
      for (int kk = 0; kk < SB; ++kk) {
        const float* bp = PT + kk * SLD + jb * SB;
#pragma unroll
        for (int c4 = 0; c4 < SB; c4 += 4) {
          const float4 b = *reinterpret_cast<const float4*>(bp + c4);
          acc[c4 + 0] = fmaf(-a[kk], b.x, acc[c4 + 0]);
          acc[c4 + 1] = fmaf(-a[kk], b.y, acc[c4 + 1]);
          acc[c4 + 2] = fmaf(-a[kk], b.z, acc[c4 + 2]);
          acc[c4 + 3] = fmaf(-a[kk], b.w, acc[c4 + 3]);
        }
      }
      st_row32(rowp + k0 + SB + jb * SB, acc);
    }
    __syncthreads();
  }

  // ---- inverse: diagonal blocks (warp per block, lane = column j solves L_kk x = e_j) ----
  for (int k = wid; k < nblk; k += kSmallWarps) {
    const int k0 = k * SB;
    const int j = lane;
    float x[SB];
    // fully unrolled over the row i (static triangle: 496 FMAs); x[c] = 0 for c < j by construction
#pragma unroll
    for (int i = 0; i < SB; ++i) {
      const float* lrow = Ls + (k0 + i) * SLD + k0;
      float s0 = (i == j) ? 1.f : 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < i; c4 += 4) {
        const float4 l = *reinterpret_cast<const float4*>(lrow + c4);
        s0 = fmaf(-l.x, x[c4 + 0], s0);
        if (c4 + 1 < i) s1 = fmaf(-l.y, x[c4 + 1], s1);
        if (c4 + 2 < i) s2 = fmaf(-l.z, x[c4 + 2], s2);
        if (c4 + 3 < i) s3 = fmaf(-l.w, x[c4 + 3], s3);
      }
      const float xi = (i >= j) ? ((s0 + s1) + (s2 + s3)) * dinv[k0 + i] : 0.f;
      x[i] = xi;
      Ws[(k0 + i) * SLD + k0 + lane] = xi;               // row i of the block, lanes = columns (zero above the diagonal)
    }
  }
  __syncthreads();
  // ---- inverse: block sub-diagonals ----
  float* scr = PT + wid * (SB * SCR_LD);
  for (int d = 1; d < nblk; ++d) {
    const int units = (nblk - d) * 4;
    for (int u = wid; u < units; u += kSmallWarps) {
      const int jb = u >> 2, q = u & 3, ib = jb + d;
      const int c0 = jb * SB + q * 8;
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
      for (int l = jb; l < ib; ++l) {
        float a[SB];
        ld_row32(Ls + (ib * SB + lane) * SLD + l * SB, a);
#pragma unroll
        for (int kk = 0; kk < SB; ++kk) {
          const float* bp = Ws + (l * SB + kk) * SLD + c0;
          const float4 b0 = *reinterpret_cast<const float4*>(bp), b1 = *reinterpret_cast<const float4*>(bp + 4);
          acc[0] = fmaf(a[kk], b0.x, acc[0]); acc[1] = fmaf(a[kk], b0.y, acc[1]);
          acc[2] = fmaf(a[kk], b0.z, acc[2]); acc[3] = fmaf(a[kk], b0.w, acc[3]);
          acc[4] = fmaf(a[kk], b1.x, acc[4]); acc[5] = fmaf(a[kk], b1.y, acc[5]);
          acc[6] = fmaf(a[kk], b1.z, acc[6]); acc[7] = fmaf(a[kk], b1.w, acc[7]);
        }
      }
      __syncwarp();
      *reinterpret_cast<float4*>(scr + lane * SCR_LD) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(scr + lane * SCR_LD + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      __syncwarp();
      float a[SB];
      ld_row32(Ws + (ib * SB + lane) * SLD + ib * SB, a);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
      for (int kk = 0; kk < SB; ++kk) {
        const float4 b0 = *reinterpret_cast<const float4*>(scr + kk * SCR_LD);
        const float4 b1 = *reinterpret_cast<const float4*>(scr + kk * SCR_LD + 4);
        acc[0] = fmaf(-a[kk], b0.x, acc[0]); acc[1] = fmaf(-a[kk], b0.y, acc[1]);
        acc[2] = fmaf(-a[kk], b0.z, acc[2]); acc[3] = fmaf(-a[kk], b0.w, acc[3]);
        acc[4] = fmaf(-a[kk], b1.x, acc[4]); acc[5] = fmaf(-a[kk], b1.y, acc[5]);
        acc[6] = fmaf(-a[kk], b1.z, acc[6]); acc[7] = fmaf(-a[kk], b1.w, acc[7]);
      }
      float* wp = Ws + (ib * SB + lane) * SLD + c0;
      *reinterpret_cast<float4*>(wp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(wp + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
  }

  // ---- store (coalesced along j; strict upper triangles zero-filled) ----
  float* L = Lout + (int64_t)blockIdx.x * l_bs;
  float* W = Wout + (int64_t)blockIdx.x * w_bs;
  for (int i = wid; i < n; i += kSmallWarps) {
    for (int j = lane; j < n; j += 32) {
      const bool low = j <= i;
      L[(int64_t)i * l_ld + j] = low ? Ls[i * SLD + j] : 0.f;
      W[(int64_t)i * w_ld + j] = low ? Ws[i * SLD + j] : 0.f;
    }
  }
  if (tid == 0 && info) {
    const int si = *s_info;
    if (!accumulate) info[blockIdx.x] = si ? si + info_base : 0;
    else if (si && info[blockIdx.x] == 0) info[blockIdx.x] = si + info_base;
  }
}

}  // namespace vargp

using namespace vargp;

// L = chol(A + jitter I), W = L^-1 for n <= 128 (A may alias W: the matrix is consumed before anything is stored).
extern "C" int vargp_chol_inv_small(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                                    float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                                    int32_t* info, int64_t info_base, int accumulate, void* stream) {
  if (!A || !L || !W || n < 1 || batch < 1 || a_ld < n || l_ld < n || w_ld < n) return VARGP_ERR_ARG;
  if (n > SMAX) return VARGP_ERR_UNSUPPORTED;
  if (L == W || A == L) return VARGP_ERR_ARG;
  static bool attr_set = false;
  const int dyn = kSmallSmemFloats * (int)sizeof(float);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(potrf_inv_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  launch_k(potrf_inv_small_kernel, dim3((unsigned)batch), dim3(kSmallThreads), dyn, (cudaStream_t)stream, 
      A, a_ld, a_bs, L, l_ld, l_bs, W, w_ld, w_bs, (int)n, jitter, info, (int)info_base, accumulate);
  return launch_status();
}
