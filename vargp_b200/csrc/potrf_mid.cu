// Shared-memory-resident Cholesky + triangular inverse for mid-sized matrices (128 < n <= 320), one CTA per matrix:
//     L = chol(A + jitter I),   W = L^-1                       (var_gp/gp_utils.py:5-11 and the solves behind it)
//
// Written in round 2 for the P = 300 case of the benched step, where the blocked driver (potrf_blocked.cu) spends 13
// dependent launches (3 x the n <= 128 kernel, 8 tensor-core GEMMs, 2 helpers).  MEASURED (B200, batch 30): it wins for
// 128 < n <= 192 (n = 129: 97 vs 134 us, n = 180: 131 vs 138 us) and LOSES above (n = 240: 215 vs 172 us, n = 300:
// 354 vs 244 us): without room for padded rows the 32 x 32 block products run at ~0.55 us each on one SM (issue- and
// latency-bound: 12 LDS + ~25 address ops per 32 FMAs, 4 warps per scheduler), and their count grows with n^3.
// vargp_chol_inv therefore routes only 128 < n <= 192 here (vargp_chol_mid_config).  The whole lower triangle lives in
// ONE CTA's shared memory as packed 32 x 32 blocks (55 blocks = 220 KB at n = 320) and nothing leaves the SM between the
// load and the two stores:
//
//   load   blocks <- tril(A) + jitter I, padded with the identity up to a multiple of 32
//   for each block column k (right-looking):
//     (a) warp 0 factors the 32 x 32 diagonal block (lane = row in registers, column broadcast through smem)
//     (b) panel: one thread per row below solves x L_kk^T = a by forward substitution (L_kk read as broadcasts)
//     (c) trailing update B(i,j) -= B(i,k) B(j,k)^T: one warp per block pair, 8 x 4 register tile per lane
//   store L;  then the inverse IN PLACE over L:
//     diagonal blocks: one warp per block, lane = column, forward substitution
//     log2 levels of merging adjacent inverted blocks  W21 = -W22 (L21 W11): each 32 x 32 output block is one
//     warp's register tile, products summed over the inner blocks, written back after a CTA barrier (no scratch)
//   store W
//
// Block (i, j) element (r, c) sits at r * 32 + ((c + r) & 31): row-wise, column-wise and broadcast accesses are all
// bank-conflict free without padding (there is no room for padding: 227 KB is the limit).
// Roofline: latency of the 32-column diagonal steps + shared-memory bandwidth of one SM (the block products run at
// 12 LDS per 32 FMA); n^3 / 3 + n^3 / 3 flops per matrix.
#include "common.cuh"

namespace vargp {

constexpr int MB = 32;
constexpr int kMidThreads = 512;
constexpr int kMidWarps = kMidThreads / 32;
constexpr int kMidMaxBlk = 10;                        // n <= 320

__device__ __forceinline__ int bidx(int i, int j) { return (i * (i + 1) / 2 + j) * (MB * MB); }
__device__ __forceinline__ int sk(int r, int c) { return r * MB + ((c + r) & (MB - 1)); }

// acc[u][v] (+)= sign * sum_kk A(r_u, kk) * B'(kk, c_v);  BT: B' = B^T (element (c, kk) of the block), else element (kk, c)
template <bool BT>
__device__ __forceinline__ void tile_mm(const float* __restrict__ A, const float* __restrict__ B, int ly, int lx,
                                        float (&acc)[8][4], float sign) {
#pragma unroll 4
  for (int kk = 0; kk < MB; ++kk) {
    float a[8], b[4];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = sign * A[sk(ly + 4 * u, kk)];
#pragma unroll
    for (int v = 0; v < 4; ++v) b[v] = BT ? B[sk(lx + 8 * v, kk)] : B[sk(kk, lx + 8 * v)];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
  }
}

__device__ __forceinline__ void tile_zero(float (&acc)[8][4]) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
}
__device__ __forceinline__ void tile_load(const float* blk, int ly, int lx, float (&acc)[8][4]) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = blk[sk(ly + 4 * u, lx + 8 * v)];
}
__device__ __forceinline__ void tile_store(float* blk, int ly, int lx, const float (&acc)[8][4]) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) blk[sk(ly + 4 * u, lx + 8 * v)] = acc[u][v];
}

// lower triangle held in the packed blocks -> global (row-major, strict upper triangle zero-filled), coalesced along j
__device__ __forceinline__ void store_tri(const float* sm, float* __restrict__ out, int64_t ld, int n, int nblk) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = wid; i < n; i += kMidWarps) {
    const int bi = i / MB, r = i % MB;
    for (int bj = 0; bj < nblk; ++bj) {
      const int j = bj * MB + lane;
      if (j < n) out[(int64_t)i * ld + j] = (j <= i) ? sm[bidx(bi, bj) + sk(r, lane)] : 0.f;     // bj <= bi whenever j <= i
    }
  }
}

__global__ void __launch_bounds__(kMidThreads, 1)
potrf_inv_mid_kernel(const float* __restrict__ Ain, int64_t a_ld, int64_t a_bs, float* __restrict__ Lout, int64_t l_ld,
                     int64_t l_bs, float* __restrict__ Wout, int64_t w_ld, int64_t w_bs, int n, float jitter,
                     int32_t* __restrict__ info) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  const int nblk = (n + MB - 1) / MB;
  const int ntot = nblk * (nblk + 1) / 2;
  float* colj = sm + ntot * MB * MB;            // [32]
  float* dinv = colj + MB;                      // [nblk * 32]  1 / L_ii
  int* s_info = reinterpret_cast<int*>(dinv + nblk * MB);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ly = lane >> 3, lx = lane & 7;
  const float* A = Ain + (int64_t)blockIdx.x * a_bs;
  if (tid == 0) *s_info = 0;

  // ---- load: (block, row) items dealt to the warps, lane = column; 8 independent loads in flight per lane ----
  {
    const int items = ntot * MB;
    for (int it0 = wid * 8; it0 < items; it0 += kMidWarps * 8) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int it = it0 + q;
        int b = it / MB, bi = 0;
        const int r = it % MB;
        while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;          // b -> (bi, bj)
        const int bj = b - bi * (bi + 1) / 2;
        const int gi = bi * MB + r, gj = bj * MB + lane;
        v[q] = (gi == gj) ? 1.f : 0.f;
        if (it < items && gi < n && gj <= gi) v[q] = A[(int64_t)gi * a_ld + gj] + ((gi == gj) ? jitter : 0.f);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int it = it0 + q;
        if (it < items) sm[(it / MB) * (MB * MB) + sk(it % MB, lane)] = v[q];
      }
    }
  }
  __syncthreads();

  // ---- factorisation ----
  for (int k = 0; k < nblk; ++k) {
    float* Bkk = sm + bidx(k, k);
    if (wid == 0) {
      // (a) diagonal block: lane i owns row i in registers; every register index is static (fully unrolled)
      float acc[MB];
#pragma unroll
      for (int c = 0; c < MB; ++c) acc[c] = Bkk[sk(lane, c)];
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        const float d = __shfl_sync(0xffffffffu, acc[j], j);
        if (!(d > 0.f) && lane == 0 && k * MB + j < n && *s_info == 0) *s_info = k * MB + j + 1;
        const float dj = sqrtf(d);
        const float inv = 1.f / dj;
        const float lij = (lane == j) ? dj : ((lane > j) ? acc[j] * inv : 0.f);
        acc[j] = lij;
        if (lane == j) dinv[k * MB + j] = inv;
        if (j + 1 < MB) {
          colj[lane] = lij;
          __syncwarp();
#pragma unroll
          for (int c = j + 1; c < MB; ++c)
            if (c <= lane) acc[c] = fmaf(-lij, colj[c], acc[c]);
          __syncwarp();
        }
      }
#pragma unroll
      for (int c = 0; c < MB; ++c) Bkk[sk(lane, c)] = (c <= lane) ? acc[c] : 0.f;
    }
    __syncthreads();
    const int nrem = nblk - k - 1;
    if (nrem == 0) break;
    // (b) panel: thread per row below the diagonal block; x L_kk^T = a by forward substitution
    if (tid < nrem * MB) {
      float* Bik = sm + bidx(k + 1 + tid / MB, k);
      const int r = tid % MB;
      float acc[MB];
#pragma unroll
      for (int c = 0; c < MB; ++c) acc[c] = Bik[sk(r, c)];
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        float s0 = acc[j], s1 = 0.f;
#pragma unroll
        for (int l = 0; l < j; l += 2) {
          s0 = fmaf(-acc[l], Bkk[sk(j, l)], s0);
          if (l + 1 < j) s1 = fmaf(-acc[l + 1], Bkk[sk(j, l + 1)], s1);
        }
        acc[j] = (s0 + s1) * dinv[k * MB + j];
      }
#pragma unroll
      for (int c = 0; c < MB; ++c) Bik[sk(r, c)] = acc[c];
    }
    __syncthreads();
    // (c) trailing update over the block pairs (i, j), k < j <= i
    const int npairs = nrem * (nrem + 1) / 2;
    for (int p = wid; p < npairs; p += kMidWarps) {
      int rb = 0, rem = p;
      while (rem > rb) { rem -= rb + 1; ++rb; }
      const int i = k + 1 + rb, j = k + 1 + rem;
      float acc[8][4];
      float* Bij = sm + bidx(i, j);
      tile_load(Bij, ly, lx, acc);
      tile_mm<true>(sm + bidx(i, k), sm + bidx(j, k), ly, lx, acc, -1.f);
      tile_store(Bij, ly, lx, acc);
    }
    __syncthreads();
  }

  store_tri(sm, Lout + (int64_t)blockIdx.x * l_bs, l_ld, n, nblk);
  __syncthreads();

  // ---- inverse, in place: diagonal blocks (warp per block, lane = column j solves L_kk x = e_j) ----
  for (int k = wid; k < nblk; k += kMidWarps) {
    float* Bkk = sm + bidx(k, k);
    float x[MB];
#pragma unroll
    for (int i = 0; i < MB; ++i) {
      float s0 = (i == lane) ? 1.f : 0.f, s1 = 0.f;
#pragma unroll
      for (int l = 0; l < i; l += 2) {
        s0 = fmaf(-Bkk[sk(i, l)], x[l], s0);
        if (l + 1 < i) s1 = fmaf(-Bkk[sk(i, l + 1)], x[l + 1], s1);
      }
      x[i] = (i >= lane) ? (s0 + s1) * dinv[k * MB + i] : 0.f;
    }
    __syncwarp();                                  // every lane has read the whole block
#pragma unroll
    for (int i = 0; i < MB; ++i) Bkk[sk(i, lane)] = x[i];
  }
  __syncthreads();
  // ---- inverse: merge adjacent inverted block ranges [lo, mid) and [mid, hi):  W21 = -W22 (L21 W11) ----
  for (int s = 1; s < nblk; s *= 2) {
    // tasks of this level: (pair, i in [mid, hi), j in [lo, mid)); a warp computes one 32 x 32 output block in registers
    int ntask = 0;
    for (int lo = 0; lo + s < nblk; lo += 2 * s) ntask += (min(lo + 2 * s, nblk) - (lo + s)) * s;
    for (int phase = 0; phase < 2; ++phase) {
      // phase 0: tmp(i,j) = sum_{l=j}^{mid-1} L(i,l) W(l,j): overwrites L(i,j), which tasks (i, j' < j) still read
      //          -> batches in ascending j.   phase 1: W(i,j) = -sum_{l=mid}^{i} W(i,l) tmp(l,j): overwrites tmp(i,j),
      //          which tasks (i' > i, j) still read -> batches in descending i.
      for (int t0 = 0; t0 < ntask; t0 += kMidWarps) {
        const int t = t0 + wid;
        float acc[8][4];
        float* dst = nullptr;
        if (t < ntask) {
          // decode t: tasks are ordered by the batch key (j ascending / i descending) first, pairs interleaved
          int lo = 0, i = 0, j = 0, tt = t;
          // key index q runs over s columns (phase 0) or up to s rows (phase 1); within a key: pairs, then the other index
          bool found = false;
          const int nkey = s;
          for (int q = 0; q < nkey && !found; ++q) {
            for (int l0 = 0; l0 + s < nblk && !found; l0 += 2 * s) {
              const int mid = l0 + s, hi = min(l0 + 2 * s, nblk), rows = hi - mid;
              int cnt;
              if (phase == 0) cnt = rows;                          // column j = l0 + q, all rows of the pair
              else cnt = (q < rows) ? s : 0;                       // row i = hi - 1 - q, all columns of the pair
              if (tt < cnt) {
                lo = l0;
                if (phase == 0) { j = l0 + q; i = mid + tt; }
                else { i = hi - 1 - q; j = l0 + tt; }
                found = true;
              } else {
                tt -= cnt;
              }
            }
          }
          const int mid = lo + s;
          tile_zero(acc);
          if (phase == 0) {
            for (int l = j; l < mid; ++l) tile_mm<false>(sm + bidx(i, l), sm + bidx(l, j), ly, lx, acc, 1.f);
          } else {
            for (int l = mid; l <= i; ++l) tile_mm<false>(sm + bidx(i, l), sm + bidx(l, j), ly, lx, acc, -1.f);
          }
          dst = sm + bidx(i, j);
        }
        __syncthreads();                           // all reads of this batch are done
        if (dst) tile_store(dst, ly, lx, acc);
        __syncthreads();
      }
    }
  }

  store_tri(sm, Wout + (int64_t)blockIdx.x * w_bs, w_ld, n, nblk);
  if (tid == 0 && info) info[blockIdx.x] = *s_info;
}

}  // namespace vargp

using namespace vargp;

// L = chol(A + jitter I), W = L^-1 for 1 <= n <= 320, whole matrix resident in one CTA's shared memory.
extern "C" int vargp_chol_inv_mid(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                                  float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                                  int32_t* info, void* stream) {
  if (!A || !L || !W || n < 1 || batch < 1 || a_ld < n || l_ld < n || w_ld < n) return VARGP_ERR_ARG;
  if (n > kMidMaxBlk * MB) return VARGP_ERR_UNSUPPORTED;
  if (L == W || A == L || A == W) return VARGP_ERR_ARG;
  const int nblk = (int)ceil_div(n, MB);
  const int dyn = (nblk * (nblk + 1) / 2 * MB * MB + MB + nblk * MB + 4) * (int)sizeof(float);
  static int attr_dyn = 0;
  if (dyn > attr_dyn) {
    cudaError_t e = cudaFuncSetAttribute(potrf_inv_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return (int)e;
    attr_dyn = dyn;
  }
  launch_k(potrf_inv_mid_kernel, dim3((unsigned)batch), dim3(kMidThreads), dyn, (cudaStream_t)stream, A, a_ld, a_bs, L, l_ld, l_bs,
           W, w_ld, w_bs, (int)n, jitter, info);
  return launch_status();
}
