// Batched Cholesky L = chol(A + jitter I) and triangular inverse W = L^-1, one CTA per matrix.
//
// Left-looking over 32-wide block columns.  Thread r owns row r of the current panel: its 32 panel
// entries live in registers, the 32x32 block shared by all rows is staged in shared memory and read as
// warp-broadcast float4 (one wavefront per 4 FMAs x 32 lanes), the diagonal block is factored by one
// warp with register shuffles.  The matrices of the Split/Permuted-MNIST shapes (P <= 1000, 30 of them)
// stay L2-resident; the bound is FMA issue + latency, not HBM (AI = P/24 flop/B).
// Larger P is driven block-wise from the host schedule (GEMM-based trailing updates), which calls these
// kernels on the diagonal blocks only.
#include "common.cuh"

namespace vargp {

constexpr int NB = 32;
constexpr int kCholThreads = 512;

// Thread-per-row kernels need "lane r <- 32 consecutive floats of row r".  Doing that directly costs 32 cache
// lines per warp request; instead the warp reads row after row fully coalesced (one 128 B line per request)
// into its private 32 x 33 smem tile and each lane then picks up its own row (stride 33: conflict-free).
//   base: element (row 0 of the warp's 32-row group, first column); rows_ok rows and cols_ok columns are valid.
__device__ __forceinline__ void warp_load_rows(const float* base, int64_t ld, int rows_ok, int cols_ok,
                                               float (*tile)[NB + 1], float (&a)[NB]) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
#pragma unroll 8
  for (int i = 0; i < NB; ++i)
    tile[i][lane] = (i < rows_ok && lane < cols_ok) ? base[(int64_t)i * ld + lane] : 0.f;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NB; ++c) a[c] = tile[lane][c];
}
__device__ __forceinline__ void warp_store_rows(float* base, int64_t ld, int rows_ok, int cols_ok,
                                                float (*tile)[NB + 1], const float (&a)[NB]) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NB; ++c) tile[lane][c] = a[c];
  __syncwarp();
#pragma unroll 8
  for (int i = 0; i < NB; ++i)
    if (i < rows_ok && lane < cols_ok) base[(int64_t)i * ld + lane] = tile[i][lane];
}

__global__ void __launch_bounds__(kCholThreads)
chol_kernel(const float* Ain, int64_t a_ld, int64_t a_bs, float* Lout, int64_t l_ld,
            int64_t l_bs, int n, float jitter, int32_t* __restrict__ info, int info_base, int accumulate) {
  pdl_enter();
  __shared__ __align__(16) float Bs[NB][NB + 4];   // Bs[kk][c] = L[k0 + c][kc + kk]   (transposed block)
  __shared__ __align__(16) float Ds[NB][NB + 4];   // factored diagonal block, Ds[j][l] = Lkk[j][l]
  __shared__ __align__(16) float colj[NB];         // column j of the diagonal block during its factorisation
  extern __shared__ float dyn_tiles[];             // per-warp 32 x 33 transpose tiles
  float (*tile)[NB + 1] = reinterpret_cast<float (*)[NB + 1]>(dyn_tiles + (threadIdx.x >> 5) * NB * (NB + 1));
  __shared__ float Dinv[NB];
  __shared__ int s_info;

  const float* A = Ain + (int64_t)blockIdx.x * a_bs;
  float* L = Lout + (int64_t)blockIdx.x * l_bs;
  const int tid = threadIdx.x;
  if (tid == 0) s_info = 0;

  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nbk = min(NB, n - k0);     // live columns of this block column
    const int R = n - k0;                // live rows (diag block first)
    for (int r0 = 0; r0 < R; r0 += kCholThreads) {
      const int r = r0 + tid;
      const bool live = r < R;
      const int wrow0 = r0 + (tid & ~31);                   // first panel row of this warp's 32-row group
      const int wrows = min(NB, R - wrow0);                 // <= 0 when the whole warp is past the matrix
      float acc[NB];
      warp_load_rows(A + (int64_t)(k0 + wrow0) * a_ld + k0, a_ld, wrows, nbk, tile, acc);
      if (live && r < NB) {                                 // diagonal entry (static indexing keeps acc in registers)
#pragma unroll
        for (int c = 0; c < NB; ++c)
          if (c == r) acc[c] += jitter;
      }
      // ---- left-looking update: acc[c] -= sum_{k<k0} L[k0+r][k] * L[k0+c][k] ----
      for (int kc = 0; kc < k0; kc += NB) {
        __syncthreads();
        for (int e = tid; e < NB * NB; e += kCholThreads) {
          const int c = e / NB, kk = e % NB;
          Bs[kk][c] = (c < nbk) ? L[(int64_t)(k0 + c) * l_ld + kc + kk] : 0.f;
        }
        __syncthreads();
        if (wrows > 0) {
          float a[NB];
          warp_load_rows(L + (int64_t)(k0 + wrow0) * l_ld + kc, l_ld, wrows, NB, tile, a);
#pragma unroll
          for (int kk = 0; kk < NB; ++kk) {
#pragma unroll
            for (int c4 = 0; c4 < NB; c4 += 4) {
              const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][c4]);
              acc[c4 + 0] = fmaf(-a[kk], b.x, acc[c4 + 0]);
              acc[c4 + 1] = fmaf(-a[kk], b.y, acc[c4 + 1]);
              acc[c4 + 2] = fmaf(-a[kk], b.z, acc[c4 + 2]);
              acc[c4 + 3] = fmaf(-a[kk], b.w, acc[c4 + 3]);
            }
          }
        }
      }
      if (r0 == 0) {
        // ---- factor the 32x32 diagonal block in warp 0: lane i keeps row i in registers.  The column loop
        //      is a runtime loop; element j of a row is read / written with fully unrolled selects, so every
        //      register index is static (dynamic indexing would push the row into local memory).  This block is
        //      the sequential critical path of the kernel: 15 warps wait for it at the barrier below. ----
        if (tid < NB) {
          const int lane = tid;
          if (lane >= nbk) {                                   // virtual identity rows past the matrix edge
#pragma unroll
            for (int c = 0; c < NB; ++c) acc[c] = (c == lane) ? 1.f : 0.f;
          }
#pragma unroll 1
          for (int j = 0; j < NB; ++j) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < NB; ++c) e = (c == j) ? acc[c] : e;
            const float d = __shfl_sync(0xffffffffu, e, j);
            if (!(d > 0.f) && lane == 0 && j < nbk && s_info == 0) s_info = k0 + j + 1;
            const float dj = sqrtf(d);
            const float inv = 1.f / dj;
            const float lij = (lane == j) ? dj : ((lane > j) ? e * inv : 0.f);
            colj[lane] = lij;
            if (lane == j) Dinv[j] = inv;
            __syncwarp();
#pragma unroll
            for (int c4 = 0; c4 < NB; c4 += 4) {
              const float4 l = *reinterpret_cast<const float4*>(&colj[c4]);
              const float lc[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int c = c4 + u;
                if (c > j && c <= lane) acc[c] = fmaf(-lij, lc[u], acc[c]);
                if (c == j) acc[c] = lij;
              }
            }
            __syncwarp();
          }
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            if (c > lane) acc[c] = 0.f;
            Ds[lane][c] = acc[c];
          }
        }
        __syncthreads();
      }
      // ---- panel solve for rows below the diagonal block: x Lkk^T = a ----
      if (live && r >= NB) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          float s = acc[j];
#pragma unroll
          for (int l4 = 0; l4 < j; l4 += 4) {
            const float4 dd = *reinterpret_cast<const float4*>(&Ds[j][l4]);
            s = fmaf(-acc[l4 + 0], dd.x, s);
            if (l4 + 1 < j) s = fmaf(-acc[l4 + 1], dd.y, s);
            if (l4 + 2 < j) s = fmaf(-acc[l4 + 2], dd.z, s);
            if (l4 + 3 < j) s = fmaf(-acc[l4 + 3], dd.w, s);
          }
          acc[j] = s * Dinv[j];
        }
      }
      if (wrows > 0) warp_store_rows(L + (int64_t)(k0 + wrow0) * l_ld + k0, l_ld, wrows, nbk, tile, acc);
    }
    // zero the strict upper part to the right of the diagonal block
    const int ncols = n - (k0 + NB);
    if (ncols > 0) {
      for (int e = tid; e < nbk * ncols; e += kCholThreads) {
        const int rr = e / ncols, cc = e % ncols;
        L[(int64_t)(k0 + rr) * l_ld + k0 + NB + cc] = 0.f;
      }
    }
    __syncthreads();
  }
  if (tid == 0 && info) {
    // blocked drivers factor one diagonal block per call: keep the FIRST failing pivot, in whole-matrix numbering
    if (!accumulate) info[blockIdx.x] = s_info ? s_info + info_base : 0;
    else if (s_info && info[blockIdx.x] == 0) info[blockIdx.x] = s_info + info_base;
  }
}

// W = L^-1 in two kernels.
//  (1) trtri_diag_kernel: one warp per 32x32 diagonal block: W_kk = L_kk^-1 written into W's diagonal blocks.
//  (2) trtri_sweep_kernel: the columns of L^-1 are independent, so ONE CTA PER 32-COLUMN BLOCK of every matrix
//      (grid = column blocks x batch: 300 CTAs at P = 300 instead of 30) walks down its block column:
//          W[k][j] = -W_kk * sum_{l = j..k-1} L[k][l] W[l][j]            (32x32 blocks)
//      The 16 warps split the sum over l; lane c owns column c of the block (W read coalesced, the L block
//      staged per warp in smem and read as broadcasts), partial sums are reduced through shared memory.
constexpr int kDiagWarps = 4;

__global__ void __launch_bounds__(kDiagWarps * 32)
trtri_diag_kernel(const float* Lin, int64_t l_ld, int64_t l_bs, float* Wout, int64_t w_ld, int64_t w_bs, int n,
                  int nblk) {
  pdl_enter();
  __shared__ __align__(16) float Ls[kDiagWarps][NB][NB + 4];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk = blockIdx.x * kDiagWarps + wid;
  if (blk >= nblk) return;
  const float* L = Lin + (int64_t)blockIdx.y * l_bs;
  float* W = Wout + (int64_t)blockIdx.y * w_bs;
  const int k0 = blk * NB, nbk = min(NB, n - k0);
  for (int i = 0; i < NB; ++i) {
    float v = (i == lane) ? 1.f : 0.f;
    if (i < nbk && lane < nbk && lane <= i) v = L[(int64_t)(k0 + i) * l_ld + k0 + lane];
    Ls[wid][i][lane] = v;
  }
  __syncwarp();
  // lane j solves L_kk x = e_j by forward substitution; x lives in registers (static indices only)
  const int j = lane;
  float x[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) x[c] = 0.f;
#pragma unroll 1
  for (int i = 0; i < NB; ++i) {
    float s0 = (i == j) ? 1.f : 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < NB; c4 += 4) {
      const float4 l = *reinterpret_cast<const float4*>(&Ls[wid][i][c4]);
      s0 = fmaf(-((c4 + 0 < i) ? l.x : 0.f), x[c4 + 0], s0);
      s1 = fmaf(-((c4 + 1 < i) ? l.y : 0.f), x[c4 + 1], s1);
      s2 = fmaf(-((c4 + 2 < i) ? l.z : 0.f), x[c4 + 2], s2);
      s3 = fmaf(-((c4 + 3 < i) ? l.w : 0.f), x[c4 + 3], s3);
    }
    const float xi = (i >= j) ? ((s0 + s1) + (s2 + s3)) / Ls[wid][i][i] : 0.f;
#pragma unroll
    for (int c = 0; c < NB; ++c) x[c] = (c == i) ? xi : x[c];
  }
  // lane j holds column j; write rows coalesced through the (now free) tile
  __syncwarp();
#pragma unroll
  for (int i = 0; i < NB; ++i) Ls[wid][i][lane] = x[i];
  __syncwarp();
  for (int i = 0; i < nbk; ++i)
    if (lane < nbk) W[(int64_t)(k0 + i) * w_ld + k0 + lane] = (lane <= i) ? Ls[wid][i][lane] : 0.f;
}

constexpr int kSweepWarps = 16;

__global__ void __launch_bounds__(kSweepWarps * 32)
trtri_sweep_kernel(const float* Lin, int64_t l_ld, int64_t l_bs, float* Wout, int64_t w_ld, int64_t w_bs, int n,
                   int nblk) {
  pdl_enter();
  extern __shared__ float dyn[];
  float (*tile)[NB + 1] = reinterpret_cast<float (*)[NB + 1]>(dyn + (threadIdx.x >> 5) * NB * (NB + 1));  // per warp
  float (*red)[NB][NB + 1] = reinterpret_cast<float (*)[NB][NB + 1]>(dyn + kSweepWarps * NB * (NB + 1));  // [warp][r][c]
  __shared__ float Ps[NB][NB + 1];
  __shared__ float Wk[NB][NB + 1];

  const float* L = Lin + (int64_t)blockIdx.y * l_bs;
  float* W = Wout + (int64_t)blockIdx.y * w_bs;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int jb = blockIdx.x, j0 = jb * NB, ncb = min(NB, n - j0);

  // this CTA owns columns [j0, j0 + ncb): zero everything above its diagonal block
  for (int e = tid; e < j0 * ncb; e += kSweepWarps * 32) {
    const int i = e / ncb, c = e % ncb;
    W[(int64_t)i * w_ld + j0 + c] = 0.f;
  }

  for (int kb = jb + 1; kb < nblk; ++kb) {
    const int k0 = kb * NB, nbk = min(NB, n - k0);
    __syncthreads();                       // rows written in the previous iteration are visible
    // ---- partial products: warp w takes l = jb + w, jb + w + 16, ... ----
    float acc[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) acc[r] = 0.f;
    for (int l = jb + wid; l < kb; l += kSweepWarps) {
      const int l0 = l * NB;
      // stage L[kb][l] (rows k0.., cols l0..) coalesced into the warp's tile: tile[r][kk]
      __syncwarp();
      for (int r = 0; r < NB; ++r) tile[r][lane] = (r < nbk) ? L[(int64_t)(k0 + r) * l_ld + l0 + lane] : 0.f;
      float wv[NB];
#pragma unroll
      for (int kk = 0; kk < NB; ++kk) wv[kk] = (lane < ncb) ? W[(int64_t)(l0 + kk) * w_ld + j0 + lane] : 0.f;
      __syncwarp();
#pragma unroll
      for (int r = 0; r < NB; ++r) {
        float a = acc[r];
#pragma unroll
        for (int kk = 0; kk < NB; ++kk) a = fmaf(tile[r][kk], wv[kk], a);     // tile[r][kk]: warp-wide broadcast
        acc[r] = a;
      }
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) red[wid][r][lane] = acc[r];
    // W_kk (written by trtri_diag_kernel) -> smem
    for (int e = tid; e < NB * NB; e += kSweepWarps * 32) {
      const int r = e / NB, c = e % NB;
      Wk[r][c] = (r < nbk && c <= r) ? W[(int64_t)(k0 + r) * w_ld + k0 + c] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += kSweepWarps * 32) {
      const int r = e / NB, c = e % NB;
      float sacc = 0.f;
#pragma unroll
      for (int w = 0; w < kSweepWarps; ++w) sacc += red[w][r][c];
      Ps[r][c] = sacc;
    }
    __syncthreads();
    // W[kb][jb] = -W_kk * P
    for (int e = tid; e < NB * NB; e += kSweepWarps * 32) {
      const int r = e / NB, c = e % NB;
      if (r < nbk && c < ncb) {
        float sacc = 0.f;
        for (int rp = 0; rp <= r; ++rp) sacc = fmaf(Wk[r][rp], Ps[rp][c], sacc);
        W[(int64_t)(k0 + r) * w_ld + j0 + c] = -sacc;
      }
    }
  }
}

}  // namespace vargp

using namespace vargp;

extern "C" int vargp_chol_ex(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                             int64_t n, int64_t batch, float jitter, int32_t* info, int64_t info_base, int accumulate,
                             void* stream) {
  if (!A || !L || n < 1 || batch < 1 || a_ld < n || l_ld < n) return VARGP_ERR_ARG;
  if (n > (1 << 20)) return VARGP_ERR_UNSUPPORTED;
  static bool attr_set = false;
  const int dyn = (kCholThreads / 32) * NB * (NB + 1) * (int)sizeof(float);     // 67.6 KB of transpose tiles
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  launch_k(chol_kernel, dim3((unsigned)batch), dim3(kCholThreads), dyn, (cudaStream_t)stream, A, a_ld, a_bs, L, l_ld, l_bs, (int)n,
                                                                            jitter, info, (int)info_base, accumulate);
  return launch_status();
}

extern "C" int vargp_chol(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                          int64_t n, int64_t batch, float jitter, int32_t* info, void* stream) {
  return vargp_chol_ex(A, a_ld, a_bs, L, l_ld, l_bs, n, batch, jitter, info, 0, 0, stream);
}

extern "C" int vargp_trtri(const float* L, int64_t l_ld, int64_t l_bs, float* W, int64_t w_ld, int64_t w_bs,
                           int64_t n, int64_t batch, void* stream) {
  if (!L || !W || n < 1 || batch < 1 || l_ld < n || w_ld < n) return VARGP_ERR_ARG;
  if (L == W) return VARGP_ERR_ARG;
  const int nblk = (int)ceil_div(n, NB);
  if (batch > 65535) return VARGP_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  launch_k(trtri_diag_kernel, dim3(dim3((unsigned)ceil_div(nblk, kDiagWarps), (unsigned)batch)), dim3(kDiagWarps * 32), 0, s, 
      L, l_ld, l_bs, W, w_ld, w_bs, (int)n, nblk);
  int rc = launch_status();
  if (rc) return rc;
  static bool attr_set = false;
  const int dyn = (kSweepWarps * NB * (NB + 1) * 2) * (int)sizeof(float);       // per-warp tiles + reduction buffer
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(trtri_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  launch_k(trtri_sweep_kernel, dim3(dim3((unsigned)nblk, (unsigned)batch)), dim3(kSweepWarps * 32), dyn, s, L, l_ld, l_bs, W, w_ld, w_bs,
                                                                                        (int)n, nblk);
  return launch_status();
}
