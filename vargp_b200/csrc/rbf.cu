// ARD-RBF kernel: operand scaling / norms (forward) and the adjoint passes that turn Kbar into
// gradients for z, x and theta.  All of these are streaming (HBM / L2 bound) SIMT kernels; the only
// dense contractions of the RBF path (x.z^T forward, W.X backward) are GEMMs (gemm_simt.cu / gemm_tc.cu).
#include <cstdlib>

#include "common.cuh"

namespace vargp {

// ---------------------------------------------------------------------------------------------
// dst[h][r][d] = src[r][d] * exp(-theta[h][d]); norms[h][r] = sum_d dst^2     (kernels.py:41-44,50,54)
// grid (row blocks, H); one warp per row, lanes stride over d (coalesced 128 B per request).
// Algorithmic bytes: 4*(R*D read + H*R*D written + H*R).
// ---------------------------------------------------------------------------------------------
constexpr int kScaleWarps = 8;
constexpr int kScaleRowsPerWarp = 4;

__global__ void __launch_bounds__(kScaleWarps * 32)
scale_rows_kernel(const float* __restrict__ src, int64_t R, int64_t D, int64_t src_rs,
                  const float* __restrict__ theta, int64_t theta_rs,
                  float* __restrict__ dst, float* __restrict__ norms) {
  pdl_enter();
  extern __shared__ __align__(16) float s_isig[];
  const int h = blockIdx.y;
  for (int64_t d = threadIdx.x; d < D; d += blockDim.x) s_isig[d] = expf(-theta[h * theta_rs + d]);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t row0 = ((int64_t)blockIdx.x * kScaleWarps + wid) * kScaleRowsPerWarp;
  const bool vec4 = (D % 4 == 0) && (src_rs % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0);
#pragma unroll
  for (int rr = 0; rr < kScaleRowsPerWarp; ++rr) {
    const int64_t r = row0 + rr;
    if (r >= R) break;
    const float* sp = src + r * src_rs;
    float* dp = dst + ((int64_t)h * R + r) * D;
    float acc = 0.f;
    if (vec4) {
      for (int64_t d = lane * 4; d < D; d += 128) {
        float4 v = *reinterpret_cast<const float4*>(sp + d);
        const float4 s = *reinterpret_cast<const float4*>(s_isig + d);
        v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
        *reinterpret_cast<float4*>(dp + d) = v;
      }
    } else {
      for (int64_t d = lane; d < D; d += 32) {
        const float v = sp[d] * s_isig[d];
        acc = fmaf(v, v, acc);
        dp[d] = v;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) norms[(int64_t)h * R + r] = acc;
  }
}

// Same contract for D <= 1024 (16 B-aligned rows): the source row is read ONCE for all H hyper samples and all of its 16 B
// chunks are in flight together -- the kernel above re-reads it per sample and walks a row in D / 128 dependent round trips
// (7 at D = 784), which made the z-side scaling a 23 us latency chain at the head of the Split-MNIST step.
constexpr int kScaleChunks = 8;           // 8 x 128 floats per row
constexpr int kFusedRows = 2;             // rows per warp (16 per CTA: 188 CTAs for the 3000 inducing rows of the Split-MNIST step)
__global__ void __launch_bounds__(kScaleWarps * 32)
scale_rows_fused_kernel(const float* __restrict__ src, int64_t R, int64_t D, int64_t src_rs,
                        const float* __restrict__ theta, int64_t theta_rs, int H,
                        float* __restrict__ dst, float* __restrict__ norms) {
  pdl_enter();
  extern __shared__ __align__(16) float s_isig[];          // [H][D]
  for (int64_t i = threadIdx.x; i < (int64_t)H * D; i += blockDim.x) {
    const int64_t hh = i / D, d = i - hh * D;
    s_isig[i] = expf(-theta[hh * theta_rs + d]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t row0 = ((int64_t)blockIdx.x * kScaleWarps + wid) * kFusedRows;
#pragma unroll 1
  for (int rr = 0; rr < kFusedRows; ++rr) {
    const int64_t r = row0 + rr;
    if (r >= R) break;
    const float* sp = src + r * src_rs;
    float4 v[kScaleChunks];
#pragma unroll
    for (int i = 0; i < kScaleChunks; ++i) {
      const int64_t d = lane * 4 + 128 * i;
      v[i] = d < D ? *reinterpret_cast<const float4*>(sp + d) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int hh = 0; hh < H; ++hh) {
      float* dp = dst + ((int64_t)hh * R + r) * D;
      const float* sg = s_isig + (int64_t)hh * D;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < kScaleChunks; ++i) {
        const int64_t d = lane * 4 + 128 * i;
        if (d < D) {
          const float4 s = *reinterpret_cast<const float4*>(sg + d);
          float4 w = v[i];
          w.x *= s.x; w.y *= s.y; w.z *= s.z; w.w *= s.w;
          acc = fmaf(w.x, w.x, acc); acc = fmaf(w.y, w.y, acc); acc = fmaf(w.z, w.z, acc); acc = fmaf(w.w, w.w, acc);
          *reinterpret_cast<float4*>(dp + d) = w;
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) norms[(int64_t)hh * R + r] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kbar <- Kbar * K; rsum[g][i] = sum_j; csum[h][j] += sum_{c,i}            (W = Kbar (.) K)
// grid (column tiles, row chunks, G).  A block owns `rows_blk` rows x (256 * VEC) columns; a thread keeps VEC
// columns: its column sums stay in registers over the whole row chunk (one atomic per column and block), the row
// sums go warp-shuffle -> shared memory -> one atomic per row and block.  Rows are processed four at a time so that
// 8 independent 16 B loads per thread are in flight (this kernel streams 3 x 16 GB at the scaled config).
// Algorithmic bytes: 4 * 3 * G * Pa * Pb.
// ---------------------------------------------------------------------------------------------
constexpr int kPrepThreads = 256;
constexpr int kPrepRU = 4;           // rows in flight
constexpr int kPrepRS = 32;          // rows per shared-memory reduction round

template <int VEC>
__global__ void __launch_bounds__(kPrepThreads)
rbf_bwd_prep_kernel(float* __restrict__ Kbar, const float* __restrict__ K, int64_t C, int64_t Pa, int64_t Pb,
                    int rows_blk, float* __restrict__ rsum, float* __restrict__ csum, float* __restrict__ dsum) {
  pdl_enter();
  __shared__ float s_red[kPrepRS][kPrepThreads / 32 + 1];
  const int64_t g = blockIdx.z;
  const int64_t h = g / C;
  const int64_t j0 = ((int64_t)blockIdx.x * kPrepThreads + threadIdx.x) * VEC;
  const int64_t i_lo = (int64_t)blockIdx.y * rows_blk;
  const int64_t i_hi = min(Pa, i_lo + rows_blk);
  const bool live = j0 < Pb;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float cacc[VEC];
#pragma unroll
  for (int u = 0; u < VEC; ++u) cacc[u] = 0.f;

  for (int64_t ib = i_lo; ib < i_hi; ib += kPrepRS) {
    const int nr = (int)min((int64_t)kPrepRS, i_hi - ib);
#pragma unroll 1
    for (int r0 = 0; r0 < nr; r0 += kPrepRU) {
      float w[kPrepRU][VEC];
#pragma unroll
      for (int r = 0; r < kPrepRU; ++r) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) w[r][u] = 0.f;
        if (live && r0 + r < nr) {
          const int64_t o = (g * Pa + ib + r0 + r) * Pb + j0;
          if (VEC == 4) {
            const float4 a = *reinterpret_cast<const float4*>(Kbar + o);
            const float4 k = *reinterpret_cast<const float4*>(K + o);
            w[r][0] = a.x * k.x; w[r][1] = a.y * k.y; w[r][2] = a.z * k.z; w[r][3] = a.w * k.w;
          } else {
            w[r][0] = Kbar[o] * K[o];
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kPrepRU; ++r) {
        const int64_t i = ib + r0 + r;
        float rpart = 0.f;
        if (live && r0 + r < nr) {
          if (dsum) {                 // symmetric Gram: the diagonal only carries the gamma gradient
#pragma unroll
            for (int u = 0; u < VEC; ++u)
              if (j0 + u == i) {
                dsum[g * Pa + i] = w[r][u];
                w[r][u] = 0.f;
              }
          }
          const int64_t o = (g * Pa + i) * Pb + j0;
          if (VEC == 4) *reinterpret_cast<float4*>(Kbar + o) = make_float4(w[r][0], w[r][1], w[r][2], w[r][3]);
          else Kbar[o] = w[r][0];
#pragma unroll
          for (int u = 0; u < VEC; ++u) {
            cacc[u] += w[r][u];
            rpart += w[r][u];
          }
        }
        rpart = warp_sum(rpart);
        if (lane == 0 && r0 + r < nr) s_red[r0 + r][wid] = rpart;
      }
    }
    __syncthreads();
    if (threadIdx.x < nr) {
      float v = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < kPrepThreads / 32; ++w8) v += s_red[threadIdx.x][w8];
      if (gridDim.x == 1) rsum[g * Pa + ib + threadIdx.x] = v;
      else atomicAdd(rsum + g * Pa + ib + threadIdx.x, v);
    }
    __syncthreads();
  }
  if (csum && live) {
#pragma unroll
    for (int u = 0; u < VEC; ++u) atomicAdd(csum + h * Pb + j0 + u, cacc[u]);
  }
}

// Same contract for narrow matrices (Pb <= 32 * VEC * kRowsK columns: the Split / Permuted-MNIST shapes).  The
// column-per-thread kernel above leaves most lanes idle there (Pb = 300: the second 256-column tile is 83 % empty),
// needs atomics + a memset for the row sums and walks its rows serially.  Here a WARP owns a row (lanes stride over
// float4 column groups), row sums are one warp reduction and a plain store (deterministic), and the column sums of the
// CTA's 32 rows are combined in shared memory before they go out as one atomic per column and CTA.
// grid (row blocks of 32, G); 8 warps x 4 rows.
constexpr int kRowsK = 8;            // column groups per lane
template <int VEC>
__global__ void __launch_bounds__(256)
rbf_bwd_prep_rows_kernel(float* __restrict__ Kbar, const float* __restrict__ K, int64_t C, int64_t Pa, int64_t Pb,
                         float* __restrict__ rsum, float* __restrict__ csum, float* __restrict__ dsum) {
  pdl_enter();
  extern __shared__ __align__(16) float s_c[];      // [8][Pb] when csum != nullptr
  const int64_t g = blockIdx.y;
  const int64_t h = g / C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float cacc[kRowsK][VEC];
#pragma unroll
  for (int k = 0; k < kRowsK; ++k)
#pragma unroll
    for (int u = 0; u < VEC; ++u) cacc[k][u] = 0.f;
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
    const int64_t i = (int64_t)blockIdx.x * 32 + wid * 4 + r;
    if (i >= Pa) break;                                 // warp-uniform
    float* kb = Kbar + (g * Pa + i) * Pb;
    const float* kk = K + (g * Pa + i) * Pb;
    float w[kRowsK][VEC];
#pragma unroll
    for (int k = 0; k < kRowsK; ++k) {
      const int64_t j0 = ((int64_t)k * 32 + lane) * VEC;
#pragma unroll
      for (int u = 0; u < VEC; ++u) w[k][u] = 0.f;
      if (j0 < Pb) {
        if (VEC == 4) {
          const float4 a = *reinterpret_cast<const float4*>(kb + j0);
          const float4 b = *reinterpret_cast<const float4*>(kk + j0);
          w[k][0] = a.x * b.x; w[k][1] = a.y * b.y; w[k][2] = a.z * b.z; w[k][3] = a.w * b.w;
        } else {
          w[k][0] = kb[j0] * kk[j0];
        }
      }
    }
    float rpart = 0.f;
#pragma unroll
    for (int k = 0; k < kRowsK; ++k) {
      const int64_t j0 = ((int64_t)k * 32 + lane) * VEC;
      if (j0 < Pb) {
        if (dsum) {                 // symmetric Gram: the diagonal only carries the gamma gradient
#pragma unroll
          for (int u = 0; u < VEC; ++u)
            if (j0 + u == i) {
              dsum[g * Pa + i] = w[k][u];
              w[k][u] = 0.f;
            }
        }
        if (VEC == 4) *reinterpret_cast<float4*>(kb + j0) = make_float4(w[k][0], w[k][1], w[k][2], w[k][3]);
        else kb[j0] = w[k][0];
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
          cacc[k][u] += w[k][u];
          rpart += w[k][u];
        }
      }
    }
    rpart = warp_sum(rpart);
    if (lane == 0) rsum[g * Pa + i] = rpart;
  }
  if (csum) {
#pragma unroll
    for (int k = 0; k < kRowsK; ++k) {
      const int64_t j0 = ((int64_t)k * 32 + lane) * VEC;
      if (j0 < Pb) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) s_c[wid * Pb + j0 + u] = cacc[k][u];
      }
    }
    __syncthreads();
    for (int64_t j = threadIdx.x; j < Pb; j += 256) {
      float v = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) v += s_c[w8 * Pb + j];
      atomicAdd(csum + h * Pb + j, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// zs_bar = -(r1 + 2 r2) zs + Gz1 + 2 Gz2 ; Zbar[c][i][d] = sum_h zs_bar exp(-theta[h][d])
// theta_bar[h][d] += sum_{c,i} (-zs zs_bar - zs Gz1) ; theta_bar[h][D] += 2 sum_{c,i} (r1 + r2)
// grid (D tiles, row tiles over C*P); thread = one d, loops h (outer) and the block's rows (inner).
// ---------------------------------------------------------------------------------------------
constexpr int kFinRows = 8;       // rows held in registers at a time
constexpr int kFinChunks = 1;     // row chunks per CTA (3 was measured 2x slower at the Split-MNIST shape: the kernel is latency-bound, it needs the CTAs)
constexpr int kFinThreads = 128;

__global__ void __launch_bounds__(kFinThreads)
rbf_bwd_finish_kernel(const float* __restrict__ zs, const float* __restrict__ Gz1, const float* __restrict__ Gz2,
                      const float* __restrict__ r1, const float* __restrict__ r2, const float* __restrict__ dg,
                      const float* __restrict__ theta, int64_t theta_rs, int64_t H, int64_t R, int64_t D,
                      float* __restrict__ Zbar, float* __restrict__ theta_bar) {
  pdl_enter();
  const int64_t d = (int64_t)blockIdx.x * kFinThreads + threadIdx.x;
  const int64_t blk0 = (int64_t)blockIdx.y * kFinRows * kFinChunks;
  constexpr int kMaxH = 4;          // per-h theta accumulators kept across the chunks when H is small
  float tsum[kMaxH];
#pragma unroll
  for (int h = 0; h < kMaxH; ++h) tsum[h] = 0.f;
  for (int ch = 0; ch < kFinChunks; ++ch) {
    const int64_t row0 = blk0 + (int64_t)ch * kFinRows;
    if (row0 >= R) break;
    const int rows = (int)min((int64_t)kFinRows, R - row0);
    float zacc[kFinRows];
#pragma unroll
    for (int r = 0; r < kFinRows; ++r) zacc[r] = 0.f;
    if (d < D) {
      for (int64_t h = 0; h < H; ++h) {
        const float isig = expf(-theta[h * theta_rs + d]);
        float tacc = 0.f;
#pragma unroll
        for (int r = 0; r < kFinRows; ++r) {
          if (r < rows) {
            const int64_t row = h * R + row0 + r;
            const float z = zs[row * D + d];
            const float g1 = Gz1 ? Gz1[row * D + d] : 0.f;
            const float g2 = Gz2 ? Gz2[row * D + d] : 0.f;
            const float rr = (r1 ? r1[row] : 0.f) + 2.f * (r2 ? r2[row] : 0.f);
            const float zb = fmaf(-rr, z, g1 + 2.f * g2);
            zacc[r] = fmaf(zb, isig, zacc[r]);
            tacc -= z * (zb + g1);
          }
        }
        if (H <= kMaxH) {
#pragma unroll
          for (int hh = 0; hh < kMaxH; ++hh)
            if (hh == h) tsum[hh] += tacc;
        } else {
          atomicAdd(theta_bar + h * (D + 1) + d, tacc);
        }
      }
#pragma unroll
      for (int r = 0; r < kFinRows; ++r)
        if (r < rows) Zbar[(row0 + r) * D + d] = zacc[r];
    }
    if (blockIdx.x == 0 && threadIdx.x < rows) {
      for (int64_t h = 0; h < H; ++h) {
        const int64_t row = h * R + row0 + threadIdx.x;
        atomicAdd(theta_bar + h * (D + 1) + D, 2.f * ((r1 ? r1[row] : 0.f) + (r2 ? r2[row] : 0.f) + (dg ? dg[row] : 0.f)));
      }
    }
  }
  if (d < D && H <= kMaxH) {
#pragma unroll
    for (int hh = 0; hh < kMaxH; ++hh)
      if (hh < H) atomicAdd(theta_bar + hh * (D + 1) + d, tsum[hh]);
  }
}

// Same contract, float4 along d, for H <= 4 and D % 4 == 0.  The kernel above issues one atomic per (thread, h) on
// theta_bar: 375 CTAs hammer the same 3 x 784 addresses at the Split-MNIST shape (50 us for 85 MB that sit in L2).
// Here a CTA covers 32 rows x 128 d (8 warps x 4 rows, lanes x float4), reduces its theta partial sums over the 8 warps
// in shared memory and issues one atomic per (h, d) and CTA.
// grid (D tiles of 128, row blocks of 32).
constexpr int kFin4Rows = 4;       // rows per warp
__global__ void __launch_bounds__(256)
rbf_bwd_finish_v4_kernel(const float* __restrict__ zs, const float* __restrict__ Gz1, const float* __restrict__ Gz2,
                         const float* __restrict__ r1, const float* __restrict__ r2, const float* __restrict__ dg,
                         const float* __restrict__ theta, int64_t theta_rs, int H, int64_t R, int64_t D,
                         float* __restrict__ Zbar, float* __restrict__ theta_bar) {
  pdl_enter();
  __shared__ float4 s_t[8][4][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t d = (int64_t)blockIdx.x * 128 + lane * 4;
  const int64_t row0 = (int64_t)blockIdx.y * 32 + wid * kFin4Rows;
  const bool dl = d < D;
  float4 zacc[kFin4Rows], tacc[4];
#pragma unroll
  for (int r = 0; r < kFin4Rows; ++r) zacc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int h = 0; h < 4; ++h) tacc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (dl) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      if (h < H) {
        const float* th = theta + h * theta_rs + d;          // rows of theta are D + 1 long: not 16 B aligned
        const float4 isig = make_float4(expf(-__ldg(th)), expf(-__ldg(th + 1)), expf(-__ldg(th + 2)), expf(-__ldg(th + 3)));
        float4 z[kFin4Rows], g1[kFin4Rows], g2[kFin4Rows];
        float rr[kFin4Rows];
#pragma unroll
        for (int r = 0; r < kFin4Rows; ++r) {
          const int64_t row = (int64_t)h * R + row0 + r;
          const bool ok = row0 + r < R;
          const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
          z[r] = ok ? *reinterpret_cast<const float4*>(zs + row * D + d) : zero;
          g1[r] = (ok && Gz1) ? *reinterpret_cast<const float4*>(Gz1 + row * D + d) : zero;
          g2[r] = (ok && Gz2) ? *reinterpret_cast<const float4*>(Gz2 + row * D + d) : zero;
          rr[r] = ok ? ((r1 ? __ldg(r1 + row) : 0.f) + 2.f * (r2 ? __ldg(r2 + row) : 0.f)) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < kFin4Rows; ++r) {
          float4 zb;
          zb.x = fmaf(-rr[r], z[r].x, g1[r].x + 2.f * g2[r].x);
          zb.y = fmaf(-rr[r], z[r].y, g1[r].y + 2.f * g2[r].y);
          zb.z = fmaf(-rr[r], z[r].z, g1[r].z + 2.f * g2[r].z);
          zb.w = fmaf(-rr[r], z[r].w, g1[r].w + 2.f * g2[r].w);
          zacc[r].x = fmaf(zb.x, isig.x, zacc[r].x); zacc[r].y = fmaf(zb.y, isig.y, zacc[r].y);
          zacc[r].z = fmaf(zb.z, isig.z, zacc[r].z); zacc[r].w = fmaf(zb.w, isig.w, zacc[r].w);
          tacc[h].x -= z[r].x * (zb.x + g1[r].x); tacc[h].y -= z[r].y * (zb.y + g1[r].y);
          tacc[h].z -= z[r].z * (zb.z + g1[r].z); tacc[h].w -= z[r].w * (zb.w + g1[r].w);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kFin4Rows; ++r)
      if (row0 + r < R) *reinterpret_cast<float4*>(Zbar + (row0 + r) * D + d) = zacc[r];
  }
#pragma unroll
  for (int h = 0; h < 4; ++h) s_t[wid][h][lane] = tacc[h];
  __syncthreads();
  if (wid < H && dl) {                   // warp h sums the 8 partials of hyper sample h
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) {
      const float4 v = s_t[w8][wid][lane];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    float* tb = theta_bar + (int64_t)wid * (D + 1) + d;
    atomicAdd(tb, a.x); atomicAdd(tb + 1, a.y); atomicAdd(tb + 2, a.z); atomicAdd(tb + 3, a.w);
  }
  if (blockIdx.x == 0 && wid >= 4 && wid - 4 < H) {      // gamma part: 2 sum_rows (r1 + r2 + dg), one atomic per (CTA, h)
    const int h = wid - 4;
    const int64_t rloc = (int64_t)blockIdx.y * 32 + lane;
    float v = 0.f;
    if (rloc < R) {
      const int64_t row = (int64_t)h * R + rloc;
      v = (r1 ? r1[row] : 0.f) + (r2 ? r2[row] : 0.f) + (dg ? dg[row] : 0.f);
    }
    v = warp_sum(v);
    if (lane == 0) atomicAdd(theta_bar + (int64_t)h * (D + 1) + D, 2.f * v);
  }
}

// theta_bar[h][d] += sum_j csum[h][j] xs[h][j][d]^2     grid (D tiles, j tiles, H)
constexpr int kXsRows = 16;
__global__ void __launch_bounds__(128)
rbf_bwd_xside_theta_kernel(const float* __restrict__ xs, const float* __restrict__ csum, int64_t B, int64_t D,
                           float* __restrict__ theta_bar) {
  pdl_enter();
  const int64_t d = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int64_t h = blockIdx.z;
  const int64_t j0 = (int64_t)blockIdx.y * kXsRows;
  const int64_t j1 = min(B, j0 + kXsRows);
  if (d >= D) return;
  float acc = 0.f;
#pragma unroll 8
  for (int64_t j = j0; j < j1; ++j) {
    const float v = xs[(h * B + j) * D + d];
    acc = fmaf(csum[h * B + j] * v, v, acc);
  }
  atomicAdd(theta_bar + h * (D + 1) + d, acc);
}

// xbar[j][d] = sum_h (-csum[h][j] xs[h][j][d] + sum_c Gx[h][c][j][d]) exp(-theta[h][d])
__global__ void __launch_bounds__(128)
rbf_bwd_xside_x_kernel(const float* __restrict__ xs, const float* __restrict__ csum, const float* __restrict__ Gx,
                       const float* __restrict__ theta, int64_t theta_rs, int64_t H, int64_t C, int64_t B, int64_t D,
                       float* __restrict__ xbar) {
  pdl_enter();
  const int64_t d = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (d >= D) return;
  float acc = 0.f;
  for (int64_t h = 0; h < H; ++h) {
    float v = -csum[h * B + j] * xs[(h * B + j) * D + d];
    for (int64_t c = 0; c < C; ++c) v += Gx[(((h * C + c) * B) + j) * D + d];
    acc = fmaf(v, expf(-theta[h * theta_rs + d]), acc);
  }
  xbar[j * D + d] = acc;
}

}  // namespace vargp

using namespace vargp;

extern "C" int vargp_scale_rows(const float* src, int64_t R, int64_t D, int64_t src_rs,
                                const float* theta, int64_t H, int64_t theta_rs,
                                float* dst, float* norms, void* stream) {
  if (!src || !theta || !dst || !norms || R < 0 || D < 1 || H < 1) return VARGP_ERR_ARG;
  if (R == 0) return 0;
  if (D * sizeof(float) > 48 * 1024) return VARGP_ERR_UNSUPPORTED;
  const bool vec4 = (D % 4 == 0) && (src_rs % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0);
  if (vec4 && D <= 128 * kScaleChunks && (size_t)H * D * sizeof(float) <= 48 * 1024) {
    dim3 fgrid((unsigned)ceil_div(R, kScaleWarps * kFusedRows), 1u);
    launch_k(scale_rows_fused_kernel, dim3(fgrid), dim3(kScaleWarps * 32), (size_t)H * D * sizeof(float), (cudaStream_t)stream,
             src, R, D, src_rs, theta, theta_rs, (int)H, dst, norms);
    return launch_status();
  }
  dim3 grid((unsigned)ceil_div(R, kScaleWarps * kScaleRowsPerWarp), (unsigned)H);
  launch_k(scale_rows_kernel, dim3(grid), dim3(kScaleWarps * 32), D * sizeof(float), (cudaStream_t)stream, 
      src, R, D, src_rs, theta, theta_rs, dst, norms);
  return launch_status();
}

extern "C" int vargp_rbf_bwd_prep(float* Kbar, const float* K, int64_t H, int64_t C, int64_t Pa, int64_t Pb,
                                  float* rsum, float* csum, float* dsum, void* stream) {
  if (!Kbar || !K || !rsum || H < 1 || C < 1 || Pa < 1 || Pb < 1) return VARGP_ERR_ARG;
  if (H * C > 65535) return VARGP_ERR_UNSUPPORTED;
  if (dsum && Pa != Pb) return VARGP_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t G = H * C;
  static int rows_on = -1;
  if (rows_on < 0) { const char* e = getenv("VARGP_RBF_ROWS"); rows_on = e ? atoi(e) : 1; }
  if (rows_on) {   // narrow matrices: warp-per-row kernel (no memset, no row-sum atomics)
    const bool a16 = ((reinterpret_cast<uintptr_t>(Kbar) | reinterpret_cast<uintptr_t>(K)) % 16 == 0) && Pb % 4 == 0;
    const int64_t cap = 32 * kRowsK * (a16 ? 4 : 1);
    if (Pb <= cap && ceil_div(Pa, 32) <= 65535 && G * Pa * Pb <= (int64_t)64 << 20) {
      dim3 grid((unsigned)ceil_div(Pa, 32), (unsigned)G);
      const size_t smem = csum ? sizeof(float) * 8 * Pb : 0;
      if (a16) launch_k((rbf_bwd_prep_rows_kernel<4>), dim3(grid), dim3(256), smem, s, Kbar, K, C, Pa, Pb, rsum, csum, dsum);
      else launch_k((rbf_bwd_prep_rows_kernel<1>), dim3(grid), dim3(256), smem, s, Kbar, K, C, Pa, Pb, rsum, csum, dsum);
      return launch_status();
    }
  }
  const bool vec4 = Pb % 4 == 0 && Pb >= 2048 &&
                    ((reinterpret_cast<uintptr_t>(Kbar) | reinterpret_cast<uintptr_t>(K)) % 16 == 0);
  const int64_t ctile = kPrepThreads * (vec4 ? 4 : 1);
  const int64_t ctiles = ceil_div(Pb, ctile);
  // row chunks: few (column-sum atomics scale with them), but enough blocks for ~8 full waves of 8 blocks per SM so
  // that the tail wave does not matter (measured at the scaled config: 1.6 waves ran at 61 % of HBM peak)
  int64_t chunks = ceil_div(148 * 8 * 8, ctiles * G);
  int64_t rows_blk = ceil_div(ceil_div(Pa, chunks), kPrepRS) * kPrepRS;
  chunks = ceil_div(Pa, rows_blk);
  if (chunks > 65535) return VARGP_ERR_UNSUPPORTED;
  if (ctiles > 1) {                       // several column tiles accumulate into rsum
    cudaError_t e = cudaMemsetAsync(rsum, 0, sizeof(float) * G * Pa, s);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid((unsigned)ctiles, (unsigned)chunks, (unsigned)G);
  if (vec4) launch_k((rbf_bwd_prep_kernel<4>), dim3(grid), dim3(kPrepThreads), 0, s, Kbar, K, C, Pa, Pb, (int)rows_blk, rsum, csum, dsum);
  else launch_k((rbf_bwd_prep_kernel<1>), dim3(grid), dim3(kPrepThreads), 0, s, Kbar, K, C, Pa, Pb, (int)rows_blk, rsum, csum, dsum);
  return launch_status();
}

extern "C" int vargp_rbf_bwd_finish(const float* zs, const float* Gz1, const float* Gz2, const float* r1,
                                    const float* r2, const float* dg, const float* theta, int64_t theta_rs, int64_t H, int64_t C,
                                    int64_t P, int64_t D, float* Zbar, float* theta_bar, void* stream) {
  if (!zs || !theta || !Zbar || !theta_bar) return VARGP_ERR_ARG;
  if ((Gz1 != nullptr) != (r1 != nullptr) || (Gz2 != nullptr) != (r2 != nullptr)) return VARGP_ERR_ARG;
  const int64_t R = C * P;
  static int v4_on = -1;
  if (v4_on < 0) { const char* e = getenv("VARGP_RBF_FIN4"); v4_on = e ? atoi(e) : 1; }
  if (v4_on) {
    uintptr_t bits = reinterpret_cast<uintptr_t>(zs) | reinterpret_cast<uintptr_t>(Zbar);
    if (Gz1) bits |= reinterpret_cast<uintptr_t>(Gz1);
    if (Gz2) bits |= reinterpret_cast<uintptr_t>(Gz2);
    if (H <= 4 && D % 4 == 0 && bits % 16 == 0 && ceil_div(R, 32) <= 65535) {
      dim3 grid((unsigned)ceil_div(D, 128), (unsigned)ceil_div(R, 32));
      launch_k(rbf_bwd_finish_v4_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, zs, Gz1, Gz2, r1, r2, dg, theta, theta_rs,
               (int)H, R, D, Zbar, theta_bar);
      return launch_status();
    }
  }
  if (ceil_div(R, kFinRows * kFinChunks) > 65535) return VARGP_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(D, kFinThreads), (unsigned)ceil_div(R, kFinRows * kFinChunks));
  launch_k(rbf_bwd_finish_kernel, dim3(grid), dim3(kFinThreads), 0, (cudaStream_t)stream, zs, Gz1, Gz2, r1, r2, dg, theta, theta_rs, H, R,
                                                                         D, Zbar, theta_bar);
  return launch_status();
}

extern "C" int vargp_rbf_bwd_xside(const float* xs, const float* csum, const float* Gx, const float* theta,
                                   int64_t theta_rs, int64_t H, int64_t C, int64_t B, int64_t D,
                                   float* theta_bar, float* xbar, void* stream) {
  if (!xs || !csum || !theta || !theta_bar) return VARGP_ERR_ARG;
  if ((xbar != nullptr) != (Gx != nullptr)) return VARGP_ERR_ARG;
  if (ceil_div(B, kXsRows) > 65535 || H > 65535 || B > (int64_t)65535 * kXsRows) return VARGP_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(D, 128), (unsigned)ceil_div(B, kXsRows), (unsigned)H);
  launch_k(rbf_bwd_xside_theta_kernel, dim3(grid), dim3(128), 0, s, xs, csum, B, D, theta_bar);
  int rc = launch_status();
  if (rc || !xbar) return rc;
  if (B > 65535) return VARGP_ERR_UNSUPPORTED;
  dim3 grid2((unsigned)ceil_div(D, 128), (unsigned)B);
  launch_k(rbf_bwd_xside_x_kernel, dim3(grid2), dim3(128), 0, s, xs, csum, Gx, theta, theta_rs, H, C, B, D, xbar);
  return launch_status();
}
