// Cluster-cooperative Cholesky + triangular inverse for the Split-MNIST sizes (32 < n <= 320), one thread-block CLUSTER
// (2 or 4 CTAs on 2 or 4 SMs) per matrix:
//     L = chol(A + jitter I),   W = L^-1            (var_gp/gp_utils.py:5-11 and every triangular_solve behind it)
//
// Why: the one-CTA kernel (potrf_small.cu) uses 30 of 148 SMs for the 30 = H*C matrices of a step and the blocked driver
// (potrf_blocked.cu) needs 3 such launches + 8 dependent tensor-core GEMMs at P = 300 (265 us of a 990 us step).  Here
// the whole factorisation AND the inverse are one launch on 120 SMs, nothing leaves the chip between the load of A and
// the store of L and W, and the inverse rides on the factorisation sweep instead of following it.
//
// Layout: 32 x 32 blocks; block row i lives in the shared memory of CTA (i mod CS) of the cluster, lower triangle only
// (blocks j <= i), once for L (first A) and once for X -> W (first I).  Right-looking sweep over block columns k:
//   (A) the owner of row k factors the diagonal block out of registers (lane = row, pivots through warp shuffles,
//       rsqrt + one Newton step) and inverts it (lane = column, forward substitution); D_k^T is pushed into every CTA
//       through distributed shared memory.  This happens one step AHEAD, in the shadow of phase (C) of step k-1.
//   (B) panel: every CTA forms L_ik = A_ik D_k^T for its rows i > k and pushes L_ik^T to the CTAs that need it; the owner
//       of k also finishes row k of the inverse, W_kj = D_k X_kj (j <= k), and pushes it.
//   (C) trailing update, one warp per block:  A_ij -= L_ik L_jk^T (k < j <= i)  and  X_ij -= L_ik W_kj (j <= k)
//       -- the same 32x32x32 register-tiled product against a broadcast operand; every block of a row is touched at every
//       step, so the factorisation (shrinking) and the inverse (growing) add up to a constant load per row.
// Two cluster barriers per step.  Roofline: fp32 FMA issue + the sequential diagonal chain (~2 us per 32 columns).
#include <cooperative_groups.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace vargp {

constexpr int CB = 32;                 // block edge
constexpr int BLD = 36;                // row stride of a stored block: a lane reading ITS row as float4 is conflict-free
constexpr int BLK = CB * BLD;          // floats per block
constexpr int kClThreads = 512;
constexpr int kClWarps = kClThreads / 32;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void cl_ld_row(const float* p, float (&a)[CB]) {
#pragma unroll
  for (int c4 = 0; c4 < CB; c4 += 4) {
    const float4 v = *reinterpret_cast<const float4*>(p + c4);
    a[c4] = v.x; a[c4 + 1] = v.y; a[c4 + 2] = v.z; a[c4 + 3] = v.w;
  }
}
__device__ __forceinline__ void cl_st_row(float* p, const float (&a)[CB]) {
#pragma unroll
  for (int c4 = 0; c4 < CB; c4 += 4) *reinterpret_cast<float4*>(p + c4) = make_float4(a[c4], a[c4 + 1], a[c4 + 2], a[c4 + 3]);
}

// Register micro-tile of the 32x32x32 block products: lane (rg, cg) = (lane / 4, lane % 4) owns R rows and the columns
// 8 cg .. 8 cg + 7 of the result.  R = 4: one warp per block (rows 4 rg ..): per k-index one LDS.128 of the (k-major) left
// operand and two of the right operand feed 32 FMAs -- a shared-memory read costs its RETURN width (an LDS.128 is 4 cycles
// of the 128 B/clk pipe even when every lane reads the same address), so the lane-per-row form with 8 LDS.128 per 32 FMAs
// was shared-memory bound.  R = 2 / 1: the block is split over 2 / 4 warps (16 / 8 rows each) when a phase has fewer
// blocks than warps: a lone warp needs ~1.5 us for a whole block and those phases are on the critical path.
//   acc[a][b] (+/-)= sum_m At[m][a] * B[m][b]        (At, B already offset to the lane's rows / columns)
// Blackwell packed fp32 FMA (SASS FFMA2): two IEEE fp32 FMAs per lane and instruction, one operand may be a broadcast
// scalar with a free negation.  The plain FFMA issues every other cycle per scheduler here, so the block products were
// FMA-issue bound; the packed form halves the instruction count at identical results (each half is a round-to-nearest fma).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int R> struct mk_vec;
template <> struct mk_vec<4> { using t = float4; };
template <> struct mk_vec<2> { using t = float2; };
template <> struct mk_vec<1> { using t = float; };
template <int R> __device__ __forceinline__ void mk_unpack(const typename mk_vec<R>::t& v, float (&a)[R]);
template <> __device__ __forceinline__ void mk_unpack<4>(const float4& v, float (&a)[4]) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
template <> __device__ __forceinline__ void mk_unpack<2>(const float2& v, float (&a)[2]) { a[0] = v.x; a[1] = v.y; }
template <> __device__ __forceinline__ void mk_unpack<1>(const float& v, float (&a)[1]) { a[0] = v; }
template <int R> __device__ __forceinline__ typename mk_vec<R>::t mk_pack(const float (&a)[R]);
template <> __device__ __forceinline__ float4 mk_pack<4>(const float (&a)[4]) { return make_float4(a[0], a[1], a[2], a[3]); }
template <> __device__ __forceinline__ float2 mk_pack<2>(const float (&a)[2]) { return make_float2(a[0], a[1]); }
template <> __device__ __forceinline__ float mk_pack<1>(const float (&a)[1]) { return a[0]; }

template <int R, bool NEG>
__device__ __forceinline__ void mk_fma(float (&acc)[R][8], const float* __restrict__ At, const float* __restrict__ B) {
  // unrolled by 4 only: the step executes each of these products ONCE per warp, so fully unrolled bodies (24 KB each, four
  // of them plus the diagonal block) were streamed from L2 through the 32 KB instruction cache at every step
  f32x2 c2[R][4];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c2[i][j] = pk2(acc[i][2 * j], acc[i][2 * j + 1]);
#pragma unroll 4
  for (int m = 0; m < CB; ++m) {
    float a[R];
    mk_unpack<R>(*reinterpret_cast<const typename mk_vec<R>::t*>(At + m * BLD), a);
    const float4 b0 = *reinterpret_cast<const float4*>(B + m * BLD);
    const float4 b1 = *reinterpret_cast<const float4*>(B + m * BLD + 4);
    const f32x2 b2[4] = {pk2(b0.x, b0.y), pk2(b0.z, b0.w), pk2(b1.x, b1.y), pk2(b1.z, b1.w)};
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float ai = NEG ? -a[i] : a[i];
      const f32x2 a2 = pk2(ai, ai);
#pragma unroll
      for (int j = 0; j < 4; ++j) c2[i][j] = fma2(a2, b2[j], c2[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) upk2(c2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);
}
// same with a ROW-major left operand: acc[a][b] += sum_m A[a][m] * B[m][b]
template <int R>
__device__ __forceinline__ void mk_fma_rowA(float (&acc)[R][8], const float* __restrict__ A, const float* __restrict__ B) {
  f32x2 c2[R][4];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c2[i][j] = pk2(acc[i][2 * j], acc[i][2 * j + 1]);
#pragma unroll 1
  for (int m4 = 0; m4 < CB; m4 += 4) {
    float4 ar[R];
#pragma unroll
    for (int i = 0; i < R; ++i) ar[i] = *reinterpret_cast<const float4*>(A + i * BLD + m4);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int m = m4 + u;
      const float4 b0 = *reinterpret_cast<const float4*>(B + m * BLD);
      const float4 b1 = *reinterpret_cast<const float4*>(B + m * BLD + 4);
      const f32x2 b2[4] = {pk2(b0.x, b0.y), pk2(b0.z, b0.w), pk2(b1.x, b1.y), pk2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const float a = (u == 0) ? ar[i].x : (u == 1) ? ar[i].y : (u == 2) ? ar[i].z : ar[i].w;
        const f32x2 a2 = pk2(a, a);
#pragma unroll
        for (int j = 0; j < 4; ++j) c2[i][j] = fma2(a2, b2[j], c2[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) upk2(c2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);
}
template <int R>
__device__ __forceinline__ void mk_zero(float (&acc)[R][8]) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}
// T points at the lane's first element (row-major block)
template <int R>
__device__ __forceinline__ void mk_load(float (&acc)[R][8], const float* T) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const float4 v0 = *reinterpret_cast<const float4*>(T + i * BLD), v1 = *reinterpret_cast<const float4*>(T + i * BLD + 4);
    acc[i][0] = v0.x; acc[i][1] = v0.y; acc[i][2] = v0.z; acc[i][3] = v0.w;
    acc[i][4] = v1.x; acc[i][5] = v1.y; acc[i][6] = v1.z; acc[i][7] = v1.w;
  }
}
template <int R>
__device__ __forceinline__ void mk_store(float* T, const float (&acc)[R][8]) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
    *reinterpret_cast<float4*>(T + i * BLD) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(T + i * BLD + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}
// transposed: Tt points at element (column c0, row r) of the block that receives the transpose
template <int R>
__device__ __forceinline__ void mk_store_t(float* Tt, const float (&acc)[R][8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float col[R];
#pragma unroll
    for (int i = 0; i < R; ++i) col[i] = acc[i][j];
    *reinterpret_cast<typename mk_vec<R>::t*>(Tt + j * BLD) = mk_pack<R>(col);
  }
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float rsqrt_fast(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

#define CL_STAMP(slot) do { if (dbgk && lane == 0) dbgk[slot] = clock64(); } while (0)

// Diagonal block (one warp): L_kk in place, D = L_kk^-1 into Wkk (row-major) and D^T into the DT buffer of every CTA.
// Lane i keeps row i of the block AND column i of the inverse in registers.  Per pivot j: the lane that owns the pivot
// has it one FMA after the previous column (no shuffle on the chain), rsqrt + one Newton step, then the column of L is
// handed round by shuffles and every shuffled value feeds two FMAs: the rank-1 update of the factor and the forward
// substitution of the inverse (x_c[i] -= l_ij D[j][c]), which therefore costs no extra communication.  Lanes above the
// diagonal compute garbage that never reaches a valid entry; it is zeroed at the end.
template <int CS>
__device__ __forceinline__ void cl_diag_block(float* Lkk, float* Wkk, float* DTloc, float* const (&rDT)[CS], int lane,
                                              int k0, int n, int* s_info, long long* dbgk) {
  float acc[CB];
  f32x2 ax[CB];                            // (row `lane` of the block, column `lane` of the inverse), element c
  CL_STAMP(8);
  cl_ld_row(Lkk + lane * BLD, acc);
#pragma unroll
  for (int i = 0; i < CB; ++i) ax[i] = pk2(acc[i], (i == lane) ? 1.f : 0.f);
  float dn = acc[0];                       // next pivot, valid in the lane that owns it
  int bad = 0;
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    float r = rsqrt_fast(dn);
    r = r * fmaf(-0.5f * dn * r, r, 1.5f);
    const float inv = __shfl_sync(FULL, r, j);
    const float d = __shfl_sync(FULL, dn, j);
    if (!(d > 0.f) && bad == 0 && k0 + j < n) bad = k0 + j + 1;
    float aj, xj;
    upk2(ax[j], aj, xj);
    const float lij = aj * inv;            // lanes >= j; lane j: d * rsqrt(d)
    xj *= inv;                             // D[j][lane]
    ax[j] = pk2(lij, xj);
    if (j + 1 < CB) {
      float an, xn;
      upk2(ax[j + 1], an, xn);
      dn = fmaf(-lij, lij, an);            // lane j+1: its own diagonal after this step, no shuffle on the critical chain
      const f32x2 m2 = pk2(-lij, -xj);
#pragma unroll
      for (int c = j + 1; c < CB; ++c) {
        const float lc = __shfl_sync(FULL, lij, c);
        ax[c] = fma2(pk2(lc, lc), m2, ax[c]);      // acc[c] -= lij l_cj ;  x[c] -= l_cj D[j][lane]
      }
    }
  }
  float x[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) upk2(ax[c], acc[c], x[c]);
  CL_STAMP(9);
  if (bad && lane == 0 && *s_info == 0) *s_info = bad;
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    if (c > lane) acc[c] = 0.f;
    if (c < lane) x[c] = 0.f;
  }
  cl_st_row(Lkk + lane * BLD, acc);
#pragma unroll
  for (int i = 0; i < CB; ++i) Wkk[i * BLD + lane] = x[i];
  CL_STAMP(10);
  cl_st_row(DTloc + lane * BLD, x);                                   // DT[m = lane][c] = D[c][m]
  __syncwarp();
#pragma unroll
  for (int t = 0; t < BLK / 4 / 32; ++t) {                           // whole 512 B lines to the other CTAs
    const float4 v = reinterpret_cast<const float4*>(DTloc)[t * 32 + lane];
#pragma unroll
    for (int q = 0; q < CS; ++q)
      if (rDT[q] != DTloc) reinterpret_cast<float4*>(rDT[q])[t * 32 + lane] = v;
  }
  CL_STAMP(11);
}

template <int CS>
__global__ void __launch_bounds__(kClThreads, 1)
potrf_inv_cluster_kernel(const float* __restrict__ Ain, int64_t a_ld, int64_t a_bs, float* __restrict__ Lout, int64_t l_ld,
                         int64_t l_bs, float* __restrict__ Wout, int64_t w_ld, int64_t w_bs, int n, float jitter,
                         int32_t* __restrict__ info, int info_base, int accumulate, int maxblk, long long* dbg) {
  pdl_enter();
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int mat = blockIdx.x / CS;
  extern __shared__ __align__(16) float sm[];
  const int nblk = (n + CB - 1) / CB;
  float* Lst = sm;
  float* Wst = Lst + maxblk * BLK;
  float* slots = Wst + maxblk * BLK;       // per step: slot j > k = L_jk^T, slot j <= k = W_kj
  float* DT = slots + nblk * BLK;
  int* s_info = reinterpret_cast<int*>(DT + BLK);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int c0 = (lane & 3) * 8;                              // micro-tile column origin inside a block
  float* rslots[CS];
  float* rDT[CS];
#pragma unroll
  for (int q = 0; q < CS; ++q) {
    rslots[q] = (q == rank) ? slots : cluster.map_shared_rank(slots, q);
    rDT[q] = (q == rank) ? DT : cluster.map_shared_rank(DT, q);
  }
  if (tid == 0) *s_info = 0;
  // block (i, j) of this CTA (i mod CS == rank): rows before it hold rank+1, rank+1+CS, ... blocks
  auto rowoff = [&](int i) { const int li = i / CS; return li * (rank + 1) + CS * (li * (li - 1) / 2); };
  // the last block row of rank q (or < 0)
  auto lastrow = [&](int q) { return (nblk - 1 >= q) ? q + ((nblk - 1 - q) / CS) * CS : -1; };

  // ---- load the own block rows: L storage <- tril(A) + jitter I (identity padding), X storage <- I ----
  const float* A = Ain + (int64_t)mat * a_bs;
  for (int i = rank; i < nblk; i += CS) {
    const int items = CB * (i + 1);
    float* Lrow = Lst + rowoff(i) * BLK;
    float* Wrow = Wst + rowoff(i) * BLK;
    for (int it0 = wid * 8; it0 < items; it0 += kClWarps * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int it = it0 + u;
        const int r = it / (i + 1), jb = it % (i + 1);
        const int gi = i * CB + r, gj = jb * CB + lane;
        v[u] = (gi == gj) ? 1.f : 0.f;
        if (it < items && gi < n && gj <= gi) v[u] = A[(int64_t)gi * a_ld + gj] + ((gi == gj) ? jitter : 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int it = it0 + u;
        if (it < items) {
          const int r = it / (i + 1), jb = it % (i + 1);
          Lrow[jb * BLK + r * BLD + lane] = v[u];
          Wrow[jb * BLK + r * BLD + lane] = (i * CB + r == jb * CB + lane) ? 1.f : 0.f;
        }
      }
    }
  }
  cluster.sync();                                     // every CTA of the cluster runs: its shared memory may be written
  // debug stamps (vargp_chol_cluster_debug): matrix 0 only, 16 clock64 slots per (rank, step); warps 0 and 1
  long long* dbg0 = (dbg && mat == 0 && wid < 2) ? dbg + (int64_t)rank * 16 * 16 + (wid ? 4 : 0) : nullptr;
  if (dbg0 && lane == 0 && wid == 0) dbg0[15] = clock64();

  for (int k = -1; k < nblk; ++k) {                   // k = -1: only the diagonal block of step 0
    const int owner = (k + CS) % CS;
    long long* dbgk = dbg0 ? dbg0 + k * 16 : nullptr;
    if (k >= 0) {
    cluster.sync();                                   // D_k^T everywhere; everybody is done with step k-1
    CL_STAMP(0);
    // ---- (B) panel rows i > k of this CTA; the owner also finishes row k of the inverse.  Blocks are split over
    //      S = 4 / R warps when there are fewer blocks than warps ----
    {
      int i0 = rank + (k / CS) * CS;
      if (i0 <= k) i0 += CS;
      const int nrow = (i0 < nblk) ? (nblk - 1 - i0) / CS + 1 : 0;
      const int Ub = nrow + ((rank == owner) ? k + 1 : 0);
      // split factor: a block costs 384 / 640 / 1152 cycles of the shared-memory pipe when 1 / 2 / 4 warps share it and
      // 2048 / 1024 / 512 cycles of FMA issue per warp: two warps per block while there are at most two blocks, else one
      const int S = (Ub <= 2) ? 2 : 1;
      float* Wkk = Wst + (rowoff(k) + k) * BLK;
      auto item = [&](auto rtag, int u, int sub) {
        constexpr int R = decltype(rtag)::value;
        const int rr = sub * (8 * R) + (lane >> 2) * R;           // the lane's first row inside the block
        float acc[R][8];
        constexpr int S_ = 4 / R, PER = (BLK / 4) / S_;             // float4s of a block that each of the S_ warps pushes
        auto unit_sync = [&]() { if (R == 4) __syncwarp(); else named_bar(1 + (u & 7), S_ * 32); };
        // the finished block sits in the LOCAL slot; hand it to the other CTAs as whole 512 B lines (a lane-strided
        // remote store is one distributed-shared-memory packet per lane: measured ~0.5 us per block that way)
        auto push = [&](int slot, auto need) {
          const float4* src = reinterpret_cast<const float4*>(slots + slot * BLK) + sub * PER;
          for (int t = lane; t < PER; t += 32) {
            const float4 v = src[t];
#pragma unroll
            for (int q = 0; q < CS; ++q)
              if (q != rank && need(q)) reinterpret_cast<float4*>(rslots[q] + slot * BLK)[sub * PER + t] = v;
          }
        };
        if (u < nrow) {                                            // panel block of row i: L_ik = A_ik D_k^T
          const int i = i0 + u * CS;
          float* Lik = Lst + (rowoff(i) + k) * BLK;
          if (wid == 0) CL_STAMP(12);
          mk_zero<R>(acc);
          mk_fma_rowA<R>(acc, Lik + rr * BLD, DT + c0);
          __syncwarp();
          if (wid == 0) CL_STAMP(13);
          mk_store<R>(Lik + rr * BLD + c0, acc);
          mk_store_t<R>(slots + i * BLK + c0 * BLD + rr, acc);     // slot i = L_ik^T
          unit_sync();
          if (wid == 0) CL_STAMP(14);
          push(i, [&](int q) { return lastrow(q) >= i; });
        } else {                                                   // row k of the inverse: W_kj = D_k X_kj
          const int j = u - nrow;
          float* Xkj = Wst + (rowoff(k) + j) * BLK;
          if (j < k) {
            mk_zero<R>(acc);
            mk_fma<R, false>(acc, DT + rr, Xkj + c0);            // D_k[r][m] = DT[m][r]
          } else {
            mk_load<R>(acc, Wkk + rr * BLD + c0);
          }
          mk_store<R>(slots + j * BLK + rr * BLD + c0, acc);
          unit_sync();                                             // X_kj is read by every warp of the unit before anybody
          if (j < k) mk_store<R>(Xkj + rr * BLD + c0, acc);       // overwrites it in place
          push(j, [&](int q) { return lastrow(q) > k; });
        }
      };
      for (int it = wid; it < Ub * S; it += kClWarps) {
        if (S == 4) item(std::integral_constant<int, 1>(), it >> 2, it & 3);
        else if (S == 2) item(std::integral_constant<int, 2>(), it >> 1, it & 1);
        else item(std::integral_constant<int, 4>(), it, 0);
      }
    }
    CL_STAMP(1);
    if (k + 1 == nblk) break;
    cluster.sync();                                   // panel and inverse row visible
    CL_STAMP(2);
    }
    // ---- (C) trailing update of the own rows:  A_ij -= L_ik L_jk^T (j > k),  X_ij -= L_ik W_kj (j <= k).  The owner of
    //      k+1 factors its diagonal block one step ahead: the four warps of scheduler 0 update that block together, warp 0
    //      goes on with the factorisation and warps 4, 8, 12 stay out of its way ----
    {
      const bool ahead = (rank == (k + 1) % CS);
      auto trail = [&](auto rtag, int i, int j, int sub) {
        constexpr int R = decltype(rtag)::value;
        const int rr = sub * (8 * R) + (lane >> 2) * R;
        float* T = (((j > k) ? Lst : Wst) + (rowoff(i) + j) * BLK) + rr * BLD + c0;
        float acc[R][8];
        mk_load<R>(acc, T);
        mk_fma<R, true>(acc, slots + i * BLK + rr, slots + j * BLK + c0);      // slot i = L_ik^T, pushed here in (B)
        mk_store<R>(T, acc);
      };
      // the diagonal block of the next step is the critical path: its update runs FIRST and alone (the other warps of
      // this CTA wait at barrier 14 instead of competing for the shared-memory pipe: 2.0 -> 0.6 us)
      if (ahead && k >= 0) {
        if ((wid & 3) == 0) trail(std::integral_constant<int, 1>(), k + 1, k + 1, wid >> 2);
        named_bar(14, kClThreads);                    // one call site for all 16 warps
      }
      if (ahead && (wid & 3) == 0) {
        if (wid == 0)
          cl_diag_block<CS>(Lst + (rowoff(k + 1) + k + 1) * BLK, Wst + (rowoff(k + 1) + k + 1) * BLK, DT, rDT, lane,
                            (k + 1) * CB, n, s_info, dbgk ? dbgk + 16 : nullptr);
      } else if (k >= 0) {
        const int nw = ahead ? kClWarps - kClWarps / 4 : kClWarps;
        const int w = ahead ? wid - 1 - (wid >> 2) : wid;
        int i0 = rank + (k / CS) * CS;
        if (i0 <= k) i0 += CS;
        int Uc = 0;
        for (int i = i0; i < nblk; i += CS) Uc += i + 1;
        if (ahead) Uc -= 1;                           // block (k+1, k+1), the last one of the first row, is warp 0's
        const int S = (Uc <= 3) ? 2 : 1;              // same trade as in (B)
        for (int it = w; it < Uc * S; it += nw) {
          int u = (S == 4) ? it >> 2 : ((S == 2) ? it >> 1 : it);
          const int sub = (S == 4) ? it & 3 : ((S == 2) ? it & 1 : 0);
          int i = i0;
          for (;;) {
            const int len = (ahead && i == k + 1) ? i : i + 1;
            if (u < len) break;
            u -= len;
            i += CS;
          }
          if (S == 4) trail(std::integral_constant<int, 1>(), i, u, sub);
          else if (S == 2) trail(std::integral_constant<int, 2>(), i, u, sub);
          else trail(std::integral_constant<int, 4>(), i, u, sub);
        }
      }
      if (k >= 0) CL_STAMP(3);
    }
  }

  // ---- store the own rows (coalesced along j; strict upper triangles zero-filled), gather the first bad pivot ----
  cluster.sync();
  if (dbg0 && lane == 0 && wid == 0) dbg0[15 * 16 + 14] = clock64();
  if (rank == 0 && tid == 0 && info) {
    int best = 0;
#pragma unroll
    for (int q = 0; q < CS; ++q) {
      const int v = (q == 0) ? *s_info : *cluster.map_shared_rank(s_info, q);
      if (v && (best == 0 || v < best)) best = v;
    }
    if (!accumulate) info[mat] = best ? best + info_base : 0;
    else if (best && info[mat] == 0) info[mat] = best + info_base;
  }
  float* L = Lout + (int64_t)mat * l_bs;
  float* W = Wout + (int64_t)mat * w_bs;
  for (int i = rank; i < nblk; i += CS) {
    const float* Lrow = Lst + rowoff(i) * BLK;
    const float* Wrow = Wst + rowoff(i) * BLK;
    for (int r = wid; r < CB; r += kClWarps) {
      const int gi = i * CB + r;
      if (gi >= n) break;
      for (int gj = lane; gj < n; gj += 32) {
        const bool low = gj <= gi;
        const int o = (gj >> 5) * BLK + r * BLD + lane;
        L[(int64_t)gi * l_ld + gj] = low ? Lrow[o] : 0.f;
        W[(int64_t)gi * w_ld + gj] = low ? Wrow[o] : 0.f;
      }
    }
  }
  cluster.sync();                                     // nobody leaves while rank 0 may still read its s_info
  if (dbg0 && lane == 0 && wid == 0) dbg0[15 * 16 + 15] = clock64();
}

static long long* g_cl_dbg = nullptr;                // device buffer of 4 * 16 * 16 int64 stamps, or null
static int g_cl_min_n = 33, g_cl_max_n = 320;        // vargp_chol_inv routes min_n <= n <= max_n here (0 / 0 disables)

template <int CS>
static int launch_cluster(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs, float* W,
                          int64_t w_ld, int64_t w_bs, int n, int64_t batch, float jitter, int32_t* info, int info_base,
                          int accumulate, cudaStream_t stream) {
  const int nblk = (n + CB - 1) / CB;
  int maxblk = 0;
  for (int q = 0; q < CS; ++q) {
    int s = 0;
    for (int i = q; i < nblk; i += CS) s += i + 1;
    if (s > maxblk) maxblk = s;
  }
  const size_t dyn = (size_t)(2 * maxblk + nblk + 1) * BLK * sizeof(float) + 16;
  if (dyn > 227 * 1024) return VARGP_ERR_UNSUPPORTED;
  static size_t attr_set = 0;
  if (dyn > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(potrf_inv_cluster_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return (int)e;
    attr_set = dyn;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(batch * CS));
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, potrf_inv_cluster_kernel<CS>, A, a_ld, a_bs, L, l_ld, l_bs, W, w_ld, w_bs, n, jitter, info,
                     info_base, accumulate, maxblk, g_cl_dbg);
  return launch_status();
}

}  // namespace vargp

using namespace vargp;

// Routing window of vargp_chol_inv for the cluster kernel; negative arguments only query.  Returns (min_n << 32) | max_n.
extern "C" int64_t vargp_chol_cluster_config(int64_t min_n, int64_t max_n) {
  const int64_t old = ((int64_t)g_cl_min_n << 32) | (int64_t)g_cl_max_n;
  if (min_n >= 0 && max_n >= 0) {
    g_cl_min_n = (int)min_n;
    g_cl_max_n = (int)(max_n > 320 ? 320 : max_n);
  }
  return old;
}

// clock64 stamps of matrix 0 (phase boundaries per rank and step) into a device buffer of 4*16*16 int64; null turns it off
extern "C" void vargp_chol_cluster_debug(void* buf) { g_cl_dbg = reinterpret_cast<long long*>(buf); }

// 1 if vargp_chol_inv would take the cluster kernel for n-row matrices
extern "C" int vargp_chol_cluster_wants(int64_t n) {
  static bool env_read = false;
  if (!env_read) {
    env_read = true;
    const char* e = getenv("VARGP_CHOL_CLUSTER");
    if (e && atoi(e) == 0) g_cl_min_n = g_cl_max_n = 0;
  }
  return (g_cl_max_n > 0 && n >= g_cl_min_n && n <= g_cl_max_n && n > CB) ? 1 : 0;
}

// L = chol(A + jitter I), W = L^-1 for 32 < n <= 320 on one cluster per matrix.  A may alias L or W (a CTA reads its block
// rows completely before anybody stores); L and W must differ.
extern "C" int vargp_chol_inv_cluster_ex(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                                         float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                                         int32_t* info, int64_t info_base, int accumulate, void* stream) {
  if (!A || !L || !W || n < 1 || batch < 1 || a_ld < n || l_ld < n || w_ld < n || L == W) return VARGP_ERR_ARG;
  if (n <= CB || n > 320 || batch > (1 << 20)) return VARGP_ERR_UNSUPPORTED;
  const int nblk = (int)((n + CB - 1) / CB);
  if (nblk <= 3)
    return launch_cluster<2>(A, a_ld, a_bs, L, l_ld, l_bs, W, w_ld, w_bs, (int)n, batch, jitter, info, (int)info_base,
                             accumulate, (cudaStream_t)stream);
  return launch_cluster<4>(A, a_ld, a_bs, L, l_ld, l_bs, W, w_ld, w_bs, (int)n, batch, jitter, info, (int)info_base,
                           accumulate, (cudaStream_t)stream);
}

extern "C" int vargp_chol_inv_cluster(const float* A, int64_t a_ld, int64_t a_bs, float* L, int64_t l_ld, int64_t l_bs,
                                      float* W, int64_t w_ld, int64_t w_bs, int64_t n, int64_t batch, float jitter,
                                      int32_t* info, void* stream) {
  return vargp_chol_inv_cluster_ex(A, a_ld, a_bs, L, l_ld, l_bs, W, w_ld, w_bs, n, batch, jitter, info, 0, 0, stream);
}
