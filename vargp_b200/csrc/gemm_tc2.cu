// Persistent 2-CTA tcgen05 / TMA batched GEMM with 3xTF32 splitting for the large problems of the hot path
// (Kxz / Kzz Grams, the TRSM-replacement products W*Kzx ..., their adjoints).  Same descriptor contract as
// gemm_tc.cu; chosen by vargp_gemm_tc when the problem fills 256 x 256 tiles.
//
// Why a second kernel: 3xTF32 issues three MMAs per product, and every operand element is re-read from shared
// memory by each of them.  With 128 x 128 tiles on one CTA the kernel is bound by the 128 B/clk shared-memory
// port, not by the tensor pipe (measured: 88 TFLOP/s = 38 % of the 3xTF32 peak; model: 43 %).  Here a CTA PAIR
// (cta_group::2) owns a 256 x 256 tile: each CTA stages only its 128 rows of A and its 128 columns of B, and
// the UMMA of M = 256, N = 256 reads the other half from the peer's shared memory, which halves the bytes per
// flop.  The split itself is also cheaper: kind::tf32 ignores the 13 low mantissa bits, so the RAW fp32 slab is
// the `hi` operand as it lands from TMA and only  lo = x - trunc_tf32(x)  is written (8 B instead of 12 B of
// shared-memory traffic per element for the split).
//
// Per CTA (512 threads, 1 CTA / SM, cluster of 2, persistent over tiles):
//   warp 0       TMA producer   raw fp32 slabs (128 rows x 32 k) of A and B -> 3-stage ring, local mbarrier
//   warp 1       MMA issuer     (leader CTA only) per 32-wide slab: 8 cross-term UMMAs (lo*hi, hi*lo) then
//                               4 hi*hi UMMAs into one of two 256-column TMEM buffers; tcgen05.commit multicast
//                               frees the stage in both CTAs and hands the buffer to both epilogues
//   warps 4-7    splitter       lo = x - trunc(x) into the second buffer of the stage, fence.proxy.async,
//                               arrive on the LEADER's mbarrier (remote arrive from the peer)
//   warps 8-15   epilogue       every slab: tcgen05.ld the slab sum and add it to 128 fp32 registers per thread
//                               (the tensor core accumulates with truncation: short 12-step sums + fp32
//                               promotion keep fp32-grade accuracy, see gemm_tc.cu); per tile: alpha / beta /
//                               triangle mask / RBF exp epilogue and the store
// Registers: setmaxnreg moves the budget to the epilogue warpgroups (200 / thread) from the others (56).
// Roofline: tensor pipe.  Algorithmic flops per launch: 2*M*N*K*batch (x 1/2 per triangular flag).
#include <cstdlib>

#include "tc_common.cuh"

namespace vargp {

constexpr int T2_STAGES = 3;
constexpr int T2_THREADS = 512;
constexpr int T2_BM = 256, T2_BN = 256;                // cluster tile
constexpr int T2_SMEM_BYTES = 4 * T2_STAGES * TC_TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int T2_EPI_WARPS = 8;                        // per CTA
constexpr int T2_SPLIT_WARPS = 4;                      // per CTA
constexpr int T2_TMEM_COLS = 512;                      // two 256-column slab accumulators

// ---------------------------------------------------------------------------------------------
// cluster / 2-CTA PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP_CL:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE_CL;\n\t"
      "bra.uni WAIT_LOOP_CL;\n\t"
      "WAIT_DONE_CL:\n\t"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// tile schedule (identical in every role)
// ---------------------------------------------------------------------------------------------
struct T2Sched {
  int n_mt, n_nt;
  int64_t total;
};
struct T2Tile {
  int64_t m0, n0;
  int i0, i1, i2;
  int kb_lo, nk;
};

__device__ __forceinline__ T2Tile t2_tile(const TcParams& p, const T2Sched& sc, int64_t t) {
  T2Tile tl;
  int mt = (int)(t % sc.n_mt);
  int64_t r = t / sc.n_mt;
  int nt = (int)(r % sc.n_nt);
  int64_t z = r / sc.n_nt;
  // heaviest tiles first for the triangular k-ranges (static round-robin over clusters ~ LPT)
  if (p.tri_a == VARGP_TRI_LOWER) mt = sc.n_mt - 1 - mt;
  if (p.tri_b == VARGP_TRI_UPPER) nt = sc.n_nt - 1 - nt;
  tl.m0 = (int64_t)mt * T2_BM;
  tl.n0 = (int64_t)nt * T2_BN;
  tl.i2 = (int)(z % p.nb[2]); z /= p.nb[2];
  tl.i1 = (int)(z % p.nb[1]);
  tl.i0 = (int)(z / p.nb[1]);
  bool dead = false;
  if (p.tri_c == VARGP_TRI_LOWER && tl.n0 > tl.m0 + T2_BM - 1) dead = true;
  if (p.tri_c == VARGP_TRI_UPPER && tl.m0 > tl.n0 + T2_BN - 1) dead = true;
  int64_t k_lo = 0, k_hi = p.K;
  if (p.tri_a == VARGP_TRI_LOWER) k_hi = min(k_hi, tl.m0 + T2_BM);
  if (p.tri_a == VARGP_TRI_UPPER) k_lo = max(k_lo, tl.m0);
  if (p.tri_b == VARGP_TRI_LOWER) k_lo = max(k_lo, tl.n0);
  if (p.tri_b == VARGP_TRI_UPPER) k_hi = min(k_hi, tl.n0 + T2_BN);
  tl.kb_lo = (int)(k_lo / TC_BK);
  tl.nk = (dead || k_hi <= k_lo) ? 0 : (int)((k_hi + TC_BK - 1) / TC_BK) - tl.kb_lo;
  return tl;
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// T2_RAW_HI: hi = the raw fp32 slab (hardware truncation to tf32), lo = x - trunc(x); otherwise hi = rna(x) is
// rewritten in place like in gemm_tc.cu (4 B more shared-memory traffic per element, unbiased split).
template <bool T2_RAW_HI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p,
                const int c_vec4) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA_hi = smem;
  uint8_t* sA_lo = sA_hi + T2_STAGES * TC_TILE_BYTES;
  uint8_t* sB_hi = sA_lo + T2_STAGES * TC_TILE_BYTES;
  uint8_t* sB_lo = sB_hi + T2_STAGES * TC_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB_lo + T2_STAGES * TC_TILE_BYTES);
  uint64_t* full_bar = bars;                      // [S] own CTA: TMA landed
  uint64_t* conv_bar = bars + T2_STAGES;          // [S] leader: lo written in both CTAs
  uint64_t* empty_bar = bars + 2 * T2_STAGES;     // [S] own CTA: MMAs that read the stage retired (multicast commit)
  uint64_t* accf_bar = bars + 3 * T2_STAGES;      // [2] own CTA: slab sum complete (multicast commit)
  uint64_t* acce_bar = bars + 3 * T2_STAGES + 2;  // [2] leader: buffer drained by the epilogues of both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * T2_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  T2Sched sc;
  sc.n_mt = (int)ceil_div(p.M, T2_BM);
  sc.n_nt = (int)ceil_div(p.N, T2_BN);
  sc.total = (int64_t)sc.n_mt * sc.n_nt * p.nb[0] * p.nb[1] * p.nb[2];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < T2_STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&conv_bar[s], 2 * T2_SPLIT_WARPS);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&accf_bar[b], 1);
        mbar_init(&acce_bar[b], 2 * T2_EPI_WARPS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(T2_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                              // peer's barriers are initialised before any remote arrive
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ================= TMA producer (each CTA loads its own 128 rows of A and 128 columns of B) ==========
      if (lane == 0) {
        uint32_t g = 0;
        for (int64_t t = cid; t < sc.total; t += ncl) {
          const T2Tile tl = t2_tile(p, sc, t);
          const int ca2 = p.a_b[2] ? tl.i2 : 0, ca1 = p.a_b[1] ? tl.i1 : 0, ca0 = p.a_b[0] ? tl.i0 : 0;
          const int cb2 = p.b_b[2] ? tl.i2 : 0, cb1 = p.b_b[1] ? tl.i1 : 0, cb0 = p.b_b[0] ? tl.i0 : 0;
          const int am = (int)tl.m0 + (int)rank * TC_ROWS, bn = (int)tl.n0 + (int)rank * TC_ROWS;
          for (int it = 0; it < tl.nk; ++it, ++g) {
            const int s = g % T2_STAGES;
            const uint32_t ph = (g / T2_STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_expect_tx(&full_bar[s], 2 * TC_TILE_BYTES);
            const int k0 = (tl.kb_lo + it) * TC_BK;
            uint8_t* da = sA_hi + s * TC_TILE_BYTES;
            uint8_t* db = sB_hi + s * TC_TILE_BYTES;
            if (!p.a_mn) {
              tma_load_5d(&tmA, &full_bar[s], da, k0, am, ca2, ca1, ca0);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) tma_load_5d(&tmA, &full_bar[s], da + c * 4096, am + 32 * c, k0, ca2, ca1, ca0);
            }
            if (!p.b_mn) {
              tma_load_5d(&tmB, &full_bar[s], db, k0, bn, cb2, cb1, cb0);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) tma_load_5d(&tmB, &full_bar[s], db + c * 4096, bn + 32 * c, k0, cb2, cb1, cb0);
            }
          }
        }
      }
    } else if (warp == 1) {
      // ================= MMA issuer (leader CTA; one elected lane) =================
      if (rank == 0) {
        // instruction descriptor: D=f32, A=B=tf32, majors, N>>3 at bit 17, M>>4 at bit 24 (M = 256 over the pair)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                               ((uint32_t)(T2_BN >> 3) << 17) | ((uint32_t)(T2_BM >> 4) << 24);
        const uint32_t lbo_a = p.a_mn ? 4096 : 16, lbo_b = p.b_mn ? 4096 : 16;
        const uint32_t sbo_a = p.a_mn ? 512 : 1024, sbo_b = p.b_mn ? 512 : 1024;
        const uint64_t la = p.a_mn ? 1 : 2, lb = p.b_mn ? 1 : 2;
        const uint32_t step_a = p.a_mn ? 1024 : 32, step_b = p.b_mn ? 1024 : 32;
        uint32_t g = 0;
        for (int64_t t = cid; t < sc.total; t += ncl) {
          const T2Tile tl = t2_tile(p, sc, t);
          for (int it = 0; it < tl.nk; ++it, ++g) {
            const int s = g % T2_STAGES;
            const uint32_t ph = (g / T2_STAGES) & 1;
            const int buf = g & 1;
            mbar_wait_cl(&conv_bar[s], ph);
            mbar_wait_cl(&acce_bar[buf], ((g >> 1) & 1) ^ 1);     // both epilogues have drained this buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
              const uint32_t a_hi = smem_u32(sA_hi + s * TC_TILE_BYTES), a_lo = smem_u32(sA_lo + s * TC_TILE_BYTES);
              const uint32_t b_hi = smem_u32(sB_hi + s * TC_TILE_BYTES), b_lo = smem_u32(sB_lo + s * TC_TILE_BYTES);
              const uint32_t t_acc = tmem_base + (uint32_t)(buf * T2_BN);
              // cross terms first (accumulator still small: their truncation error is negligible), then hi*hi
#pragma unroll
              for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
                const uint64_t dah = make_desc(a_hi + k8 * step_a, lbo_a, sbo_a, la);
                const uint64_t dal = make_desc(a_lo + k8 * step_a, lbo_a, sbo_a, la);
                const uint64_t dbh = make_desc(b_hi + k8 * step_b, lbo_b, sbo_b, lb);
                const uint64_t dbl = make_desc(b_lo + k8 * step_b, lbo_b, sbo_b, lb);
                umma_tf32_2cta(t_acc, dal, dbh, idesc, k8 > 0 ? 1u : 0u);
                umma_tf32_2cta(t_acc, dah, dbl, idesc, 1u);
              }
#pragma unroll
              for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
                const uint64_t dah = make_desc(a_hi + k8 * step_a, lbo_a, sbo_a, la);
                const uint64_t dbh = make_desc(b_hi + k8 * step_b, lbo_b, sbo_b, lb);
                umma_tf32_2cta(t_acc, dah, dbh, idesc, 1u);
              }
              umma_commit_2cta(&empty_bar[s]);              // stage reusable in both CTAs once these MMAs retire
              umma_commit_2cta(&accf_bar[buf]);             // slab sum ready for promotion in both CTAs
            }
            __syncwarp();
          }
        }
      }
    } else if (warp >= 4) {
      // ================= splitter (128 threads): lo = x - trunc_tf32(x) =================
      const int tI = threadIdx.x - 128;
      const uint32_t conv_remote = mapa_u32(smem_u32(&conv_bar[0]), 0);     // leader's barrier array
      uint32_t g = 0;
      for (int64_t t = cid; t < sc.total; t += ncl) {
        const T2Tile tl = t2_tile(p, sc, t);
        for (int it = 0; it < tl.nk; ++it, ++g) {
          const int s = g % T2_STAGES;
          const uint32_t ph = (g / T2_STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          float4* ah = reinterpret_cast<float4*>(sA_hi + s * TC_TILE_BYTES);
          float4* al = reinterpret_cast<float4*>(sA_lo + s * TC_TILE_BYTES);
          float4* bh = reinterpret_cast<float4*>(sB_hi + s * TC_TILE_BYTES);
          float4* bl = reinterpret_cast<float4*>(sB_lo + s * TC_TILE_BYTES);
#pragma unroll 4
          for (int e = 0; e < TC_TILE_BYTES / 16 / 128; ++e) {
            const int idx = e * 128 + tI;
            float4 va = ah[idx], vb = bh[idx], h, l;
            if (T2_RAW_HI) {
              h.x = __uint_as_float(__float_as_uint(va.x) & 0xFFFFE000u); h.y = __uint_as_float(__float_as_uint(va.y) & 0xFFFFE000u);
              h.z = __uint_as_float(__float_as_uint(va.z) & 0xFFFFE000u); h.w = __uint_as_float(__float_as_uint(va.w) & 0xFFFFE000u);
            } else {
              h.x = to_tf32_rna(va.x); h.y = to_tf32_rna(va.y); h.z = to_tf32_rna(va.z); h.w = to_tf32_rna(va.w);
              ah[idx] = h;
            }
            l.x = va.x - h.x; l.y = va.y - h.y; l.z = va.z - h.z; l.w = va.w - h.w;
            al[idx] = l;
            if (T2_RAW_HI) {
              h.x = __uint_as_float(__float_as_uint(vb.x) & 0xFFFFE000u); h.y = __uint_as_float(__float_as_uint(vb.y) & 0xFFFFE000u);
              h.z = __uint_as_float(__float_as_uint(vb.z) & 0xFFFFE000u); h.w = __uint_as_float(__float_as_uint(vb.w) & 0xFFFFE000u);
            } else {
              h.x = to_tf32_rna(vb.x); h.y = to_tf32_rna(vb.y); h.z = to_tf32_rna(vb.z); h.w = to_tf32_rna(vb.w);
              bh[idx] = h;
            }
            l.x = vb.x - h.x; l.y = vb.y - h.y; l.z = vb.z - h.z; l.w = vb.w - h.w;
            bl[idx] = l;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(conv_remote + 8u * s);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // ================= epilogue (warps 8..15: TMEM lane quadrant = warp % 4, column half = (warp - 8) / 4) ====
    const int quad = warp & 3;
    const int half = (warp - 8) >> 2;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 128);
    const uint32_t acce_remote = mapa_u32(smem_u32(&acce_bar[0]), 0);       // leader's barrier array
    uint32_t g = 0;
    for (int64_t t = cid; t < sc.total; t += ncl) {
      const T2Tile tl = t2_tile(p, sc, t);
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int it = 0; it < tl.nk; ++it, ++g) {
        const int buf = g & 1;
        mbar_wait(&accf_bar[buf], (g >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + (uint32_t)(buf * T2_BN + c * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(acce_remote + 8u * buf);
      }
      // ---- tile epilogue: this thread owns row m, columns n0 + half*128 .. +128 ----
      const int64_t m = tl.m0 + (int64_t)rank * TC_ROWS + quad * 32 + lane;
      if (m < p.M) {
        float* Crow = p.C + tl.i0 * p.c_bs[0] + tl.i1 * p.c_bs[1] + tl.i2 * p.c_bs[2] + m * p.c_rs;
        float gamma2 = 1.f, rown = 0.f;
        const float* e_col = nullptr;
        if (p.epi != VARGP_EPI_NONE) {
          gamma2 = expf(2.f * p.e_theta[tl.i0 * p.e_theta_bs[0] + tl.i1 * p.e_theta_bs[1] + tl.i2 * p.e_theta_bs[2] + p.e_D]);
          const float* e_row = p.e_row + tl.i0 * p.e_row_bs[0] + tl.i1 * p.e_row_bs[1] + tl.i2 * p.e_row_bs[2];
          e_col = p.e_col + tl.i0 * p.e_col_bs[0] + tl.i1 * p.e_col_bs[1] + tl.i2 * p.e_col_bs[2];
          rown = 0.5f * e_row[m];
        }
        const int64_t nbase = tl.n0 + half * 128;
#pragma unroll
        for (int j4 = 0; j4 < 128; j4 += 4) {
          const int64_t n = nbase + j4;
          if (n >= p.N) break;
          float v[4];
          bool keep[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t nn = n + u;
            keep[u] = !((p.tri_c == VARGP_TRI_LOWER && nn > m) || (p.tri_c == VARGP_TRI_UPPER && nn < m));
            float x = acc[j4 + u];
            if (p.epi != VARGP_EPI_NONE) {
              x = (nn < p.N) ? gamma2 * expf(x - rown - 0.5f * e_col[nn]) : 0.f;
              if (p.epi == VARGP_EPI_RBF_SYM && m == nn) x = gamma2;
            }
            v[u] = x * p.alpha;
          }
          if (c_vec4 && n + 3 < p.N) {
            float4* cp = reinterpret_cast<float4*>(Crow + n);
            float4 o;
            if (p.beta != 0.f) {
              o = *cp;
              o.x = keep[0] ? fmaf(p.beta, o.x, v[0]) : o.x; o.y = keep[1] ? fmaf(p.beta, o.y, v[1]) : o.y;
              o.z = keep[2] ? fmaf(p.beta, o.z, v[2]) : o.z; o.w = keep[3] ? fmaf(p.beta, o.w, v[3]) : o.w;
            } else {
              o.x = keep[0] ? v[0] : 0.f; o.y = keep[1] ? v[1] : 0.f; o.z = keep[2] ? v[2] : 0.f; o.w = keep[3] ? v[3] : 0.f;
            }
            *cp = o;
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (n + u >= p.N) break;
              float* cp = Crow + (n + u) * p.c_cs;
              if (!keep[u]) {
                if (p.beta == 0.f) *cp = 0.f;
                continue;
              }
              *cp = (p.beta != 0.f) ? fmaf(p.beta, *cp, v[u]) : v[u];
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();                              // no remote arrive / peer smem read may target an exited CTA
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T2_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int g_t2_clusters = 74;      // CTA pairs that can be co-resident (SM count / 2)
static int64_t g_t2_min_tiles = 24;
static int64_t g_t2_launches = 0;
static bool g_t2_rna = false;

int tc2_init() {
  cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  const char* rna = getenv("VARGP_TC2_RNA");
  if (rna) g_t2_rna = atoi(rna) != 0;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  g_t2_clusters = sms / 2 > 0 ? sms / 2 : 1;
  const char* env = getenv("VARGP_TC2_MIN_TILES");
  if (env) g_t2_min_tiles = atoll(env);
  return 0;
}

// the 2-CTA kernel pays off once the problem fills 256 x 256 tiles on most of the chip
bool tc2_wants(const vargp_gemm_t* g) {
  if (g_t2_min_tiles < 0) return false;
  if (g_t2_min_tiles <= 1) return true;            // forced (tests): anything vargp_gemm_tc accepts
  if (g->M < 192 || g->N < 192 || g->K < 64) return false;
  const int64_t tiles = ceil_div(g->M, T2_BM) * ceil_div(g->N, T2_BN) * g->nb[0] * g->nb[1] * g->nb[2];
  if (tiles < g_t2_min_tiles) return false;
  // padding waste of the 256-wide tiles must stay moderate
  const double eff = (double)(g->M * g->N) / (double)(ceil_div(g->M, T2_BM) * T2_BM * ceil_div(g->N, T2_BN) * T2_BN);
  return eff >= 0.7;
}

int tc2_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t stream) {
  const int64_t tiles = ceil_div(p.M, T2_BM) * ceil_div(p.N, T2_BN) * p.nb[0] * p.nb[1] * p.nb[2];
  const int64_t ncl = tiles < g_t2_clusters ? tiles : g_t2_clusters;
  int c_vec4 = (p.c_cs == 1 && p.c_rs % 4 == 0 && reinterpret_cast<uintptr_t>(p.C) % 16 == 0) ? 1 : 0;
  for (int i = 0; i < 3; ++i)
    if (p.nb[i] > 1 && p.c_bs[i] % 4 != 0) c_vec4 = 0;
  if (g_t2_rna)
    gemm_tc2_kernel<false><<<dim3((unsigned)(2 * ncl)), T2_THREADS, T2_SMEM_BYTES, stream>>>(tmA, tmB, p, c_vec4);
  else
    gemm_tc2_kernel<true><<<dim3((unsigned)(2 * ncl)), T2_THREADS, T2_SMEM_BYTES, stream>>>(tmA, tmB, p, c_vec4);
  ++g_t2_launches;
  return launch_status();
}

}  // namespace vargp

using namespace vargp;

extern "C" int64_t vargp_tc2_config(int64_t min_tiles) {
  const int64_t old = g_t2_min_tiles;
  if (min_tiles != INT64_MIN) g_t2_min_tiles = min_tiles;
  return old;
}

extern "C" int64_t vargp_tc2_launch_count(void) { return g_t2_launches; }
