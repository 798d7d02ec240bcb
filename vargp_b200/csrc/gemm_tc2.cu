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
constexpr int T2_EPI_WARPS = 8;                        // per CTA
constexpr int T2_SMEM_BYTES = 4 * T2_STAGES * TC_TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                              T2_EPI_WARPS * 32 * 33 * 4 /*per-warp transpose tiles of the store*/;
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
// Remote arrive with the default (.release.cta) semantics: what crosses the CTA boundary here is never generic-proxy
// data -- the consumers are tcgen05.mma (async proxy; the writer issues fence.proxy.async first) and TMEM reuse
// (ordered by tcgen05.fence).  A .release.cluster arrive compiles to MEMBAR.ALL.GPU + ERRBAR and was measured as
// THE bottleneck of this kernel (thousands of cycles per slab on the drain -> MMA hand-back).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP_CL:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
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
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
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
  int64_t nbatch;
  int64_t pairs;          // live (m-tile, n-tile) pairs
  int64_t dead;           // pairs culled by the output triangle; visited last and only to be zero-filled (beta == 0)
  int64_t total;          // (pairs + dead) * nbatch
};
struct T2Tile {
  int64_t m0, n0;
  int i0, i1, i2;
  int kb_lo, nk;
};

// number of live n-tiles of m-tile row `mt` (square 256 x 256 tiles: n0 > m0 + 255 <=> nt > mt)
__device__ __forceinline__ int t2_row_live(const TcParams& p, const T2Sched& sc, int mt) {
  if (p.tri_c == VARGP_TRI_LOWER) return min(mt + 1, sc.n_nt);
  if (p.tri_c == VARGP_TRI_UPPER) return max(sc.n_nt - mt, 0);
  return sc.n_nt;
}

__device__ __forceinline__ T2Sched t2_sched(const TcParams& p) {
  T2Sched sc;
  sc.n_mt = (int)ceil_div(p.M, T2_BM);
  sc.n_nt = (int)ceil_div(p.N, T2_BN);
  sc.nbatch = p.nb[0] * p.nb[1] * p.nb[2];
  sc.pairs = 0;
  for (int mt = 0; mt < sc.n_mt; ++mt) sc.pairs += t2_row_live(p, sc, mt);
  sc.dead = (p.beta == 0.f) ? (int64_t)sc.n_mt * sc.n_nt - sc.pairs : 0;
  sc.total = (sc.pairs + sc.dead) * sc.nbatch;
  return sc;
}

// Tile order: the batch index runs fastest, so the ~74 tiles that the clusters work on at any time have the same
// (m-tile, n-tile) and therefore the same k-range; the pairs are visited heaviest first (longest k-range of a
// triangular operand), which makes the static round-robin over clusters an LPT schedule.  (With m-tile fastest,
// 74 mod 8 = 2 made half of the clusters own only the odd, i.e. heavier, m-tiles: 25 % imbalance, measured
// 194 vs 256 TFLOP/s between triangular and dense products.)
__device__ __forceinline__ T2Tile t2_tile(const TcParams& p, const T2Sched& sc, int64_t t) {
  T2Tile tl;
  int64_t z = t % sc.nbatch;
  int64_t q = t / sc.nbatch;
  // k-range grows with mt for a lower-triangular A (k_hi = m0 + 256) and shrinks with it for an upper-triangular A
  const bool m_desc = !(p.tri_a == VARGP_TRI_UPPER);
  int mt = 0, nt = 0;
  bool dead = false;
  if (q < sc.pairs) {
    for (int i = 0; i < sc.n_mt; ++i) {
      mt = m_desc ? sc.n_mt - 1 - i : i;
      const int live = t2_row_live(p, sc, mt);
      if (q < live) {
        nt = (int)q;
        break;
      }
      q -= live;
    }
    if (p.tri_c == VARGP_TRI_UPPER) nt += mt;                 // live n-tiles of this row start at the diagonal
    else if (p.tri_b == VARGP_TRI_UPPER) nt = t2_row_live(p, sc, mt) - 1 - nt;   // heaviest n first
  } else {                                                    // culled tile: nothing to compute, zero-fill only
    dead = true;
    q -= sc.pairs;
    for (mt = 0; mt < sc.n_mt; ++mt) {
      const int live = t2_row_live(p, sc, mt), nd = sc.n_nt - live;
      if (q < nd) {
        nt = (p.tri_c == VARGP_TRI_LOWER) ? live + (int)q : (int)q;
        break;
      }
      q -= nd;
    }
  }
  tl.m0 = (int64_t)mt * T2_BM;
  tl.n0 = (int64_t)nt * T2_BN;
  tl.i2 = (int)(z % p.nb[2]); z /= p.nb[2];
  tl.i1 = (int)(z % p.nb[1]);
  tl.i0 = (int)(z / p.nb[1]);
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
                const int c_vec4, const int dbg) {
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
  float* stage_s = reinterpret_cast<float*>(bars + 32);       // [T2_EPI_WARPS][32][33] transpose tiles of the store

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const T2Sched sc = t2_sched(p);

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
  pdl_wait();

  if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
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
          const uint32_t ah = smem_u32(sA_hi + s * TC_TILE_BYTES) + 16u * tI, al = smem_u32(sA_lo + s * TC_TILE_BYTES) + 16u * tI;
          const uint32_t bh = smem_u32(sB_hi + s * TC_TILE_BYTES) + 16u * tI, bl = smem_u32(sB_lo + s * TC_TILE_BYTES) + 16u * tI;
#pragma unroll 4
          for (int e = 0; e < ((dbg & 2) ? 0 : TC_TILE_BYTES / 16 / 128); ++e) {
            const uint32_t off = (uint32_t)e * 2048u;
            const float4 va = lds128(ah + off), vb = lds128(bh + off);
            float4 l;
            if (T2_RAW_HI) {
              l.x = tf32_lo_of(va.x); l.y = tf32_lo_of(va.y); l.z = tf32_lo_of(va.z); l.w = tf32_lo_of(va.w);
            } else {
              float4 h;
              h.x = to_tf32_rna(va.x); h.y = to_tf32_rna(va.y); h.z = to_tf32_rna(va.z); h.w = to_tf32_rna(va.w);
              sts128(ah + off, h);
              l.x = va.x - h.x; l.y = va.y - h.y; l.z = va.z - h.z; l.w = va.w - h.w;
            }
            sts128(al + off, l);
            if (T2_RAW_HI) {
              l.x = tf32_lo_of(vb.x); l.y = tf32_lo_of(vb.y); l.z = tf32_lo_of(vb.z); l.w = tf32_lo_of(vb.w);
            } else {
              float4 h;
              h.x = to_tf32_rna(vb.x); h.y = to_tf32_rna(vb.y); h.z = to_tf32_rna(vb.z); h.w = to_tf32_rna(vb.w);
              sts128(bh + off, h);
              l.x = vb.x - h.x; l.y = vb.y - h.y; l.z = vb.z - h.z; l.w = vb.w - h.w;
            }
            sts128(bl + off, l);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(conv_remote + 8u * s);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
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
          if (dbg & 4) break;
          uint32_t r[32];
          tmem_ld32(t_row + (uint32_t)(buf * T2_BN + c * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(acce_remote + 8u * buf);
      }
      // ---- tile epilogue.  After the drain a thread owns ROW m (TMEM lane) and 128 consecutive columns.  Stored
      //      like that, every store instruction of a warp touches 32 different rows (measured: the store phase
      //      took 44 % of the kernel).  For row-major C each 32 x 32 sub-block is therefore transposed through a
      //      padded per-warp smem tile, so that a lane owns a COLUMN and a warp writes one full 128 B line per
      //      instruction.  Column-major C (c_rs == 1) is already coalesced along the lanes and stored directly. ----
      if (dbg & 1) continue;
      const int64_t mrow0 = tl.m0 + (int64_t)rank * TC_ROWS + quad * 32;       // first row of this warp
      const int64_t m = mrow0 + lane;
      const int64_t nbase = tl.n0 + half * 128;
      float* Cb = p.C + tl.i0 * p.c_bs[0] + tl.i1 * p.c_bs[1] + tl.i2 * p.c_bs[2];
      float gamma2 = 1.f, rown = 0.f;
      const float* e_col = nullptr;
      if (p.epi != VARGP_EPI_NONE) {
        gamma2 = expf(2.f * p.e_theta[tl.i0 * p.e_theta_bs[0] + tl.i1 * p.e_theta_bs[1] + tl.i2 * p.e_theta_bs[2] + p.e_D]);
        const float* e_row = p.e_row + tl.i0 * p.e_row_bs[0] + tl.i1 * p.e_row_bs[1] + tl.i2 * p.e_row_bs[2];
        e_col = p.e_col + tl.i0 * p.e_col_bs[0] + tl.i1 * p.e_col_bs[1] + tl.i2 * p.e_col_bs[2];
        if (m < p.M) rown = 0.5f * e_row[m];
      }
      {
        // explicit shared-space accesses (a generic LD/ST here costs address translation on every element) and
        // rolled loops behind the staging tile: this code runs once per tile, so straight-line unrolled code only
        // misses the instruction cache (measured: stall_no_inst was the top stall of the store phase)
        const uint32_t tile = smem_u32(stage_s + (warp - 8) * (32 * 33));
        const int Mi = (int)p.M, Ni = (int)p.N, mrow = (int)mrow0;
        const bool sym = p.epi == VARGP_EPI_RBF_SYM, rbf = p.epi != VARGP_EPI_NONE;
        const bool row_major = p.c_cs == 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j)      // the RBF row term is subtracted here, where a thread still owns one row
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(tile + 4u * (lane * 33 + j)), "f"(acc[c * 32 + j] - rown) : "memory");
          __syncwarp();
          const int nc0 = (int)nbase + c * 32;
          if (nc0 >= Ni) break;                                                // warp-uniform
          if (row_major) {
            const int n = nc0 + lane;                                          // this lane's column after the transpose
            const bool n_ok = n < Ni;
            float coln = 0.f;
            if (rbf && n_ok) coln = 0.5f * e_col[n];
            float* cp0 = Cb + (int64_t)mrow * p.c_rs + n;
            // interior 32 x 32 blocks (no ragged edge, no triangle boundary, beta == 0) skip every per-element predicate
            const bool kept = p.tri_c == VARGP_TRI_NONE || (p.tri_c == VARGP_TRI_LOWER && nc0 + 31 <= mrow) ||
                              (p.tri_c == VARGP_TRI_UPPER && nc0 >= mrow + 31);
            if (kept && !sym && mrow + 31 < Mi && nc0 + 31 < Ni) {
              const float scale = gamma2 * p.alpha;
              const bool accum = p.beta != 0.f;                                // warp-uniform
#pragma unroll 1
              for (int r0 = 0; r0 < 32; r0 += 8) {
                float xv[8], cv[8];
                if (accum) {                       // 8 independent coalesced 128 B reads in flight per warp
#pragma unroll
                  for (int u = 0; u < 8; ++u) cv[u] = cp0[(int64_t)(r0 + u) * p.c_rs];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xv[u]) : "r"(tile + 4u * ((r0 + u) * 33 + lane)));
#pragma unroll
                for (int u = 0; u < 8; ++u) xv[u] = rbf ? scale * expf(xv[u] - coln) : xv[u] * p.alpha;
                if (accum) {
#pragma unroll
                  for (int u = 0; u < 8; ++u) xv[u] = fmaf(p.beta, cv[u], xv[u]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) cp0[(int64_t)(r0 + u) * p.c_rs] = xv[u];
              }
              continue;
            }
#pragma unroll 1
            for (int r0 = 0; r0 < 32; r0 += 8) {
              float xv[8];
#pragma unroll
              for (int u = 0; u < 8; ++u)
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xv[u]) : "r"(tile + 4u * ((r0 + u) * 33 + lane)));
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                if (rbf) {
                  xv[u] = gamma2 * expf(xv[u] - coln);
                  if (sym && mrow + r0 + u == n) xv[u] = gamma2;
                }
                xv[u] *= p.alpha;
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int mm = mrow + r0 + u;
                float* cp = cp0 + (int64_t)(r0 + u) * p.c_rs;
                const bool ok = n_ok && mm < Mi;
                const bool masked = (p.tri_c == VARGP_TRI_LOWER && n > mm) || (p.tri_c == VARGP_TRI_UPPER && n < mm);
                if (p.beta != 0.f) {
                  if (ok && !masked) *cp = fmaf(p.beta, *cp, xv[u]);
                } else if (ok) {
                  *cp = masked ? 0.f : xv[u];
                }
              }
            }
          } else {
            // C not row-major (e.g. written through a transposed view): lanes run along m, which is then the
            // coalesced direction already; the staging tile only serves as a dynamically indexable copy of acc[]
            const int mm = (int)m;
            float* cp0 = Cb + (int64_t)mm * p.c_rs;
#pragma unroll 2
            for (int j = 0; j < 32; ++j) {
              const int n = nc0 + j;
              float x;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(tile + 4u * (lane * 33 + j)));
              if (rbf) {
                x = gamma2 * expf(x - ((n < Ni) ? 0.5f * e_col[n] : 0.f));
                if (sym && mm == n) x = gamma2;
              }
              x *= p.alpha;
              float* cp = cp0 + (int64_t)n * p.c_cs;
              const bool ok = n < Ni && mm < Mi;
              const bool masked = (p.tri_c == VARGP_TRI_LOWER && n > mm) || (p.tri_c == VARGP_TRI_UPPER && n < mm);
              if (p.beta != 0.f) {
                if (ok && !masked) *cp = fmaf(p.beta, *cp, x);
              } else if (ok) {
                *cp = masked ? 0.f : x;
              }
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
static int g_t2_dbg = 0;           // VARGP_TC2_DBG: timing experiments only (results are wrong): 1 no store, 2 no split, 4 no drain

int tc2_init() {
  cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  const char* rna = getenv("VARGP_TC2_RNA");
  if (rna) g_t2_rna = atoi(rna) != 0;
  const char* dbg = getenv("VARGP_TC2_DBG");
  if (dbg) g_t2_dbg = atoi(dbg);
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
  int64_t ncl = tiles < g_t2_clusters ? tiles : g_t2_clusters;
  if (p.sm_limit >= 2 && ncl > p.sm_limit / 2) ncl = p.sm_limit / 2;        // leave SMs to the chain this product runs beside
  int c_vec4 = (p.c_cs == 1 && p.c_rs % 4 == 0 && reinterpret_cast<uintptr_t>(p.C) % 16 == 0) ? 1 : 0;
  for (int i = 0; i < 3; ++i)
    if (p.nb[i] > 1 && p.c_bs[i] % 4 != 0) c_vec4 = 0;
  if (g_t2_rna)
    launch_k((gemm_tc2_kernel<false>), dim3(dim3((unsigned)(2 * ncl))), dim3(T2_THREADS), T2_SMEM_BYTES, stream, tmA, tmB, p, c_vec4, g_t2_dbg);
  else
    launch_k((gemm_tc2_kernel<true>), dim3(dim3((unsigned)(2 * ncl))), dim3(T2_THREADS), T2_SMEM_BYTES, stream, tmA, tmB, p, c_vec4, g_t2_dbg);
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
