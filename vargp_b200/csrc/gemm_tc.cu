// tcgen05 / TMA batched GEMM with fp32-grade accuracy through 3xTF32 splitting (sm_100a).
//
//   C = alpha * A * B (+ beta * C), same descriptor contract as vargp_gemm (gemm_simt.cu), including the
//   triangular k-range skipping, the lower/upper output mask and the fused RBF epilogue.
//
// Pipeline per CTA (one 128 x 128 output tile, K marched in 32-wide slabs, 3 smem stages):
//   warp 0      TMA producer : cp.async.bulk.tensor (SWIZZLE_128B) of the raw fp32 A / B slabs -> smem
//   warps 2-5   splitter     : hi = rna.tf32(x) (in place), lo = x - hi (second buffer); fence.proxy.async
//   warp 1      MMA issuer   : one elected lane issues tcgen05.mma.kind::tf32 for lo*hi, hi*lo, hi*hi
//                              (12 UMMAs of 128x128x8 per slab); tcgen05.commit releases the smem stage and
//                              hands the slab's accumulator to the epilogue
//   warps 6-13  epilogue     : per slab, tcgen05.ld the hi*hi partial sum and add it in fp32 registers (round to
//                              nearest); at the end add the lo accumulator, apply alpha/beta/exp, store
// Accuracy: the tensor core adds into its fp32 accumulator with truncation, which drifts by ~1/4 ulp per MMA
// step (measured: 2.6e-5 relative on a K=784 Gram with a single accumulator).  Therefore the hi*hi products
// of each 32-wide slab go to a fresh ping-pong TMEM buffer (4 accumulation steps only) and are promoted to
// registers, while the 2^-11-times smaller cross terms share one long-running TMEM accumulator.
// Both operand majors are supported: K-contiguous (K-major UMMA descriptor, one 128 B swizzle atom along K)
// and M/N-contiguous (MN-major descriptor, four 32-element chunks along M/N, LBO = chunk stride).
// Structural zeros of triangular operands must be PHYSICALLY zero in memory (TMA cannot mask); callers that
// only "declare" a triangle stay on the SIMT kernel.
//
// Roofline: tensor pipe (3 TF32 MMAs per product); the split doubles the smem footprint of a slab instead of
// the HBM/L2 traffic.  Algorithmic flops per launch: 2*M*N*K*batch (x 1/2 per triangular flag).
//
// A 128 x 64-tile, two-CTAs-per-SM instantiation (<64, 2, 2>: 96 KiB of shared memory, 256 TMEM columns) was measured in
// round 2 at the Split-MNIST shape and LOST (744 vs 755 steps/s: halving the tile doubles the shared-memory reads of
// the A operand per flop, and the kernel is bound by the 128 B/clk shared-memory port); it was removed.
#include "tc_common.cuh"

namespace vargp {

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 448;                                  // 14 warps: TMA, MMA, 4 split, 8 epilogue
constexpr int TC_EPI_WARPS = 8;

template <int BN, int STAGES>
struct TcCfg {
  static constexpr int B_TILE = BN * TC_BK * 4;                  // bytes of one B slab (TC_TILE_BYTES for the A slab)
  static constexpr int SMEM_BYTES = STAGES * 2 * (TC_TILE_BYTES + B_TILE) + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN == 128 ? 512 : 256;        // main0 | main1 | lo | (unused), BN columns each
  static constexpr int EC = BN / 2;                              // output columns per epilogue warp
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int TC_BN, int TC_STAGES, int MINB>
__global__ void __launch_bounds__(TC_THREADS, MINB)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const TcParams p) {
  using Cfg = TcCfg<TC_BN, TC_STAGES>;
  constexpr int B_TILE = Cfg::B_TILE;
  constexpr int TC_TMEM_COLS = Cfg::TMEM_COLS;
  constexpr int EC = Cfg::EC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA_hi = smem;
  uint8_t* sA_lo = sA_hi + TC_STAGES * TC_TILE_BYTES;
  uint8_t* sB_hi = sA_lo + TC_STAGES * TC_TILE_BYTES;
  uint8_t* sB_lo = sB_hi + TC_STAGES * B_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB_lo + TC_STAGES * B_TILE);
  uint64_t* full_bar = bars;                      // TMA landed
  uint64_t* conv_bar = bars + TC_STAGES;          // hi/lo split done
  uint64_t* empty_bar = bars + 2 * TC_STAGES;     // MMAs that read the stage retired
  uint64_t* accf_bar = bars + 3 * TC_STAGES;      // [2] hi*hi partial sum of a slab complete
  uint64_t* acce_bar = bars + 3 * TC_STAGES + 2;  // [2] partial-sum buffer drained by the epilogue
  uint64_t* lo_bar = bars + 3 * TC_STAGES + 4;    // cross-term accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 5);

  pdl_launch_dependents();     // the next kernel's CTAs may be scheduled as soon as every CTA of this grid has started
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define TC_STAMP(i) do { if (stamp) p.dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0) TC_STAMP(0);
  const int64_t m0 = (int64_t)blockIdx.y * TC_BM, n0 = (int64_t)blockIdx.x * TC_BN;
  int64_t z = blockIdx.z;
  const int i2 = (int)(z % p.nb[2]); z /= p.nb[2];
  const int i1 = (int)(z % p.nb[1]);
  const int i0 = (int)(z / p.nb[1]);

  // k-slab range implied by structural zeros / output-triangle culling (warp-uniform)
  bool dead = false;
  if (p.tri_c == VARGP_TRI_LOWER && n0 > m0 + TC_BM - 1) dead = true;
  if (p.tri_c == VARGP_TRI_UPPER && m0 > n0 + TC_BN - 1) dead = true;
  int64_t k_lo = 0, k_hi = p.K;
  if (p.tri_a == VARGP_TRI_LOWER) k_hi = min(k_hi, m0 + TC_BM);
  if (p.tri_a == VARGP_TRI_UPPER) k_lo = max(k_lo, m0);
  if (p.tri_b == VARGP_TRI_LOWER) k_lo = max(k_lo, n0);
  if (p.tri_b == VARGP_TRI_UPPER) k_hi = min(k_hi, n0 + TC_BN);
  const int kb_lo = (int)(k_lo / TC_BK);
  const int nk = (dead || k_hi <= k_lo) ? 0 : (int)((k_hi + TC_BK - 1) / TC_BK) - kb_lo;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < TC_STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&conv_bar[s], 4);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&accf_bar[b], 1);
        mbar_init(&acce_bar[b], TC_EPI_WARPS);
      }
      mbar_init(lo_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // barriers, TMEM and tensor maps are set up; from here on global memory is touched
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int ca2 = p.a_b[2] ? i2 : 0, ca1 = p.a_b[1] ? i1 : 0, ca0 = p.a_b[0] ? i0 : 0;
      const int cb2 = p.b_b[2] ? i2 : 0, cb1 = p.b_b[1] ? i1 : 0, cb0 = p.b_b[0] ? i0 : 0;
      for (int it = 0; it < nk; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], TC_TILE_BYTES + B_TILE);
        const int k0 = (kb_lo + it) * TC_BK;
        uint8_t* da = sA_hi + s * TC_TILE_BYTES;
        uint8_t* db = sB_hi + s * B_TILE;
        if (!p.a_mn) {
          tma_load_5d(&tmA, &full_bar[s], da, k0, (int)m0, ca2, ca1, ca0);
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_5d(&tmA, &full_bar[s], da + c * 4096, (int)m0 + 32 * c, k0, ca2, ca1, ca0);
        }
        if (!p.b_mn) {
          tma_load_5d(&tmB, &full_bar[s], db, k0, (int)n0, cb2, cb1, cb0);       // box of TC_BN rows (make_map)
        } else {
#pragma unroll
          for (int c = 0; c < TC_BN / 32; ++c)
            tma_load_5d(&tmB, &full_bar[s], db + c * 4096, (int)n0 + 32 * c, k0, cb2, cb1, cb0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // instruction descriptor: D=f32, A=B=tf32, majors, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                           ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    for (int it = 0; it < nk; ++it) {
      const int s = it % TC_STAGES;
      const uint32_t ph = (it / TC_STAGES) & 1;
      const int buf = it & 1;
      mbar_wait(&conv_bar[s], ph);
      mbar_wait(&acce_bar[buf], ((it >> 1) & 1) ^ 1);      // epilogue has drained this partial-sum buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(sA_hi + s * TC_TILE_BYTES), a_lo = smem_u32(sA_lo + s * TC_TILE_BYTES);
        const uint32_t b_hi = smem_u32(sB_hi + s * B_TILE), b_lo = smem_u32(sB_lo + s * B_TILE);
        const uint32_t t_main = tmem_base + (uint32_t)(buf * TC_BN), t_lo = tmem_base + 2u * TC_BN;
#pragma unroll
        for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
          // K-major (SWIZZLE_128B): 8 tf32 = 32 B further along the 128 B swizzle row; SBO = 8 rows x 128 B.
          // MN-major tf32 only exists as SWIZZLE_128B_BASE32B: atoms of 4 k-rows x 128 B (SBO = 512 B between
          // k atoms, LBO = 4 KiB between the 32-element M/N chunks); 8 k per UMMA = 1 KiB further.
          const uint32_t oa = p.a_mn ? k8 * 1024 : k8 * 32;
          const uint32_t ob = p.b_mn ? k8 * 1024 : k8 * 32;
          const uint32_t lbo_a = p.a_mn ? 4096 : 16, lbo_b = p.b_mn ? 4096 : 16;
          const uint32_t sbo_a = p.a_mn ? 512 : 1024, sbo_b = p.b_mn ? 512 : 1024;
          const uint64_t la = p.a_mn ? 1 : 2, lb = p.b_mn ? 1 : 2;
          const uint64_t dah = make_desc(a_hi + oa, lbo_a, sbo_a, la), dal = make_desc(a_lo + oa, lbo_a, sbo_a, la);
          const uint64_t dbh = make_desc(b_hi + ob, lbo_b, sbo_b, lb), dbl = make_desc(b_lo + ob, lbo_b, sbo_b, lb);
          umma_tf32(t_lo, dal, dbh, idesc, (it > 0 || k8 > 0) ? 1u : 0u);   // cross terms: one long accumulator
          umma_tf32(t_lo, dah, dbl, idesc, 1u);
          umma_tf32(t_main, dah, dbh, idesc, k8 > 0 ? 1u : 0u);             // hi*hi: fresh partial sum per slab
        }
        umma_commit(&empty_bar[s]);                 // smem stage reusable once these MMAs retire
        umma_commit(&accf_bar[buf]);                // slab partial sum ready for promotion
        if (it == nk - 1) umma_commit(lo_bar);
        if (it == 0) TC_STAMP(3);
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ================= hi / lo splitter (128 threads) =================
    const int t = threadIdx.x - 64;
    for (int it = 0; it < nk; ++it) {
      const int s = it % TC_STAGES;
      const uint32_t ph = (it / TC_STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      if (it == 0 && t == 0) TC_STAMP(2);
      float4* ah = reinterpret_cast<float4*>(sA_hi + s * TC_TILE_BYTES);
      float4* al = reinterpret_cast<float4*>(sA_lo + s * TC_TILE_BYTES);
      float4* bh = reinterpret_cast<float4*>(sB_hi + s * B_TILE);
      float4* bl = reinterpret_cast<float4*>(sB_lo + s * B_TILE);
      // kind::tf32 ignores the 13 low mantissa bits, so the RAW slab is the `hi` operand as it landed; only
      // lo = x - trunc_tf32(x) is written (8 B instead of 12 B of shared-memory traffic per element)
      if constexpr (TC_BN == 128) {
#pragma unroll 4
        for (int e = 0; e < TC_TILE_BYTES / 16 / 128; ++e) {
          const int idx = e * 128 + t;
          const float4 va = ah[idx], vb = bh[idx];
          float4 l;
          l.x = tf32_lo_of(va.x); l.y = tf32_lo_of(va.y); l.z = tf32_lo_of(va.z); l.w = tf32_lo_of(va.w);
          al[idx] = l;
          l.x = tf32_lo_of(vb.x); l.y = tf32_lo_of(vb.y); l.z = tf32_lo_of(vb.z); l.w = tf32_lo_of(vb.w);
          bl[idx] = l;
        }
      } else {
#pragma unroll 4
        for (int e = 0; e < TC_TILE_BYTES / 16 / 128; ++e) {
          const int idx = e * 128 + t;
          const float4 va = ah[idx];
          float4 l;
          l.x = tf32_lo_of(va.x); l.y = tf32_lo_of(va.y); l.z = tf32_lo_of(va.z); l.w = tf32_lo_of(va.w);
          al[idx] = l;
        }
#pragma unroll 4
        for (int e = 0; e < B_TILE / 16 / 128; ++e) {
          const int idx = e * 128 + t;
          const float4 vb = bh[idx];
          float4 l;
          l.x = tf32_lo_of(vb.x); l.y = tf32_lo_of(vb.y); l.z = tf32_lo_of(vb.z); l.w = tf32_lo_of(vb.w);
          bl[idx] = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv_bar[s]);
    }
  } else {
    // ================= epilogue (warps 6..13: TMEM lane quadrant = warp % 4, column half = (warp - 6) / 4) ====
    const int quad = warp & 3;
    const int half = (warp - 6) >> 2;
    const int64_t m = m0 + quad * 32 + lane;
    float* Cb = p.C + i0 * p.c_bs[0] + i1 * p.c_bs[1] + i2 * p.c_bs[2];
    float gamma2 = 1.f, rown = 0.f;
    const float* e_col = nullptr;
    if (p.epi != VARGP_EPI_NONE) {
      gamma2 = expf(2.f * p.e_theta[i0 * p.e_theta_bs[0] + i1 * p.e_theta_bs[1] + i2 * p.e_theta_bs[2] + p.e_D]);
      const float* e_row = p.e_row + i0 * p.e_row_bs[0] + i1 * p.e_row_bs[1] + i2 * p.e_row_bs[2];
      e_col = p.e_col + i0 * p.e_col_bs[0] + i1 * p.e_col_bs[1] + i2 * p.e_col_bs[2];
      if (m < p.M) rown = 0.5f * e_row[m];
    }
    float acc[EC];
#pragma unroll
    for (int j = 0; j < EC; ++j) acc[j] = 0.f;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * EC);
    auto drain = [&](uint32_t taddr) {
#pragma unroll
      for (int c = 0; c < EC / 32; ++c) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr + (uint32_t)(c * 32)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
      }
    };
    for (int it = 0; it < nk; ++it) {
      const int buf = it & 1;
      mbar_wait(&accf_bar[buf], (it >> 1) & 1);
      if (it == 0 && warp == 6 && lane == 0) TC_STAMP(4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      drain(t_row + (uint32_t)(buf * TC_BN));
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acce_bar[buf]);
    }
    if (nk > 0) {
      mbar_wait(lo_bar, 0);
      if (warp == 6 && lane == 0) TC_STAMP(5);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      drain(t_row + 2u * TC_BN);
    }
    // ---- store.  After the drain a thread owns ROW m and 64 consecutive columns; stored like that, each store
    //      instruction of a warp touches 32 different rows (ncu: long-scoreboard stalls on the beta read-modify-write
    //      and 32 sectors per request).  Row-major C: every 32 x 32 block is transposed through a padded per-warp
    //      tile carved out of the (now idle: all MMAs have retired) operand ring, so that a lane owns a COLUMN and a
    //      warp writes one 128 B line per instruction, eight rows in flight.  Other layouts keep lanes along m. ----
    if (p.tma_store) {
      // ---- TMA store.  A thread owns ROW m; it applies the epilogue to its 32-column chunks and lays them down in a
      //      per-warp staging box (32 rows x 128 B, SWIZZLE_128B: 16 B group g of row r sits at (g ^ (r & 7)), which also
      //      makes the 32 float4 stores of a warp bank-conflict free), carved out of the idle operand ring.  One elected
      //      lane hands the box to the TMA engine (cp.async.bulk.tensor store, or cp.reduce ... add for beta == 1); rows /
      //      columns beyond M / N are clipped by the tensor map.  No transposition pass, no per-thread global stores:
      //      the tail of a tile shrinks from ~4.5 us to the time it takes to fill the staging boxes. ----
      const int Mi = (int)p.M, Ni = (int)p.N;
      const int mrow = (int)m0 + quad * 32;
      const bool rbf = p.epi != VARGP_EPI_NONE, sym = p.epi == VARGP_EPI_RBF_SYM;
      const bool add = p.tma_store == 2;
      uint8_t* box0 = smem + (size_t)(warp - 6) * (EC / 32) * 4096;
      bool any = false;
      if (!(add && nk == 0)) {             // culled tile of an accumulating product: nothing to add
#pragma unroll
        for (int c = 0; c < EC / 32; ++c) {
          const int nc0 = (int)n0 + half * EC + c * 32;
          if (nc0 >= Ni || mrow >= Mi) break;                                  // warp-uniform
          float cn = 0.f;
          if (rbf && nc0 + lane < Ni) cn = 0.5f * e_col[nc0 + lane];
          const uint32_t rowb = smem_u32(box0 + c * 4096) + (uint32_t)lane * 128u;
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = g4 * 4 + u;
              const int64_t n = nc0 + j;
              float x = acc[c * 32 + j];
              if (rbf) {
                const float coln = __shfl_sync(0xffffffffu, cn, j);
                x = gamma2 * expf(x - rown - coln);
                if (sym && m == n) x = gamma2;
              }
              x *= p.alpha;
              if ((p.tri_c == VARGP_TRI_LOWER && n > m) || (p.tri_c == VARGP_TRI_UPPER && n < m)) x = 0.f;
              v[u] = x;
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + (uint32_t)((g4 ^ (lane & 7)) << 4)),
                         "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
          }
          any = true;
        }
      }
      if (any) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < EC / 32; ++c) {
            const int nc0 = (int)n0 + half * EC + c * 32;
            if (nc0 >= Ni) break;
            tma_store_5d(&tmC, box0 + c * 4096, nc0, mrow, i2, i1, i0, add);
          }
          tma_store_commit_wait();
        }
      }
    } else {
      float* tile_f = reinterpret_cast<float*>(smem) + (warp - 6) * (32 * 33);
      const uint32_t tile = smem_u32(tile_f);
      const int Mi = (int)p.M, Ni = (int)p.N;
      const int mrow = (int)m0 + quad * 32;
      const bool rbf = p.epi != VARGP_EPI_NONE, sym = p.epi == VARGP_EPI_RBF_SYM;
      const bool row_major = p.c_cs == 1;
      const bool accum = p.beta != 0.f;
#pragma unroll
      for (int c = 0; c < EC / 32; ++c) {
        const int nc0 = (int)n0 + half * EC + c * 32;
        if (nc0 >= Ni || mrow >= Mi) break;                                    // warp-uniform
        if (row_major) {
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(tile + 4u * (lane * 33 + j)), "f"(acc[c * 32 + j] - rown) : "memory");
          __syncwarp();
          const int n = nc0 + lane;
          const bool n_ok = n < Ni;
          float coln = 0.f;
          if (rbf && n_ok) coln = 0.5f * e_col[n];
          float* cp0 = Cb + (int64_t)mrow * p.c_rs + n;
#pragma unroll 1
          for (int r0 = 0; r0 < 32; r0 += 8) {
            if (mrow + r0 >= Mi) break;
            float xv[8], cv[8];
            bool ok[8], masked[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int mm = mrow + r0 + u;
              ok[u] = n_ok && mm < Mi;
              masked[u] = (p.tri_c == VARGP_TRI_LOWER && n > mm) || (p.tri_c == VARGP_TRI_UPPER && n < mm);
              cv[u] = 0.f;
              if (accum && ok[u] && !masked[u]) cv[u] = cp0[(int64_t)(r0 + u) * p.c_rs];
            }
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
              if (accum) xv[u] = fmaf(p.beta, cv[u], xv[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float* cp = cp0 + (int64_t)(r0 + u) * p.c_rs;
              if (accum) {
                if (ok[u] && !masked[u]) *cp = xv[u];
              } else if (ok[u]) {
                *cp = masked[u] ? 0.f : xv[u];
              }
            }
          }
        } else if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int64_t n = nc0 + j;
            if (n >= p.N) break;
            const bool masked = (p.tri_c == VARGP_TRI_LOWER && n > m) || (p.tri_c == VARGP_TRI_UPPER && n < m);
            float* cp = Cb + m * p.c_rs + n * p.c_cs;
            if (masked) {
              if (p.beta == 0.f) *cp = 0.f;
              continue;
            }
            float v = acc[c * 32 + j];
            if (rbf) {
              v = gamma2 * expf(v - rown - 0.5f * e_col[n]);
              if (sym && m == n) v = gamma2;
            }
            v *= p.alpha;
            if (p.beta != 0.f) v = fmaf(p.beta, *cp, v);
            *cp = v;
          }
        }
      }
    }
    if (warp == 6 && lane == 0) TC_STAMP(6);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) TC_STAMP(7);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

long long* g_tc_dbg = nullptr;
bool g_tma_store = true;         // VARGP_TMA_STORE=0: keep the shared-memory transposition + st.global epilogue
EncodeTiledFn g_encode = nullptr;
bool g_tc_ready = false;

int tcp_init();                                                                    // gemm_tcp.cu (persistent 1-CTA form)
bool tcp_wants(const TcParams& p, int64_t ntiles);
int tcp_config(int mode);
int tcp_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const TcParams& p, int64_t gx,
               int64_t gy, int64_t nbatch, cudaStream_t stream);
int tc2_init();                                                                    // gemm_tc2.cu
int tc2_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t stream);
bool tc2_wants(const vargp_gemm_t* g);

}  // namespace vargp

using namespace vargp;

int vargp_tc_init() {
  if (g_tc_ready) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return e != cudaSuccess ? (int)e : VARGP_ERR_NOT_INIT;
  g_encode = (EncodeTiledFn)fn;
  e = cudaFuncSetAttribute(gemm_tc_kernel<128, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128, 3>::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  const char* ts = getenv("VARGP_TMA_STORE");
  if (ts) g_tma_store = atoi(ts) != 0;
  int rc = tc2_init();
  if (rc) return rc;
  rc = tcp_init();
  if (rc) return rc;
  g_tc_ready = true;
  return 0;
}

extern "C" int vargp_gemm_tc(const vargp_gemm_t* g, void* stream) {
  if (!g || !g->A || !g->B || !g->C) return VARGP_ERR_ARG;
  if (!g_tc_ready) return VARGP_ERR_UNSUPPORTED;
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return VARGP_ERR_UNSUPPORTED;
  // operand majors
  const bool a_k = g->a_cs == 1, a_mn = g->a_rs == 1;
  const bool b_k = g->b_rs == 1, b_mn = g->b_cs == 1;
  if (!(a_k || a_mn) || !(b_k || b_mn)) return VARGP_ERR_UNSUPPORTED;
  // too small to pay for the pipeline prologue: leave to the SIMT kernel
  if (g->M * g->N < 32 * 32 || g->K < 32) return VARGP_ERR_UNSUPPORTED;
  const int64_t nbatch = g->nb[0] * g->nb[1] * g->nb[2];
  if (nbatch > 65535 || g->M > (1ll << 30) || g->N > (1ll << 30) || g->K > (1ll << 30)) return VARGP_ERR_UNSUPPORTED;

  TcParams p;
  p.C = g->C; p.M = g->M; p.N = g->N; p.K = g->K; p.c_rs = g->c_rs; p.c_cs = g->c_cs;
  for (int i = 0; i < 3; ++i) {
    p.nb[i] = g->nb[i]; p.c_bs[i] = g->c_bs[i];
    p.e_row_bs[i] = g->e_row_bs[i]; p.e_col_bs[i] = g->e_col_bs[i]; p.e_theta_bs[i] = g->e_theta_bs[i];
  }
  p.alpha = g->alpha; p.beta = g->beta;
  p.tri_a = g->tri_a; p.tri_b = g->tri_b; p.tri_c = g->tri_c; p.epi = g->epi;
  p.e_row = g->e_row; p.e_col = g->e_col; p.e_theta = g->e_theta; p.e_D = g->e_D;
  p.dbg = g_tc_dbg;
  p.sm_limit = (int32_t)(g->sm_limit > 0 && g->sm_limit < (1 << 20) ? g->sm_limit : 0);
  // prefer the K-major form when a dimension of size 1 makes both strides look contiguous
  p.a_mn = (a_k && !(a_mn && g->K == 1)) ? 0 : 1;
  p.b_mn = (b_k && !(b_mn && g->K == 1)) ? 0 : 1;

  const bool big = tc2_wants(g);

  alignas(64) CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, g->A, g->M, g->K, g->a_rs, g->a_cs, g->nb, g->a_bs, p.a_mn, p.a_b);
  if (rc) return rc;
  // B(k, n): "rows" of the operand are n; row stride = b_cs, k stride = b_rs
  rc = make_map(&tmB, g->B, g->N, g->K, g->b_cs, g->b_rs, g->nb, g->b_bs, p.b_mn, p.b_b, TC_ROWS);
  if (rc) return rc;

  // large problems: persistent 2-CTA kernel with 256 x 256 tiles (gemm_tc2.cu)
  p.tma_store = 0;
  if (big) return tc2_launch(tmA, tmB, p, (cudaStream_t)stream);

  // C through the TMA engine when it is row-major with 16 B-aligned rows / batches and beta is 0 or 1
  alignas(64) CUtensorMap tmC = tmA;
  if (g_tma_store && g->c_cs == 1 && (g->beta == 0.f || g->beta == 1.f)) {
    int32_t use_c[3];
    if (make_map(&tmC, g->C, g->M, g->N, g->c_rs, 1, g->nb, g->c_bs, false, use_c, 32) == 0) {
      bool ok = true;
      for (int i = 0; i < 3; ++i) ok = ok && (use_c[i] || g->nb[i] == 1);       // a broadcast C cannot be stored to
      if (ok) p.tma_store = g->beta == 0.f ? 1 : 2;
    }
    if (!p.tma_store) tmC = tmA;
  }

  // more than one wave of tiles: the persistent form (gemm_tcp.cu) overlaps a tile's ramp with the previous tile's store
  {
    const int64_t gx = ceil_div(g->N, 128), gy = ceil_div(g->M, TC_BM);
    // sm_limit < 0: a product that runs BESIDE a critical chain on a lower-priority stream stays one tile per CTA -- resident
    // persistent CTAs cannot be preempted, the chain's next launch would wait for the whole product
    if (g->sm_limit >= 0 && tcp_wants(p, gx * gy * nbatch))
      return tcp_launch(tmA, tmB, tmC, p, gx, gy, nbatch, (cudaStream_t)stream);
  }
  // (a 1-D grid in longest-tile-first order was measured here: products alone 15-20 % faster at triangular operands, the step
  // 2.6 % SLOWER on the same box -- 1242 -> 1210 steps/s, profiles/r3_schedule_ab.txt -- and removed)
  dim3 grid((unsigned)ceil_div(g->N, 128), (unsigned)ceil_div(g->M, TC_BM), (unsigned)nbatch);
  launch_k(gemm_tc_kernel<128, 3, 1>, dim3(grid), dim3(TC_THREADS), TcCfg<128, 3>::SMEM_BYTES, (cudaStream_t)stream, tmA, tmB, tmC, p);
  return launch_status();
}

/* routing of the persistent form of the 1-CTA kernel (gemm_tcp.cu): 0 off, 1 launches of more than one wave of tiles
 * (default; VARGP_TC_PERSIST), 2 every TMA-store launch, 3 launches of at least four waves; < 0 only queries.
 * Returns the previous mode. */
extern "C" int64_t vargp_tc_persist_config(int64_t mode) { return tcp_config((int)mode); }

/* profiling aid: device buffer of >= 8 int64 that CTA (0,0,0) of every following vargp_gemm_tc launch (1-CTA kernel)
 * fills with clock64() stamps: entry, setup done, first slab landed, first slab issued, first partial sum ready,
 * all MMAs retired, stored, exit.  NULL switches it off. */
extern "C" void vargp_tc_debug(long long* buf) { g_tc_dbg = buf; }
