// Persistent form of the 1-CTA tcgen05 / TMA 3xTF32 GEMM (gemm_tc.cu) for the products whose C goes out through the TMA
// engine: one CTA per SM walks a static round-robin list of 128 x 128 output tiles.
//
// Why: at the Split / Permuted-MNIST shapes a product is 180 ... 630 tiles = 1.2 ... 4.3 waves of 148 SMs, and a tile costs
// ~3 us of ramp (barriers, TMEM, first TMA round trip, first hi/lo split) + 0.74 us per 32-deep slab + ~2 us of store tail
// (clock64 stamps of gemm_tc.cu).  With one tile per CTA every wave pays ramp and tail; here
//   * barriers, TMEM and tensor-map prefetch are set up once per CTA,
//   * the TMA producer and the splitter run ahead into the NEXT tile's slabs while the epilogue warps store the current one
//     (the store staging boxes are their own 32 KB, not the operand ring, which is never idle here),
//   * the hi*hi partial sums have THREE TMEM buffers instead of two (main0 | main1 | main2 | lo = 512 columns): the epilogue
//     warps drain a tile's cross-term accumulator right after its last MMA and only then store, so the MMA issuer waits
//     ~0.7 us for `lo` and then has three slabs of the next tile to run under the ~2.5 us store (two buffers + a
//     double-buffered `lo` left it idle for ~1.9 us per tile).
// Same contract, numerics and per-slab promotion scheme as gemm_tc_kernel (every partial sum is produced by the same
// instruction sequence: results are bit-identical); only launches with p.tma_store != 0 come here.
#include "tc_common.cuh"

namespace vargp {

constexpr int TP_BM = 128, TP_BN = 128, TP_STAGES = 3;
constexpr int TP_THREADS = 448;                                  // 14 warps: TMA, MMA, 4 split, 8 epilogue
constexpr int TP_EPI_WARPS = 8;
constexpr int TP_B_TILE = TP_BN * TC_BK * 4;
constexpr int TP_RING = TP_STAGES * 2 * (TC_TILE_BYTES + TP_B_TILE);
constexpr int TP_BOXES = TP_EPI_WARPS * 4096;                    // one 32 x 32 fp32 staging box per epilogue warp
constexpr int TP_SMEM_BYTES = TP_RING + TP_BOXES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TP_EC = TP_BN / 2;                                 // output columns per epilogue warp
constexpr int TP_NQ = 4;                                         // depth of the per-CTA tile queue
constexpr int TP_NACC = 3;                                       // hi*hi partial-sum buffers in TMEM (main0 | main1 | main2 | lo)

struct TpTile {
  int64_t m0, n0;
  int i0, i1, i2, kb_lo, nk;
};

// Tile order: TcOrder (tc_common.cuh), positions drawn in ascending order from the launch-wide counter: longest tiles first.

__device__ __forceinline__ TpTile tp_tile(const TcParams& p, const TcOrder& od, int q, int gx) {
  TpTile ti;
  int idx;
  int64_t z;
  if (od.T <= 64) {
    idx = od.perm[q / od.nbatch];
    z = q % od.nbatch;
  } else {
    idx = q % od.T;
    z = q / od.T;
  }
  const int bx = idx % gx, by = idx / gx;
  ti.m0 = (int64_t)by * TP_BM;
  ti.n0 = (int64_t)bx * TP_BN;
  ti.i2 = (int)(z % p.nb[2]); z /= p.nb[2];
  ti.i1 = (int)(z % p.nb[1]);
  ti.i0 = (int)(z / p.nb[1]);
  // k-slab range implied by structural zeros / output-triangle culling
  bool dead = false;
  if (p.tri_c == VARGP_TRI_LOWER && ti.n0 > ti.m0 + TP_BM - 1) dead = true;
  if (p.tri_c == VARGP_TRI_UPPER && ti.m0 > ti.n0 + TP_BN - 1) dead = true;
  int64_t k_lo = 0, k_hi = p.K;
  if (p.tri_a == VARGP_TRI_LOWER) k_hi = min(k_hi, ti.m0 + TP_BM);
  if (p.tri_a == VARGP_TRI_UPPER) k_lo = max(k_lo, ti.m0);
  if (p.tri_b == VARGP_TRI_LOWER) k_lo = max(k_lo, ti.n0);
  if (p.tri_b == VARGP_TRI_UPPER) k_hi = min(k_hi, ti.n0 + TP_BN);
  ti.kb_lo = (int)(k_lo / TC_BK);
  ti.nk = (dead || k_hi <= k_lo) ? 0 : (int)((k_hi + TC_BK - 1) / TC_BK) - ti.kb_lo;
  return ti;
}


// Tiles are handed out by a launch-wide counter (work stealing): a persistent grid is only partly resident when it shares the
// chip with other kernels of the step, and statically assigned tiles of the CTAs that got no SM would wait for them.  The
// producer thread draws the positions (one draw ahead, so that the round trip to L2 is hidden) and publishes them through a
// small shared-memory queue; each of the 13 consumer warps reads an entry and releases it at once.
#define TP_NEXT_TILE(q)                                         \
  {                                                             \
    const int slot_ = nq & (TP_NQ - 1);                         \
    mbar_wait(&tqf_bar[slot_], (nq / TP_NQ) & 1);               \
    q = *reinterpret_cast<volatile int*>(&tq[slot_]);           \
    __syncwarp();                                               \
    if (lane == 0) mbar_arrive(&tqe_bar[slot_]);                \
    ++nq;                                                       \
  }

__global__ void __launch_bounds__(TP_THREADS, 1)
gemm_tcp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const TcParams p, const __grid_constant__ TcOrder od, int gx,
                int ntiles, unsigned* ctr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA_hi = smem;
  uint8_t* sA_lo = sA_hi + TP_STAGES * TC_TILE_BYTES;
  uint8_t* sB_hi = sA_lo + TP_STAGES * TC_TILE_BYTES;
  uint8_t* sB_lo = sB_hi + TP_STAGES * TP_B_TILE;
  uint8_t* boxes = sB_lo + TP_STAGES * TP_B_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(boxes + TP_BOXES);
  uint64_t* full_bar = bars;                       // TMA landed
  uint64_t* conv_bar = bars + TP_STAGES;           // hi/lo split done
  uint64_t* empty_bar = bars + 2 * TP_STAGES;      // MMAs that read the stage retired
  uint64_t* accf_bar = bars + 3 * TP_STAGES;       // [TP_NACC] hi*hi partial sum of a slab complete
  uint64_t* acce_bar = bars + 3 * TP_STAGES + 3;   // [TP_NACC] partial-sum buffer drained by the epilogue
  uint64_t* lof_bar = bars + 3 * TP_STAGES + 6;    // cross-term accumulator of a tile complete
  uint64_t* loe_bar = bars + 3 * TP_STAGES + 7;    // ... drained by the epilogue
  uint64_t* tqf_bar = bars + 3 * TP_STAGES + 8;    // [TP_NQ] tile queue entry published by the producer thread
  uint64_t* tqe_bar = bars + 3 * TP_STAGES + 8 + TP_NQ;   // [TP_NQ] ... read by the 13 consumer warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TP_STAGES + 8 + 2 * TP_NQ);
  int* tq = reinterpret_cast<int*>(tmem_slot + 1);  // [TP_NQ] positions handed out by the launch-wide counter, -1 = end

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool stamp = p.dbg != nullptr && blockIdx.x == 0;
#define TP_STAMP(i) do { if (stamp) p.dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0) TP_STAMP(0);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < TP_STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&conv_bar[s], 4);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < TP_NACC; ++b) {
        mbar_init(&accf_bar[b], 1);
        mbar_init(&acce_bar[b], TP_EPI_WARPS);
      }
      mbar_init(lof_bar, 1);
      mbar_init(loe_bar, TP_EPI_WARPS);
      for (int i = 0; i < TP_NQ; ++i) {
        mbar_init(&tqf_bar[i], 1);
        mbar_init(&tqe_bar[i], 1 + 4 + TP_EPI_WARPS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // barriers, TMEM and tensor maps are set up; from here on global memory is touched
  if (threadIdx.x == 0) TP_STAMP(1);

  if (warp == 0) {
    // ================= TMA producer: runs up to TP_STAGES slabs ahead, across tile boundaries =================
    if (lane == 0) {
      uint32_t g = 0;
      int qn = (int)atomicAdd(ctr, 1u);
      for (int nq = 0;; ++nq) {
        const int q = qn;
        const int slot = nq & (TP_NQ - 1);
        mbar_wait(&tqe_bar[slot], ((nq / TP_NQ) & 1) ^ 1);      // every consumer warp has read the entry this one replaces
        *reinterpret_cast<volatile int*>(&tq[slot]) = q < ntiles ? q : -1;
        mbar_arrive(&tqf_bar[slot]);
        if (q >= ntiles) {
          __threadfence();
          if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) {        // the last CTA of the launch re-arms the counter pair
            ctr[0] = 0;
            ctr[1] = 0;
          }
          break;
        }
        qn = (int)atomicAdd(ctr, 1u);                           // next draw: in flight while this tile is produced
        const TpTile ti = tp_tile(p, od, q, gx);
        const int ca2 = p.a_b[2] ? ti.i2 : 0, ca1 = p.a_b[1] ? ti.i1 : 0, ca0 = p.a_b[0] ? ti.i0 : 0;
        const int cb2 = p.b_b[2] ? ti.i2 : 0, cb1 = p.b_b[1] ? ti.i1 : 0, cb0 = p.b_b[0] ? ti.i0 : 0;
        for (int it = 0; it < ti.nk; ++it, ++g) {
          const int s = g % TP_STAGES;
          const uint32_t ph = (g / TP_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], TC_TILE_BYTES + TP_B_TILE);
          const int k0 = (ti.kb_lo + it) * TC_BK;
          uint8_t* da = sA_hi + s * TC_TILE_BYTES;
          uint8_t* db = sB_hi + s * TP_B_TILE;
          if (!p.a_mn) {
            tma_load_5d(&tmA, &full_bar[s], da, k0, (int)ti.m0, ca2, ca1, ca0);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_load_5d(&tmA, &full_bar[s], da + c * 4096, (int)ti.m0 + 32 * c, k0, ca2, ca1, ca0);
          }
          if (!p.b_mn) {
            tma_load_5d(&tmB, &full_bar[s], db, k0, (int)ti.n0, cb2, cb1, cb0);
          } else {
#pragma unroll
            for (int c = 0; c < TP_BN / 32; ++c)
              tma_load_5d(&tmB, &full_bar[s], db + c * 4096, (int)ti.n0 + 32 * c, k0, cb2, cb1, cb0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                           ((uint32_t)(TP_BN >> 3) << 17) | ((uint32_t)(TP_BM >> 4) << 24);
    uint32_t g = 0, lt = 0;
    for (int nq = 0;;) {
      int q;
      TP_NEXT_TILE(q);
      if (q < 0) break;
      const TpTile ti = tp_tile(p, od, q, gx);
      if (ti.nk == 0) continue;
      mbar_wait(loe_bar, (lt & 1) ^ 1);                     // the previous tile's cross terms have been drained (right after its
                                                            // last MMA, BEFORE its store: the store overlaps three slabs here)
      for (int it = 0; it < ti.nk; ++it, ++g) {
        const int s = g % TP_STAGES;
        const uint32_t ph = (g / TP_STAGES) & 1;
        const int buf = g % TP_NACC;
        mbar_wait(&conv_bar[s], ph);
        mbar_wait(&acce_bar[buf], ((g / TP_NACC) & 1) ^ 1); // epilogue has drained this partial-sum buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a_hi = smem_u32(sA_hi + s * TC_TILE_BYTES), a_lo = smem_u32(sA_lo + s * TC_TILE_BYTES);
          const uint32_t b_hi = smem_u32(sB_hi + s * TP_B_TILE), b_lo = smem_u32(sB_lo + s * TP_B_TILE);
          const uint32_t t_main = tmem_base + (uint32_t)(buf * TP_BN), t_lo = tmem_base + (uint32_t)(TP_NACC * TP_BN);
#pragma unroll
          for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
            const uint32_t oa = p.a_mn ? k8 * 1024 : k8 * 32;
            const uint32_t ob = p.b_mn ? k8 * 1024 : k8 * 32;
            const uint32_t lbo_a = p.a_mn ? 4096 : 16, lbo_b = p.b_mn ? 4096 : 16;
            const uint32_t sbo_a = p.a_mn ? 512 : 1024, sbo_b = p.b_mn ? 512 : 1024;
            const uint64_t la = p.a_mn ? 1 : 2, lb = p.b_mn ? 1 : 2;
            const uint64_t dah = make_desc(a_hi + oa, lbo_a, sbo_a, la), dal = make_desc(a_lo + oa, lbo_a, sbo_a, la);
            const uint64_t dbh = make_desc(b_hi + ob, lbo_b, sbo_b, lb), dbl = make_desc(b_lo + ob, lbo_b, sbo_b, lb);
            umma_tf32(t_lo, dal, dbh, idesc, (it > 0 || k8 > 0) ? 1u : 0u);   // cross terms: one accumulator per tile
            umma_tf32(t_lo, dah, dbl, idesc, 1u);
            umma_tf32(t_main, dah, dbh, idesc, k8 > 0 ? 1u : 0u);             // hi*hi: fresh partial sum per slab
          }
          umma_commit(&empty_bar[s]);
          umma_commit(&accf_bar[buf]);
          if (it == ti.nk - 1) umma_commit(lof_bar);
          if (g == 0) TP_STAMP(3);
        }
        __syncwarp();
      }
      ++lt;
    }
  } else if (warp < 6) {
    // ================= hi / lo splitter (128 threads) =================
    const int tt = threadIdx.x - 64;
    uint32_t g = 0;
    for (int nq = 0;;) {
      int q;
      TP_NEXT_TILE(q);
      if (q < 0) break;
      const TpTile ti = tp_tile(p, od, q, gx);
      for (int it = 0; it < ti.nk; ++it, ++g) {
        const int s = g % TP_STAGES;
        const uint32_t ph = (g / TP_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        if (g == 0 && tt == 0) TP_STAMP(2);
        float4* ah = reinterpret_cast<float4*>(sA_hi + s * TC_TILE_BYTES);
        float4* al = reinterpret_cast<float4*>(sA_lo + s * TC_TILE_BYTES);
        float4* bh = reinterpret_cast<float4*>(sB_hi + s * TP_B_TILE);
        float4* bl = reinterpret_cast<float4*>(sB_lo + s * TP_B_TILE);
#pragma unroll 4
        for (int e = 0; e < TC_TILE_BYTES / 16 / 128; ++e) {
          const int idx = e * 128 + tt;
          const float4 va = ah[idx], vb = bh[idx];
          float4 l;
          l.x = tf32_lo_of(va.x); l.y = tf32_lo_of(va.y); l.z = tf32_lo_of(va.z); l.w = tf32_lo_of(va.w);
          al[idx] = l;
          l.x = tf32_lo_of(vb.x); l.y = tf32_lo_of(vb.y); l.z = tf32_lo_of(vb.z); l.w = tf32_lo_of(vb.w);
          bl[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[s]);
      }
    }
  } else {
    // ================= epilogue (warps 6..13: TMEM lane quadrant = warp % 4, column half = (warp - 6) / 4) ====
    const int quad = warp & 3;
    const int half = (warp - 6) >> 2;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * TP_EC);
    uint8_t* box = boxes + (size_t)(warp - 6) * 4096;
    const int Mi = (int)p.M, Ni = (int)p.N;
    const bool rbf = p.epi != VARGP_EPI_NONE, sym = p.epi == VARGP_EPI_RBF_SYM;
    const bool add = p.tma_store == 2;
    uint32_t g = 0, lt = 0;
    for (int nq = 0;;) {
      int q;
      TP_NEXT_TILE(q);
      if (q < 0) break;
      const TpTile ti = tp_tile(p, od, q, gx);
      const int64_t m = ti.m0 + quad * 32 + lane;
      float gamma2 = 1.f, rown = 0.f;
      const float* e_col = nullptr;
      if (rbf) {
        gamma2 = expf(2.f * p.e_theta[ti.i0 * p.e_theta_bs[0] + ti.i1 * p.e_theta_bs[1] + ti.i2 * p.e_theta_bs[2] + p.e_D]);
        const float* e_row = p.e_row + ti.i0 * p.e_row_bs[0] + ti.i1 * p.e_row_bs[1] + ti.i2 * p.e_row_bs[2];
        e_col = p.e_col + ti.i0 * p.e_col_bs[0] + ti.i1 * p.e_col_bs[1] + ti.i2 * p.e_col_bs[2];
        if (m < p.M) rown = 0.5f * e_row[m];
      }
      float acc[TP_EC];
#pragma unroll
      for (int j = 0; j < TP_EC; ++j) acc[j] = 0.f;
      auto drain = [&](uint32_t taddr) {
#pragma unroll
        for (int c = 0; c < TP_EC / 32; ++c) {
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
      if (ti.nk > 0) {
        for (int it = 0; it < ti.nk; ++it, ++g) {
          const int buf = g % TP_NACC;
          mbar_wait(&accf_bar[buf], (g / TP_NACC) & 1);
          if (g == 0 && warp == 6 && lane == 0) TP_STAMP(4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          drain(t_row + (uint32_t)(buf * TP_BN));
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&acce_bar[buf]);
        }
        mbar_wait(lof_bar, lt & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        drain(t_row + (uint32_t)(TP_NACC * TP_BN));
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(loe_bar);
        if (lt == 0 && warp == 6 && lane == 0) TP_STAMP(5);
        ++lt;
      }
      // ---- TMA store (see gemm_tc.cu): one staging box per warp, re-used chunk after chunk once the engine has read it;
      //      the mainloop warps are already on the next tile ----
      const int mrow = (int)ti.m0 + quad * 32;
      if (!(add && ti.nk == 0)) {
#pragma unroll
        for (int c = 0; c < TP_EC / 32; ++c) {
          const int nc0 = (int)ti.n0 + half * TP_EC + c * 32;
          if (nc0 >= Ni || mrow >= Mi) break;                                  // warp-uniform
          float cn = 0.f;
          if (rbf && nc0 + lane < Ni) cn = 0.5f * e_col[nc0 + lane];
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          const uint32_t rowb = smem_u32(box) + (uint32_t)lane * 128u;
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
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic-proxy writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tmC, box, nc0, mrow, ti.i2, ti.i1, ti.i0, add);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
    }
    if (warp == 6 && lane == 0) TP_STAMP(6);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every store of this warp is complete
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) TP_STAMP(7);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

extern int g_device;             // api.cu: the device of vargp_init
static int g_tp_sms = 0;
static unsigned* g_tp_ctr = nullptr;   // pool of (next position, CTAs done) pairs, one per launch in flight; self re-arming
static unsigned g_tp_seq = 0;
constexpr unsigned TP_CTR_SLOTS = 1024;
int g_tp_mode = 1;               // VARGP_TC_PERSIST: 0 off, 1 on for launches of at least four waves of tiles, 2 always, 3 above one wave

int tcp_init() {
  cudaError_t e = cudaFuncSetAttribute(gemm_tcp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  int cur = 0;
  cudaGetDevice(&cur);
  const int dev = g_device >= 0 ? g_device : cur;         // the device vargp_init was given, whatever is current
  e = cudaDeviceGetAttribute(&g_tp_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess || g_tp_sms <= 0) return e != cudaSuccess ? (int)e : VARGP_ERR_NOT_INIT;
  if (!g_tp_ctr) {
    if (dev != cur) cudaSetDevice(dev);
    e = cudaMalloc(&g_tp_ctr, TP_CTR_SLOTS * 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(g_tp_ctr, 0, TP_CTR_SLOTS * 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();      // callers launch on non-blocking streams: the zeros must be there
    if (dev != cur) cudaSetDevice(cur);
    if (e != cudaSuccess) return (int)e;
  }
  const char* m = getenv("VARGP_TC_PERSIST");
  if (m) g_tp_mode = atoi(m);
  return 0;
}

int tcp_config(int mode) {
  const int old = g_tp_mode;
  if (mode >= 0 && mode <= 3) g_tp_mode = mode;
  return old;
}

bool tcp_wants(const TcParams& p, int64_t ntiles) {
  if (!p.tma_store || g_tp_mode == 0 || g_tp_sms <= 0) return false;
  // Resident persistent CTAs cannot be preempted and assume that all of them ARE resident: beside the capped side-branch
  // products, the cluster factorisation or the whitening kernels of the Split-MNIST step only part of the grid gets an SM and
  // the statically assigned tiles of the rest wait (products alone 10-25 % faster, the step 3 % slower: 1214 -> 1181 steps/s).
  // Default: only the long products (>= 4 waves), which run when the chip is theirs.
  if (g_tp_mode == 2) return true;
  if (g_tp_mode == 3) return ntiles >= 4 * (int64_t)g_tp_sms;
  return ntiles > g_tp_sms;
}

int tcp_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const TcParams& p, int64_t gx,
               int64_t gy, int64_t nbatch, cudaStream_t stream) {
  const int64_t ntiles = gx * gy * nbatch;
  if (ntiles > (1ll << 30)) return VARGP_ERR_UNSUPPORTED;
  const TcOrder od = make_order(p, gx, gy, nbatch, TP_BM, TP_BN);
  const unsigned grid = (unsigned)(ntiles < g_tp_sms ? ntiles : g_tp_sms);
  unsigned* ctr = g_tp_ctr + 2 * (g_tp_seq++ % TP_CTR_SLOTS);
  launch_k(gemm_tcp_kernel, dim3(grid), dim3(TP_THREADS), TP_SMEM_BYTES, stream, tmA, tmB, tmC, p, od, (int)gx, (int)ntiles,
           ctr);
  return launch_status();
}

}  // namespace vargp
