// Shared pieces of the tcgen05 / TMA GEMM kernels (gemm_tc.cu: 1-CTA 128x128 tiles; gemm_tc2.cu: persistent
// 2-CTA 256x256 tiles): kernel parameter block, PTX wrappers (mbarrier, TMA, UMMA descriptors, tcgen05.ld),
// and the host-side tensor-map builder.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace vargp {

constexpr int TC_BK = 32;                                        // K slab: 32 fp32 = one 128 B swizzle row
constexpr int TC_ROWS = 128;                                     // operand rows (M or N) per CTA and slab
constexpr int TC_TILE_BYTES = TC_ROWS * TC_BK * 4;               // 16 KiB per operand slab

struct TcParams {
  float* C;
  int64_t M, N, K;
  int64_t c_rs, c_cs;
  int64_t nb[3];
  int64_t c_bs[3];
  float alpha, beta;
  int32_t tri_a, tri_b, tri_c, epi;
  const float* e_row;
  const float* e_col;
  int64_t e_row_bs[3], e_col_bs[3];
  const float* e_theta;
  int64_t e_theta_bs[3], e_D;
  int32_t a_mn, b_mn;            // 1: operand is M/N-contiguous (MN-major), 0: K-contiguous
  int32_t a_b[3], b_b[3];        // 1 if the operand really varies along that batch dim (else coordinate 0)
  long long* dbg;                // optional: 8 clock64() stamps of CTA 0's pipeline (vargp_tc_debug; profiling only)
  int32_t sm_limit;              // 2-CTA kernel: at most this many CTAs (0 = one per SM)
  int32_t tma_store;             // 1-CTA kernel: C goes out through cp.async.bulk.tensor stores (tmC valid); 2 = reduce-add (beta == 1)
};

// Tile order of the 1-CTA kernels.  The tiles of one matrix differ a lot in length when operands are triangular (2 ... 10 slabs
// at P = 300), and whoever draws two long ones sets the makespan of the launch.  Position q of the launch-wide list =
// (class q / nbatch of the tiles of a matrix, sorted by DESCENDING slab count on the host, batch q % nbatch): the hardware
// block scheduler hands positions to SMs as they free up (longest-processing-time-first list scheduling); tiles culled by the
// output triangle (zero fill only) come last.
struct TcOrder {
  uint8_t perm[64];            // sorted position -> tile of the matrix (by * gx + bx); unused when gx * gy > 64
  int T;                       // gx * gy
  int nbatch;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4) : "memory");
}

// smem (32 x 32 fp32 box, SWIZZLE_128B) -> global through the tensor map; `add`: C += box (cp.reduce, fp32 add at L2)
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3, int c4,
                                             bool add) {
  if (!add) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
  } else {
    asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
  }
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the staging tile may be released (CTA exit)
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
  // UMMA shared-memory matrix descriptor (version 1); layout 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;          // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// x -> (hi, lo): hi = x truncated to tf32 (what the tensor core sees when it is fed x itself), lo = x - hi rounded to
// tf32 (round-half-up on the magnitude, 2 integer ops: left to the hardware, lo would be TRUNCATED, a one-sided error)
__device__ __forceinline__ float tf32_lo_of(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
extern EncodeTiledFn g_encode;
extern bool g_tc_ready;

inline TcOrder make_order(const TcParams& p, int64_t gx, int64_t gy, int64_t nbatch, int bm, int bn) {
  TcOrder od = {};
  od.T = (int)(gx * gy);
  od.nbatch = (int)nbatch;
  if (od.T > 64) return od;
  static int lpt = -1;                     // VARGP_TC_LPT=0: row-major tile order (A/B)
  if (lpt < 0) {
    const char* e = getenv("VARGP_TC_LPT");
    lpt = (e && atoi(e) == 0) ? 0 : 1;
  }
  int nk[64];
  for (int idx = 0; idx < od.T; ++idx) {
    const int64_t m0 = (idx / gx) * bm, n0 = (idx % gx) * bn;
    const bool dead = (p.tri_c == VARGP_TRI_LOWER && n0 > m0 + bm - 1) || (p.tri_c == VARGP_TRI_UPPER && m0 > n0 + bn - 1);
    int64_t k_lo = 0, k_hi = p.K;
    if (p.tri_a == VARGP_TRI_LOWER) k_hi = k_hi < m0 + bm ? k_hi : m0 + bm;
    if (p.tri_a == VARGP_TRI_UPPER) k_lo = k_lo > m0 ? k_lo : m0;
    if (p.tri_b == VARGP_TRI_LOWER) k_lo = k_lo > n0 ? k_lo : n0;
    if (p.tri_b == VARGP_TRI_UPPER) k_hi = k_hi < n0 + bn ? k_hi : n0 + bn;
    nk[idx] = (dead || k_hi <= k_lo) ? 0 : (int)((k_hi + TC_BK - 1) / TC_BK - k_lo / TC_BK);
    od.perm[idx] = (uint8_t)idx;
  }
  if (!lpt) return od;
  for (int a = 1; a < od.T; ++a) {         // insertion sort, descending, stable
    const uint8_t v = od.perm[a];
    int b = a;
    while (b > 0 && nk[od.perm[b - 1]] < nk[v]) { od.perm[b] = od.perm[b - 1]; --b; }
    od.perm[b] = v;
  }
  return od;
}

// operand (rows x K) described by (row stride rs, k stride cs): build a 5-D map (inner, outer, b2, b1, b0)
// box_rows: operand rows per TMA box of a K-contiguous operand (an M/N-contiguous one is fetched in 32-row chunks)
inline int make_map(CUtensorMap* tm, const float* base, int64_t rows, int64_t K, int64_t rs, int64_t cs,
                    const int64_t* nb, const int64_t* bs, bool mn_major, int32_t* use_b, int box_rows = TC_ROWS) {
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  cuuint32_t box[5] = {32, 1, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (!mn_major) {            // K contiguous: dims (K, rows)
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows;
    strides[0] = (cuuint64_t)rs * 4;
    box[1] = (cuuint32_t)box_rows;
  } else {                    // rows (M or N) contiguous: dims (rows, K)
    dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K;
    strides[0] = (cuuint64_t)cs * 4;
    box[1] = TC_BK;
  }
  for (int i = 0; i < 3; ++i) {            // map dim 2 <- nb[2] (fastest batch), dim 4 <- nb[0]
    const int b = 2 - i;
    const bool varies = nb[b] > 1 && bs[b] != 0;
    use_b[b] = varies ? 1 : 0;
    dims[2 + i] = varies ? (cuuint64_t)nb[b] : 1;
    // a unit dim still needs a legal (multiple of 16 B, non-zero) stride
    strides[1 + i] = varies ? (cuuint64_t)bs[b] * 4 : strides[0] * dims[1];
  }
  for (int i = 0; i < 4; ++i)
    if (strides[i] % 16 != 0 || strides[i] == 0 || strides[i] >= (1ull << 40)) return VARGP_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return VARGP_ERR_UNSUPPORTED;
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : VARGP_ERR_UNSUPPORTED;
}


}  // namespace vargp
