// Data-parallel tail of the ELBO training step as ONE compute + collective pair over NVLink peer memory:
//     flat_g <- sum over ranks of flat_g (gradient all-reduce, experiments-side data parallelism of SURVEY.md section 8e)
//     Yogi update of the replicated parameters with the summed gradient (the optimizer of experiments/vargp.py:23)
//
// The NCCL route (dist.all_reduce + vargp_yogi_step) costs ~25..45 us per step on 2..8 B200s for a 1.9 MB bucket: the
// collective is latency-bound, and it can only start when the last gradient kernel has finished.  Here every rank owns
// a buffer in SYMMETRIC memory (allocated and exchanged by torch.distributed._symmetric_memory; this library only sees
// the peers' device pointers): [staging parity 0 | staging parity 1 | flags].
//
//   kernel 1 (stage):   copy the local gradient into the rank's staging half of this step's parity; the CTA that draws the
//                       last ticket publishes the step number to every peer's flag word (st.release.sys over NVLink)
//   kernel 2 (reduce):  every CTA waits until all peers' flags show this step (ld.acquire.sys on local memory), then each
//                       thread reads its elements from ALL ranks' staging buffers in rank order -- one-shot all-reduce, the
//                       sum is bit-identical on every rank -- applies the Yogi update in registers and writes p, m, v and
//                       the summed gradient.  The last CTA advances the bias-correction powers and the step counter.
//
// One cross-GPU synchronisation per step: the two staging halves alternate, so a rank can only overwrite a half after
// every peer has passed the NEXT step's flag wait, i.e. long after they finished reading it.  No NCCL call, no host
// involvement: both kernels are plain graph nodes of the captured step.
// Roofline: NVLink reads of (R - 1) * n floats per rank + HBM for p, m, v; at n = 0.5 M parameters it is latency-sized.
#include "common.cuh"

namespace vargp {

constexpr int kPeerMax = 8;
struct PeerBufs { float* buf[kPeerMax]; };

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {          // peer memory: never through a (non-coherent) cache
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// in-switch reduction over the multicast mapping of the symmetric buffer (NVLS): the sum of every rank's 16 bytes
__device__ __forceinline__ float4 ld_reduce4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// ctr[0] = steps completed so far, ctr[1] / ctr[2] = ticket counters of the two kernels (zero between launches)
__global__ void __launch_bounds__(256)
peer_stage_kernel(PeerBufs pb, int world, int rank, int64_t n4, int64_t npad, const float* __restrict__ g, unsigned* ctr) {
  pdl_enter();
  __shared__ bool s_last;
  const unsigned done = ctr[0];
  float* mine = pb.buf[rank] + (done & 1u) * npad;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(mine);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) m4[i] = g4[i];
  __threadfence();                              // device scope here; the CTA that publishes the flag adds the system-scope
  __syncthreads();                              // fence after it has seen every CTA's ticket (fences are cumulative)
  if (threadIdx.x == 0) s_last = atomicAdd(&ctr[1], 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if (threadIdx.x < world) {
      unsigned* flags = reinterpret_cast<unsigned*>(pb.buf[threadIdx.x] + 2 * npad);
      st_release_sys(flags + rank, done + 1u);  // "rank `rank` has staged step done + 1" into every rank's flag words
    }
    if (threadIdx.x == 0) ctr[1] = 0u;
  }
}

__global__ void __launch_bounds__(256)
peer_reduce_yogi_kernel(PeerBufs pb, const float* mc, int world, int rank, int64_t n4, int64_t npad, float* __restrict__ gsum,
                        float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float lr, float b1, float b2,
                        float eps, float* pows, unsigned* ctr) {
  pdl_enter();
  __shared__ bool s_last;
  const unsigned done = ctr[0];
  const unsigned* flags = reinterpret_cast<const unsigned*>(pb.buf[rank] + 2 * npad);
  if (threadIdx.x < world) {
    while (ld_acquire_sys(flags + threadIdx.x) < done + 1u) { __nanosleep(20); }
  }
  __syncthreads();
  const float pw0 = pows[0] * b1, pw1 = pows[1] * b2;           // this step's beta^t (advanced below by the last CTA)
  const float bc1 = 1.f - pw0, bc2 = 1.f - pw1;
  const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const int64_t off = (done & 1u) * npad;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  float4* g4 = reinterpret_cast<float4*>(gsum);
  // kPU elements per thread and pass: all of their loads (own + every peer's, over NVLink) are issued before the first
  // use, so a pass costs ONE NVLink round trip instead of one per element
  constexpr int kPU = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += kPU * stride) {
    float4 t[kPU][kPeerMax], pp[kPU], mm[kPU], vv[kPU];
#pragma unroll
    for (int q = 0; q < kPU; ++q) {
      const int64_t i = i0 + q * stride;
      if (mc) {                     // NVLS: ONE load, the NVSwitch sums the ranks' staging buffers (multicast address)
        if (i < n4) t[q][0] = ld_reduce4(mc + off + 4 * i);
      } else {
#pragma unroll
        for (int r = 0; r < kPeerMax; ++r)
          if (r < world && i < n4) t[q][r] = ld_peer4(pb.buf[r] + off + 4 * i);
      }
      if (i < n4) { pp[q] = p4[i]; mm[q] = m4[i]; vv[q] = v4[i]; }
    }
#pragma unroll
    for (int q = 0; q < kPU; ++q) {
      const int64_t i = i0 + q * stride;
      if (i >= n4) break;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mc) {
        s = t[q][0];
      } else {
#pragma unroll
        for (int r = 0; r < kPeerMax; ++r)
          if (r < world) { s.x += t[q][r].x; s.y += t[q][r].y; s.z += t[q][r].z; s.w += t[q][r].w; }
      }
      g4[i] = s;
      const float gs[4] = {s.x, s.y, s.z, s.w};
      float* pe[4] = {&pp[q].x, &pp[q].y, &pp[q].z, &pp[q].w};
      float* me[4] = {&mm[q].x, &mm[q].y, &mm[q].z, &mm[q].w};
      float* ve[4] = {&vv[q].x, &vv[q].y, &vv[q].z, &vv[q].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float gi = gs[e];
        const float g2 = gi * gi;
        const float mi = fmaf(b1, *me[e], (1.f - b1) * gi);
        float vi = *ve[e];
        const float d = vi - g2;
        const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        vi = fmaf(-(1.f - b2) * sg, g2, vi);
        *me[e] = mi;
        *ve[e] = vi;
        *pe[e] -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
      }
      p4[i] = pp[q]; m4[i] = mm[q]; v4[i] = vv[q];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&ctr[2], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {             // every CTA has read pows / ctr[0]: advance them for the next step
    pows[0] = pw0;
    pows[1] = pw1;
    ctr[2] = 0u;
    __threadfence();
    ctr[0] = done + 1u;
  }
}

}  // namespace vargp

using namespace vargp;

// floats of symmetric memory every rank must provide for an n-element gradient (n padded to a multiple of 4, two staging
// halves, 32 flag words); the buffer must be zero-filled once before the first step
extern "C" int64_t vargp_peer_buffer_floats(int64_t n) { return 2 * ((n + 3) / 4 * 4) + 32; }

extern "C" int vargp_peer_allreduce_yogi_nvls(float* const* peer_bufs, const float* multicast, int world, int rank, int64_t n,
                                              float* flat_g, float* p, float* m, float* v, float lr, float b1, float b2,
                                              float eps, float* pows, uint32_t* ctr, void* stream) {
  if (!peer_bufs || world < 1 || world > kPeerMax || rank < 0 || rank >= world || n < 1 || !flat_g || !p || !m || !v || !pows || !ctr)
    return VARGP_ERR_ARG;
  if (n % 4 != 0) return VARGP_ERR_UNSUPPORTED;        // the caller pads the flat buffers to a multiple of 4 elements
  const uintptr_t bits = reinterpret_cast<uintptr_t>(flat_g) | reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v);
  if (bits % 16 != 0) return VARGP_ERR_UNSUPPORTED;
  PeerBufs pb;
  for (int r = 0; r < kPeerMax; ++r) pb.buf[r] = r < world ? peer_bufs[r] : nullptr;
  for (int r = 0; r < world; ++r)
    if (!pb.buf[r] || reinterpret_cast<uintptr_t>(pb.buf[r]) % 16 != 0) return VARGP_ERR_ARG;
  const int64_t n4 = n / 4, npad = n;
  int64_t blocks = ceil_div(n4, 256 * 4);
  if (blocks > 148) blocks = 148;                      // at most one CTA per SM: every CTA spins on the flags
  if (blocks < 1) blocks = 1;
  cudaStream_t s = (cudaStream_t)stream;
  launch_k(peer_stage_kernel, dim3((unsigned)blocks), dim3(256), 0, s, pb, world, rank, n4, npad, (const float*)flat_g,
           reinterpret_cast<unsigned*>(ctr));
  int rc = launch_status();
  if (rc) return rc;
  if (multicast && reinterpret_cast<uintptr_t>(multicast) % 16 != 0) return VARGP_ERR_ARG;
  launch_k(peer_reduce_yogi_kernel, dim3((unsigned)blocks), dim3(256), 0, s, pb, multicast, world, rank, n4, npad, flat_g, p, m, v,
           lr, b1, b2, eps, pows, reinterpret_cast<unsigned*>(ctr));
  return launch_status();
}

extern "C" int vargp_peer_allreduce_yogi(float* const* peer_bufs, int world, int rank, int64_t n, float* flat_g, float* p,
                                         float* m, float* v, float lr, float b1, float b2, float eps, float* pows,
                                         uint32_t* ctr, void* stream) {
  return vargp_peer_allreduce_yogi_nvls(peer_bufs, nullptr, world, rank, n, flat_g, p, m, v, lr, b1, b2, eps, pows, ctr, stream);
}
