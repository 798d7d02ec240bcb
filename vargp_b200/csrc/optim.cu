// Fused Yogi update over one flat fp32 parameter buffer (the optimizer of experiments/vargp.py:23), one
// launch per step instead of ~10 foreach launches; the flat gradient buffer doubles as the NCCL bucket.
// Streaming kernel: reads p, g, m, v once, writes p, m, v once (28 B per parameter) -> HBM bound.
#include "common.cuh"

namespace vargp {

// pows[0] = beta1^t, pows[1] = beta2^t kept on the device so the step is CUDA-graph replayable
__global__ void yogi_advance_kernel(float* pows, float b1, float b2) {
  pdl_enter();
  pows[0] *= b1;
  pows[1] *= b2;
}

__global__ void __launch_bounds__(256)
yogi_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float lr, float b1, float b2, float eps, const float* __restrict__ pows) {
  pdl_enter();
  const float bc1 = 1.f - pows[0], bc2 = 1.f - pows[1];
  const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float g2 = gi * gi;
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    float vi = v[i];
    const float d = vi - g2;
    const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
    vi = fmaf(-(1.f - b2) * sg, g2, vi);              // v <- v - (1 - b2) sign(v - g^2) g^2
    m[i] = mi;
    v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

}  // namespace vargp

using namespace vargp;

extern "C" int vargp_yogi_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1,
                               float b2, float eps, float* pows, void* stream) {
  if (!p || !g || !m || !v || !pows || n < 0) return VARGP_ERR_ARG;
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  launch_k(yogi_advance_kernel, dim3(1), dim3(1), 0, s, pows, b1, b2);
  int rc = launch_status();
  if (rc) return rc;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(yogi_step_kernel, dim3((unsigned)blocks), dim3(256), 0, s, p, g, m, v, n, lr, b1, b2, eps, pows);
  return launch_status();
}
