// Kernel hyper-parameters: reparameterised sample of the log-normal variational posterior and its KL to the
// hyper-prior, forward and adjoint, one launch each (var_gp/kernels.py:62-77).  785 elements: pure launch-latency
// work -- as ~30 separate elementwise launches (sample, KL, and their autograd) it was 4 % of the Split-MNIST step.
#include "common.cuh"

namespace vargp {

// theta[h][d] = m[d] + exp(lv[d] / 2) eps[h][d]
// kl = sum_d (exp(lv - plv) + (m - pm)^2 exp(-plv) - 1 - (lv - plv)) / 2        (one block: deterministic sum)
__global__ void __launch_bounds__(256)
hyper_fwd_kernel(const float* __restrict__ m, const float* __restrict__ lv, const float* __restrict__ pm,
                 const float* __restrict__ plv, const float* __restrict__ eps, int64_t H, int64_t D1,
                 float* __restrict__ theta, float* __restrict__ kl) {
  pdl_enter();
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int64_t d = threadIdx.x; d < D1; d += blockDim.x) {
    const float md = m[d], lvd = lv[d], sd = expf(0.5f * lvd);
    for (int64_t h = 0; h < H; ++h) theta[h * D1 + d] = fmaf(sd, eps[h * D1 + d], md);
    if (kl) {
      const float dl = lvd - plv[d], dm = md - pm[d];
      acc += 0.5f * (expf(dl) + dm * dm * expf(-plv[d]) - 1.f - dl);
    }
  }
  if (kl) {
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) kl[0] = acc;
  }
}

// m_bar[d]  = sum_h theta_bar[h][d] + g_kl (m - pm) exp(-plv)
// lv_bar[d] = exp(lv / 2) / 2 * sum_h theta_bar[h][d] eps[h][d] + g_kl (exp(lv - plv) - 1) / 2
__global__ void __launch_bounds__(256)
hyper_bwd_kernel(const float* __restrict__ m, const float* __restrict__ lv, const float* __restrict__ pm,
                 const float* __restrict__ plv, const float* __restrict__ eps, const float* __restrict__ theta_bar,
                 const float* __restrict__ g_kl, int64_t H, int64_t D1, float* __restrict__ m_bar,
                 float* __restrict__ lv_bar) {
  pdl_enter();
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D1) return;
  float sm = 0.f, sl = 0.f;
  if (theta_bar) {
    for (int64_t h = 0; h < H; ++h) {
      const float t = theta_bar[h * D1 + d];
      sm += t;
      sl = fmaf(t, eps[h * D1 + d], sl);
    }
  }
  const float lvd = lv[d];
  float mb = sm, lb = 0.5f * expf(0.5f * lvd) * sl;
  if (g_kl) {
    const float g = g_kl[0];
    mb = fmaf(g, (m[d] - pm[d]) * expf(-plv[d]), mb);
    lb = fmaf(0.5f * g, expf(lvd - plv[d]) - 1.f, lb);
  }
  m_bar[d] = mb;
  lv_bar[d] = lb;
}

}  // namespace vargp

using namespace vargp;

extern "C" int vargp_hyper_fwd(const float* log_mean, const float* log_logvar, const float* prior_log_mean,
                               const float* prior_log_logvar, const float* eps, int64_t H, int64_t D1, float* theta,
                               float* kl, void* stream) {
  if (!log_mean || !log_logvar || !eps || !theta || H < 1 || D1 < 1) return VARGP_ERR_ARG;
  if (kl && (!prior_log_mean || !prior_log_logvar)) return VARGP_ERR_ARG;
  launch_k(hyper_fwd_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps, H,
                                                        D1, theta, kl);
  return launch_status();
}

extern "C" int vargp_hyper_bwd(const float* log_mean, const float* log_logvar, const float* prior_log_mean,
                               const float* prior_log_logvar, const float* eps, const float* theta_bar,
                               const float* g_kl, int64_t H, int64_t D1, float* log_mean_bar, float* log_logvar_bar,
                               void* stream) {
  if (!log_mean || !log_logvar || !eps || !log_mean_bar || !log_logvar_bar || H < 1 || D1 < 1) return VARGP_ERR_ARG;
  if (g_kl && (!prior_log_mean || !prior_log_logvar)) return VARGP_ERR_ARG;
  launch_k(hyper_bwd_kernel, dim3((unsigned)ceil_div(D1, 256)), dim3(256), 0, (cudaStream_t)stream, 
      log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps, theta_bar, g_kl, H, D1, log_mean_bar, log_logvar_bar);
  return launch_status();
}
