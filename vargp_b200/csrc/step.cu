// Prologue / epilogue of the fused ELBO training step (vargp_b200/fused_step.py): the parameter plumbing that the
// autograd path spends ~20 small torch launches on (tril unpack + three `cat`s before the forward pass; two sums over
// hyper samples, five slice / accumulate kernels, the tril and hyper adjoints after the backward pass) as ONE launch
// each.  Both are a few hundred KB of streaming work: pure launch-latency savings on the critical path.
#include "common.cuh"

namespace vargp {

__device__ __forceinline__ float softplus_s(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus

// Zcat[c][Q + i][:] = z[c][i][:];  m_last[c][i] = u_mean[c][i];  Lu_last[c] = tril(unpack(u_tril_vec[c])) with a
// softplus diagonal (var_gp/gp_utils.py:22-49; var_gp/vargp.py:52-59 concatenates the same operands with torch.cat)
__global__ void __launch_bounds__(256)
step_assemble_kernel(const float* __restrict__ z, const float* __restrict__ u_mean, const float* __restrict__ u_tril_vec,
                     int64_t C, int64_t M, int64_t D, int64_t P, int64_t Q, float* __restrict__ Zcat,
                     float* __restrict__ m_last, float* __restrict__ Lu_last) {
  pdl_enter();
  const int64_t n1 = C * M * D, n2 = C * M * M, n3 = C * M, T = M * (M + 1) / 2;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n1 + n2 + n3; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < n1) {
      const int64_t c = e / (M * D), r = e - c * M * D;
      Zcat[(c * P + Q) * D + r] = z[e];
    } else if (e < n1 + n2) {
      const int64_t f = e - n1, c = f / (M * M), ij = f - c * M * M, i = ij / M, j = ij - i * M;
      float v = 0.f;
      if (j <= i) {
        v = u_tril_vec[c * T + i * (i + 1) / 2 + j];
        if (i == j) v = softplus_s(v);
      }
      Lu_last[f] = v;
    } else {
      m_last[e - n1 - n2] = u_mean[e - n1 - n2];
    }
  }
}

// z_g[c][i][:]    = Zbar[c][Q + i][:]
// um_g[c][i]      = sum_h mbar[h][c][i]
// ut_g[c][t(i,j)] = (sum_h Lubar[h][c][i][j] - [i == j] g_kl_u / Lu[c][i][i]) * ([i == j] ? sigmoid(vec) : 1)
//                   (adjoint of the KL term -sum_i log Lu_ii and of vec2tril, var_gp/vargp.py:182-190, gp_utils.py:22-49)
// lm_g, llv_g     = adjoint of the hyper sample + KL(q(theta) || p(theta))         (var_gp/kernels.py:62-77)
__global__ void __launch_bounds__(256)
step_grad_finish_kernel(const float* __restrict__ Zbar, const float* __restrict__ mbar, const float* __restrict__ Lubar,
                        const float* __restrict__ Lu, const float* __restrict__ u_tril_vec, const float* __restrict__ g_kl_u,
                        const float* __restrict__ lm, const float* __restrict__ llv, const float* __restrict__ pm,
                        const float* __restrict__ plv, const float* __restrict__ eps, const float* __restrict__ theta_bar,
                        const float* __restrict__ g_kl_h,
                        int64_t H, int64_t C, int64_t M, int64_t D, int64_t P, int64_t Q, int64_t mbar_hs, int64_t Lubar_hs,
                        float* __restrict__ z_g, float* __restrict__ um_g, float* __restrict__ ut_g,
                        float* __restrict__ lm_g, float* __restrict__ llv_g) {
  pdl_enter();
  const int64_t T = M * (M + 1) / 2, D1 = D + 1;
  const int64_t n1 = C * M * D, n2 = C * M * M, n3 = C * M, n4 = D1;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n1 + n2 + n3 + n4; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < n1) {
      const int64_t c = e / (M * D), r = e - c * M * D;
      z_g[e] = Zbar[(c * P + Q) * D + r];
    } else if (e < n1 + n2) {
      const int64_t f = e - n1, c = f / (M * M), ij = f - c * M * M, i = ij / M, j = ij - i * M;
      if (j <= i) {
        float v = 0.f;
        for (int64_t h = 0; h < H; ++h) v += Lubar[h * Lubar_hs + f];
        const int64_t o = c * T + i * (i + 1) / 2 + j;
        if (i == j) {
          if (g_kl_u) v -= g_kl_u[0] / Lu[f];
          v *= 1.f / (1.f + expf(-u_tril_vec[o]));
        }
        ut_g[o] = v;
      }
    } else if (e < n1 + n2 + n3) {
      const int64_t f = e - n1 - n2;
      float v = 0.f;
      for (int64_t h = 0; h < H; ++h) v += mbar[h * mbar_hs + f];
      um_g[f] = v;
    } else {
      const int64_t d = e - n1 - n2 - n3;
      float sm = 0.f, sl = 0.f;
      for (int64_t h = 0; h < H; ++h) {
        const float t = theta_bar[h * D1 + d];
        sm += t;
        sl = fmaf(t, eps[h * D1 + d], sl);
      }
      const float lvd = llv[d], g = g_kl_h[0];
      lm_g[d] = fmaf(g, (lm[d] - pm[d]) * expf(-plv[d]), sm);
      llv_g[d] = fmaf(0.5f * g, expf(lvd - plv[d]) - 1.f, 0.5f * expf(0.5f * lvd) * sl);
    }
  }
}

}  // namespace vargp

using namespace vargp;

extern "C" int vargp_step_assemble(const float* z, const float* u_mean, const float* u_tril_vec, int64_t C, int64_t M,
                                   int64_t D, int64_t P, float* Zcat, float* m_last, float* Lu_last, void* stream) {
  if (!z || !u_mean || !u_tril_vec || !Zcat || !m_last || !Lu_last || C < 1 || M < 1 || D < 1 || P < M) return VARGP_ERR_ARG;
  const int64_t n = C * M * (D + M + 1);
  const int64_t blocks = ceil_div(n, 256) > 148 * 8 ? 148 * 8 : ceil_div(n, 256);
  launch_k(step_assemble_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, z, u_mean, u_tril_vec, C, M, D, P, P - M,
           Zcat, m_last, Lu_last);
  return launch_status();
}

extern "C" int vargp_step_grad_finish(const float* Zbar, const float* mbar, int64_t mbar_hs, const float* Lubar,
                                      int64_t Lubar_hs, const float* Lu, const float* u_tril_vec, const float* g_kl_u,
                                      const float* log_mean, const float* log_logvar, const float* prior_log_mean,
                                      const float* prior_log_logvar, const float* eps, const float* theta_bar,
                                      const float* g_kl_h, int64_t H, int64_t C, int64_t M, int64_t D, int64_t P,
                                      float* z_grad, float* u_mean_grad, float* u_tril_vec_grad, float* log_mean_grad,
                                      float* log_logvar_grad, void* stream) {
  if (!Zbar || !mbar || !Lubar || !Lu || !u_tril_vec || !log_mean || !log_logvar || !prior_log_mean || !prior_log_logvar ||
      !eps || !theta_bar || !g_kl_h || !z_grad || !u_mean_grad || !u_tril_vec_grad || !log_mean_grad || !log_logvar_grad)
    return VARGP_ERR_ARG;
  if (H < 1 || C < 1 || M < 1 || D < 1 || P < M) return VARGP_ERR_ARG;
  const int64_t n = C * M * (D + M + 1) + D + 1;
  const int64_t blocks = ceil_div(n, 256) > 148 * 8 ? 148 * 8 : ceil_div(n, 256);
  launch_k(step_grad_finish_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, Zbar, mbar, Lubar, Lu, u_tril_vec,
           g_kl_u, log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps, theta_bar, g_kl_h, H, C, M, D, P, P - M,
           mbar_hs, Lubar_hs, z_grad, u_mean_grad, u_tril_vec_grad, log_mean_grad, log_logvar_grad);
  return launch_status();
}
