// Predictive-marginal reductions, KL(u), packed-triangle unpack and the small adjoint helpers.
// All streaming SIMT kernels (HBM / L2 bound): threads run along the contiguous minibatch axis.
#include "common.cuh"

namespace vargp {

// f_mean[g][b] = sum_p nu[g][p] V[g][p][b]; f_var = gamma2 - sum V^2 + sum TV^2 + jitter sum A^2
// grid (B tiles, G); algorithmic bytes 4*(3*G*P*B + 2*G*B).          (gp_utils.py:178-186)
__global__ void __launch_bounds__(128)
marginal_reduce_kernel(const float* __restrict__ V, const float* __restrict__ TV, const float* __restrict__ A,
                       const float* __restrict__ nu, const float* __restrict__ theta, int64_t theta_rs, int64_t D,
                       int64_t C, int64_t P, int64_t B, float jitter,
                       float* __restrict__ f_mean, float* __restrict__ f_var) {
  const int64_t g = blockIdx.y;
  const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (b >= B) return;
  const float* v = V + g * P * B + b;
  const float* tv = TV + g * P * B + b;
  const float* a = A + g * P * B + b;
  const float* nug = nu + g * P;
  float m = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
  for (int64_t p = 0; p < P; ++p) {
    const float vv = v[p * B], tt = tv[p * B], aa = a[p * B];
    m = fmaf(__ldg(nug + p), vv, m);
    q1 = fmaf(vv, vv, q1);
    q2 = fmaf(tt, tt, q2);
    q3 = fmaf(aa, aa, q3);
  }
  const float gamma2 = expf(2.f * theta[(g / C) * theta_rs + D]);
  f_mean[g * B + b] = m;
  f_var[g * B + b] = gamma2 - q1 + q2 + jitter * q3;
}

// Vbar = nu gm^T - 2 V gv; A *= 2 jitter gv; TV *= 2 gv; theta_bar[h][D] += 2 gamma2 sum_{c,b} gv
// grid (B tiles, P chunks, G)
constexpr int kPrepPch = 16;
__global__ void __launch_bounds__(128)
marginal_bwd_prep_kernel(const float* __restrict__ V, float* __restrict__ TV, float* __restrict__ A,
                         const float* __restrict__ nu, const float* __restrict__ g_mean,
                         const float* __restrict__ g_var, const float* __restrict__ theta, int64_t theta_rs,
                         int64_t D, int64_t C, int64_t P, int64_t B, float jitter,
                         float* __restrict__ Vbar, float* __restrict__ theta_bar) {
  __shared__ float scratch[32];
  const int64_t g = blockIdx.z;
  const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int64_t p0 = (int64_t)blockIdx.y * kPrepPch, p1 = min(P, p0 + kPrepPch);
  const bool live = b < B;
  const float gm = live ? g_mean[g * B + b] : 0.f;
  const float gv = live ? g_var[g * B + b] : 0.f;
  if (live) {
    const float two_gv = 2.f * gv, eps_gv = 2.f * jitter * gv;
    for (int64_t p = p0; p < p1; ++p) {
      const int64_t o = (g * P + p) * B + b;
      Vbar[o] = fmaf(__ldg(nu + g * P + p), gm, -two_gv * V[o]);
      A[o] *= eps_gv;
      TV[o] *= two_gv;
    }
  }
  if (blockIdx.y == 0) {
    const float s = block_sum(gv, scratch);
    if (threadIdx.x == 0) {
      const int64_t h = g / C;
      atomicAdd(theta_bar + h * (D + 1) + D, 2.f * expf(2.f * theta[h * theta_rs + D]) * s);
    }
  }
}

// X <- (Phi(X) + Phi(X)^T)/2 : Xi_ij = Xi_ji = X_ij / 2 for i >= j
__global__ void sym_phi_kernel(float* __restrict__ X, int64_t n) {
  float* x = X + (int64_t)blockIdx.z * n * n;
  const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
  if (i >= n || j >= n || j > i) return;
  const float v = 0.5f * x[i * n + j];
  x[i * n + j] = v;
  x[j * n + i] = v;
}

// KL(u) forward, two deterministic stages (bit-reproducible: no float atomics):
//   stage 1: grid (kKlChunks, H*C): chunk k of (h, c) sums its slice of |T_t|_F^2 (chunk 0 also the O(M) terms)
//            -> part[g][k];  stage 2: one warp sums part[] in a fixed order.                       (vargp.py:182-190)
constexpr int kKlChunks = VARGP_KL_CHUNKS;

__global__ void __launch_bounds__(256)
kl_fwd_part_kernel(const float* __restrict__ W, const float* __restrict__ T, const float* __restrict__ nu,
                   const float* __restrict__ Lu, int64_t C, int64_t P, int64_t M, float* __restrict__ part) {
  __shared__ float scratch[32];
  const int64_t g = blockIdx.y, c = g % C, S = P / M, Q = P - M;
  const float* w = W + g * P * P;
  const float* t = T + (g * S + (S - 1)) * M * M;
  const float* nug = nu + g * P + Q;
  const float* lu = Lu + c * M * M;
  float acc = 0.f;
  if (blockIdx.x == 0) {
    for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
      acc -= logf(w[(Q + i) * P + (Q + i)]);
      acc -= logf(lu[i * M + i]);
      const float v = nug[i];
      acc += 0.5f * (v * v - 1.f);
    }
  }
  // rows of T_t are dealt round-robin to the chunks; only the lower triangle is non-zero
  for (int64_t i = blockIdx.x; i < M; i += kKlChunks) {
    const float* row = t + i * M;
    for (int64_t j = threadIdx.x; j <= i; j += blockDim.x) {
      const float v = row[j];
      acc = fmaf(0.5f * v, v, acc);
    }
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) part[g * kKlChunks + blockIdx.x] = acc;
}

__global__ void __launch_bounds__(32)
kl_fwd_sum_kernel(const float* __restrict__ part, int64_t n, int64_t H, float* __restrict__ kl) {
  float acc = 0.f;
  for (int64_t g = threadIdx.x; g < n; g += 32) acc += part[g];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) kl[0] += acc / (float)H;
}

// grid (row chunks, H*C)
__global__ void __launch_bounds__(256)
kl_bwd_kernel(const float* __restrict__ W, const float* __restrict__ T, const float* __restrict__ nu,
              const float* __restrict__ g_kl, int64_t H, int64_t C, int64_t P, int64_t M,
              float* __restrict__ Wbar, float* __restrict__ Tbar, float* __restrict__ nubar) {
  const int64_t g = blockIdx.y, S = P / M, Q = P - M;
  const float s = g_kl[0] / (float)H;
  const float* t = T + (g * S + (S - 1)) * M * M;
  float* tb = Tbar + (g * S + (S - 1)) * M * M;
  for (int64_t i = blockIdx.x; i < M; i += gridDim.x)
    for (int64_t j = threadIdx.x; j <= i; j += blockDim.x) tb[i * M + j] = fmaf(s, t[i * M + j], tb[i * M + j]);
  if (blockIdx.x == 0) {
    for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
      nubar[g * P + Q + i] = fmaf(s, nu[g * P + Q + i], nubar[g * P + Q + i]);
      const int64_t o = g * P * P + (Q + i) * P + (Q + i);
      Wbar[o] -= s / W[o];
    }
  }
}

__global__ void kl_bwd_lu_kernel(const float* __restrict__ Lu, const float* __restrict__ g_kl, int64_t C, int64_t M,
                                 float* __restrict__ Lubar) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= C * M) return;
  const int64_t c = e / M, i = e % M;
  const int64_t o = c * M * M + i * M + i;
  Lubar[o] -= g_kl[0] / Lu[o];
}

__device__ __forceinline__ float softplus_f(float x) {   // torch.nn.functional.softplus, threshold 20
  return x > 20.f ? x : log1pf(expf(x));
}

// packed row-major lower triangle -> dense, softplus on the diagonal      (gp_utils.py:22-49)
__global__ void tril_unpack_kernel(const float* __restrict__ vec, int64_t M, float* __restrict__ out) {
  const int64_t c = blockIdx.z;
  const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
  if (i >= M || j >= M) return;
  const int64_t T = M * (M + 1) / 2;
  float v = 0.f;
  if (j <= i) {
    v = vec[c * T + i * (i + 1) / 2 + j];
    if (i == j) v = softplus_f(v);
  }
  out[(c * M + i) * M + j] = v;
}

__global__ void tril_unpack_bwd_kernel(const float* __restrict__ Lbar, const float* __restrict__ vec, int64_t M,
                                       float* __restrict__ vec_bar) {
  const int64_t c = blockIdx.z;
  const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
  if (i >= M || j > i) return;
  const int64_t T = M * (M + 1) / 2;
  const int64_t o = c * T + i * (i + 1) / 2 + j;
  float gvl = Lbar[(c * M + i) * M + j];
  if (i == j) gvl *= 1.f / (1.f + expf(-vec[o]));
  vec_bar[o] = gvl;
}

}  // namespace vargp

using namespace vargp;

extern "C" int vargp_marginal_reduce(const float* V, const float* TV, const float* A, const float* nu,
                                     const float* theta, int64_t theta_rs, int64_t D, int64_t H, int64_t C,
                                     int64_t P, int64_t B, float jitter, float* f_mean, float* f_var,
                                     void* stream) {
  if (!V || !TV || !A || !nu || !theta || !f_mean || !f_var) return VARGP_ERR_ARG;
  if (H * C > 65535) return VARGP_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  dim3 grid((unsigned)ceil_div(B, 128), (unsigned)(H * C));
  marginal_reduce_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(V, TV, A, nu, theta, theta_rs, D, C, P, B, jitter,
                                                                  f_mean, f_var);
  return launch_status();
}

extern "C" int vargp_marginal_bwd_prep(const float* V, float* TV, float* A, const float* nu, const float* g_mean,
                                       const float* g_var, const float* theta, int64_t theta_rs, int64_t D,
                                       int64_t H, int64_t C, int64_t P, int64_t B, float jitter, float* Vbar,
                                       float* theta_bar, void* stream) {
  if (!V || !TV || !A || !nu || !g_mean || !g_var || !theta || !Vbar || !theta_bar) return VARGP_ERR_ARG;
  if (H * C > 65535 || ceil_div(P, kPrepPch) > 65535) return VARGP_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  dim3 grid((unsigned)ceil_div(B, 128), (unsigned)ceil_div(P, kPrepPch), (unsigned)(H * C));
  marginal_bwd_prep_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(V, TV, A, nu, g_mean, g_var, theta, theta_rs, D,
                                                                    C, P, B, jitter, Vbar, theta_bar);
  return launch_status();
}

extern "C" int vargp_sym_phi(float* X, int64_t n, int64_t batch, void* stream) {
  if (!X || n < 1 || batch < 1) return VARGP_ERR_ARG;
  if (batch > 65535) return VARGP_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(n, 8), (unsigned)batch);
  sym_phi_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(X, n);
  return launch_status();
}

extern "C" int vargp_kl_fwd(const float* W, const float* T, const float* nu, const float* Lu, int64_t H, int64_t C,
                            int64_t P, int64_t M, float* kl, float* work, void* stream) {
  if (!W || !T || !nu || !Lu || !kl || !work || M < 1 || P % M) return VARGP_ERR_ARG;
  if (H * C > 65535) return VARGP_ERR_UNSUPPORTED;
  kl_fwd_part_kernel<<<dim3(kKlChunks, (unsigned)(H * C)), 256, 0, (cudaStream_t)stream>>>(W, T, nu, Lu, C, P, M, work);
  int rc = launch_status();
  if (rc) return rc;
  kl_fwd_sum_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(work, H * C * kKlChunks, H, kl);
  return launch_status();
}

extern "C" int vargp_kl_bwd(const float* W, const float* T, const float* nu, const float* g_kl, int64_t H,
                            int64_t C, int64_t P, int64_t M, float* Wbar, float* Tbar, float* nubar, void* stream) {
  if (!W || !T || !nu || !g_kl || !Wbar || !Tbar || !nubar || M < 1 || P % M) return VARGP_ERR_ARG;
  if (H * C > 65535) return VARGP_ERR_UNSUPPORTED;
  const unsigned chunks = (unsigned)(M < 64 ? 1 : (M / 32 > 128 ? 128 : M / 32));
  kl_bwd_kernel<<<dim3(chunks, (unsigned)(H * C)), 256, 0, (cudaStream_t)stream>>>(W, T, nu, g_kl, H, C, P, M, Wbar,
                                                                                   Tbar, nubar);
  return launch_status();
}

extern "C" int vargp_kl_bwd_lu(const float* Lu, const float* g_kl, int64_t C, int64_t M, float* Lubar,
                               void* stream) {
  if (!Lu || !g_kl || !Lubar) return VARGP_ERR_ARG;
  kl_bwd_lu_kernel<<<(unsigned)ceil_div(C * M, 128), 128, 0, (cudaStream_t)stream>>>(Lu, g_kl, C, M, Lubar);
  return launch_status();
}

extern "C" int vargp_tril_unpack(const float* vec, int64_t C, int64_t M, float* out, void* stream) {
  if (!vec || !out || C < 1 || M < 1) return VARGP_ERR_ARG;
  if (C > 65535) return VARGP_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(M, 32), (unsigned)ceil_div(M, 8), (unsigned)C);
  tril_unpack_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(vec, M, out);
  return launch_status();
}

extern "C" int vargp_tril_unpack_bwd(const float* Lbar, const float* vec, int64_t C, int64_t M, float* vec_bar,
                                     void* stream) {
  if (!Lbar || !vec || !vec_bar || C < 1 || M < 1) return VARGP_ERR_ARG;
  if (C > 65535) return VARGP_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(M, 32), (unsigned)ceil_div(M, 8), (unsigned)C);
  tril_unpack_bwd_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(Lbar, vec, M, vec_bar);
  return launch_status();
}
