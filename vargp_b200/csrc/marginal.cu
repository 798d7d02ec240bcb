// Predictive-marginal reductions, KL(u), packed-triangle unpack and the small adjoint helpers.
// All streaming SIMT kernels (HBM / L2 bound): threads run along the contiguous minibatch axis.
#include "common.cuh"

namespace vargp {

// f_mean[g][b] = sum_p nu[g][p] V[g][p][b];  f_var[g][b] = gamma2 + sum_p V (NV - V)
// (NV = N V with N = blockdiag(T_s T_s^T) + jitter W W^T: gamma2 - |V|^2 + |T^T V|^2 + jitter |W^T V|^2)
// grid (B tiles, G); thread = one minibatch column, 4 independent accumulator chains over p.
// Algorithmic bytes 4*(2*G*P*B + 2*G*B).                                           (gp_utils.py:178-186)
__global__ void __launch_bounds__(128)
marginal_reduce_kernel(const float* __restrict__ V, const float* __restrict__ NV, const float* __restrict__ nu,
                       const float* __restrict__ theta, int64_t theta_rs, int64_t D, int64_t C, int64_t P,
                       int64_t B, float* __restrict__ f_mean, float* __restrict__ f_var) {
  pdl_enter();
  const int64_t g = blockIdx.y;
  const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (b >= B) return;
  const float* v = V + g * P * B + b;
  const float* nv = NV + g * P * B + b;
  const float* nug = nu + g * P;
  float m[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  int64_t p = 0;
  for (; p + 4 <= P; p += 4) {
    float vv[4], nn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vv[u] = v[(p + u) * B];
      nn[u] = nv[(p + u) * B];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      m[u] = fmaf(__ldg(nug + p + u), vv[u], m[u]);
      q[u] = fmaf(vv[u], nn[u] - vv[u], q[u]);
    }
  }
  for (; p < P; ++p) {
    const float vv = v[p * B], nn = nv[p * B];
    m[0] = fmaf(__ldg(nug + p), vv, m[0]);
    q[0] = fmaf(vv, nn - vv, q[0]);
  }
  const float gamma2 = expf(2.f * theta[(g / C) * theta_rs + D]);
  f_mean[g * B + b] = (m[0] + m[1]) + (m[2] + m[3]);
  f_var[g * B + b] = gamma2 + ((q[0] + q[1]) + (q[2] + q[3]));
}

// Same contract for SMALL problems (Split / Permuted-MNIST shapes: a few MB that sit in L2).  One thread per column
// walking all P rows is latency-bound there (P / 4 dependent round trips to L2 on 120 CTAs: 37 us at P = 300); here a
// CTA owns 32 columns and its 8 warps split the rows (P / 32 round trips, 4 x as many CTAs), then reduce through smem.
__global__ void __launch_bounds__(256)
marginal_reduce_split_kernel(const float* __restrict__ V, const float* __restrict__ NV, const float* __restrict__ nu,
                             const float* __restrict__ theta, int64_t theta_rs, int64_t D, int64_t C, int64_t P,
                             int64_t B, float* __restrict__ f_mean, float* __restrict__ f_var) {
  pdl_enter();
  __shared__ float s_m[8][33], s_q[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t g = blockIdx.y;
  const int64_t b = (int64_t)blockIdx.x * 32 + lane;
  const bool live = b < B;
  const float* v = V + g * P * B + b;
  const float* nv = NV + g * P * B + b;
  const float* nug = nu + g * P;
  float m[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t p0 = w; p0 < P; p0 += 32) {
    float vv[4], nn[4], nn_[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t p = p0 + 8 * u;
      const bool ok = live && p < P;
      vv[u] = ok ? v[p * B] : 0.f;
      nn[u] = ok ? nv[p * B] : 0.f;
      nn_[u] = p < P ? __ldg(nug + p) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      m[u] = fmaf(nn_[u], vv[u], m[u]);
      q[u] = fmaf(vv[u], nn[u] - vv[u], q[u]);
    }
  }
  s_m[w][lane] = (m[0] + m[1]) + (m[2] + m[3]);
  s_q[w][lane] = (q[0] + q[1]) + (q[2] + q[3]);
  __syncthreads();
  if (w == 0 && live) {
    float ms = 0.f, qs = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { ms += s_m[k][lane]; qs += s_q[k][lane]; }
    const float gamma2 = expf(2.f * theta[(g / C) * theta_rs + D]);
    f_mean[g * B + b] = ms;
    f_var[g * B + b] = gamma2 + qs;
  }
}

// Vbar = nu gm^T + 2 gv (NV - V)  (Vbar may alias NV);  Vg = gv V;  theta_bar[h][D] += 2 gamma2 sum_{c,b} gv
// grid (B tiles of 128*VEC, P chunks of kPrepPch rows, G); VEC = 4: float4 along the minibatch axis.
// Algorithmic bytes 4*(4*G*P*B + 2*G*B).
constexpr int kPrepPch = 8;
template <int VEC>
__global__ void __launch_bounds__(128)
marginal_bwd_prep_kernel(const float* __restrict__ V, const float* NV, const float* __restrict__ nu,
                         const float* __restrict__ g_mean, const float* __restrict__ g_var,
                         const float* __restrict__ theta, int64_t theta_rs, int64_t D, int64_t C, int64_t P,
                         int64_t B, float* Vbar, float* __restrict__ Vg, float* __restrict__ theta_bar) {
  pdl_enter();
  __shared__ float scratch[32];
  const int64_t g = blockIdx.z;
  const int64_t b = ((int64_t)blockIdx.x * 128 + threadIdx.x) * VEC;
  const int64_t p0 = (int64_t)blockIdx.y * kPrepPch;
  const int rows = (int)min((int64_t)kPrepPch, P - p0);
  const bool live = b < B;
  float gm[VEC], gv[VEC];
#pragma unroll
  for (int u = 0; u < VEC; ++u) gm[u] = gv[u] = 0.f;
  if (live) {
    if (VEC == 4) {
      const float4 a = *reinterpret_cast<const float4*>(g_mean + g * B + b);
      const float4 c = *reinterpret_cast<const float4*>(g_var + g * B + b);
      gm[0] = a.x; gm[1] = a.y; gm[2] = a.z; gm[3] = a.w;
      gv[0] = c.x; gv[1] = c.y; gv[2] = c.z; gv[3] = c.w;
    } else {
      gm[0] = g_mean[g * B + b];
      gv[0] = g_var[g * B + b];
    }
    float vv[kPrepPch][VEC], nn[kPrepPch][VEC];
#pragma unroll
    for (int r = 0; r < kPrepPch; ++r) {
      if (r < rows) {
        const int64_t o = (g * P + p0 + r) * B + b;
        if (VEC == 4) {
          const float4 a = *reinterpret_cast<const float4*>(V + o);
          const float4 c = *reinterpret_cast<const float4*>(NV + o);
          vv[r][0] = a.x; vv[r][1] = a.y; vv[r][2] = a.z; vv[r][3] = a.w;
          nn[r][0] = c.x; nn[r][1] = c.y; nn[r][2] = c.z; nn[r][3] = c.w;
        } else {
          vv[r][0] = V[o];
          nn[r][0] = NV[o];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kPrepPch; ++r) {
      if (r < rows) {
        const int64_t o = (g * P + p0 + r) * B + b;
        const float nur = __ldg(nu + g * P + p0 + r);
        float vb[VEC], vg[VEC];
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
          vb[u] = fmaf(nur, gm[u], 2.f * gv[u] * (nn[r][u] - vv[r][u]));
          vg[u] = gv[u] * vv[r][u];
        }
        if (VEC == 4) {
          *reinterpret_cast<float4*>(Vbar + o) = make_float4(vb[0], vb[1], vb[2], vb[3]);
          *reinterpret_cast<float4*>(Vg + o) = make_float4(vg[0], vg[1], vg[2], vg[3]);
        } else {
          Vbar[o] = vb[0];
          Vg[o] = vg[0];
        }
      }
    }
  }
  if (blockIdx.y == 0) {
    float t = 0.f;
#pragma unroll
    for (int u = 0; u < VEC; ++u) t += gv[u];
    const float s = block_sum(t, scratch);
    if (threadIdx.x == 0) {
      const int64_t h = g / C;
      atomicAdd(theta_bar + h * (D + 1) + D, 2.f * expf(2.f * theta[h * theta_rs + D]) * s);
    }
  }
}

// mirror == 0: X <- (Phi(X) + Phi(X)^T)/2, i.e. Xi_ij = Xi_ji = X_ij / 2 for i >= j;  mirror != 0: X_ji <- X_ij (i >= j)
// grid (tile column bj, tile row bi >= bj, batch), 64 x 64 tiles through shared memory so that both the lower tile
// and its mirror image are written with full 128 B lines.
constexpr int kSymT = 64;
__global__ void __launch_bounds__(256)
sym_phi_kernel(float* __restrict__ X, int64_t n, float scale) {
  pdl_enter();
  __shared__ float tile[kSymT][kSymT + 1];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  float* x = X + (int64_t)blockIdx.z * n * n;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
  const int64_t i0 = (int64_t)bi * kSymT, j0 = (int64_t)bj * kSymT;
  // all 16 loads of a thread in flight (unrolled by 4 this was four dependent round trips to L2 per tile, 12 ... 25 us per call
  // on the critical chain); the mirror-only form (scale == 1) does not rewrite the lower triangle
  float vals[kSymT / 4];
#pragma unroll
  for (int q = 0; q < kSymT / 4; ++q) {
    const int64_t i = i0 + ty + 4 * q, j = j0 + tx;
    vals[q] = (i < n && j < n && j <= i) ? x[i * n + j] : 0.f;
  }
#pragma unroll
  for (int q = 0; q < kSymT / 4; ++q) {
    const int r = ty + 4 * q;
    const int64_t i = i0 + r, j = j0 + tx;
    const float v = scale * vals[q];
    if (scale != 1.f && i < n && j < n && j <= i) x[i * n + j] = v;
    tile[r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < kSymT; r += 4) {
    // mirrored element (row j0 + r, col i0 + tx) = lower element (i0 + tx, j0 + r)
    const int64_t jj = j0 + r, ii = i0 + tx;
    if (ii < n && jj < n && jj < ii) x[jj * n + ii] = tile[tx][r];
  }
}

// KL(u) forward, two deterministic stages (bit-reproducible: no float atomics):
//   stage 1: grid (kKlChunks, H*C): chunk k of (h, c) sums its slice of |T_t|_F^2 (chunk 0 also the O(M) terms)
//            -> part[g][k];  stage 2: one warp sums part[] in a fixed order.                       (vargp.py:182-190)
constexpr int kKlChunks = VARGP_KL_CHUNKS;

__global__ void __launch_bounds__(256)
kl_fwd_part_kernel(const float* __restrict__ W, const float* __restrict__ T, const float* __restrict__ nu,
                   const float* __restrict__ Lu, int64_t C, int64_t P, int64_t M, float* __restrict__ part) {
  pdl_enter();
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
  pdl_enter();
  float acc = 0.f;
  for (int64_t g = threadIdx.x; g < n; g += 32) acc += part[g];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) kl[0] += acc / (float)H;
}

// grid (element chunks of 256 over the M x M block, H*C): one thread per element of T_t (a row loop per CTA was
// latency-bound: 30 us at M = 60); chunk 0 also handles the O(M) terms
__global__ void __launch_bounds__(256)
kl_bwd_kernel(const float* __restrict__ W, const float* __restrict__ T, const float* __restrict__ nu,
              const float* __restrict__ g_kl, int64_t H, int64_t C, int64_t P, int64_t M,
              float* __restrict__ Wbar, float* __restrict__ Tbar, float* __restrict__ nubar) {
  pdl_enter();
  const int64_t g = blockIdx.y, S = P / M, Q = P - M;
  const float s = g_kl[0] / (float)H;
  const float* t = T + (g * S + (S - 1)) * M * M;
  float* tb = Tbar + (g * S + (S - 1)) * M * M;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * M; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / M, j = e - i * M;
    if (j <= i) tb[e] = fmaf(s, t[e], tb[e]);
  }
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
  pdl_enter();
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
  pdl_enter();
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
  pdl_enter();
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

extern "C" int vargp_marginal_reduce(const float* V, const float* NV, const float* nu, const float* theta,
                                     int64_t theta_rs, int64_t D, int64_t H, int64_t C, int64_t P, int64_t B,
                                     float* f_mean, float* f_var, void* stream) {
  if (!V || !NV || !nu || !theta || !f_mean || !f_var) return VARGP_ERR_ARG;
  if (H * C > 65535) return VARGP_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  if (ceil_div(B, 128) * H * C < 148 * 8) {     // less than one full wave of the streaming kernel: split the rows too
    dim3 grid((unsigned)ceil_div(B, 32), (unsigned)(H * C));
    launch_k(marginal_reduce_split_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, V, NV, nu, theta, theta_rs, D, C, P, B, f_mean, f_var);
    return launch_status();
  }
  dim3 grid((unsigned)ceil_div(B, 128), (unsigned)(H * C));
  launch_k(marginal_reduce_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, V, NV, nu, theta, theta_rs, D, C, P, B, f_mean, f_var);
  return launch_status();
}

extern "C" int vargp_marginal_bwd_prep(const float* V, const float* NV, const float* nu, const float* g_mean,
                                       const float* g_var, const float* theta, int64_t theta_rs, int64_t D,
                                       int64_t H, int64_t C, int64_t P, int64_t B, float* Vbar, float* Vg,
                                       float* theta_bar, void* stream) {
  if (!V || !NV || !nu || !g_mean || !g_var || !theta || !Vbar || !Vg || !theta_bar) return VARGP_ERR_ARG;
  if (Vg == V || Vg == NV || Vbar == V) return VARGP_ERR_ARG;
  if (H * C > 65535 || ceil_div(P, kPrepPch) > 65535) return VARGP_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(V) | reinterpret_cast<uintptr_t>(NV) |
                         reinterpret_cast<uintptr_t>(Vbar) | reinterpret_cast<uintptr_t>(Vg) |
                         reinterpret_cast<uintptr_t>(g_mean) | reinterpret_cast<uintptr_t>(g_var);
  if (B % 4 == 0 && bits % 16 == 0) {
    dim3 grid((unsigned)ceil_div(B, 512), (unsigned)ceil_div(P, kPrepPch), (unsigned)(H * C));
    launch_k((marginal_bwd_prep_kernel<4>), dim3(grid), dim3(128), 0, (cudaStream_t)stream, V, NV, nu, g_mean, g_var, theta, theta_rs, D, C,
                                                                         P, B, Vbar, Vg, theta_bar);
  } else {
    dim3 grid((unsigned)ceil_div(B, 128), (unsigned)ceil_div(P, kPrepPch), (unsigned)(H * C));
    launch_k((marginal_bwd_prep_kernel<1>), dim3(grid), dim3(128), 0, (cudaStream_t)stream, V, NV, nu, g_mean, g_var, theta, theta_rs, D, C,
                                                                         P, B, Vbar, Vg, theta_bar);
  }
  return launch_status();
}

extern "C" int vargp_sym_phi(float* X, int64_t n, int64_t batch, int mirror, void* stream) {
  if (!X || n < 1 || batch < 1) return VARGP_ERR_ARG;
  if (batch > 65535) return VARGP_ERR_UNSUPPORTED;
  if (ceil_div(n, kSymT) > 65535) return VARGP_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(n, kSymT), (unsigned)ceil_div(n, kSymT), (unsigned)batch);
  launch_k(sym_phi_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, X, n, mirror ? 1.f : 0.5f);
  return launch_status();
}

extern "C" int vargp_kl_fwd(const float* W, const float* T, const float* nu, const float* Lu, int64_t H, int64_t C,
                            int64_t P, int64_t M, float* kl, float* work, void* stream) {
  if (!W || !T || !nu || !Lu || !kl || !work || M < 1 || P % M) return VARGP_ERR_ARG;
  if (H * C > 65535) return VARGP_ERR_UNSUPPORTED;
  launch_k(kl_fwd_part_kernel, dim3(dim3(kKlChunks, (unsigned)(H * C))), dim3(256), 0, (cudaStream_t)stream, W, T, nu, Lu, C, P, M, work);
  int rc = launch_status();
  if (rc) return rc;
  launch_k(kl_fwd_sum_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, work, H * C * kKlChunks, H, kl);
  return launch_status();
}

extern "C" int vargp_kl_bwd(const float* W, const float* T, const float* nu, const float* g_kl, int64_t H,
                            int64_t C, int64_t P, int64_t M, float* Wbar, float* Tbar, float* nubar, void* stream) {
  if (!W || !T || !nu || !g_kl || !Wbar || !Tbar || !nubar || M < 1 || P % M) return VARGP_ERR_ARG;
  if (H * C > 65535) return VARGP_ERR_UNSUPPORTED;
  const int64_t want = ceil_div(M * M, 256);
  const unsigned chunks = (unsigned)(want > 256 ? 256 : want);
  launch_k(kl_bwd_kernel, dim3(dim3(chunks, (unsigned)(H * C))), dim3(256), 0, (cudaStream_t)stream, W, T, nu, g_kl, H, C, P, M, Wbar,
                                                                                   Tbar, nubar);
  return launch_status();
}

extern "C" int vargp_kl_bwd_lu(const float* Lu, const float* g_kl, int64_t C, int64_t M, float* Lubar,
                               void* stream) {
  if (!Lu || !g_kl || !Lubar) return VARGP_ERR_ARG;
  launch_k(kl_bwd_lu_kernel, dim3((unsigned)ceil_div(C * M, 128)), dim3(128), 0, (cudaStream_t)stream, Lu, g_kl, C, M, Lubar);
  return launch_status();
}

extern "C" int vargp_tril_unpack(const float* vec, int64_t C, int64_t M, float* out, void* stream) {
  if (!vec || !out || C < 1 || M < 1) return VARGP_ERR_ARG;
  if (C > 65535) return VARGP_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(M, 32), (unsigned)ceil_div(M, 8), (unsigned)C);
  launch_k(tril_unpack_kernel, dim3(grid), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, vec, M, out);
  return launch_status();
}

extern "C" int vargp_tril_unpack_bwd(const float* Lbar, const float* vec, int64_t C, int64_t M, float* vec_bar,
                                     void* stream) {
  if (!Lbar || !vec || !vec_bar || C < 1 || M < 1) return VARGP_ERR_ARG;
  if (C > 65535) return VARGP_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(M, 32), (unsigned)ceil_div(M, 8), (unsigned)C);
  launch_k(tril_unpack_bwd_kernel, dim3(grid), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, Lbar, vec, M, vec_bar);
  return launch_status();
}
