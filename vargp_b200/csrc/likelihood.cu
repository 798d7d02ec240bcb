// Monte-Carlo softmax likelihood: forward NLL reduction fused with its adjoint, and the predict epilogue.
// Streaming kernels: the (H, F, C, B) noise tensor is read exactly once, coalesced along B.
// Algorithmic bytes (nll): 4*(H*F*C*B + 4*H*C*B) + 8*B.                    (likelihoods.py:13-63)
#include "common.cuh"

namespace vargp {

template <int CMAX>
__global__ void __launch_bounds__(128)
softmax_nll_kernel(const float* __restrict__ f_mean, const float* __restrict__ f_var,
                   const float* __restrict__ eps, const int64_t* __restrict__ y,
                   int64_t H, int64_t F, int64_t C, int64_t B,
                   float* __restrict__ nll, float* __restrict__ g_mean, float* __restrict__ g_var, float gscale,
                   float* __restrict__ work) {
  pdl_enter();
  __shared__ float scratch[32];
  __shared__ bool s_last;
  const int64_t h = blockIdx.y;
  const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const bool live = b < B;
  float acc = 0.f;
  if (live) {
    float mu[CMAX], sd[CMAX], gm[CMAX], gs[CMAX];
    const int yb = (int)y[b];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      gm[c] = 0.f; gs[c] = 0.f; mu[c] = 0.f; sd[c] = 0.f;
      if (c < C) {
        mu[c] = f_mean[(h * C + c) * B + b];
        sd[c] = sqrtf(f_var[(h * C + c) * B + b]);
      }
    }
    for (int64_t f = 0; f < F; ++f) {
      float e[CMAX], v[CMAX];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          e[c] = eps[((h * F + f) * C + c) * B + b];
          v[c] = fmaf(sd[c], e[c], mu[c]);
          mx = fmaxf(mx, v[c]);
        }
      }
      float sum = 0.f, vy = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          if (c == yb) vy = v[c];
          v[c] = expf(v[c] - mx);
          sum += v[c];
        }
      }
      acc -= (vy - mx) - logf(sum);
      const float inv = 1.f / sum;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          const float gsm = v[c] * inv - (c == yb ? 1.f : 0.f);
          gm[c] += gsm;
          gs[c] = fmaf(gsm, e[c], gs[c]);
        }
      }
    }
    const float scale = 1.f / (float)(H * F);
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        g_mean[(h * C + c) * B + b] = gm[c] * (scale * gscale);
        g_var[(h * C + c) * B + b] = gs[c] * (scale * gscale) / (2.f * sd[c]);
      }
    }
    acc *= scale;
    if ((unsigned)yb >= (unsigned)C) acc = __int_as_float(0x7fc00000);   // label outside [0, C): F.nll_loss raises; here the loss is NaN
  }
  // deterministic two-stage sum (no float atomics): every CTA leaves its partial in work[1 + cta]; the CTA that draws
  // the last ticket adds them up in a fixed order.  work[0] is the ticket counter: zero at launch, zero again at exit.
  acc = block_sum(acc, scratch);
  const unsigned nparts = gridDim.x * gridDim.y;
  unsigned* ticket = reinterpret_cast<unsigned*>(work);
  float* part = work + 1;
  if (threadIdx.x == 0) {
    part[blockIdx.y * gridDim.x + blockIdx.x] = acc;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == nparts - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float t = 0.f;
    for (unsigned i = threadIdx.x; i < nparts; i += 128) t += __ldcg(part + i);
    t = block_sum(t, scratch);
    if (threadIdx.x == 0) {
      nll[0] += t;
      *ticket = 0u;
    }
  }
}

template <int CMAX>
__global__ void __launch_bounds__(128)
softmax_predict_kernel(const float* __restrict__ f_mean, const float* __restrict__ f_var,
                       const float* __restrict__ eps, int64_t H, int64_t F, int64_t C, int64_t B,
                       float* __restrict__ probs) {
  pdl_enter();
  const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (b >= B) return;
  float p[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) p[c] = 0.f;
  for (int64_t h = 0; h < H; ++h) {
    float mu[CMAX], sd[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      mu[c] = 0.f; sd[c] = 0.f;
      if (c < C) {
        mu[c] = f_mean[(h * C + c) * B + b];
        sd[c] = sqrtf(f_var[(h * C + c) * B + b]);
      }
    }
    for (int64_t f = 0; f < F; ++f) {
      float v[CMAX];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          v[c] = fmaf(sd[c], eps[((h * F + f) * C + c) * B + b], mu[c]);
          mx = fmaxf(mx, v[c]);
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) { v[c] = expf(v[c] - mx); sum += v[c]; }
      }
      const float inv = 1.f / sum;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) p[c] = fmaf(v[c], inv, p[c]);
    }
  }
  const float scale = 1.f / (float)(H * F);
#pragma unroll
  for (int c = 0; c < CMAX; ++c)
    if (c < C) probs[b * C + c] = p[c] * scale;
}

}  // namespace vargp

using namespace vargp;

extern "C" int64_t vargp_softmax_nll_work(int64_t H, int64_t B) { return 1 + H * ceil_div(B, 128); }

extern "C" int vargp_softmax_nll(const float* f_mean, const float* f_var, const float* eps, const int64_t* y,
                                 int64_t H, int64_t F, int64_t C, int64_t B, float* nll, float* g_mean,
                                 float* g_var, float gscale, float* work, void* stream) {
  if (!f_mean || !f_var || !eps || !y || !nll || !g_mean || !g_var || !work) return VARGP_ERR_ARG;
  if (H < 1 || F < 1 || C < 1 || B < 0) return VARGP_ERR_ARG;
  if (C > 32 || H > 65535) return VARGP_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  dim3 grid((unsigned)ceil_div(B, 128), (unsigned)H);
  cudaStream_t s = (cudaStream_t)stream;
  if (C <= 4) launch_k((softmax_nll_kernel<4>), dim3(grid), dim3(128), 0, s, f_mean, f_var, eps, y, H, F, C, B, nll, g_mean, g_var, gscale, work);
  else if (C <= 10) launch_k((softmax_nll_kernel<10>), dim3(grid), dim3(128), 0, s, f_mean, f_var, eps, y, H, F, C, B, nll, g_mean, g_var, gscale, work);
  else if (C <= 16) launch_k((softmax_nll_kernel<16>), dim3(grid), dim3(128), 0, s, f_mean, f_var, eps, y, H, F, C, B, nll, g_mean, g_var, gscale, work);
  else launch_k((softmax_nll_kernel<32>), dim3(grid), dim3(128), 0, s, f_mean, f_var, eps, y, H, F, C, B, nll, g_mean, g_var, gscale, work);
  return launch_status();
}

extern "C" int vargp_softmax_predict(const float* f_mean, const float* f_var, const float* eps, int64_t H,
                                     int64_t F, int64_t C, int64_t B, float* probs, void* stream) {
  if (!f_mean || !f_var || !eps || !probs) return VARGP_ERR_ARG;
  if (H < 1 || F < 1 || C < 1 || B < 0) return VARGP_ERR_ARG;
  if (C > 32) return VARGP_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  dim3 grid((unsigned)ceil_div(B, 128));
  cudaStream_t s = (cudaStream_t)stream;
  if (C <= 4) launch_k((softmax_predict_kernel<4>), dim3(grid), dim3(128), 0, s, f_mean, f_var, eps, H, F, C, B, probs);
  else if (C <= 10) launch_k((softmax_predict_kernel<10>), dim3(grid), dim3(128), 0, s, f_mean, f_var, eps, H, F, C, B, probs);
  else if (C <= 16) launch_k((softmax_predict_kernel<16>), dim3(grid), dim3(128), 0, s, f_mean, f_var, eps, H, F, C, B, probs);
  else launch_k((softmax_predict_kernel<32>), dim3(grid), dim3(128), 0, s, f_mean, f_var, eps, H, F, C, B, probs);
  return launch_status();
}
