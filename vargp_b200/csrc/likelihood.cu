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

// Same contract for SMALL minibatches (Split / Permuted-MNIST steps: B = 512).  One thread per column walking its F likelihood
// samples is a chain of F dependent L2 round trips on H * B / 128 = 12 CTAs (16 us in the step, on the critical path); here a
// thread owns ONE (column, sample) pair -- block (32 columns, F samples) -- the F-fold sums of the adjoints go through shared
// memory in a fixed order, the NLL through the same deterministic two-stage sum.
constexpr int kNllSmallFC = 160;          // F * C limit: 2 x 160 x 32 floats of shared memory
template <int CMAX>
__global__ void __launch_bounds__(512)
softmax_nll_small_kernel(const float* __restrict__ f_mean, const float* __restrict__ f_var,
                         const float* __restrict__ eps, const int64_t* __restrict__ y,
                         int64_t H, int64_t F, int64_t C, int64_t B,
                         float* __restrict__ nll, float* __restrict__ g_mean, float* __restrict__ g_var, float gscale,
                         float* __restrict__ work) {
  pdl_enter();
  __shared__ float sg[2][kNllSmallFC][32];
  __shared__ float scratch[32];
  __shared__ bool s_last;
  const int bl = threadIdx.x, f = threadIdx.y;
  const int tid = f * 32 + bl, nthr = 32 * (int)F;
  const int64_t h = blockIdx.y;
  const int64_t b = (int64_t)blockIdx.x * 32 + bl;
  const bool live = b < B;
  const float scale = 1.f / (float)(H * F);
  float acc = 0.f;
  if (live) {
    float e[CMAX], v[CMAX];
    const int yb = (int)y[b];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        e[c] = eps[((h * F + f) * C + c) * B + b];
        v[c] = fmaf(sqrtf(f_var[(h * C + c) * B + b]), e[c], f_mean[(h * C + c) * B + b]);
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
    acc = -((vy - mx) - logf(sum)) * scale;
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        const float gsm = v[c] * inv - (c == yb ? 1.f : 0.f);
        sg[0][f * C + c][bl] = gsm;
        sg[1][f * C + c][bl] = gsm * e[c];
      }
    }
    if ((unsigned)yb >= (unsigned)C) acc = __int_as_float(0x7fc00000);   // label outside [0, C): F.nll_loss raises; here the loss is NaN
  }
  __syncthreads();
  for (int item = tid; item < (int)C * 32; item += nthr) {          // adjoints: sum over the samples in a fixed order
    const int c = item >> 5, bb = item & 31;
    const int64_t b2 = (int64_t)blockIdx.x * 32 + bb;
    if (b2 < B) {
      float gm = 0.f, gs = 0.f;
      for (int ff = 0; ff < (int)F; ++ff) {
        gm += sg[0][ff * C + c][bb];
        gs += sg[1][ff * C + c][bb];
      }
      const float sd = sqrtf(f_var[(h * C + c) * B + b2]);
      g_mean[(h * C + c) * B + b2] = gm * (scale * gscale);
      g_var[(h * C + c) * B + b2] = gs * (scale * gscale) / (2.f * sd);
    }
  }
  // deterministic two-stage sum of the NLL, as in softmax_nll_kernel
  acc = warp_sum(acc);
  if (bl == 0) scratch[f] = acc;
  __syncthreads();
  const unsigned nparts = gridDim.x * gridDim.y;
  unsigned* ticket = reinterpret_cast<unsigned*>(work);
  float* part = work + 1;
  if (tid == 0) {
    float t = 0.f;
    for (int ff = 0; ff < (int)F; ++ff) t += scratch[ff];
    part[blockIdx.y * gridDim.x + blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == nparts - 1;
  }
  __syncthreads();
  if (s_last && f == 0) {                                            // one warp adds the partials in a fixed order
    __threadfence();
    float t = 0.f;
    for (unsigned i = bl; i < nparts; i += 32) t += __ldcg(part + i);
    t = warp_sum(t);
    if (bl == 0) {
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

extern "C" int64_t vargp_softmax_nll_work(int64_t H, int64_t B) { return 1 + H * ceil_div(B, 32); }

extern "C" int vargp_softmax_nll(const float* f_mean, const float* f_var, const float* eps, const int64_t* y,
                                 int64_t H, int64_t F, int64_t C, int64_t B, float* nll, float* g_mean,
                                 float* g_var, float gscale, float* work, void* stream) {
  if (!f_mean || !f_var || !eps || !y || !nll || !g_mean || !g_var || !work) return VARGP_ERR_ARG;
  if (H < 1 || F < 1 || C < 1 || B < 0) return VARGP_ERR_ARG;
  if (C > 32 || H > 65535) return VARGP_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (ceil_div(B, 128) * H < 148 && F <= 16 && F * C <= kNllSmallFC && C <= 16) {     // less than a wave of the streaming kernel
    dim3 sgrid((unsigned)ceil_div(B, 32), (unsigned)H), sblock(32, (unsigned)F);
    if (C <= 4) launch_k((softmax_nll_small_kernel<4>), dim3(sgrid), dim3(sblock), 0, s, f_mean, f_var, eps, y, H, F, C, B, nll, g_mean, g_var, gscale, work);
    else if (C <= 10) launch_k((softmax_nll_small_kernel<10>), dim3(sgrid), dim3(sblock), 0, s, f_mean, f_var, eps, y, H, F, C, B, nll, g_mean, g_var, gscale, work);
    else launch_k((softmax_nll_small_kernel<16>), dim3(sgrid), dim3(sblock), 0, s, f_mean, f_var, eps, y, H, F, C, B, nll, g_mean, g_var, gscale, work);
    return launch_status();
  }
  dim3 grid((unsigned)ceil_div(B, 128), (unsigned)H);
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
