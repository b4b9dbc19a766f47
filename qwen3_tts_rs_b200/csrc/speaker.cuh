// ECAPA-TDNN speaker encoder (voice-clone front end): the small kernels around the 1x1 convolutions,
// which run on the vocoder's tensor-core conv kernels (a k = 1 causal conv is a "same" conv).  F32 throughout, one utterance
// at a time ([C][T] tensors), as SpeakerEncoder::forward (src/models/speaker.rs:448-476) is called.  One-shot per voice: these
// are plain coalesced kernels, not a hot path.
#pragma once
#include "common.cuh"

// reflect index of speaker.rs:26-53 (PyTorch padding_mode = "reflect"): position p of the padded signal, p in [-left, T + right)
__device__ __forceinline__ int spk_reflect(int p, int T) {
  if (p < 0) p = -p;
  if (p >= T) p = 2 * (T - 1) - p;
  return p;
}

// ReflectPadConv1d + bias + ReLU (TimeDelayNetBlock, speaker.rs:68-139) for the convs with k > 1: the initial TDNN and the
// Res2Net branches.  y[co][t] = relu(b[co] + sum_{ci,j} w[co][ci][j] * in[ci][reflect(t - left + j * dil)]),
// in = xa (+ xb when given: the Res2Net cascade adds the previous branch's output, speaker.rs:176-181).
// grid (ceil(T / 128), Cout); block 128: thread = one output position; weights of the output channel staged in shared memory.
__global__ void __launch_bounds__(128) spk_reflect_conv_relu_kernel(const float* __restrict__ xa, const float* __restrict__ xb,
                                                                     const float* __restrict__ w, const float* __restrict__ bias,
                                                                     float* __restrict__ y, int Cin, int k, int dil, int T) {
  extern __shared__ float spk_w[];                  // [Cin * k]
  const int co = blockIdx.y;
  for (int i = threadIdx.x; i < Cin * k; i += blockDim.x) spk_w[i] = w[(size_t)co * Cin * k + i];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int left = (dil * (k - 1)) / 2;
  float acc = bias[co];
  for (int ci = 0; ci < Cin; ++ci) {
    const float* ra = xa + (size_t)ci * T;
    const float* rb = xb ? xb + (size_t)ci * T : nullptr;
    for (int j = 0; j < k; ++j) {
      const int p = spk_reflect(t - left + j * dil, T);
      float v = ra[p];
      if (rb) v += rb[p];
      acc = fmaf(spk_w[ci * k + j], v, acc);
    }
  }
  y[(size_t)co * T + t] = fmaxf(acc, 0.f);
}

// mean over T of every channel (SE squeeze, speaker.rs:216); with `std_out`: also sqrt(mean((x - mean)^2) + 1e-5) (ASP global
// statistics, speaker.rs:299-303).  One block per channel, fixed-order tree reduction.
__global__ void __launch_bounds__(256) spk_channel_stats_kernel(const float* __restrict__ x, float* __restrict__ mean_out,
                                                                 float* __restrict__ std_out, int T) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  const float* row = x + (size_t)c * T;
  float s = 0.f;
  for (int t = threadIdx.x; t < T; t += 256) s += row[t];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  const float mean = red[0] / (float)T;
  __syncthreads();
  if (std_out != nullptr) {
    float v = 0.f;
    for (int t = threadIdx.x; t < T; t += 256) { const float d = row[t] - mean; v = fmaf(d, d, v); }
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) std_out[c] = sqrtf(red[0] / (float)T + 1e-5f);
  }
  if (threadIdx.x == 0) mean_out[c] = mean;
}

// SE excitation + residual (speaker.rs:217-221, 266): y = x * sigmoid(s[c]) + res, sigmoid = 1 / (exp(-s) + 1)
__global__ void spk_se_apply_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ res,
                                    float* __restrict__ y, int C, int T) {
  const size_t n = (size_t)C * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = 1.0f / (expf(-s[i / T]) + 1.0f);
    y[i] = x[i] * g + res[i];
  }
}

// ASP attention input (speaker.rs:305-309): rows [C, 2C) = mean[c], rows [2C, 3C) = std[c], broadcast over T (rows [0, C) are x,
// copied by the caller)
__global__ void spk_broadcast_stats_kernel(const float* __restrict__ mean, const float* __restrict__ stdv, float* __restrict__ out,
                                           int C, int T) {
  const size_t n = (size_t)2 * C * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t c = i / T;
    out[i] = c < (size_t)C ? mean[c] : stdv[c - C];
  }
}

__global__ void spk_tanh_kernel(float* __restrict__ x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = tanhf(x[i]);
}

// ASP pooling (speaker.rs:314-346): softmax over T of the attention logits of channel c, weighted mean and weighted std of x.
// out[c] = w_mean, out[C + c] = sqrt(sum((x - w_mean)^2 * a) + 1e-5).  One block per channel.
__global__ void __launch_bounds__(256) spk_asp_pool_kernel(const float* __restrict__ x, const float* __restrict__ logit,
                                                            float* __restrict__ out, int C, int T) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  const float* xr = x + (size_t)c * T;
  const float* lr = logit + (size_t)c * T;
  auto reduce = [&](float v, bool is_max) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] = is_max ? fmaxf(red[threadIdx.x], red[threadIdx.x + o]) : red[threadIdx.x] + red[threadIdx.x + o];
      __syncthreads();
    }
    const float r = red[0];
    __syncthreads();
    return r;
  };
  float m = -INFINITY;
  for (int t = threadIdx.x; t < T; t += 256) m = fmaxf(m, lr[t]);
  m = reduce(m, true);
  float s = 0.f;
  for (int t = threadIdx.x; t < T; t += 256) s += expf(lr[t] - m);
  s = reduce(s, false);
  const float inv = 1.0f / s;
  float wm = 0.f;
  for (int t = threadIdx.x; t < T; t += 256) wm = fmaf(xr[t], expf(lr[t] - m) * inv, wm);
  wm = reduce(wm, false);
  float wv = 0.f;
  for (int t = threadIdx.x; t < T; t += 256) { const float d = xr[t] - wm; wv = fmaf(d * d, expf(lr[t] - m) * inv, wv); }
  wv = reduce(wv, false);
  if (threadIdx.x == 0) { out[c] = wm; out[C + c] = sqrtf(wv + 1e-5f); }
}
