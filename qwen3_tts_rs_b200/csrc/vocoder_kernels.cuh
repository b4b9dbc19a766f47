// F32 kernels of the 12 Hz acoustic-codec decoder (vocoder).  ref: src/models/codec/decoder_12hz.rs:411-699
// and causal_conv.rs / causal_trans_conv.rs / convnext_block.rs / snake_beta.rs / decoder_block.rs.
// Everything is F32 as in the reference (src/lib.rs:344-345).  Activations are channel-major
// [B][C][T] (time contiguous) from end to end, so the reference's transposes disappear and every
// conv / 1x1 projection is one implicit GEMM  Y[co][t] = sum_{ci,j} W[co][ci][j] X[ci][t-(k-1-j)*dil]
// with the SnakeBeta activation fused into the operand load and bias / layer-scale / residual / GELU /
// clamp fused into the epilogue.
#pragma once
#include "common.cuh"

// ---- RVQ lookup: E_first[b][c][t] = first_cb[codes[b][0][t] % size][c];  E_rest = sum_q rest_cb[q][code]
// (sum order q = 0..14 starting from zero, decoder_12hz.rs:438-446).
__global__ void voc_rvq_gather_kernel(const long long* __restrict__ codes, const float* __restrict__ first_cb,
                                      const float* __restrict__ rest_cb, int nq, int cb_size, int vq_dim, int T,
                                      float* __restrict__ e_first, float* __restrict__ e_rest) {
  const int b = blockIdx.y, t = blockIdx.x;
  const long long* cb = codes + (size_t)b * nq * T;
  for (int c = threadIdx.x; c < vq_dim; c += blockDim.x) {
    long long c0 = cb[t] % cb_size;                       // decoder_12hz.rs:423-429
    if (c0 < 0) c0 += cb_size;
    e_first[((size_t)b * vq_dim + c) * T + t] = first_cb[(size_t)c0 * vq_dim + c];
    float acc = 0.f;
    for (int q = 1; q < nq; ++q) {
      long long cq = cb[(size_t)q * T + t];
      cq = cq < 0 ? 0 : (cq >= cb_size ? cb_size - 1 : cq);
      acc = acc + rest_cb[((size_t)(q - 1) * cb_size + cq) * vq_dim + c];
    }
    e_rest[((size_t)b * vq_dim + c) * T + t] = acc;
  }
}

// u32 frame-major codes [B][frames_cap][16] (+ per-row lengths) -> i64 [B][16][T] (codes_to_tensor, lib.rs:1417-1431)
__global__ void voc_codes_to_tensor_kernel(const uint32_t* __restrict__ frames, int frames_cap, int f0, int T,
                                           long long* __restrict__ out) {
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 16 * T; i += gridDim.x * blockDim.x) {
    int q = i / T, f = i - q * T;
    out[(size_t)b * 16 * T + i] = (long long)frames[((size_t)b * frames_cap + f0 + f) * 16 + q];
  }
}

// ---- implicit-GEMM causal conv1d ----------------------------------------------------------------------
enum ConvEpi { CEPI_NONE = 0, CEPI_GELU = 1, CEPI_CLAMP = 2, CEPI_RELU = 3 };   // RELU: speaker encoder TDNN blocks
struct ConvArgs {
  const float* x;       // [B][Cin][T]
  const float* w;       // re-packed at load: [Cin*k][Cout]  (row = ci*k + j)
  const float* bias;    // [Cout] or null
  const float* snake_a; // [Cin] exp(alpha) or null  (prologue: x + sin^2(x*a) * inv_b)
  const float* snake_ib;// [Cin] 1/(exp(beta)+1e-9)
  const float* res;     // [B][Cout][T] or null: y = res + scale*(conv + bias)
  const float* scale;   // [Cout] or null
  float* y;             // [B][Cout][T]
  int B, Cin, Cout, T, k, dil;
  int epi;
};

constexpr int CV_BM = 64, CV_BN = 64, CV_BK = 16;

// sin with one explicit 2*pi range reduction (two FMAs, Cody-Waite split) followed by the SFU sine: absolute
// error ~1e-6 for |x| up to a few thousand, at a fraction of sinf()'s instruction count.  The vocoder spends a
// large share of its non-tensor time here (every conv operand goes through SnakeBeta).
__device__ __forceinline__ float sin_reduced(float x) {
  const float k = rintf(x * 0.15915494309189535f);          // x / (2*pi)
  float r = fmaf(k, -6.28318548202514648f, x);              // 2*pi high part (f32)
  r = fmaf(k, 1.74845553e-7f, r);                           // 2*pi low part: 2*pi = 6.28318548.. - 1.748e-7
  return __sinf(r);
}
__device__ __forceinline__ float snake_f(float x, float a, float ib) {
  float s = sin_reduced(x * a);
  return x + (s * s) * ib;
}

// Final conv of the vocoder: Cout == 1 (96 -> 1, k = 7), SnakeBeta prologue, clamp epilogue.  HBM-bound
// (reads C*T floats once, writes T): each block stages snake(x) for 256 positions + halo, 32 channels at a
// time, and every thread owns one output position.
__global__ void __launch_bounds__(256) voc_conv_cout1_kernel(const float* __restrict__ x, const float* __restrict__ w /*[Cin*k]*/,
                                                             const float* __restrict__ bias, const float* __restrict__ snake_a,
                                                             const float* __restrict__ snake_ib, float* __restrict__ y, int Cin,
                                                             int T, int k, int clamp) {
  extern __shared__ float sm_c1[];
  const int halo = k - 1, W = 256 + halo;
  float* xs = sm_c1;                    // [32][W]
  float* ws = sm_c1 + 32 * W;           // [32][k]
  const int b = blockIdx.y, t0 = blockIdx.x * 256, tid = threadIdx.x;
  const float* xb = x + (size_t)b * Cin * T;
  float acc = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += 32) {
    __syncthreads();
    for (int i = tid; i < 32 * W; i += 256) {
      const int ci = i / W, p = i - ci * W, t = t0 - halo + p;
      float v = 0.f;
      if (c0 + ci < Cin && t >= 0 && t < T) {
        v = xb[(size_t)(c0 + ci) * T + t];
        if (snake_a) v = snake_f(v, snake_a[c0 + ci], snake_ib[c0 + ci]);
      }
      xs[i] = v;
    }
    for (int i = tid; i < 32 * k; i += 256) ws[i] = (c0 + i / k < Cin) ? w[(size_t)c0 * k + i] : 0.f;
    __syncthreads();
    for (int ci = 0; ci < 32; ++ci)
      for (int j = 0; j < k; ++j) acc = fmaf(ws[ci * k + j], xs[ci * W + tid + j], acc);
  }
  const int t = t0 + tid;
  if (t < T) {
    float v = acc + (bias ? bias[0] : 0.f);
    if (clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
    y[(size_t)b * T + t] = v;
  }
}

// 256 threads; thread (ty = tid/16, tx = tid%16) computes rows ty*4..+3 (co) x cols tx, tx+16, tx+32, tx+48 (t).
__global__ void __launch_bounds__(256) voc_conv1d_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) float sm_conv[];
  const int k = a.k, dil = a.dil, halo = (k - 1) * dil;
  const int xw = CV_BN + halo;                         // staged input width
  float* Ws = sm_conv;                                 // [CV_BK*k][CV_BM]
  float* Xs = sm_conv + CV_BK * k * CV_BM;             // [CV_BK][xw]
  const int b = blockIdx.z;
  const int co0 = blockIdx.y * CV_BM, t0 = blockIdx.x * CV_BN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float* xb = a.x + (size_t)b * a.Cin * a.T;
  for (int c0 = 0; c0 < a.Cin; c0 += CV_BK) {
    // stage weights: Ws[(ci*k + j)][co]
    for (int i = tid; i < CV_BK * k * CV_BM; i += 256) {
      int rem = i / CV_BM, co = i - rem * CV_BM;                 // rem = ci*k + j; co fastest (coalesced)
      int ci = rem / k;
      float v = 0.f;
      if (co0 + co < a.Cout && c0 + ci < a.Cin) v = a.w[((size_t)c0 * k + rem) * a.Cout + co0 + co];
      Ws[i] = v;
    }
    // stage inputs with left zero padding (causal) and the SnakeBeta prologue
    for (int i = tid; i < CV_BK * xw; i += 256) {
      int ci = i / xw, p = i - ci * xw;
      int t = t0 - halo + p;
      float v = 0.f;
      if (c0 + ci < a.Cin && t >= 0 && t < a.T) {
        v = xb[(size_t)(c0 + ci) * a.T + t];
        if (a.snake_a) v = snake_f(v, a.snake_a[c0 + ci], a.snake_ib[c0 + ci]);
      }
      Xs[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < CV_BK; ++ci) {
      for (int j = 0; j < k; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(&Ws[(ci * k + j) * CV_BM + ty * 4]);
        const float* xp = &Xs[ci * xw + tx + j * dil];
        const float x0 = xp[0], x1 = xp[16], x2 = xp[32], x3 = xp[48];
        acc[0][0] = fmaf(wv.x, x0, acc[0][0]); acc[0][1] = fmaf(wv.x, x1, acc[0][1]);
        acc[0][2] = fmaf(wv.x, x2, acc[0][2]); acc[0][3] = fmaf(wv.x, x3, acc[0][3]);
        acc[1][0] = fmaf(wv.y, x0, acc[1][0]); acc[1][1] = fmaf(wv.y, x1, acc[1][1]);
        acc[1][2] = fmaf(wv.y, x2, acc[1][2]); acc[1][3] = fmaf(wv.y, x3, acc[1][3]);
        acc[2][0] = fmaf(wv.z, x0, acc[2][0]); acc[2][1] = fmaf(wv.z, x1, acc[2][1]);
        acc[2][2] = fmaf(wv.z, x2, acc[2][2]); acc[2][3] = fmaf(wv.z, x3, acc[2][3]);
        acc[3][0] = fmaf(wv.w, x0, acc[3][0]); acc[3][1] = fmaf(wv.w, x1, acc[3][1]);
        acc[3][2] = fmaf(wv.w, x2, acc[3][2]); acc[3][3] = fmaf(wv.w, x3, acc[3][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= a.Cout) continue;
    const float bv = a.bias ? a.bias[co] : 0.f;
    const float sc = a.scale ? a.scale[co] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx + 16 * j;
      if (t >= a.T) continue;
      const size_t o = ((size_t)b * a.Cout + co) * a.T + t;
      float v = acc[i][j] + bv;
      if (a.epi == CEPI_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));   // erf GELU
      if (a.epi == CEPI_RELU) v = fmaxf(v, 0.f);
      if (a.scale) v = v * sc;
      if (a.res) v = a.res[o] + v;
      if (a.epi == CEPI_CLAMP) v = fminf(fmaxf(v, -1.0f), 1.0f);
      a.y[o] = v;
    }
  }
}

static size_t conv_smem_bytes(int k, int dil) {
  return (size_t)(CV_BK * k * CV_BM + CV_BK * (CV_BN + (k - 1) * dil)) * sizeof(float);
}

// ---- causal transposed conv (kernel k <= 2*stride), right-trimmed to T*stride outputs ------------------
// ref: causal_trans_conv.rs:63-100.  y[co][s*q + r] = bias + sum_ci x[ci][q] w[ci][co][r] + x[ci][q-1] w[ci][co][r+s]
struct TConvArgs {
  const float* x;        // [B][Cin][T]
  const float* w;        // re-packed at load: [Cin*k][Cout]  (row = ci*k + j)
  const float* bias;     // [Cout]
  const float* snake_a;  // [Cin] or null
  const float* snake_ib;
  float* y;              // [B][Cout][T*stride]
  int B, Cin, Cout, T, k, stride;
};
constexpr int TC_BM = 64, TC_BN = 64, TC_BK = 8;

__global__ void __launch_bounds__(256) voc_tconv1d_kernel(const TConvArgs a) {
  extern __shared__ __align__(16) float sm_tconv[];
  const int k = a.k, s = a.stride, To = a.T * s;
  const int b = blockIdx.z, co0 = blockIdx.y * TC_BM, t0 = blockIdx.x * TC_BN;
  const int q0 = t0 / s - 1;                                  // first staged input position (may be -1)
  const int nq = (t0 + TC_BN - 1) / s - q0 + 1;               // staged input positions
  float* Ws = sm_tconv;                                       // [TC_BK][k][TC_BM]
  float* Xs = sm_tconv + TC_BK * k * TC_BM;                   // [TC_BK][nq]  (nq <= TC_BN/s + 3)
  const int xw = TC_BN + 3;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  int qq[4], rr[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int t = t0 + tx + 16 * j;
    qq[j] = t / s - q0;          // index into Xs of x[q]; x[q-1] is qq-1 >= 0
    rr[j] = t - (t / s) * s;
  }
  const float* xb = a.x + (size_t)b * a.Cin * a.T;
  for (int c0 = 0; c0 < a.Cin; c0 += TC_BK) {
    for (int i = tid; i < TC_BK * k * TC_BM; i += 256) {
      int rem = i / TC_BM, co = i - rem * TC_BM;              // rem = ci*k + j; co fastest (coalesced)
      int ci = rem / k;
      float v = 0.f;
      if (c0 + ci < a.Cin && co0 + co < a.Cout) v = a.w[((size_t)c0 * k + rem) * a.Cout + co0 + co];
      Ws[i] = v;
    }
    for (int i = tid; i < TC_BK * nq; i += 256) {
      int ci = i / nq, p = i - ci * nq;
      int q = q0 + p;
      float v = 0.f;
      if (c0 + ci < a.Cin && q >= 0 && q < a.T) {
        v = xb[(size_t)(c0 + ci) * a.T + q];
        if (a.snake_a) v = snake_f(v, a.snake_a[c0 + ci], a.snake_ib[c0 + ci]);
      }
      Xs[ci * xw + p] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < TC_BK; ++ci) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x1 = Xs[ci * xw + qq[j]];
        const float4 w1 = *reinterpret_cast<const float4*>(&Ws[(ci * k + rr[j]) * TC_BM + ty * 4]);
        acc[0][j] = fmaf(w1.x, x1, acc[0][j]); acc[1][j] = fmaf(w1.y, x1, acc[1][j]);
        acc[2][j] = fmaf(w1.z, x1, acc[2][j]); acc[3][j] = fmaf(w1.w, x1, acc[3][j]);
        if (rr[j] + s < k) {
          const float x0 = Xs[ci * xw + qq[j] - 1];
          const float4 w0 = *reinterpret_cast<const float4*>(&Ws[(ci * k + rr[j] + s) * TC_BM + ty * 4]);
          acc[0][j] = fmaf(w0.x, x0, acc[0][j]); acc[1][j] = fmaf(w0.y, x0, acc[1][j]);
          acc[2][j] = fmaf(w0.z, x0, acc[2][j]); acc[3][j] = fmaf(w0.w, x0, acc[3][j]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= a.Cout) continue;
    const float bv = a.bias ? a.bias[co] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx + 16 * j;
      if (t < To) a.y[((size_t)b * a.Cout + co) * To + t] = acc[i][j] + bv;
    }
  }
}
static size_t tconv_smem_bytes(int k) { return (size_t)(TC_BK * k * TC_BM + TC_BK * (TC_BN + 3)) * sizeof(float); }

// ---- depthwise causal conv k (ConvNeXt dwconv, groups == C) ---------------------------------------------
__global__ void voc_dwconv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ y, int C, int T, int k) {
  const int bc = blockIdx.y;                    // b*C + c
  const int c = bc % C;
  const float* xr = x + (size_t)bc * T;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
      int ti = t - (k - 1) + j;
      if (ti >= 0) acc = fmaf(w[c * k + j], xr[ti], acc);
    }
    y[(size_t)bc * T + t] = acc + (bias ? bias[c] : 0.f);
  }
}

// ---- normalisation over channels of a [B][C][T] tensor -----------------------------------------------------
// mode 0: RMSNorm  x / sqrt(mean(x^2) + eps) * w            (decoder_12hz.rs:675-679)
// mode 1: LayerNorm (x - mean) / sqrt(var + eps) * w + b    (convnext_block.rs:119-120, eps 1e-6)
// block (32, 8): threadIdx.x = time within a 32-wide tile (coalesced), threadIdx.y strides channels.
__global__ void __launch_bounds__(256) voc_channel_norm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ y,
                                                               int C, int T, float eps, int mode) {
  __shared__ float red[8][33];
  __shared__ float s_mean[32], s_inv[32];
  const int b = blockIdx.y, t = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
  const float* xb = x + (size_t)b * C * T;
  const bool ok = t < T;
  float s = 0.f;
  if (mode == 1) {
    for (int c = ty; c < C; c += 8) s += ok ? xb[(size_t)c * T + t] : 0.f;
    red[ty][threadIdx.x] = s;
    __syncthreads();
    if (ty == 0) {
      float m = 0.f;
      for (int i = 0; i < 8; ++i) m += red[i][threadIdx.x];
      s_mean[threadIdx.x] = m / (float)C;
    }
    __syncthreads();
  }
  const float mean = mode == 1 ? s_mean[threadIdx.x] : 0.f;
  float v = 0.f;
  for (int c = ty; c < C; c += 8) {
    float d = (ok ? xb[(size_t)c * T + t] : 0.f) - mean;
    v = fmaf(d, d, v);
  }
  __syncthreads();
  red[ty][threadIdx.x] = v;
  __syncthreads();
  if (ty == 0) {
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m += red[i][threadIdx.x];
    s_inv[threadIdx.x] = 1.0f / sqrtf(m / (float)C + eps);
  }
  __syncthreads();
  if (!ok) return;
  const float inv = s_inv[threadIdx.x];
  float* yb = y + (size_t)b * C * T;
  for (int c = ty; c < C; c += 8) {
    float o = (xb[(size_t)c * T + t] - mean) * inv * w[c];
    if (mode == 1 && bias) o += bias[c];
    yb[(size_t)c * T + t] = o;
  }
}

// ---- vocoder transformer attention ------------------------------------------------------------------------
// RoPE (rotate-half within head_dim, decoder_12hz.rs:682-691) + relayout [B][H*D][T] -> [B][H][To][D] at row to0 + t.
// pos0: absolute position of column 0 (0 for a whole utterance; the number of frames already decoded for a streamed chunk,
// whose keys and values are appended to the session's cache: To = cache capacity, to0 = pos0).
__global__ void voc_rope_relayout_kernel(const float* __restrict__ x, float* __restrict__ out, int H, int D, int T,
                                         float theta, int apply_rope, int pos0, int To, int to0) {
  const int b = blockIdx.z, h = blockIdx.y, t = blockIdx.x;
  const float* xb = x + ((size_t)b * H * D + (size_t)h * D) * T;
  float* ob = out + (((size_t)b * H + h) * To + to0 + t) * D;
  const int half = D / 2;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = xb[(size_t)d * T + t];
    if (apply_rope) {
      int i = d % half;
      float inv = 1.0f / powf(theta, (float)(2 * i) / (float)D);
      float ang = (float)(pos0 + t) * inv;
      float cs = cosf(ang), sn = sinf(ang);
      float rot = d < half ? -xb[(size_t)(d + half) * T + t] : xb[(size_t)(d - half) * T + t];
      v = v * cs + rot * sn;
    }
    ob[d] = v;
  }
}

// causal attention, one warp per query; q: [B][H][T][D] (D <= 128), k,v: [B][H][Tk][D] holding positions 0 .. pos0 + T - 1
// (Tk = T, pos0 = 0 for a whole utterance; the session's cache for a streamed chunk); query t sits at position pos0 + t and
// attends to positions <= pos0 + t; out: channel-major [B][H*D][T].  The four queries of a block share every 32-key tile of K
// through shared memory (coalesced loads; a lane then owns one key row, padded against bank conflicts); per (query, key)
// the arithmetic and its order do not depend on how the utterance is cut into chunks, so a streamed chunk reproduces the
// whole-utterance result bit for bit.  Shared memory: voc_attn_smem_floats(pos0 + T, D) floats.
__host__ __device__ inline size_t voc_attn_smem_floats(int L, int D) { return (size_t)4 * L + (size_t)32 * (D + 1) + (size_t)4 * D; }
__global__ void __launch_bounds__(128) voc_attn_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                       const float* __restrict__ v, float* __restrict__ out, int H, int D,
                                                       int T, float scale, int Tk, int pos0) {
  extern __shared__ float sm_vattn[];             // [4 warps][pos0 + T] scores | K tile [32][D + 1] | q [4][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, t = blockIdx.x * 4 + warp;
  const bool active = t < T;
  const int last = pos0 + (active ? t : 0);
  const int block_last = pos0 + min(T - 1, (int)blockIdx.x * 4 + 3);
  float* sc = sm_vattn + (size_t)warp * (pos0 + T);
  float* Ks = sm_vattn + (size_t)4 * (pos0 + T);
  float* qs = Ks + 32 * (D + 1) + warp * D;
  const float* kb = k + ((size_t)b * H + h) * Tk * D;
  const float* vb = v + ((size_t)b * H + h) * Tk * D;
  if (active) {
    const float* qv = q + (((size_t)b * H + h) * T + t) * D;
    for (int e = lane; e < D; e += 32) qs[e] = qv[e];
  }
  float m = -INFINITY;
  for (int j0 = 0; j0 <= block_last; j0 += 32) {
    __syncthreads();                              // the previous tile has been consumed (first pass: q is in place)
    for (int i = threadIdx.x; i < 32 * D; i += 128) {
      const int r = i / D, e = i - r * D;
      Ks[r * (D + 1) + e] = (j0 + r <= block_last) ? kb[(size_t)(j0 + r) * D + e] : 0.f;
    }
    __syncthreads();
    const int j = j0 + lane;
    if (active && j <= last) {
      const float* kr = Ks + lane * (D + 1);
      float d = 0.f;
      for (int e = 0; e < D; ++e) d = fmaf(qs[e], kr[e], d);
      d *= scale;                                 // scale applied after QK^T (decoder_12hz.rs:636-641)
      sc[j] = d;
      m = fmaxf(m, d);
    }
  }
  if (!active) return;
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j <= last; j += 32) {
    float e = expf(sc[j] - m);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum_xor(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  for (int d0 = lane; d0 < D; d0 += 32) {
    float acc = 0.f;
    for (int j = 0; j <= last; ++j) acc = fmaf(sc[j], vb[(size_t)j * D + d0], acc);
    out[((size_t)b * H * D + (size_t)h * D + d0) * T + t] = acc * inv;
  }
}

__global__ void voc_silu_mul_kernel(const float* __restrict__ g, const float* __restrict__ u, float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a = g[i];
    y[i] = (a / (1.0f + expf(-a))) * u[i];
  }
}

__global__ void voc_prep_snake_kernel(const float* __restrict__ alpha, const float* __restrict__ beta, float* __restrict__ ea,
                                      float* __restrict__ inv_b, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    ea[i] = expf(alpha[i]);
    inv_b[i] = 1.0f / (expf(beta[i]) + 1e-9f);       // snake_beta.rs:72-75: recip(beta + eps)
  }
}

__global__ void voc_prep_codebook_kernel(const float* __restrict__ emb_sum, const float* __restrict__ usage,
                                         float* __restrict__ out, int rows, int dim) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)rows * dim) {
    float u = fmaxf(usage[i / dim], 1e-7f);           // decoder_12hz.rs:199-225
    out[i] = emb_sum[i] / u;
  }
}
