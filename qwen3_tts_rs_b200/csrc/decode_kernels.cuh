// Non-GEMM kernels of the decode step: fused residual+RMSNorm, per-head QK-norm + RoPE + KV append,
// decode attention over the in-place KV cache, embedding gathers / frame bookkeeping, and the
// on-device sampler.  All are HBM/latency-bound elementwise or reduction kernels; they use 128-bit
// accesses where the layout allows and keep every per-frame scalar (positions, frame index, tokens,
// RNG state, EOS flags) in device memory so that one frame can be replayed as a CUDA graph with no
// host round trip (the reference syncs 4 bytes to the host every frame, src/lib.rs:648-649).
#pragma once
#include "common.cuh"
#include "gemv.cuh"
#include "norm.cuh"

// =================================================================================================
// Fused residual-add + RMSNorm.  ref: kernels/fused_residual_rmsnorm.cu:38-90 and
// FusedRmsNorm::forward_residual (src/models/fused_ops.rs:49-96).
//   sum = x + r (stored, rounded to T);  sumsq over the UN-rounded f32 sums;  pass 2 re-reads the
//   ROUNDED sum:  normed = (scale * sum_rounded) * w.
// One block per row; 128 threads x 128-bit accesses for ncols >= 1024, one warp for ncols < 1024.
template <typename T> struct Vec8;
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float f[8]) { unpack8(*reinterpret_cast<const uint4*>(p), f); }
  static __device__ __forceinline__ void store(bf16* p, const float f[8]) { *reinterpret_cast<uint4*>(p) = pack8(f); }
  static __device__ __forceinline__ float round(float v) { return rbf(v); }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float f[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float f[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
  static __device__ __forceinline__ float round(float v) { return v; }
};
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return bf2f(v); }
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return f2bf(v); }
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }

template <typename T>
__global__ void __launch_bounds__(128) fused_residual_rmsnorm_large(const T* __restrict__ x, const T* __restrict__ r,
                                                                    const T* __restrict__ w, T* __restrict__ out_normed,
                                                                    T* __restrict__ out_sum, int ncols, float eps) {
  __shared__ float s_part[32];
  const size_t off = (size_t)blockIdx.x * ncols;
  const int g = threadIdx.x;
  float p[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) p[e] = 0.f;
  for (int c = 8 * g; c < ncols; c += 1024) {
    float a[8], b[8], s[8];
    Vec8<T>::load(x + off + c, a);
    Vec8<T>::load(r + off + c, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] = a[e] + b[e];
      p[e] = fmaf(s[e], s[e], p[e]);
    }
    Vec8<T>::store(out_sum + off + c, s);
  }
  const float tot = sumsq_ref_large_finish(p, g, s_part, 1);
  const float sc = ref_mean_rsqrt(tot, ncols, eps);
  for (int c = 8 * g; c < ncols; c += 1024) {
    float s[8], ww[8], o[8];
    Vec8<T>::load(out_sum + off + c, s);       // the rounded sum this thread stored itself
    Vec8<T>::load(w + c, ww);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (sc * s[e]) * ww[e];
    Vec8<T>::store(out_normed + off + c, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(32) fused_residual_rmsnorm_small(const T* __restrict__ x, const T* __restrict__ r,
                                                                   const T* __restrict__ w, T* __restrict__ out_normed,
                                                                   T* __restrict__ out_sum, int ncols, float eps) {
  const size_t off = (size_t)blockIdx.x * ncols;
  const int lane = threadIdx.x;
  float tmp = 0.f;
  for (int c = lane; c < ncols; c += 32) {
    float s = to_f<T>(x[off + c]) + to_f<T>(r[off + c]);
    out_sum[off + c] = from_f<T>(s);
    tmp = fmaf(s, s, tmp);
  }
  tmp = warp_sum_xor(tmp);
  const float sc = ref_mean_rsqrt(tmp, ncols, eps);
  for (int c = lane; c < ncols; c += 32)
    out_normed[off + c] = from_f<T>((sc * to_f<T>(out_sum[off + c])) * to_f<T>(w[c]));
}

template <typename T>
static void fused_residual_rmsnorm_launch(const T* x, const T* r, const T* w, T* out_normed, T* out_sum, int rows,
                                          int cols, float eps, cudaStream_t st) {
  if (rows <= 0) return;
  if (cols >= 1024 && cols % 8 == 0)
    fused_residual_rmsnorm_large<T><<<rows, 128, 0, st>>>(x, r, w, out_normed, out_sum, cols, eps);
  else if (cols < 1024)
    fused_residual_rmsnorm_small<T><<<rows, 32, 0, st>>>(x, r, w, out_normed, out_sum, cols, eps);
  else
    throw Q3Error(Q3_ERR_INVALID, "fused_residual_rmsnorm: cols >= 1024 must be a multiple of 8");
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
}

// =================================================================================================
// Per-head QK RMSNorm + RoPE + KV append.  ref: transformer.rs:262-284 (reshape, q_norm/k_norm per
// head over D = 128, transpose, rope.apply, cache.update) and kv_cache.rs:290-310.
// One warp per (token, head).  head_dim == 128: lane owns dims lane, lane+32 (first half) and
// lane+64, lane+96 (second half) -- which is also the reference-order partial-sum ownership for
// block_size 32, and puts both members of every RoPE pair (d, d+64) in the same lane.
struct RopeArgs {
  const bf16* qkv;        // [T][(heads + 2*kv_heads) * 128]   raw projections
  bf16* q_out;            // [T][heads*128]
  bf16* k_cache;          // [B][kv_heads][max_seq][128]  (this layer)
  bf16* v_cache;
  const bf16* q_norm_w;   // [128]
  const bf16* k_norm_w;
  const bf16* cos_tab;    // [n_pos][64]  bf16(cos(pos * inv_freq))
  const bf16* sin_tab;
  const int* pos_base;    // [B] or null
  int pos_add, S, T, heads, kv_heads, max_seq;
  float eps;
};

__global__ void __launch_bounds__(128) qk_norm_rope_append_kernel(const RopeArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hh = blockIdx.x * 4 + warp;
  const int t = blockIdx.y;
  const int nh = a.heads + 2 * a.kv_heads;
  if (hh >= nh) return;
  const int b = t / a.S, s = t - b * a.S;
  const int pos = (a.pos_base ? a.pos_base[b] : 0) + a.pos_add + s;
  const bf16* src = a.qkv + (size_t)t * nh * 128 + (size_t)hh * 128;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = bf2f(src[lane + 32 * i]);
  if (hh >= a.heads + a.kv_heads) {   // V: plain append
    bf16* dst = a.v_cache + (((size_t)b * a.kv_heads + (hh - a.heads - a.kv_heads)) * a.max_seq + pos) * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[lane + 32 * i] = f2bf(v[i]);
    return;
  }
  const bool is_q = hh < a.heads;
  const bf16* nw = is_q ? a.q_norm_w : a.k_norm_w;
  float tmp = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) tmp = fmaf(v[i], v[i], tmp);
  tmp = warp_sum_xor(tmp);
  const float sc = ref_mean_rsqrt(tmp, 128, a.eps);
  float n[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) n[i] = rbf((sc * v[i]) * bf2f(nw[lane + 32 * i]));
  // RoPE (transformer.rs:42-69): pairs (d, d+64); cos/sin already rounded to bf16
  float o[4];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int d = lane + 32 * i;
    const float c = bf2f(a.cos_tab[(size_t)pos * 64 + d]), sn = bf2f(a.sin_tab[(size_t)pos * 64 + d]);
    const float x1 = n[i], x2 = n[i + 2];
    o[i] = rbf(rbf(x1 * c) - rbf(x2 * sn));
    o[i + 2] = rbf(rbf(x2 * c) + rbf(x1 * sn));
  }
  bf16* dst = is_q ? a.q_out + (size_t)t * a.heads * 128 + (size_t)hh * 128
                   : a.k_cache + (((size_t)b * a.kv_heads + (hh - a.heads)) * a.max_seq + pos) * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[lane + 32 * i] = f2bf(o[i]);
}

// =================================================================================================
// Decode attention over the KV cache (GQA, 2 query heads per KV head, D = 128).
// ref: Attention::forward matmul path, transformer.rs:347-369:
//   aw = bf16(q k^T); aw = bf16(aw * bf16(1/sqrt(D))); softmax in f32 -> bf16; out = bf16(aw v).
// One block per (kv head, token); the token attends cache positions [0, pos].
struct AttnArgs {
  const bf16* q;          // [T][heads*128]
  const bf16* k_cache;    // [B][kv_heads][max_seq][128]
  const bf16* v_cache;
  bf16* out;              // [T][heads*128]
  const int* pos_base;
  int pos_add, S, T, heads, kv_heads, max_seq;
};

__global__ void __launch_bounds__(256) attn_decode_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) float sm_attn[];
  const int kvh = blockIdx.x, t = blockIdx.y;
  const int b = t / a.S, s = t - b * a.S;
  const int pos = (a.pos_base ? a.pos_base[b] : 0) + a.pos_add + s;
  const int L = pos + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc0 = sm_attn;                 // [max_seq] scores / probabilities, head 0
  float* sc1 = sm_attn + a.max_seq;     // head 1
  float* red = sm_attn + 2 * a.max_seq; // [8][2][128] cross-warp partial outputs; also reductions
  const bf16* kbase = a.k_cache + ((size_t)b * a.kv_heads + kvh) * a.max_seq * 128;
  const bf16* vbase = a.v_cache + ((size_t)b * a.kv_heads + kvh) * a.max_seq * 128;
  const bf16* q0p = a.q + (size_t)t * a.heads * 128 + (size_t)(2 * kvh) * 128;
  float q0[4], q1[4];
  {
    uint2 u0 = *reinterpret_cast<const uint2*>(q0p + 4 * lane);
    uint2 u1 = *reinterpret_cast<const uint2*>(q0p + 128 + 4 * lane);
    q0[0] = bf_lo(u0.x); q0[1] = bf_hi(u0.x); q0[2] = bf_lo(u0.y); q0[3] = bf_hi(u0.y);
    q1[0] = bf_lo(u1.x); q1[1] = bf_hi(u1.x); q1[2] = bf_lo(u1.y); q1[3] = bf_hi(u1.y);
  }
  const float scale = rbf(0.08838834764831845f);   // 1/sqrt(128), rounded as candle's bf16 affine does
  // pass 1: scores
  for (int j = warp; j < L; j += 8) {
    uint2 u = *reinterpret_cast<const uint2*>(kbase + (size_t)j * 128 + 4 * lane);
    float k0 = bf_lo(u.x), k1 = bf_hi(u.x), k2 = bf_lo(u.y), k3 = bf_hi(u.y);
    float d0 = q0[0] * k0 + q0[1] * k1 + q0[2] * k2 + q0[3] * k3;
    float d1 = q1[0] * k0 + q1[1] * k1 + q1[2] * k2 + q1[3] * k3;
    d0 = warp_sum_xor(d0);
    d1 = warp_sum_xor(d1);
    if (lane == 0) {
      sc0[j] = rbf(rbf(d0) * scale);
      sc1[j] = rbf(rbf(d1) * scale);
    }
  }
  __syncthreads();
  // softmax (f32 internals, probabilities rounded to bf16): warps 0 and 1 take one head each
  if (warp < 2) {
    float* sc = warp == 0 ? sc0 : sc1;
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) m = fmaxf(m, sc[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      float e = expf(sc[j] - m);
      sc[j] = e;
      sum += e;
    }
    sum = warp_sum_xor(sum);
    for (int j = lane; j < L; j += 32) sc[j] = rbf(sc[j] / sum);
  }
  __syncthreads();
  // pass 2: out = P V ; warp w takes keys j = w (mod 8); lane owns dims 4*lane..4*lane+3
  float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = warp; j < L; j += 8) {
    uint2 u = *reinterpret_cast<const uint2*>(vbase + (size_t)j * 128 + 4 * lane);
    float v0 = bf_lo(u.x), v1 = bf_hi(u.x), v2 = bf_lo(u.y), v3 = bf_hi(u.y);
    const float p0 = sc0[j], p1 = sc1[j];
    o0[0] = fmaf(p0, v0, o0[0]); o0[1] = fmaf(p0, v1, o0[1]); o0[2] = fmaf(p0, v2, o0[2]); o0[3] = fmaf(p0, v3, o0[3]);
    o1[0] = fmaf(p1, v0, o1[0]); o1[1] = fmaf(p1, v1, o1[1]); o1[2] = fmaf(p1, v2, o1[2]); o1[3] = fmaf(p1, v3, o1[3]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[(warp * 2 + 0) * 128 + 4 * lane + i] = o0[i];
    red[(warp * 2 + 1) * 128 + 4 * lane + i] = o1[i];
  }
  __syncthreads();
  {
    const int h = threadIdx.x >> 7, d = threadIdx.x & 127;   // 256 threads = 2 heads x 128 dims
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += red[(w * 2 + h) * 128 + d];
    a.out[(size_t)t * a.heads * 128 + (size_t)(2 * kvh + h) * 128 + d] = f2bf(acc);
  }
}

static size_t attn_smem_bytes(int max_seq) { return (size_t)(2 * max_seq + 8 * 2 * 128) * sizeof(float); }

// =================================================================================================
// Frame bookkeeping / embedding kernels (ref: src/lib.rs:588-622, code_predictor.rs:330-345,
// 386-413, 497-519).
struct FrameState {
  // per row, device resident
  uint32_t* cur_tok;      // [B] current semantic token
  int* done;              // [B] 1 once EOS was sampled
  int* n_frames;          // [B] frames emitted
  int* token_count;       // [B] tokens sampled so far
  int* offset;            // [B] talker KV length == next position
  int* frame_idx;         // [B]
  unsigned long long* rng;   // [B] PCG state
  uint8_t* seen;          // [B][V] repetition-penalty mask
  bf16* last_hidden;      // [B][H]
  const bf16* trailing;   // [B][lt_max][H]
  const int* lt;          // [B]
  int lt_max;
  const bf16* tts_pad;    // [H]
  uint32_t* codes;        // [B][frames_cap][16]
  int frames_cap;
  unsigned long long* amax;  // [15][B] arg-max keys of the code predictor passes
  uint32_t* frame_codes;  // [B][16] codes of the frame being built
  int* host_flags;        // mapped pinned: [0] = number of rows not done
};

// CP pass 0 input: X[2b] = last_hidden[b], X[2b+1] = codec_embedding[cur_tok[b]]; clears arg-max keys.
__global__ void cp_begin_kernel(FrameState st, const bf16* __restrict__ codec_emb, bf16* __restrict__ X, int H, int B,
                                int n_ac) {
  const int b = blockIdx.x;
  const uint32_t tok = st.cur_tok[b];
  const uint4* src0 = reinterpret_cast<const uint4*>(st.last_hidden + (size_t)b * H);
  const uint4* src1 = reinterpret_cast<const uint4*>(codec_emb + (size_t)tok * H);
  uint4* d0 = reinterpret_cast<uint4*>(X + (size_t)(2 * b) * H);
  uint4* d1 = reinterpret_cast<uint4*>(X + (size_t)(2 * b + 1) * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) {
    d0[i] = src0[i];
    d1[i] = src1[i];
  }
  if (threadIdx.x < n_ac) st.amax[(size_t)threadIdx.x * B + b] = 0ull;
  if (threadIdx.x == 0) st.frame_codes[b * 16] = tok;
}

// CP pass g >= 1 input: code = argmax of pass g-1; X[b] = codec_embeddings[g-1][code].
__global__ void cp_embed_kernel(FrameState st, const bf16* __restrict__ emb_g, bf16* __restrict__ X, int H, int B, int g) {
  const int b = blockIdx.x;
  const uint32_t code = argmax_key_index(st.amax[(size_t)(g - 1) * B + b]);
  const uint4* src = reinterpret_cast<const uint4*>(emb_g + (size_t)code * H);
  uint4* d = reinterpret_cast<uint4*>(X + (size_t)b * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) d[i] = src[i];
  if (threadIdx.x == 0) st.frame_codes[b * 16 + g] = code;
}

// End of the CP frame: emit the frame, build the talker input
//   acc = E0[a0]; acc = bf16(acc + Ei[ai]) (i = 1..14, in order); summed = bf16(sem + acc);
//   step_input = bf16(summed + (frame_idx < lt ? trailing[frame_idx] : tts_pad)).
struct EmbTable { const bf16* e[15]; };
// How the threads that execute a row body synchronise: the whole block (stand-alone kernels, mega.cuh / mega2.cuh), or
// the 512 compute threads of a warp-specialised kernel whose producer warp does not take part (mega4.cuh).
struct SyncBlock {
  static __device__ __forceinline__ void sync() { __syncthreads(); }
  static __device__ __forceinline__ int threads() { return blockDim.x; }
};
struct SyncCompute512 {
  static __device__ __forceinline__ void sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
  static __device__ __forceinline__ int threads() { return 512; }
};
// body for one row b, executed by a whole block; codes_sm: 16 x u32 of shared memory
template <typename SY = SyncBlock>
__device__ __forceinline__ void frame_finish_row(const FrameState& st, const EmbTable& tab, const bf16* __restrict__ codec_emb,
                                                 bf16* __restrict__ step_input, int H, int B, int n_ac, int b,
                                                 uint32_t* codes_sm) {
  if (threadIdx.x < 16) {
    uint32_t c;
    if (threadIdx.x == 0) c = st.cur_tok[b];
    else if (threadIdx.x < n_ac) c = __ldcg(st.frame_codes + b * 16 + threadIdx.x);
    else c = argmax_key_index(__ldcg(st.amax + (size_t)(n_ac - 1) * B + b));
    codes_sm[threadIdx.x] = c;
  }
  SY::sync();
  const int fi = st.frame_idx[b];
  const bool active = !st.done[b];
  if (active && threadIdx.x < 16 && fi < st.frames_cap)
    st.codes[((size_t)b * st.frames_cap + fi) * 16 + threadIdx.x] = codes_sm[threadIdx.x];
  if (active && threadIdx.x == 0) st.n_frames[b] = fi + 1;
  const bf16* text = (fi < st.lt[b]) ? st.trailing + ((size_t)b * st.lt_max + fi) * H : st.tts_pad;
  for (int c = threadIdx.x * 8; c < H; c += SY::threads() * 8) {
    float acc[8], f[8];
    unpack8(*reinterpret_cast<const uint4*>(tab.e[0] + (size_t)codes_sm[1] * H + c), acc);
    for (int i = 1; i < n_ac; ++i) {
      unpack8(*reinterpret_cast<const uint4*>(tab.e[i] + (size_t)codes_sm[1 + i] * H + c), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = rbf(acc[e] + f[e]);
    }
    unpack8(*reinterpret_cast<const uint4*>(codec_emb + (size_t)codes_sm[0] * H + c), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = rbf(f[e] + acc[e]);
    unpack8(*reinterpret_cast<const uint4*>(text + c), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = rbf(acc[e] + f[e]);
    *reinterpret_cast<uint4*>(step_input + (size_t)b * H + c) = pack8(acc);
  }
  SY::sync();
}

__global__ void frame_finish_kernel(FrameState st, EmbTable tab, const bf16* __restrict__ codec_emb,
                                    bf16* __restrict__ step_input, int H, int B, int n_ac) {
  __shared__ uint32_t codes[16];
  frame_finish_row(st, tab, codec_emb, step_input, H, B, n_ac, blockIdx.x, codes);
}

// gather row (b, lens[b]-1) of a [B][l_max][H] tensor
__global__ void gather_last_kernel(const bf16* __restrict__ x, const int* __restrict__ lens, int l_max, int H,
                                   bf16* __restrict__ out) {
  const int b = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(x + ((size_t)b * l_max + (lens[b] - 1)) * H);
  uint4* d = reinterpret_cast<uint4*>(out + (size_t)b * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) d[i] = src[i];
}

// prompt assembly: out[t] = (text_id >= 0 ? text_proj[t] : 0) (+) (codec_id >= 0 ? codec_emb[codec_id] : 0)
// with one bf16 rounding when both are present (talker.rs:476-477, 486-487).
__global__ void assemble_embeds_kernel(const bf16* __restrict__ text_proj, const int* __restrict__ text_ids,
                                       const int* __restrict__ codec_ids, const bf16* __restrict__ codec_emb, int H,
                                       bf16* __restrict__ out) {
  const int t = blockIdx.x;
  const int ti = text_ids[t], ci = codec_ids[t];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v = 0.f;
    if (ti >= 0 && ci >= 0) v = rbf(bf2f(text_proj[(size_t)t * H + c]) + bf2f(codec_emb[(size_t)ci * H + c]));
    else if (ti >= 0) v = bf2f(text_proj[(size_t)t * H + c]);
    else if (ci >= 0) v = bf2f(codec_emb[(size_t)ci * H + c]);
    out[(size_t)t * H + c] = f2bf(v);
  }
}

// Voice-clone prompt assembly (prefill_voice_clone talker.rs:511-564, build_icl_prompt streaming overlay talker.rs:684-704,
// sum_ref_codec_embeddings lib.rs:1239-1257).  As assemble_embeds_kernel, with two more kinds of codec part:
//   codec_id == -2       : the row's continuous speaker embedding (speaker[b], bf16 [B][H])
//   codec_id <= -16      : the 16-way embedding sum of reference frame t = -16 - codec_id of row b:
//                          E_talker[c0] + E_cp0[c1] + ... + E_cp14[c15], added in that order, each add rounded to bf16
// One bf16 rounding for (text part + codec part) when both are present.  l_max = positions per row (t = b * l_max + p).
struct RefTables { const bf16* e[16]; };
__global__ void assemble_embeds_ex_kernel(const bf16* __restrict__ text_proj, const int* __restrict__ text_ids,
                                          const int* __restrict__ codec_ids, RefTables tab, const bf16* __restrict__ speaker,
                                          const uint32_t* __restrict__ ref_codes, int t_ref_max, int l_max, int H,
                                          bf16* __restrict__ out) {
  const int t = blockIdx.x, b = t / l_max;
  const int ti = text_ids[t], ci = codec_ids[t];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float cv = 0.f;
    const bool has_c = ci >= 0 || ci == -2 || ci <= -16;
    if (ci >= 0) cv = bf2f(tab.e[0][(size_t)ci * H + c]);
    else if (ci == -2) cv = bf2f(speaker[(size_t)b * H + c]);
    else if (ci <= -16) {
      const uint32_t* codes = ref_codes + ((size_t)b * t_ref_max + (-16 - ci)) * 16;
      cv = bf2f(tab.e[0][(size_t)codes[0] * H + c]);
#pragma unroll
      for (int g = 1; g < 16; ++g) cv = rbf(cv + bf2f(tab.e[g][(size_t)codes[g] * H + c]));
    }
    float v = 0.f;
    if (ti >= 0 && has_c) v = rbf(bf2f(text_proj[(size_t)t * H + c]) + cv);
    else if (ti >= 0) v = bf2f(text_proj[(size_t)t * H + c]);
    else if (has_c) v = cv;
    out[(size_t)t * H + c] = f2bf(v);
  }
}

__global__ void gather_rows_kernel(const bf16* __restrict__ table, const int* __restrict__ ids, int H, int n_rows_table,
                                   bf16* __restrict__ out) {
  const int t = blockIdx.x;
  int id = ids[t];
  if (id < 0 || id >= n_rows_table) id = 0;
  const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)id * H);
  uint4* d = reinterpret_cast<uint4*>(out + (size_t)t * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) d[i] = src[i];
}

// =================================================================================================
// Sampler: penalties -> temperature -> top-k -> top-p -> softmax -> multinomial -> state update.
// ref: apply_generation_penalties_gpu (lib.rs:1271-1322), sample/top_k_filter/top_p_filter (GPU
// tensor path)/multinomial_sample (sampling.rs:140-319), SamplingContext::rand_f32 (sampling.rs:84-94),
// update_penalty_mask (lib.rs:662-673), EOS test (lib.rs:581-585).
// One block (1024 threads) per row; vocab <= 4096.  The full row is bitonic-sorted once in shared
// memory; top-k and top-p thresholds are read from the sorted copy, and the two order-sensitive
// f32 reductions (softmax denominator, CDF) run sequentially in index order over the surviving
// entries -- the same order as the candle CPU kernels and the oracle.
struct SampleArgs {
  const float* logits;      // [B][V]
  uint8_t* seen;            // [B][V]
  unsigned long long* rng;  // [B]
  uint32_t* tok_out;        // [B]
  int* token_count;         // [B] or null -> use token_count_imm
  int token_count_imm;
  int* done;                // [B] or null
  int* offset;              // [B] or null: advanced when the row is still active
  int* frame_idx;           // [B] or null
  int* host_flags;          // or null
  int V, B;
  float inv_temp;           // (float)(1.0 / temperature)
  int use_temp, greedy;
  int top_k;
  float top_p;
  int use_top_p;
  float pen, inv_pen;       // (float)penalty, 1.0f / (float)penalty
  int use_pen;
  int eos, min_new_tokens;
  int advance;              // 1: this is a loop iteration (advance offset / frame_idx)
};

// Shared-memory scratch of one sampler row: 4096 + 4096 floats, 4096 u16, 4 scalars (40976 bytes).
struct SampleSmem {
  float xs[4096];                  // penalised / tempered logits, vocab order
  float srt[4096];                 // sorted descending; reused as kept_e once the thresholds are known
  unsigned short kept_idx[4096];
  float s_thr, s_mx;
  int s_nkept, s_tok;
};

// One row, executed by a whole block of any size that is a multiple of 32.
template <typename SY = SyncBlock>
__device__ __forceinline__ void sample_row_body(const SampleArgs& a, const int b, SampleSmem& sm) {
  float* xs = sm.xs;
  float* srt = sm.srt;
  unsigned short* kept_idx = sm.kept_idx;
  float* kept_e = srt;
  float& s_thr = sm.s_thr;
  float& s_mx = sm.s_mx;
  int& s_nkept = sm.s_nkept;
  int& s_tok = sm.s_tok;
  const int tid = threadIdx.x, V = a.V, NT = SY::threads();
  const float* lg = a.logits + (size_t)b * V;
  uint8_t* seen = a.seen + (size_t)b * V;
  const int tcount = a.token_count ? a.token_count[b] : a.token_count_imm;
  for (int i = tid; i < 4096; i += NT) {
    float x = -INFINITY;
    if (i < V) {
      x = __ldcg(lg + i);
      if (a.use_pen && seen[i]) x = x * ((x > 0.f) ? a.inv_pen : a.pen);          // sampling.rs:388-399
      if (i >= V - 1024 && i != Q3_CODEC_EOS) x = -INFINITY;                      // tts.rs:26-37 (constant EOS id, lib.rs:543-547)
      if (a.eos >= 0 && tcount < a.min_new_tokens && i == a.eos) x = -INFINITY;   // lib.rs:1304-1318
      if (a.use_temp) x = x * a.inv_temp;                                         // sampling.rs:148-152
    }
    xs[i] = x;
    srt[i] = x;
  }
  SY::sync();
  if (a.greedy) {                                                                 // sampling.rs:155-157
    // arg-max, lowest index among ties
    unsigned long long best = 0ull;
    for (int i = tid; i < V; i += NT) {
      unsigned long long k = argmax_key(xs[i], i);
      best = k > best ? k : best;
    }
    unsigned long long* red = reinterpret_cast<unsigned long long*>(kept_e);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if ((tid & 31) == 0) red[tid >> 5] = best;
    SY::sync();
    if (tid == 0) {
      for (int w = 1; w < (NT >> 5); ++w) best = red[w] > best ? red[w] : best;
      s_tok = (int)argmax_key_index(best);
    }
    SY::sync();
  } else {
    // ---- top-k threshold by RADIX SELECT (4 passes over an order-preserving 32-bit key), then a sort of the survivors only.
    // Same results as the full sort below (the threshold is the exact k-th largest value, ties keep extras, and the top-p
    // sums run over the survivors in descending order), at ~1/8 of its 78 block-wide stages.  Falls through to the full sort
    // when top-k is off or more than 256 values tie into the survivor set.
    bool have_thr = false;
    if (a.top_k > 0) {
      unsigned* hist = reinterpret_cast<unsigned*>(srt);            // [256]
      unsigned* sel = hist + 256;                                   // [0] prefix, [1] remaining, [2] survivor count
      float* sv = srt + 1024;                                       // [256] survivor values (sorted descending below)
      const int k = a.top_k < V ? a.top_k : V;
      auto okey = [](float x) -> unsigned {                          // monotone: larger float -> larger key; -inf is the smallest
        const unsigned u = __float_as_uint(x);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      };
      if (tid == 0) { sel[0] = 0u; sel[1] = (unsigned)k; }
      for (int pass = 3; pass >= 0; --pass) {
        for (int i = tid; i < 256; i += NT) hist[i] = 0u;
        SY::sync();
        const unsigned prefix = sel[0];
        for (int i = tid; i < V; i += NT) {
          const unsigned key = okey(xs[i]);
          if (pass == 3 || (key >> (8 * (pass + 1))) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
        }
        SY::sync();
        if (tid < 32) {
          // lane l owns digits 8l .. 8l+7; find the digit d with  count(digits > d) < remaining <= count(digits >= d)
          unsigned c[8], mine = 0u;
#pragma unroll
          for (int j = 0; j < 8; ++j) { c[j] = hist[8 * tid + j]; mine += c[j]; }
          unsigned above = 0u;                                       // keys in the digits of all higher lanes
          for (int l = 31; l >= 1; --l) {
            const unsigned v = __shfl_sync(0xffffffffu, mine, l);
            if (l > tid) above += v;
          }
          const unsigned remaining = sel[1];
          __syncwarp();
          if (above < remaining && remaining <= above + mine) {
            unsigned acc = above;
            int d = 7;
            for (; d >= 0; --d) {
              if (remaining <= acc + c[d]) break;
              acc += c[d];
            }
            sel[0] = (prefix << 8) | (unsigned)(8 * tid + d);
            sel[1] = remaining - acc;
          }
        }
        SY::sync();
      }
      const unsigned thr_key = sel[0];
      if (tid == 0) sel[2] = 0u;
      SY::sync();
      for (int i = tid; i < V; i += NT) {
        if (okey(xs[i]) >= thr_key) {
          const unsigned pos = atomicAdd(&sel[2], 1u);
          if (pos < 256u) sv[pos] = xs[i];
        }
      }
      SY::sync();
      const int n1 = (int)sel[2];
      if (n1 <= 256) {
        // sort the survivors descending: bitonic over the next power of two (padded with -inf); up to 64 slots (the usual
        // case: top-k 50) one warp does it with warp barriers only
        int npad = 2;
        while (npad < n1) npad <<= 1;
        for (int i = n1 + tid; i < npad; i += NT) sv[i] = -INFINITY;
        SY::sync();
        if (npad <= 64) {
          if (tid < 32) {
            for (int kk = 2; kk <= npad; kk <<= 1) {
              for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < npad; i += 32) {
                  const int ixj = i ^ j;
                  if (ixj > i) {
                    const float x = sv[i], y = sv[ixj];
                    const bool desc = ((i & kk) == 0);
                    if (desc ? (x < y) : (x > y)) { sv[i] = y; sv[ixj] = x; }
                  }
                }
                __syncwarp();
              }
            }
          }
          SY::sync();
        } else {
          for (int kk = 2; kk <= npad; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
              for (int i = tid; i < npad; i += NT) {
                const int ixj = i ^ j;
                if (ixj > i) {
                  const float x = sv[i], y = sv[ixj];
                  const bool desc = ((i & kk) == 0);
                  if (desc ? (x < y) : (x > y)) { sv[i] = y; sv[ixj] = x; }
                }
              }
              SY::sync();
            }
          }
        }
        if (tid == 0) {
          float thr = sv[k - 1];                      // sampling.rs:203-211: keep >= k-th largest (ties keep extras: all n1 survivors)
          if (a.use_top_p) {                          // sampling.rs:263-286, sums in sorted order as in the full-sort path
            const float mx = sv[0];
            float sum = 0.f;
            for (int i = 0; i < n1; ++i) sum += expf(sv[i] - mx);
            float cum = 0.f;
            float min_kept = sv[0];
            for (int i = 0; i < n1; ++i) {
              if (cum >= a.top_p) break;
              min_kept = sv[i];
              cum += expf(sv[i] - mx) / sum;
            }
            thr = fmaxf(thr, min_kept);
          }
          s_thr = thr;
          s_mx = sv[0];
        }
        have_thr = true;
        SY::sync();
      }
      if (!have_thr) {
        // restore the copy the full sort works on
        for (int i = tid; i < 4096; i += NT) srt[i] = xs[i];
        SY::sync();
      }
    }
    if (!have_thr) {
    // bitonic sort, descending, 4096 keys / 1024 threads
    for (int k = 2; k <= 4096; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < 4096; i += NT) {
          int ixj = i ^ j;
          if (ixj > i) {
            float x = srt[i], y = srt[ixj];
            bool desc = ((i & k) == 0);
            if (desc ? (x < y) : (x > y)) { srt[i] = y; srt[ixj] = x; }
          }
        }
        SY::sync();
      }
    }
    if (tid == 0) {
      float thr = -INFINITY;
      int n1 = V;                                   // survivors of top-k, as a prefix of srt
      if (a.top_k > 0) {                            // sampling.rs:203-211: keep >= k-th largest
        int k = a.top_k < V ? a.top_k : V;
        thr = srt[k - 1];
        n1 = k;
        while (n1 < V && srt[n1] >= thr) ++n1;      // ties keep extras
      }
      if (a.use_top_p) {                            // sampling.rs:263-286
        // softmax over the sorted, top-k-filtered row: sequential sum in sorted order
        const float mx = srt[0];
        float sum = 0.f;
        for (int i = 0; i < n1; ++i) sum += expf(srt[i] - mx);
        float cum = 0.f;                            // exclusive cumulative probability
        float min_kept = srt[0];
        for (int i = 0; i < n1; ++i) {
          if (cum >= a.top_p) break;                // removed from here on
          min_kept = srt[i];
          cum += expf(srt[i] - mx) / sum;
        }
        thr = fmaxf(thr, min_kept);
      }
      s_thr = thr;
      s_mx = srt[0];
    }
    SY::sync();
    }   // full-sort path
    // compact survivors (x >= thr) in index order; one warp, ballot + popc
    if (tid < 32) {
      const float thr = s_thr, mx = s_mx;
      int n = 0;
      for (int base = 0; base < V; base += 32) {
        int i = base + tid;
        bool keep = (i < V) && (xs[i] >= thr) && (xs[i] != -INFINITY);
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          int p = n + __popc(m & ((1u << tid) - 1u));
          kept_idx[p] = (unsigned short)i;
          kept_e[p] = expf(xs[i] - mx);
        }
        n += __popc(m);
      }
      if (tid == 0) s_nkept = n;
    }
    SY::sync();
    if (tid == 0) {
      const int n = s_nkept;
      float sum = 0.f;
      for (int i = 0; i < n; ++i) sum += kept_e[i];            // softmax denominator, index order
      // PCG-XSH-RR 64/32 (sampling.rs:84-94)
      unsigned long long old = a.rng[b];
      a.rng[b] = old * 6364136223846793005ull + 1442695040888963407ull;
      uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
      uint32_t rot = (uint32_t)(old >> 59u);
      uint32_t outp = (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
      const float u = __uint2float_rn(outp) / 4294967296.0f;
      // multinomial (sampling.rs:290-319): first index with inclusive cumsum >= u, else index 0
      float cum = 0.f;
      int tok = 0;
      bool found = false;
      if (!(u <= 0.f)) {
        for (int i = 0; i < n; ++i) {
          cum += kept_e[i] / sum;
          if (cum >= u) { tok = (int)kept_idx[i]; found = true; break; }
        }
      } else {
        found = true;   // cumsum[0] >= 0 always holds: vocab index 0
        tok = 0;
      }
      if (!found) tok = 0;
      s_tok = tok;
    }
    SY::sync();
  }
  if (tid == 0) {
    const int tok = s_tok;
    a.tok_out[b] = (uint32_t)tok;
    if (tok < V) seen[tok] = 1;                                 // lib.rs:662-673
    if (a.token_count) a.token_count[b] = tcount + 1;
    if (a.done) {
      const int was_done = a.done[b];
      if (a.advance && !was_done) {
        if (a.offset) a.offset[b] += 1;
        if (a.frame_idx) a.frame_idx[b] += 1;
      }
      if (a.eos >= 0 && tok == a.eos) a.done[b] = 1;            // lib.rs:581-585 (tested next iteration)
    }
  }
}

__global__ void __launch_bounds__(1024) sample_kernel(const SampleArgs a) {
  __shared__ SampleSmem sm;
  sample_row_body(a, blockIdx.x, sm);
}

// number of rows still running -> mapped host flag (polled every few frames by the host loop)
__global__ void count_active_kernel(const int* __restrict__ done, int B, int* host_flags) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int n = 0;
    for (int b = 0; b < B; ++b) n += done[b] ? 0 : 1;
    host_flags[0] = n;
  }
}
