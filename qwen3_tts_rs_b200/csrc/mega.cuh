// Persistent "frame" kernel: one cooperative launch runs whole decode frames -- the 15 dependent
// code-predictor passes, the 16-way embedding sum, the 28-layer talker step, the codec head and the
// sampler -- for all batch rows, with grid-wide barriers between dependent phases instead of ~770 kernel
// launches per frame (the reference issues ~2.5 k launches and one host sync per frame,
// docs/CUSTOM_CUDA_KERNELS_PLAN.md:5-8, src/lib.rs:648-649).
//
// Design (B200: 148 SMs, 1 CTA of 512 threads per SM, up to 227 KB shared memory):
//   * every weight matrix is streamed from HBM exactly once per phase; a phase's output rows are cut into
//     16-row tiles dealt round-robin to the CTAs, and inside a CTA the 16 warps split K (each warp issues
//     all of its 128-bit weight loads for a tile before consuming them, ~64 KB in flight per SM), then
//     combine their partial sums through shared memory in a fixed order (deterministic);
//   * the skinny GEMM  Y[t][n] = sum_k W[n][k] X[t][k]  (t <= 16 tokens) runs on the tensor cores with
//     mma.sync.m16n8k16 (bf16 x bf16 -> f32): A fragments are loaded straight from global memory -- each
//     thread's 16 contiguous bytes of a weight row ARE its fragment under a k-permutation that is applied
//     identically to the activation operand -- so weights never touch shared memory.  tcgen05 needs >= 64-row
//     tiles in shared memory; at <= 768 tiles of work per phase over 148 SMs that would force a cross-CTA
//     split-K and a second reduction phase per GEMM, so the legacy-MMA shape is the right one for this
//     HBM-bound, latency-critical step (tensor throughput is irrelevant at intensity <= 16 FLOP/B);
//   * activations of the current phase are staged once per CTA in shared memory (bf16, padded rows so the
//     B-fragment loads are bank-conflict free) together with the fused prologue: RMSNorm, or
//     residual-add + RMSNorm with the reference kernel's exact summation order (norm.cuh);
//   * epilogues (bias, SiLU, SwiGLU, residual, f32 logits, packed arg-max) are fused; rounding points are
//     the reference's (every candle op writes bf16).
#pragma once
#include "common.cuh"
#include "decode_kernels.cuh"
#include "gemv.cuh"
#include "model.h"
#include "norm.cuh"

constexpr int MEGA_THREADS = 512;
constexpr int MEGA_WARPS = 16;
constexpr int MEGA_TMAX = 16;       // tokens per phase (batch <= 8: the CP prefill pass has 2 tokens per row)

struct MegaStack { const LayerW* layers; int n_layers, H, I, heads, kv_heads; };

struct MegaArgs {
  MegaStack tk, cp;
  const bf16 *codec_emb, *t_norm, *codec_head, *cp_proj_w, *cp_proj_b, *cp_norm;
  const bf16* cp_emb[15];
  const bf16* cp_head[15];
  const bf16 *cp_cos, *cp_sin, *t_cos, *t_sin;
  int H, C, V, cpV, n_ac, B;
  float eps;
  FrameState fs;
  bf16 *tk_k, *tk_v, *cp_k, *cp_v;
  int max_seq, cp_max_seq;
  bf16 *x, *qkv, *attn, *o, *h1, *act, *step_input;
  float* logits;
  float* cp_logits;      // optional [n_ac][B][cpV]
  unsigned* bar;
  SampleArgs smp;
  int n_frames;          // loop iterations to run in this launch
  int do_cp, do_finish, do_talker, do_sample;
  const bf16* ext_step_input;   // per-op entry: talker input supplied by the caller (do_finish == 0)
};

// ---------------------------------------------------------------------------------------------------
struct GridBar {
  unsigned* ctr;
  unsigned epoch;
};
__device__ __forceinline__ void grid_sync(GridBar& gb) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned target = (gb.epoch + 1u) * gridDim.x;
    __threadfence();
    atomicAdd(gb.ctr, 1u);
    while (*((volatile unsigned*)gb.ctr) < target) {
    }
    __threadfence();
  }
  gb.epoch += 1u;
  __syncthreads();
}

__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
  return __uint_as_float(((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(p))) << 16);
}

__device__ __forceinline__ void mma_bf16_16816(float c[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------------
// activation staging
enum XMode { X_PLAIN = 0, X_RMSNORM = 1, X_RESNORM = 2, X_CP0 = 3, X_CPG = 4 };

struct GemvP {
  const bf16* W;
  const bf16* W2;        // dual (SwiGLU) partner or null
  int N, K, T;
  int xmode;
  const bf16* X;         // X_PLAIN / X_RMSNORM: [T][ldx];  X_RESNORM: the o_proj output
  int ldx;
  const bf16* X2;        // X_RESNORM: residual input x
  const bf16* norm_w;
  bf16* h1_out;          // X_RESNORM: rounded sum written by CTA 0
  bf16* xn_out;          // X_RMSNORM: normalised rows written by CTA 0 (talker last_hidden)
  const bf16* emb;       // X_CP0: talker codec embedding; X_CPG: codec_embeddings[g-1]
  int g;                 // X_CPG: pass index
  int epi;
  bf16* Y;
  int ldy;
  const bf16* bias;
  const bf16* R;
  int ldr;
  float* Yf;
  unsigned long long* amax;
};

// xs: [T8][K + 32] bf16, rows >= T zero.  Returns nothing; ends with __syncthreads().
__device__ __noinline__ void mega_stage_x(const MegaArgs& a, const GemvP& p, bf16* xs, float* s_part) {
  const int K = p.K, XS = K + 32, T = p.T, T8 = (T + 7) & ~7;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool writer = blockIdx.x == 0;
  auto src_row = [&](int t) -> const bf16* {
    if (p.xmode == X_CP0) {
      const int b = t >> 1;
      return (t & 1) ? p.emb + (size_t)a.fs.cur_tok[b] * K : a.fs.last_hidden + (size_t)b * K;
    }
    if (p.xmode == X_CPG) {
      const uint32_t code = argmax_key_index(__ldcg(a.fs.amax + (size_t)(p.g - 1) * a.B + t));
      return p.emb + (size_t)code * K;
    }
    return p.X + (size_t)t * p.ldx;
  };
  if (p.xmode == X_CPG && writer && tid < T)
    a.fs.frame_codes[tid * 16 + p.g] = argmax_key_index(__ldcg(a.fs.amax + (size_t)(p.g - 1) * a.B + tid));
  if (p.xmode == X_CP0 && writer && tid < a.B) a.fs.frame_codes[tid * 16] = a.fs.cur_tok[tid];
  // zero padding rows
  for (int i = tid; i < (T8 - T) * (K >> 3); i += MEGA_THREADS) {
    int r = i / (K >> 3), q = i - r * (K >> 3);
    *reinterpret_cast<uint4*>(xs + (size_t)(T + r) * XS + q * 8) = make_uint4(0, 0, 0, 0);
  }
  if (p.xmode == X_PLAIN || p.xmode == X_CP0 || p.xmode == X_CPG) {
    const int K8 = K >> 3;
    for (int i = tid; i < T * K8; i += MEGA_THREADS) {
      int t = i / K8, q = i - t * K8;
      *reinterpret_cast<uint4*>(xs + (size_t)t * XS + q * 8) = ldcg16(src_row(t) + q * 8);
    }
    __syncthreads();
    return;
  }
  const bool res = p.xmode == X_RESNORM;
  if (K >= 1024) {
    const int grp = tid >> 7, g = tid & 127;      // four 128-thread groups, one token each
    for (int t = grp; t < T; t += 4) {
      const bf16* xr = p.X + (size_t)t * p.ldx;
      const bf16* rr = res ? p.X2 + (size_t)t * p.ldx : nullptr;
      float pp[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) pp[e] = 0.f;
      for (int c = 8 * g; c < K; c += 1024) {
        float f[8];
        unpack8(ldcg16(xr + c), f);
        if (res) {
          float r2[8];
          unpack8(ldcg16(rr + c), r2);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = f[e] + r2[e];
          // keep the ROUNDED sum in shared memory for pass 2 (the reference re-reads its stored sum)
          uint4 pk = pack8(f);
          *reinterpret_cast<uint4*>(xs + (size_t)t * XS + c) = pk;
          if (writer) *reinterpret_cast<uint4*>(p.h1_out + (size_t)t * K + c) = pk;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) pp[e] = fmaf(f[e], f[e], pp[e]);
      }
      const float tot = sumsq_ref_large_finish(pp, g, s_part + grp * 32, 1 + grp);
      const float sc = ref_mean_rsqrt(tot, K, a.eps);
      for (int c = 8 * g; c < K; c += 1024) {
        float f[8], w[8], o[8];
        if (res) unpack8(*reinterpret_cast<const uint4*>(xs + (size_t)t * XS + c), f);
        else unpack8(ldcg16(xr + c), f);
        unpack8(*reinterpret_cast<const uint4*>(p.norm_w + c), w);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (sc * f[e]) * w[e];
        const uint4 pk = pack8(o);
        *reinterpret_cast<uint4*>(xs + (size_t)t * XS + c) = pk;
        if (p.xn_out != nullptr && writer) *reinterpret_cast<uint4*>(p.xn_out + (size_t)t * K + c) = pk;
      }
    }
  } else {
    for (int t = warp; t < T; t += MEGA_WARPS) {
      const bf16* xr = p.X + (size_t)t * p.ldx;
      const bf16* rr = res ? p.X2 + (size_t)t * p.ldx : nullptr;
      float tmp = 0.f;
      for (int c = lane; c < K; c += 32) {
        float v = ldcg_bf16(xr + c);
        if (res) {
          v = v + ldcg_bf16(rr + c);
          const bf16 rounded = f2bf(v);
          xs[(size_t)t * XS + c] = rounded;
          if (writer) p.h1_out[(size_t)t * K + c] = rounded;
        }
        tmp = fmaf(v, v, tmp);
      }
      tmp = warp_sum_xor(tmp);
      const float sc = ref_mean_rsqrt(tmp, K, a.eps);
      for (int c = lane; c < K; c += 32) {
        const float f = res ? bf2f(xs[(size_t)t * XS + c]) : ldcg_bf16(xr + c);
        const bf16 o = f2bf((sc * f) * bf2f(p.norm_w[c]));
        xs[(size_t)t * XS + c] = o;
        if (p.xn_out != nullptr && writer) p.xn_out[(size_t)t * K + c] = o;
      }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------
// One skinny-GEMM phase.  smem: xs [T8][K+32] bf16 | red [16 warps][NT][(DUAL?2:1)][16][8] f32
template <bool DUAL>
__device__ __noinline__ void mega_gemv(const MegaArgs& a, const GemvP& p, unsigned char* smem, float* s_part) {
  const int K = p.K, XS = K + 32, T = p.T, T8 = (T + 7) & ~7, NT = T8 >> 3;
  bf16* xs = reinterpret_cast<bf16*>(smem);
  float* red = reinterpret_cast<float*>(smem + (((size_t)T8 * XS * 2 + 127) & ~(size_t)127));
  // 16-row tiles; 8-row tiles when there would be fewer tiles than CTAs
  const int RT = (p.N / 16 >= (int)gridDim.x) ? 16 : 8;
  const int n_tiles = p.N / RT;
  if ((int)blockIdx.x >= n_tiles) return;          // nothing to do here (block 0 always has a tile)
  mega_stage_x(a, p, xs, s_part);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int ksteps = K >> 5;                       // 32 k per step
  constexpr int NM = DUAL ? 2 : 1;
  constexpr int U = DUAL ? 2 : 4;                  // k-steps loaded per batch
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n0 = tile * RT;
    float acc[NM][2][4];                           // [matrix][n-tile][frag]  (T8 <= 16 -> NT <= 2)
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;
    const bf16* wlo[NM];
    const bf16* whi[NM];
    wlo[0] = p.W + (size_t)(n0 + g) * K + 8 * tg;
    whi[0] = p.W + (size_t)(n0 + g + 8) * K + 8 * tg;
    if (DUAL) {
      wlo[1] = p.W2 + (size_t)(n0 + g) * K + 8 * tg;
      whi[1] = p.W2 + (size_t)(n0 + g + 8) * K + 8 * tg;
    }
    for (int ks0 = warp; ks0 < ksteps; ks0 += MEGA_WARPS * U) {
      uint4 wl[NM][U], wh[NM][U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ks = ks0 + u * MEGA_WARPS;
        if (ks < ksteps) {
#pragma unroll
          for (int m = 0; m < NM; ++m) {
            wl[m][u] = ldg_stream(reinterpret_cast<const uint4*>(wlo[m] + ks * 32));
            if (RT == 16) wh[m][u] = ldg_stream(reinterpret_cast<const uint4*>(whi[m] + ks * 32));
            else wh[m][u] = make_uint4(0, 0, 0, 0);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ks = ks0 + u * MEGA_WARPS;
        if (ks < ksteps) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            if (nt < NT) {
              const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)(nt * 8 + g) * XS + ks * 32 + 8 * tg);
#pragma unroll
              for (int m = 0; m < NM; ++m) {
                mma_bf16_16816(acc[m][nt], wl[m][u].x, wh[m][u].x, wl[m][u].y, wh[m][u].y, xv.x, xv.y);
                mma_bf16_16816(acc[m][nt], wl[m][u].z, wh[m][u].z, wl[m][u].w, wh[m][u].w, xv.z, xv.w);
              }
            }
          }
        }
      }
    }
    // partial sums -> shared memory: red[warp][nt][m][row 16][col 8]
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        if (nt < NT) {
          float* r = red + ((((size_t)warp * NT + nt) * NM + m) * 16) * 8;
          r[g * 8 + 2 * tg] = acc[m][nt][0];
          r[g * 8 + 2 * tg + 1] = acc[m][nt][1];
          r[(g + 8) * 8 + 2 * tg] = acc[m][nt][2];
          r[(g + 8) * 8 + 2 * tg + 1] = acc[m][nt][3];
        }
    __syncthreads();
    // fixed-order combine + epilogue: thread -> (row, token)
    for (int idx = tid; idx < RT * T; idx += MEGA_THREADS) {
      const int row = idx % RT, t = idx / RT, nt = t >> 3, col = t & 7;
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < MEGA_WARPS; ++w) {
        v0 += red[((((size_t)w * NT + nt) * NM + 0) * 16 + row) * 8 + col];
        if (DUAL) v1 += red[((((size_t)w * NT + nt) * NM + 1) * 16 + row) * 8 + col];
      }
      const int n = n0 + row;
      const float v = rbf(v0);
      switch (p.epi) {
        case EPI_STORE: p.Y[(size_t)t * p.ldy + n] = f2bf(v); break;
        case EPI_BIAS: p.Y[(size_t)t * p.ldy + n] = f2bf(v + bf2f(p.bias[n])); break;
        case EPI_BIAS_SILU: {
          const float y = rbf(v + bf2f(p.bias[n]));
          p.Y[(size_t)t * p.ldy + n] = f2bf(silu_f(y));
        } break;
        case EPI_RESIDUAL: {
          const float r = ldcg_bf16(p.R + (size_t)t * p.ldr + n);
          p.Y[(size_t)t * p.ldy + n] = f2bf(r + v);
        } break;
        case EPI_SWIGLU: {
          const float s = rbf(silu_f(v));
          p.Y[(size_t)t * p.ldy + n] = f2bf(s * rbf(v1));
        } break;
        case EPI_LOGITS: {
          if (p.Yf != nullptr) p.Yf[(size_t)t * p.N + n] = v;
          if (p.amax != nullptr) atomicMax(p.amax + t, argmax_key(v, n));
        } break;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// QK-norm + RoPE + KV append + attention for (row b, kv head) items; S tokens per row processed in order.
struct AttnP {
  const bf16* qkv;       // [T][(heads+2kv)*128]
  bf16* out;             // [T][heads*128]
  bf16 *k_cache, *v_cache;
  const bf16 *q_norm_w, *k_norm_w, *cos_tab, *sin_tab;
  const int* pos_base;
  int pos_add, S, B, heads, kv_heads, max_seq;
};

__device__ __noinline__ void mega_attn(const MegaArgs& a, const AttnP& p, unsigned char* smem) {
  float* sc0 = reinterpret_cast<float*>(smem);          // [max_seq]
  float* sc1 = sc0 + p.max_seq;
  float* qs = sc1 + p.max_seq;                           // [2][128] rotated queries (bf16 values)
  float* red = qs + 256;                                 // [16][2][128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nh = p.heads + 2 * p.kv_heads;
  const float scale = rbf(0.08838834764831845f);
  for (int item = blockIdx.x; item < p.B * p.kv_heads; item += gridDim.x) {
    const int b = item / p.kv_heads, kvh = item - b * p.kv_heads;
    bf16* kbase = p.k_cache + ((size_t)b * p.kv_heads + kvh) * p.max_seq * 128;
    bf16* vbase = p.v_cache + ((size_t)b * p.kv_heads + kvh) * p.max_seq * 128;
    for (int s = 0; s < p.S; ++s) {
      const int t = b * p.S + s;
      const int pos = (p.pos_base ? p.pos_base[b] : 0) + p.pos_add + s;
      const int L = pos + 1;
      // warps 0,1: q heads 2kvh, 2kvh+1; warp 2: k; warp 3: v
      if (warp < 4) {
        const int hh = warp < 2 ? 2 * kvh + warp : (warp == 2 ? p.heads + kvh : p.heads + p.kv_heads + kvh);
        const unsigned short* src = reinterpret_cast<const unsigned short*>(p.qkv + (size_t)t * nh * 128 + (size_t)hh * 128);
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(((uint32_t)__ldcg(src + lane + 32 * i)) << 16);
        if (warp == 3) {
#pragma unroll
          for (int i = 0; i < 4; ++i) vbase[(size_t)pos * 128 + lane + 32 * i] = f2bf(v[i]);
        } else {
          const bf16* nw = warp < 2 ? p.q_norm_w : p.k_norm_w;
          float tmp = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) tmp = fmaf(v[i], v[i], tmp);
          tmp = warp_sum_xor(tmp);
          const float sc = ref_mean_rsqrt(tmp, 128, a.eps);
          float n[4], o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) n[i] = rbf((sc * v[i]) * bf2f(nw[lane + 32 * i]));
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int d = lane + 32 * i;
            const float c = bf2f(p.cos_tab[(size_t)pos * 64 + d]), sn = bf2f(p.sin_tab[(size_t)pos * 64 + d]);
            o[i] = rbf(rbf(n[i] * c) - rbf(n[i + 2] * sn));
            o[i + 2] = rbf(rbf(n[i + 2] * c) + rbf(n[i] * sn));
          }
          if (warp < 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) qs[warp * 128 + lane + 32 * i] = o[i];
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) kbase[(size_t)pos * 128 + lane + 32 * i] = f2bf(o[i]);
          }
        }
      }
      __syncthreads();      // q in smem; this block's own K/V writes are visible to the block
      float q0[4], q1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { q0[i] = qs[4 * lane + i]; q1[i] = qs[128 + 4 * lane + i]; }
      for (int j = warp; j < L; j += MEGA_WARPS) {
        const uint2 u = __ldcg(reinterpret_cast<const uint2*>(kbase + (size_t)j * 128 + 4 * lane));
        const float k0 = bf_lo(u.x), k1 = bf_hi(u.x), k2 = bf_lo(u.y), k3 = bf_hi(u.y);
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
      if (warp < 2) {
        float* sc = warp == 0 ? sc0 : sc1;
        float m = -INFINITY;
        for (int j = lane; j < L; j += 32) m = fmaxf(m, sc[j]);
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < L; j += 32) {
          const float e = expf(sc[j] - m);
          sc[j] = e;
          sum += e;
        }
        sum = warp_sum_xor(sum);
        for (int j = lane; j < L; j += 32) sc[j] = rbf(sc[j] / sum);
      }
      __syncthreads();
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = warp; j < L; j += MEGA_WARPS) {
        const uint2 u = __ldcg(reinterpret_cast<const uint2*>(vbase + (size_t)j * 128 + 4 * lane));
        const float v0 = bf_lo(u.x), v1 = bf_hi(u.x), v2 = bf_lo(u.y), v3 = bf_hi(u.y);
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
      if (tid < 256) {
        const int h = tid >> 7, d = tid & 127;
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < MEGA_WARPS; ++w) acc += red[(w * 2 + h) * 128 + d];
        p.out[(size_t)t * p.heads * 128 + (size_t)(2 * kvh + h) * 128 + d] = f2bf(acc);
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// One decoder stack over T = B*S tokens, in place on a.x  (DecoderLayer::forward x layers).
__device__ __noinline__ void mega_layers(const MegaArgs& a, const MegaStack& st, int T, int S, const int* pos_base, int pos_add, bf16* kc,
                            bf16* vc, int cache_seq, const bf16* cos_tab, const bf16* sin_tab, unsigned char* smem,
                            float* s_part, GridBar& gb) {
  const int nh = st.heads + 2 * st.kv_heads;
  const size_t layer_stride = (size_t)a.B * st.kv_heads * cache_seq * 128;
  for (int l = 0; l < st.n_layers; ++l) {
    const LayerW w = st.layers[l];
    GemvP q{};
    q.W = w.wqkv; q.N = nh * 128; q.K = st.H; q.T = T; q.xmode = X_RMSNORM; q.X = a.x; q.ldx = st.H; q.norm_w = w.in_ln;
    q.epi = EPI_STORE; q.Y = a.qkv; q.ldy = nh * 128;
    mega_gemv<false>(a, q, smem, s_part);
    grid_sync(gb);
    AttnP at{};
    at.qkv = a.qkv; at.out = a.attn; at.k_cache = kc + l * layer_stride; at.v_cache = vc + l * layer_stride;
    at.q_norm_w = w.q_norm; at.k_norm_w = w.k_norm; at.cos_tab = cos_tab; at.sin_tab = sin_tab; at.pos_base = pos_base;
    at.pos_add = pos_add; at.S = S; at.B = a.B; at.heads = st.heads; at.kv_heads = st.kv_heads; at.max_seq = cache_seq;
    mega_attn(a, at, smem);
    grid_sync(gb);
    GemvP o{};
    o.W = w.wo; o.N = st.H; o.K = st.heads * 128; o.T = T; o.xmode = X_PLAIN; o.X = a.attn; o.ldx = st.heads * 128;
    o.epi = EPI_STORE; o.Y = a.o; o.ldy = st.H;
    mega_gemv<false>(a, o, smem, s_part);
    grid_sync(gb);
    GemvP gu{};
    gu.W = w.gate; gu.W2 = w.up; gu.N = st.I; gu.K = st.H; gu.T = T; gu.xmode = X_RESNORM; gu.X = a.o; gu.X2 = a.x; gu.ldx = st.H;
    gu.norm_w = w.post_ln; gu.h1_out = a.h1; gu.epi = EPI_SWIGLU; gu.Y = a.act; gu.ldy = st.I;
    mega_gemv<true>(a, gu, smem, s_part);
    grid_sync(gb);
    GemvP dn{};
    dn.W = w.down; dn.N = st.H; dn.K = st.I; dn.T = T; dn.xmode = X_PLAIN; dn.X = a.act; dn.ldx = st.I;
    dn.epi = EPI_RESIDUAL; dn.R = a.h1; dn.ldr = st.H; dn.Y = a.x; dn.ldy = st.H;
    mega_gemv<false>(a, dn, smem, s_part);
    grid_sync(gb);
  }
}

__device__ __noinline__ void mega_sample(const SampleArgs& sa, int b, SampleSmem& sm) { sample_row_body(sa, b, sm); }
__device__ __noinline__ void mega_finish(const MegaArgs& a, uint32_t* s_codes) {
  EmbTable tab{};
  for (int i = 0; i < a.n_ac; ++i) tab.e[i] = a.cp_emb[i];
  for (int b = blockIdx.x; b < a.B; b += gridDim.x)
    frame_finish_row(a.fs, tab, a.codec_emb, a.step_input, a.H, a.B, a.n_ac, b, s_codes);
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_frames_mega_kernel(const MegaArgs a) {
  extern __shared__ __align__(128) unsigned char mega_smem[];
  __shared__ float s_part[128];
  __shared__ uint32_t s_codes[16];
  GridBar gb{a.bar, 0u};
  const int B = a.B;
  for (int frame = 0; frame < a.n_frames; ++frame) {
    if (a.do_cp) {
      // ---- code predictor: 15 dependent passes (code_predictor.rs:320-416) ----
      if (blockIdx.x == 0 && threadIdx.x < a.n_ac * B) {
        for (int i = threadIdx.x; i < a.n_ac * B; i += MEGA_THREADS) a.fs.amax[i] = 0ull;
      }
      for (int g = 0; g < a.n_ac; ++g) {
        const int T = g == 0 ? 2 * B : B, S = g == 0 ? 2 : 1;
        GemvP pr{};
        pr.N = a.C; pr.K = a.H; pr.T = T; pr.xmode = g == 0 ? X_CP0 : X_CPG; pr.g = g;
        pr.emb = g == 0 ? a.codec_emb : a.cp_emb[g - 1];
        if (a.cp_proj_w) {
          pr.W = a.cp_proj_w; pr.bias = a.cp_proj_b; pr.epi = EPI_BIAS; pr.Y = a.x; pr.ldy = a.C;
          mega_gemv<false>(a, pr, mega_smem, s_part);
        } else {
          // no projection (talker hidden == CP hidden): the gathered rows are the layer input
          bf16* xs = reinterpret_cast<bf16*>(mega_smem);
          mega_stage_x(a, pr, xs, s_part);
          if (blockIdx.x == 0)
            for (int i = threadIdx.x; i < T * (a.H >> 3); i += MEGA_THREADS) {
              int t = i / (a.H >> 3), q = i - t * (a.H >> 3);
              *reinterpret_cast<uint4*>(a.x + (size_t)t * a.C + q * 8) = *reinterpret_cast<const uint4*>(xs + (size_t)t * (a.H + 32) + q * 8);
            }
        }
        grid_sync(gb);
        mega_layers(a, a.cp, T, S, nullptr, g == 0 ? 0 : g + 1, a.cp_k, a.cp_v, a.cp_max_seq, a.cp_cos, a.cp_sin, mega_smem, s_part, gb);
        GemvP hd{};
        hd.W = a.cp_head[g]; hd.N = a.cpV; hd.K = a.C; hd.T = B; hd.xmode = X_RMSNORM; hd.norm_w = a.cp_norm;
        hd.X = g == 0 ? a.x + a.C : a.x; hd.ldx = g == 0 ? 2 * a.C : a.C;
        hd.epi = EPI_LOGITS; hd.amax = a.fs.amax + (size_t)g * B;
        hd.Yf = a.cp_logits ? a.cp_logits + (size_t)g * B * a.cpV : nullptr;
        mega_gemv<false>(a, hd, mega_smem, s_part);
        grid_sync(gb);
      }
    }
    if (a.do_finish) {
      // ---- emit the frame, build the talker input (lib.rs:605-622) ----
      mega_finish(a, s_codes);
      grid_sync(gb);
    }
    if (a.do_talker) {
      // ---- talker step (talker.rs:716-736) ----
      const bf16* in = a.do_finish ? a.step_input : a.ext_step_input;
      for (int i = blockIdx.x * MEGA_THREADS + threadIdx.x; i < B * (a.H >> 3); i += gridDim.x * MEGA_THREADS)
        reinterpret_cast<uint4*>(a.x)[i] = ldcg16(reinterpret_cast<const uint4*>(in) + i);
      grid_sync(gb);
      mega_layers(a, a.tk, B, 1, a.fs.offset, 0, a.tk_k, a.tk_v, a.max_seq, a.t_cos, a.t_sin, mega_smem, s_part, gb);
      GemvP hd{};
      hd.W = a.codec_head; hd.N = a.V; hd.K = a.H; hd.T = B; hd.xmode = X_RMSNORM; hd.norm_w = a.t_norm; hd.X = a.x; hd.ldx = a.H;
      hd.xn_out = a.fs.last_hidden; hd.epi = EPI_LOGITS; hd.Yf = a.logits;
      mega_gemv<false>(a, hd, mega_smem, s_part);
      grid_sync(gb);
    }
    if (a.do_sample) {
      // ---- penalties + sampling + state update (lib.rs:639-651) ----
      SampleSmem& sm = *reinterpret_cast<SampleSmem*>(mega_smem);
      for (int b = blockIdx.x; b < B; b += gridDim.x) mega_sample(a.smp, b, sm);
      grid_sync(gb);
      // stop early once every row has sampled EOS (uniform decision: all CTAs read the same flags)
      int active = 0;
      for (int b = 0; b < B; ++b) active += __ldcg(a.fs.done + b) ? 0 : 1;
      if (active == 0) break;
    }
  }
}

static size_t mega_smem_bytes(const q3_model_desc& d, int B, int max_seq) {
  auto gemv = [](int T, int K, bool dual) {
    const int T8 = (T + 7) & ~7;
    size_t xs = (((size_t)T8 * (K + 32) * 2) + 127) & ~(size_t)127;
    return xs + (size_t)16 * (T8 / 8) * (dual ? 2 : 1) * 16 * 8 * 4;
  };
  size_t m = sizeof(SampleSmem);
  m = std::max(m, gemv(B, d.hidden, true));
  m = std::max(m, gemv(B, d.inter, false));
  m = std::max(m, gemv(B, d.heads * 128, false));
  m = std::max(m, gemv(2 * B, d.hidden, false));
  m = std::max(m, gemv(2 * B, d.cp_hidden, true));
  m = std::max(m, gemv(2 * B, d.cp_inter, false));
  m = std::max(m, gemv(2 * B, d.cp_heads * 128, false));
  m = std::max(m, (size_t)(2 * std::max(max_seq, d.cp_max_seq) + 256 + 16 * 2 * 128) * 4);
  return m;
}
