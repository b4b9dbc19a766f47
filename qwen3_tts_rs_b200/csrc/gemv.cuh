// Skinny GEMM ("GEMV with a few right-hand sides") for the decode step: Y[t][n] = sum_k W[n][k] X[t][k]
// with T = batch rows (1..32) and W streamed from HBM exactly once per token tile.  HBM-bound
// (arithmetic intensity <= T FLOP/B against a ridge of ~259): the design goals are 128-bit coalesced
// weight loads with many bytes in flight per SM and fused prologue (RMSNorm) / epilogues (bias, SiLU,
// SwiGLU, residual, f32 logits, packed arg-max) so that no activation round-trips HBM between ops.
//
// Replaces, per call site: candle Linear::forward -> cuBLAS GEMV (transformer.rs:258-260, 371,
// 408-413; talker.rs:733; code_predictor.rs:341-345, 374, 407) plus the elementwise candle ops
// that follow it.  Rounding points are the reference's: the matmul result is rounded to bf16 before
// any further arithmetic, and every following candle op rounds again.
#pragma once
#include "common.cuh"
#include "norm.cuh"

enum GemvPrologue { PRO_NONE = 0, PRO_RMSNORM = 1 };
enum GemvEpilogue {
  EPI_STORE = 0,       // Y = bf16(acc)
  EPI_BIAS = 1,        // Y = bf16(bf16(acc) + bias)
  EPI_BIAS_SILU = 2,   // Y = bf16(silu(bf16(bf16(acc) + bias)))
  EPI_RESIDUAL = 3,    // Y = bf16(R + bf16(acc))
  EPI_SWIGLU = 4,      // Y = bf16(bf16(silu(bf16(acc_gate))) * bf16(acc_up))   (DUAL kernels)
  EPI_LOGITS = 5,      // Yf = f32(bf16(acc)); optional packed arg-max
};

struct GemvArgs {
  const bf16* W;        // [N][K]
  const bf16* W2;       // [N][K] second matrix for DUAL (up_proj)
  const bf16* X;        // [T][ldx]
  int ldx;
  const bf16* norm_w;   // PRO_RMSNORM weight [K]
  float eps;
  bf16* xn_out;         // optional: normalised X, [T][K], written by block 0
  int N, K, T;
  int pro, epi;
  bf16* Y;              // [T][ldy]
  int ldy;
  const bf16* bias;     // [N]
  const bf16* R;        // [T][ldr]
  int ldr;
  float* Yf;            // [T][N]  (EPI_LOGITS; may be null when only the arg-max is wanted)
  unsigned long long* amax;   // [T] packed (ordered logit << 32 | ~n) arg-max keys, or null
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ unsigned long long argmax_key(float v, int n) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);                 // monotone float -> uint
  return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);  // ties: lowest n wins
}
__device__ __host__ __forceinline__ uint32_t argmax_key_index(unsigned long long key) {
  return 0xffffffffu - (uint32_t)(key & 0xffffffffull);
}

// Stage one token tile of X into shared memory as bf16, applying the prologue.
// xs: [TB][K] bf16.  256 threads.
template <int TB>
__device__ __forceinline__ void gemv_stage_x(const GemvArgs& a, int t0, bf16* xs, float* s_part) {
  const int K = a.K, K8 = K >> 3;
  const int tid = threadIdx.x;
  uint4* xs4 = reinterpret_cast<uint4*>(xs);
  if (a.pro == PRO_NONE) {
    for (int i = tid; i < TB * K8; i += blockDim.x) {
      int t = i / K8, q = i - t * K8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (t0 + t < a.T) v = *reinterpret_cast<const uint4*>(a.X + (size_t)(t0 + t) * a.ldx + q * 8);
      xs4[i] = v;
    }
    return;
  }
  // PRO_RMSNORM: xn = bf16((scale * x) * w), scale from the reference-order sum of squares
  if (K >= 1024) {
    const int grp = tid >> 7, g = tid & 127;                 // two 128-thread groups, one token each
    for (int tt = grp; tt < TB; tt += 2) {                   // uniform trip count per group
      const bool valid = (t0 + tt) < a.T;
      const bf16* xrow = a.X + (size_t)(valid ? (t0 + tt) : 0) * a.ldx;
      float p[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) p[e] = 0.f;
      for (int c0 = 0; c0 < K; c0 += 1024) {
        int c = c0 + 8 * g;
        if (c < K) {
          float f[8];
          unpack8(*reinterpret_cast<const uint4*>(xrow + c), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) p[e] = fmaf(f[e], f[e], p[e]);
        }
      }
      float tot = sumsq_ref_large_finish(p, g, s_part + grp * 32, 1 + grp);
      float sc = ref_mean_rsqrt(tot, K, a.eps);
      for (int c = 8 * g; c < K; c += 1024) {
        float f[8], w[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(xrow + c), f);
        unpack8(*reinterpret_cast<const uint4*>(a.norm_w + c), w);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = valid ? (sc * f[e]) * w[e] : 0.f;
        uint4 pk = pack8(o);
        xs4[tt * K8 + (c >> 3)] = pk;
        if (a.xn_out != nullptr && blockIdx.x == 0 && valid)
          *reinterpret_cast<uint4*>(a.xn_out + (size_t)(t0 + tt) * K + c) = pk;
      }
    }
  } else {
    const int warp = tid >> 5, lane = tid & 31;
    for (int tt = warp; tt < TB; tt += 8) {
      const bool valid = (t0 + tt) < a.T;
      const bf16* xrow = a.X + (size_t)(valid ? (t0 + tt) : 0) * a.ldx;
      float tot = sumsq_ref_small(K, [&](int c) { return bf2f(xrow[c]); });
      float sc = ref_mean_rsqrt(tot, K, a.eps);
      for (int c = lane; c < K; c += 32) {
        bf16 o = f2bf(valid ? (sc * bf2f(xrow[c])) * bf2f(a.norm_w[c]) : 0.f);
        xs[tt * K + c] = o;
        if (a.xn_out != nullptr && blockIdx.x == 0 && valid) a.xn_out[(size_t)(t0 + tt) * K + c] = o;
      }
    }
  }
}

// TB: tokens per tile; RPW: output rows per warp; DUAL: gate+up pair (SwiGLU).
template <int TB, int RPW, bool DUAL>
__global__ void __launch_bounds__(256) gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);
  __shared__ float s_part[64];
  constexpr int NW = DUAL ? 2 * RPW : RPW;
  const int K8 = a.K >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_block = 8 * RPW;
  const int nblocks = (a.N + rows_per_block - 1) / rows_per_block;
  const uint4* xs4 = reinterpret_cast<const uint4*>(xs);

  for (int t0 = 0; t0 < a.T; t0 += TB) {
    __syncthreads();                       // previous tile's readers are done with xs
    gemv_stage_x<TB>(a, t0, xs, s_part);
    __syncthreads();
    for (int rb = blockIdx.x; rb < nblocks; rb += gridDim.x) {
      const int n0 = rb * rows_per_block + warp * RPW;
      if (n0 >= a.N) continue;
      float acc[NW][TB];
#pragma unroll
      for (int r = 0; r < NW; ++r)
#pragma unroll
        for (int t = 0; t < TB; ++t) acc[r][t] = 0.f;
      const uint4* wp[NW];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        int n = min(n0 + r, a.N - 1);
        wp[r] = reinterpret_cast<const uint4*>(a.W + (size_t)n * a.K);
        if (DUAL) wp[RPW + r] = reinterpret_cast<const uint4*>(a.W2 + (size_t)n * a.K);
      }
      // software pipeline: the next chunk's weights are in flight while this one is consumed
      uint4 wcur[NW], wnext[NW];
      int kc = lane;
      if (kc < K8) {
#pragma unroll
        for (int r = 0; r < NW; ++r) wcur[r] = ldg_stream(wp[r] + kc);
      }
      for (; kc < K8; kc += 32) {
        const int kn = kc + 32;
        if (kn < K8) {
#pragma unroll
          for (int r = 0; r < NW; ++r) wnext[r] = ldg_stream(wp[r] + kn);
        }
        float wf[NW][8];
#pragma unroll
        for (int r = 0; r < NW; ++r) unpack8(wcur[r], wf[r]);
#pragma unroll
        for (int t = 0; t < TB; ++t) {
          float xf[8];
          unpack8(xs4[t * K8 + kc], xf);
#pragma unroll
          for (int r = 0; r < NW; ++r)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[r][t] = fmaf(wf[r][e], xf[e], acc[r][t]);
        }
#pragma unroll
        for (int r = 0; r < NW; ++r) wcur[r] = wnext[r];
      }
      // warp reduction; afterwards every lane holds every total
#pragma unroll
      for (int r = 0; r < NW; ++r)
#pragma unroll
        for (int t = 0; t < TB; ++t) acc[r][t] = warp_sum_xor(acc[r][t]);
      // lane l < RPW*TB finishes output (r = l / TB, t = l % TB)
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int t = 0; t < TB; ++t)
          if (lane == r * TB + t) {
            v0 = acc[r][t];
            if (DUAL) v1 = acc[RPW + r][t];
          }
      if (lane < RPW * TB) {
        const int r = lane / TB, t = t0 + (lane % TB), n = n0 + r;
        if (n < a.N && t < a.T) {
          float v = rbf(v0);
          switch (a.epi) {
            case EPI_STORE: a.Y[(size_t)t * a.ldy + n] = f2bf(v); break;
            case EPI_BIAS: a.Y[(size_t)t * a.ldy + n] = f2bf(v + bf2f(a.bias[n])); break;
            case EPI_BIAS_SILU: {
              float y = rbf(v + bf2f(a.bias[n]));
              a.Y[(size_t)t * a.ldy + n] = f2bf(silu_f(y));
            } break;
            case EPI_RESIDUAL: a.Y[(size_t)t * a.ldy + n] = f2bf(bf2f(a.R[(size_t)t * a.ldr + n]) + v); break;
            case EPI_SWIGLU: {
              float s = rbf(silu_f(v));
              a.Y[(size_t)t * a.ldy + n] = f2bf(s * rbf(v1));
            } break;
            case EPI_LOGITS: {
              if (a.Yf != nullptr) a.Yf[(size_t)t * a.N + n] = v;
              if (a.amax != nullptr) atomicMax(a.amax + t, argmax_key(v, n));
            } break;
          }
        }
      }
    }
  }
}

// ---- host launcher -------------------------------------------------------------------------------
template <int TB, int RPW, bool DUAL>
static void gemv_launch_inst(const GemvArgs& a, int grid, cudaStream_t st) {
  size_t smem = (size_t)TB * a.K * sizeof(bf16);
  static size_t configured = 0;
  if (smem > 32 * 1024 && smem > configured) {       // static shared memory counts against the 48 KB default too
    Q3_CHECK_CUDA(cudaFuncSetAttribute(gemv_kernel<TB, RPW, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(200 * 1024)));
    configured = 200 * 1024;
  }
  gemv_kernel<TB, RPW, DUAL><<<grid, 256, smem, st>>>(a);
  Q3_COUNT_LAUNCH();
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
      throw Q3Error(Q3_ERR_CUDA, std::string("gemv launch failed: ") + cudaGetErrorString(e) + " (TB " + std::to_string(TB) + ", RPW " +
                                     std::to_string(RPW) + ", N " + std::to_string(a.N) + ", K " + std::to_string(a.K) + ", T " +
                                     std::to_string(a.T) + ", grid " + std::to_string(grid) + ", smem " + std::to_string(smem) + ")");
  }
}

template <int TB, bool DUAL>
static void gemv_launch_tb(const GemvArgs& a, int num_sms, cudaStream_t st) {
  // rows per warp: largest that still gives every SM at least ~2 blocks
  int rpw = DUAL ? 2 : 4;
  while (rpw > 1 && ceil_div(a.N, 8 * rpw) < 2 * num_sms) rpw >>= 1;
  int grid = ceil_div(a.N, 8 * rpw);
  if (rpw == 4) {
    if constexpr (!DUAL) gemv_launch_inst<TB, 4, false>(a, grid, st);
  } else if (rpw == 2) {
    gemv_launch_inst<TB, 2, DUAL>(a, grid, st);
  } else {
    gemv_launch_inst<TB, 1, DUAL>(a, grid, st);
  }
}

static void gemv_launch(const GemvArgs& a, int num_sms, cudaStream_t st) {
  Q3_REQUIRE(a.K % 8 == 0, Q3_ERR_INVALID, "gemv: K must be a multiple of 8");
  Q3_REQUIRE(a.ldx % 8 == 0, Q3_ERR_INVALID, "gemv: ldx must be a multiple of 8");
  const bool dual = a.epi == EPI_SWIGLU;
  int tb = a.T >= 8 ? 8 : (a.T > 2 ? 4 : (a.T == 2 ? 2 : 1));
  // keep the X tile within the shared-memory budget
  while (tb > 1 && (size_t)tb * a.K * sizeof(bf16) > 160 * 1024) tb >>= 1;
#define Q3_GEMV_TB(TBV)                                  \
  if (dual) gemv_launch_tb<TBV, true>(a, num_sms, st);   \
  else gemv_launch_tb<TBV, false>(a, num_sms, st);
  switch (tb) {
    case 8: Q3_GEMV_TB(8); break;
    case 4: Q3_GEMV_TB(4); break;
    case 2: Q3_GEMV_TB(2); break;
    default: Q3_GEMV_TB(1); break;
  }
#undef Q3_GEMV_TB
}
