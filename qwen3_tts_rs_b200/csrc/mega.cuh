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
constexpr int MEGA_MAX_TILES = 4;   // 16-row tiles per CTA per phase (N <= 4 * 16 * gridDim.x rows per matrix)
constexpr int MEGA_SQ_STRIDE = 65;  // squares buffer [token][64 rows + 1]: conflict-free for row-wise writes and token-wise reads
constexpr int MEGA_MAX_OUT = (MEGA_MAX_TILES * 16 * MEGA_TMAX + 511) / 512;   // outputs per thread in the combine

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
  float *ssA, *ssB;      // per-CTA partial sums of squares [G][MEGA_TMAX] (post-attention sum / layer output)
  SampleArgs smp;
  int n_frames;          // loop iterations to run in this launch
  int do_cp, do_finish, do_talker, do_sample;
  const bf16* ext_step_input;   // per-op entry: talker input supplied by the caller (do_finish == 0)
  unsigned long long* prof;     // optional: %globaltimer stamps of block 0 (tools/profile_mega.py)
  int prof_cap;
  int bar_mode;                 // 0: release-reduction + acquire spin; 1: fence + atomic + volatile spin
  int prefetch_mode;            // 0: none, 1: own rows at phase start, 2: own + next phase's rows (bulk, thread 0),
                                // 3: next phase's rows only (bulk, thread 0, phase end), 4: next phase's rows, one
                                // prefetch.global.L2 per 128-byte line spread over all threads, right after the loads are issued
  int bench_barriers;           // > 0: run only this many grid barriers (micro-benchmark)
  int dbg;                      // unused
  int small_mode;               // 1: phases whose weights fit in registers use mega_gemv_small (Q3_SMALL=0 turns it off)
};

__device__ unsigned int g_prof_idx;
__shared__ unsigned int s_prof_idx;
__device__ __forceinline__ void prof_stamp(const MegaArgs& a, int tag) {
  if (a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned i = s_prof_idx++;
    if ((int)i < a.prof_cap) a.prof[i] = (t << 8) | (unsigned long long)(tag & 0xff);
    g_prof_idx = i + 1;
  }
}

// ---------------------------------------------------------------------------------------------------
struct GridBar {
  unsigned* ctr;
  unsigned epoch;
  int mode;
};
// Grid-wide barrier on a monotonically increasing counter (zeroed by the host before every launch; the
// launch is cooperative so all CTAs are resident).  Thread 0 publishes the CTA's writes with a release
// reduction and waits with acquire loads; bar.sync extends both to the rest of the CTA.  Data produced by
// OTHER CTAs is always read with ld.global.cg (L2), so a stale L1 line can never be observed.
__device__ __forceinline__ void grid_sync(GridBar& gb) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned target = (gb.epoch + 1u) * gridDim.x;
    if (gb.mode == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
      } while (v < target);
    } else {
      __threadfence();
      atomicAdd(gb.ctr, 1u);
      while (*((volatile unsigned*)gb.ctr) < target) {
      }
      __threadfence();
    }
  }
  gb.epoch += 1u;
  __syncthreads();
}
// Split form.  Every phase ENDS with `__syncthreads(); grid_arrive(gb);` and BEGINS with grid_wait(gb) -- placed after
// whatever the phase can do without the other CTAs' results (descriptor set-up, and in the skinny-GEMM phases the
// first weight and residual loads, which then travel while the CTA waits).  The kernel posts one arrive up front.
__device__ __forceinline__ void grid_arrive(GridBar& gb) {
  if (threadIdx.x == 0) {
    if (gb.mode == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
    } else {
      __threadfence();
      atomicAdd(gb.ctr, 1u);
    }
  }
}
__device__ __forceinline__ void grid_wait(GridBar& gb) {
  if (threadIdx.x == 0) {
    const unsigned target = (gb.epoch + 1u) * gridDim.x;
    if (gb.mode == 0) {
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
      } while (v < target);
    } else {
      while (*((volatile unsigned*)gb.ctr) < target) {
      }
      __threadfence();
    }
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
// skinny-GEMM phase descriptor
enum XMode { X_PLAIN = 0, X_NORM = 1, X_CP0 = 3, X_CPG = 4 };
enum MegaEpi { EPI_O_H1 = 16 };     // o_proj: h1 = bf16(bf16(acc) + x) -> Y, partial sum of squares of the unrounded sum

struct GemvP {
  const bf16* W;
  const bf16* W2;        // dual (SwiGLU) partner or null
  int N, K, T;
  int xmode;
  const bf16* X;         // [T][ldx]
  int ldx;
  const bf16* norm_w;    // X_NORM
  const float* ss_in;    // X_NORM: per-CTA partial sums of squares [G][MEGA_TMAX] of the rows of X, or null (computed here)
  bf16* xn_out;          // X_NORM: normalised rows written by CTA 0 (talker last_hidden)
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
  float* ss_out;         // partial sums of squares of the values written to Y, [G][MEGA_TMAX], or null
  // weights of the NEXT skinny-GEMM phase: pulled into L2 while this phase runs (they do not depend on
  // activations), so HBM keeps streaming across the grid barrier.
  const bf16* next_W;
  const bf16* next_W2;
  int next_N, next_K;
};

// Contiguous 8-row units dealt evenly to the CTAs: CTA c owns rows [r0, r1).
// c_grid_magic = ceil(2^24 / gridDim.x): exact quotient for units * gridDim.x < 2^24 (units <= 768 here).
__constant__ unsigned c_grid_magic;
__device__ __forceinline__ void mega_row_range(int N, int& r0, int& r1) {
  const int units = N >> 3, G = gridDim.x, c = blockIdx.x;
  const int base = (int)(((unsigned)units * c_grid_magic) >> 24), rem = units - base * G;   // units / G without a divide
  const int u0 = c * base + min(c, rem);
  r0 = u0 << 3;
  r1 = (u0 + base + (c < rem ? 1 : 0)) << 3;
}

__device__ __forceinline__ void l2_prefetch_bulk(const void* p, size_t bytes) {
  // cp.async.bulk.prefetch: one instruction pulls a contiguous range into L2 (TMA engine, no registers held)
  while (bytes > 0) {
    const unsigned chunk = (unsigned)(bytes > (size_t)(1u << 20) ? (1u << 20) : bytes);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(chunk) : "memory");
    p = reinterpret_cast<const char*>(p) + chunk;
    bytes -= chunk;
  }
}
__device__ __forceinline__ void mega_prefetch_rows(const bf16* W, const bf16* W2, int N, int K) {
  if (W == nullptr || threadIdx.x != 32) return;     // warp 1: thread 0 is busy with the barrier and the next descriptor
  int r0, r1;
  mega_row_range(N, r0, r1);
  if (r1 <= r0) return;
  l2_prefetch_bulk(W + (size_t)r0 * K, (size_t)(r1 - r0) * K * 2);
  if (W2 != nullptr) l2_prefetch_bulk(W2 + (size_t)r0 * K, (size_t)(r1 - r0) * K * 2);
}

constexpr int ATT_U = 8;   // cache rows per warp whose loads are in flight together (mega_attn, m2_attn)
constexpr int MEGA_MAX_LAYERS = 40;

// The first-generation kernel below (fence-based grid barriers) is history: it is compiled only into the development
// library (-DQ3_ALL_GENERATIONS, libq3tts_b200_dev.so) where the tests keep it against the oracle; the product library
// carries a stub so that the shared host code links, and q3_session_create refuses Q3_MEGA=1 there.
#ifdef Q3_ALL_GENERATIONS
// Cross-warp combine (fixed order), fused epilogue, partial sums of squares, barrier arrive: the common tail of
// the skinny-GEMM phases.  `red` holds the warps' partial sums as [nt][m][token col 0..7][warp][CTA-local row] f32
// (column stride 16*R + 4 floats, R = 16 * n_tiles): the combine reads 32 consecutive rows per warp and the
// fragment writes (lanes = 8 rows x 4 column pairs) hit 32 different banks -- both conflict-free.  (The first
// layout, [tile][warp][nt][m][row][col], made every combine read an 8-way bank conflict: ~1 us per phase.)
template <bool DUAL, int NT>
__device__ __forceinline__ void mega_gemv_tail(const MegaArgs& a, const GemvP& p, float* red, float* sqbuf,
                                               const unsigned short (&rraw)[MEGA_MAX_OUT], int r0, int r1, int n_tiles, int T,
                                               GridBar& gb) {
  constexpr int NM = DUAL ? 2 : 1;
  const int tid = threadIdx.x;
  const int red_r = n_tiles << 4, red_cs = 16 * red_r + 4;
  prof_stamp(a, 7);
  __syncthreads();
  prof_stamp(a, 8);
  // ---- fixed-order combine + epilogue for every (token, CTA-local row), spread over all 512 threads ----
#pragma unroll
  for (int it = 0; it < MEGA_MAX_OUT; ++it) {
    const int idx = tid + it * MEGA_THREADS;
    const int t = idx >> 6, rem = idx & 63;
    const int n = r0 + rem, nt = t >> 3, col = t & 7;
    const bool valid = t < T && n < r1;
    float sq = 0.f;             // contribution to the partial sum of squares of what the NEXT norm will see
    if (valid) {
      float v0 = 0.f, v1 = 0.f;
      const float* rb = red + (size_t)((nt * NM) * 8 + col) * red_cs + rem;
#pragma unroll
      for (int w = 0; w < MEGA_WARPS; ++w) {
        v0 += rb[w * red_r];
        if (DUAL) v1 += rb[8 * red_cs + w * red_r];
      }
      const float v = rbf(v0);
      switch (p.epi) {
        case EPI_STORE: p.Y[(size_t)t * p.ldy + n] = f2bf(v); break;
        case EPI_BIAS: {
          const float y = rbf(v + bf2f(p.bias[n]));
          p.Y[(size_t)t * p.ldy + n] = f2bf(y);
          sq = y * y;
        } break;
        case EPI_RESIDUAL: {
          const float y = rbf(__uint_as_float(((uint32_t)rraw[it]) << 16) + v);
          p.Y[(size_t)t * p.ldy + n] = f2bf(y);
          sq = y * y;             // the next layer's input_layernorm reads the stored (rounded) tensor
        } break;
        case EPI_O_H1: {
          const float su = __uint_as_float(((uint32_t)rraw[it]) << 16) + v;                                     // x + attn_out, un-rounded f32
          p.Y[(size_t)t * p.ldy + n] = f2bf(su);                            // h1: the rounded sum
          sq = su * su;           // fused_residual_rmsnorm.cu:60-65: sum of squares of the UN-rounded sum
        } break;
        case EPI_SWIGLU: {
          const float sg = rbf(silu_f(v));
          p.Y[(size_t)t * p.ldy + n] = f2bf(sg * rbf(v1));
        } break;
        case EPI_LOGITS: {
          if (p.Yf != nullptr) p.Yf[(size_t)t * p.N + n] = v;
          if (p.amax != nullptr) atomicMax(p.amax + t, argmax_key(v, n));
        } break;
        default: break;
      }
    }
    if (p.ss_out != nullptr && t < MEGA_TMAX) sqbuf[t * MEGA_SQ_STRIDE + rem] = sq;
  }
  if (p.ss_out != nullptr) {
    __syncthreads();
    const int w = tid >> 5, l = tid & 31;      // one warp per token sums its squares in a fixed order
    if (w < MEGA_TMAX) {
      float tot = 0.f;
      if (w < T) {
        const float s0 = l < red_r ? sqbuf[w * MEGA_SQ_STRIDE + l] : 0.f;
        const float s1 = l + 32 < red_r ? sqbuf[w * MEGA_SQ_STRIDE + l + 32] : 0.f;
        tot = warp_sum_xor(s0 + s1);
      }
      if (l == 0) p.ss_out[blockIdx.x * MEGA_TMAX + w] = tot;
    }
  }
  __syncthreads();
  grid_arrive(gb);
  prof_stamp(a, 3);
  if (a.prefetch_mode == 2 || a.prefetch_mode == 3) mega_prefetch_rows(p.next_W, p.next_W2, p.next_N, p.next_K);
}

// Same rows, one prefetch.global.L2 per 128-byte line, spread over the whole CTA (no thread is held up by the
// bulk-copy engine's issue cost).
__device__ __forceinline__ void mega_prefetch_lines(const bf16* W, const bf16* W2, int N, int K) {
  if (W == nullptr) return;
  int r0, r1;
  mega_row_range(N, r0, r1);
  if (r1 <= r0) return;
  const size_t bytes = (size_t)(r1 - r0) * K * 2;
  const char* b0 = reinterpret_cast<const char*>(W + (size_t)r0 * K);
  const char* b1 = W2 ? reinterpret_cast<const char*>(W2 + (size_t)r0 * K) : nullptr;
  for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)MEGA_THREADS * 128) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + o));
    if (b1) asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + o));
  }
}

// ---------------------------------------------------------------------------------------------------
// One skinny-GEMM phase.  No activation staging: every warp loads the B fragments of ITS k-slices straight
// from L2 (ld.global.cg, 16 B per lane) next to its weight loads, and applies the RMSNorm prologue
// xn = bf16((scale_t * x) * w_k) on the fly.  scale_t comes from per-CTA partial sums of squares that the
// PRODUCING phase left in `ss_in` (summed here in a fixed order -> deterministic), so no CTA ever re-reads
// whole activation rows.  The phase is instruction-issue bound at these sizes (16 warps x ~10^3 instructions
// per phase was 2 us of pure issue time), so the body is specialised at compile time on (DUAL, NT, NORM) and
// keeps integer divisions and predicated zero-fills off the common path.
// smem: scale[16] | sqbuf [64 rows][16] | red [tile][16 warps][NT][NM][16][8] f32
template <bool DUAL, int NT, bool NORM>
__device__ __noinline__ void mega_gemv_t(const MegaArgs& a, const GemvP& p, unsigned char* smem, GridBar& gb) {
  float* scale_s = reinterpret_cast<float*>(smem);
  float* sqbuf = scale_s + MEGA_TMAX;                       // [MEGA_MAX_TILES*16 rows][MEGA_TMAX] squares for ss_out
  float* red = sqbuf + MEGA_TMAX * MEGA_SQ_STRIDE;     // partial sums
  const int K = p.K, T = p.T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  int r0, r1;
  mega_row_range(p.N, r0, r1);
  if (r1 <= r0) {                                    // no rows here (block 0 always owns rows)
    grid_wait(gb);
    if (p.ss_out != nullptr && tid < MEGA_TMAX) p.ss_out[blockIdx.x * MEGA_TMAX + tid] = 0.f;
    __syncthreads();
    grid_arrive(gb);
    if (a.prefetch_mode == 2 || a.prefetch_mode == 3) mega_prefetch_rows(p.next_W, p.next_W2, p.next_N, p.next_K);
    if (a.prefetch_mode == 4) mega_prefetch_lines(p.next_W, p.next_W2, p.next_N, p.next_K);
    return;
  }
  if (a.prefetch_mode == 1 || a.prefetch_mode == 2) mega_prefetch_rows(p.W, p.W2, p.N, K);
  prof_stamp(a, 1);
  const bf16* xrow[NT];
  float xsc[NT];
  const bool write_xn = NORM && p.xn_out != nullptr && blockIdx.x == 0;
  const int ksteps = K >> 5;                         // 32 k per step; warp w owns k-steps w, w+16, w+32, ...
  constexpr int NM = DUAL ? 2 : 1;
  constexpr int JU = 2;                              // k-steps per chunk: all loads of a chunk are issued back to back
  const int jn = ksteps > warp ? (ksteps - warp + MEGA_WARPS - 1) / MEGA_WARPS : 0;   // this warp's k-steps
  const int n_chunks = max(1, ((ksteps + MEGA_WARPS - 1) / MEGA_WARPS + JU - 1) / JU);   // uniform over warps
  const int n_tiles = (r1 - r0 + 15) >> 4;
  const int red_r = n_tiles << 4, red_cs = 16 * red_r + 4;   // combine buffer strides (mega_gemv_tail)
  const int koff0 = warp * 32 + 8 * tg;              // element offset of this lane inside its first k-step
  const bool k_full = (ksteps % (MEGA_WARPS * JU)) == 0;   // every chunk of every warp is complete (all real models)

  // residual inputs of the epilogue, fetched now so their L2 latency hides behind the weight stream.
  // Combine mapping: output idx = tid + it*512 -> token t = idx >> 6, CTA-local row = idx & 63.
  unsigned short rraw[MEGA_MAX_OUT];                 // raw bf16 bits; converted where they are used
#pragma unroll
  for (int it = 0; it < MEGA_MAX_OUT; ++it) {
    rraw[it] = 0;
    const int idx = tid + it * MEGA_THREADS;
    const int t = idx >> 6, n = r0 + (idx & 63);
    if ((p.epi == EPI_RESIDUAL || p.epi == EPI_O_H1) && t < T && n < r1)
      rraw[it] = __ldcg(reinterpret_cast<const unsigned short*>(p.R + (size_t)t * p.ldr + n));
  }

  uint4 wl[NM][JU], wh[NM][JU], xv[NT][JU], wn[JU];
  auto load_w = [&](int tile, int c) {
    const int n0 = r0 + (tile << 4);
    const bool hi_ok = (n0 + 8) < r1;
    const size_t woff = (size_t)(n0 + g) * K + koff0 + (size_t)c * (JU * 512);
    if (k_full && hi_ok) {                           // common path: no predication, no zero fill
#pragma unroll
      for (int u = 0; u < JU; ++u) {
        wl[0][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W + woff + u * 512));
        wh[0][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W + woff + (size_t)8 * K + u * 512));
        if (DUAL) {
          wl[NM - 1][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W2 + woff + u * 512));
          wh[NM - 1][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W2 + woff + (size_t)8 * K + u * 512));
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < JU; ++u) {
        const bool ok = (c * JU + u) < jn;
        wl[0][u] = make_uint4(0, 0, 0, 0);
        wh[0][u] = make_uint4(0, 0, 0, 0);
        if (DUAL) { wl[NM - 1][u] = make_uint4(0, 0, 0, 0); wh[NM - 1][u] = make_uint4(0, 0, 0, 0); }
        if (ok) {
          wl[0][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W + woff + u * 512));
          if (hi_ok) wh[0][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W + woff + (size_t)8 * K + u * 512));
          if (DUAL) {
            wl[NM - 1][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W2 + woff + u * 512));
            if (hi_ok) wh[NM - 1][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W2 + woff + (size_t)8 * K + u * 512));
          }
        }
      }
    }
  };
  auto load_x = [&](int c) {
    const int xo = koff0 + c * (JU * 512);
#pragma unroll
    for (int u = 0; u < JU; ++u) {
      const bool ok = k_full || (c * JU + u) < jn;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        xv[nt][u] = make_uint4(0, 0, 0, 0);
        if (ok && xrow[nt] != nullptr) xv[nt][u] = ldcg16(xrow[nt] + xo + u * 512);
      }
      if (NORM) {
        wn[u] = make_uint4(0, 0, 0, 0);
        if (ok) wn[u] = *reinterpret_cast<const uint4*>(p.norm_w + xo + u * 512);
      }
    }
  };
  float acc[NM][NT][4];
  // ---- before the barrier: the first weights (and the residual rows above) do not depend on the previous phase ----
  load_w(0, 0);
  grid_wait(gb);
  prof_stamp(a, 2);
  // this lane's activation rows: token nt*8 + g of each n-tile
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int t = nt * 8 + g;
    xrow[nt] = nullptr;
    xsc[nt] = 0.f;
    if (t < T) {
      if (p.xmode == X_CP0) {
        const int b = t >> 1;
        xrow[nt] = (t & 1) ? p.emb + (size_t)__ldcg(a.fs.cur_tok + b) * K : a.fs.last_hidden + (size_t)b * K;
      } else if (p.xmode == X_CPG) {
        xrow[nt] = p.emb + (size_t)argmax_key_index(__ldcg(a.fs.amax + (size_t)(p.g - 1) * a.B + t)) * K;
      } else {
        xrow[nt] = p.X + (size_t)t * p.ldx;
      }
    }
  }
  if (blockIdx.x == 0) {
    if (p.xmode == X_CPG && tid < T)
      a.fs.frame_codes[tid * 16 + p.g] = argmax_key_index(__ldcg(a.fs.amax + (size_t)(p.g - 1) * a.B + tid));
    if (p.xmode == X_CP0 && tid < a.B) a.fs.frame_codes[tid * 16] = __ldcg(a.fs.cur_tok + tid);
  }
  load_x(0);
  if (a.prefetch_mode == 4) mega_prefetch_lines(p.next_W, p.next_W2, p.next_N, p.next_K);
  if (NORM) {
    // row scales, computed while the first chunk's loads are in flight: one warp per token
    for (int t = warp; t < T; t += MEGA_WARPS) {
      float tot = 0.f;
      if (p.ss_in != nullptr) {
        float part[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int c = lane + 32 * i;
          part[i] = c < (int)gridDim.x ? __ldcg(p.ss_in + c * MEGA_TMAX + t) : 0.f;
        }
        tot = ((part[0] + part[1]) + (part[2] + part[3])) + part[4];
        for (int c = lane + 160; c < (int)gridDim.x; c += 32) tot += __ldcg(p.ss_in + c * MEGA_TMAX + t);
      } else {
        const bf16* xr = p.X + (size_t)t * p.ldx;
        for (int c = 8 * lane; c < K; c += 256) {
          float f[8];
          unpack8(ldcg16(xr + c), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) tot = fmaf(f[e], f[e], tot);
        }
      }
      tot = warp_sum_xor(tot);
      if (lane == 0) scale_s[t] = ref_mean_rsqrt(tot, K, a.eps);
    }
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int t = nt * 8 + g;
      if (t < T) xsc[nt] = scale_s[t];
    }
  }
  prof_stamp(a, 6);
  for (int tile = 0; tile < n_tiles; ++tile) {
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
#pragma unroll
      for (int u = 0; u < JU; ++u) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint4 x4 = xv[nt][u];
          if (NORM) {
            float f[8], w[8];
            unpack8(x4, f);
            unpack8(wn[u], w);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = (xsc[nt] * f[e]) * w[e];
            x4 = pack8(f);
            if (write_xn && tile == 0 && xrow[nt] != nullptr && (c * JU + u) < jn)
              *reinterpret_cast<uint4*>(p.xn_out + (size_t)(nt * 8 + g) * K + koff0 + (c * JU + u) * 512) = x4;
          }
#pragma unroll
          for (int m = 0; m < NM; ++m) {
            mma_bf16_16816(acc[m][nt], wl[m][u].x, wh[m][u].x, wl[m][u].y, wh[m][u].y, x4.x, x4.y);
            mma_bf16_16816(acc[m][nt], wl[m][u].z, wh[m][u].z, wl[m][u].w, wh[m][u].w, x4.z, x4.w);
          }
        }
      }
      // the registers are free again: put the next chunk (possibly of the next tile) in flight right away, so it
      // overlaps the cross-warp combine below
      if (c + 1 < n_chunks) { load_w(tile, c + 1); load_x(c + 1); }
      else if (tile + 1 < n_tiles) { load_w(tile + 1, 0); load_x(0); }
    }
    // this tile's partial sums -> shared memory (combined for all tiles at once below)
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float* r = red + (size_t)((nt * NM + m) * 8 + 2 * tg) * red_cs + warp * red_r + (tile << 4) + g;
        r[0] = acc[m][nt][0]; r[red_cs] = acc[m][nt][1]; r[8] = acc[m][nt][2]; r[red_cs + 8] = acc[m][nt][3];
      }
  }
  mega_gemv_tail<DUAL, NT>(a, p, red, sqbuf, rraw, r0, r1, n_tiles, T, gb);
}


// ---------------------------------------------------------------------------------------------------
// Small phases (every code-predictor phase, and the 0.6B talker): K == CHUNKS * 1024 and at most TILES 16-row tiles
// per CTA, so ALL of a lane's weight fragments fit in registers (<= 16 x 128 bit).  They are requested before the
// grid barrier; after it only the activations (L2) are fetched once, and every MMA of the phase follows without
// another memory round trip.  Same fragment mapping, split-K and combine as mega_gemv_t -> same bits.
template <bool DUAL, int NT, bool NORM, int TILES, int CHUNKS>
__device__ __noinline__ void mega_gemv_small(const MegaArgs& a, const GemvP& p, unsigned char* smem, GridBar& gb) {
  float* scale_s = reinterpret_cast<float*>(smem);
  float* sqbuf = scale_s + MEGA_TMAX;
  float* red = sqbuf + MEGA_TMAX * MEGA_SQ_STRIDE;
  const int K = p.K, T = p.T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  constexpr int NM = DUAL ? 2 : 1;
  constexpr int JU = 2;
  int r0, r1;
  mega_row_range(p.N, r0, r1);
  if (r1 <= r0) {
    grid_wait(gb);
    if (p.ss_out != nullptr && tid < MEGA_TMAX) p.ss_out[blockIdx.x * MEGA_TMAX + tid] = 0.f;
    __syncthreads();
    grid_arrive(gb);
    if (a.prefetch_mode == 2 || a.prefetch_mode == 3) mega_prefetch_rows(p.next_W, p.next_W2, p.next_N, p.next_K);
    return;
  }
  prof_stamp(a, 1);
  const int n_tiles = (r1 - r0 + 15) >> 4;
  const int red_r = n_tiles << 4, red_cs = 16 * red_r + 4;
  const int koff0 = warp * 32 + 8 * tg;
  unsigned short rraw[MEGA_MAX_OUT];
#pragma unroll
  for (int it = 0; it < MEGA_MAX_OUT; ++it) {
    rraw[it] = 0;
    const int idx = tid + it * MEGA_THREADS;
    const int t = idx >> 6, n = r0 + (idx & 63);
    if ((p.epi == EPI_RESIDUAL || p.epi == EPI_O_H1) && t < T && n < r1)
      rraw[it] = __ldcg(reinterpret_cast<const unsigned short*>(p.R + (size_t)t * p.ldr + n));
  }
  // ---- before the barrier: every weight fragment of the phase ----
  uint4 wl[TILES][CHUNKS][NM][JU], wh[TILES][CHUNKS][NM][JU];
#pragma unroll
  for (int tile = 0; tile < TILES; ++tile) {
    const int n0 = r0 + (tile << 4);
    const bool lo_ok = tile < n_tiles, hi_ok = lo_ok && (n0 + 8) < r1;
    const size_t woff = (size_t)(n0 + g) * K + koff0;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
      for (int u = 0; u < JU; ++u) {
        const size_t o = woff + (size_t)c * (JU * 512) + u * 512;
        wl[tile][c][0][u] = make_uint4(0, 0, 0, 0);
        wh[tile][c][0][u] = make_uint4(0, 0, 0, 0);
        if (lo_ok) wl[tile][c][0][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W + o));
        if (hi_ok) wh[tile][c][0][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W + o + (size_t)8 * K));
        if (DUAL) {
          wl[tile][c][NM - 1][u] = make_uint4(0, 0, 0, 0);
          wh[tile][c][NM - 1][u] = make_uint4(0, 0, 0, 0);
          if (lo_ok) wl[tile][c][NM - 1][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W2 + o));
          if (hi_ok) wh[tile][c][NM - 1][u] = ldg_stream(reinterpret_cast<const uint4*>(p.W2 + o + (size_t)8 * K));
        }
      }
  }
  grid_wait(gb);
  prof_stamp(a, 2);
  // ---- after the barrier: activations (once), scales ----
  const bf16* xrow[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int t = nt * 8 + g;
    xrow[nt] = nullptr;
    if (t < T) {
      if (p.xmode == X_CP0) {
        const int b = t >> 1;
        xrow[nt] = (t & 1) ? p.emb + (size_t)__ldcg(a.fs.cur_tok + b) * K : a.fs.last_hidden + (size_t)b * K;
      } else if (p.xmode == X_CPG) {
        xrow[nt] = p.emb + (size_t)argmax_key_index(__ldcg(a.fs.amax + (size_t)(p.g - 1) * a.B + t)) * K;
      } else {
        xrow[nt] = p.X + (size_t)t * p.ldx;
      }
    }
  }
  if (blockIdx.x == 0) {
    if (p.xmode == X_CPG && tid < T)
      a.fs.frame_codes[tid * 16 + p.g] = argmax_key_index(__ldcg(a.fs.amax + (size_t)(p.g - 1) * a.B + tid));
    if (p.xmode == X_CP0 && tid < a.B) a.fs.frame_codes[tid * 16] = __ldcg(a.fs.cur_tok + tid);
  }
  uint4 xv[CHUNKS][NT][JU];
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
    for (int u = 0; u < JU; ++u)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        xv[c][nt][u] = make_uint4(0, 0, 0, 0);
        if (xrow[nt] != nullptr) xv[c][nt][u] = ldcg16(xrow[nt] + koff0 + c * (JU * 512) + u * 512);
      }
  if (NORM) {
    for (int t = warp; t < T; t += MEGA_WARPS) {
      float tot = 0.f;
      if (p.ss_in != nullptr) {
        float part[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int c = lane + 32 * i;
          part[i] = c < (int)gridDim.x ? __ldcg(p.ss_in + c * MEGA_TMAX + t) : 0.f;
        }
        tot = ((part[0] + part[1]) + (part[2] + part[3])) + part[4];
        for (int c = lane + 160; c < (int)gridDim.x; c += 32) tot += __ldcg(p.ss_in + c * MEGA_TMAX + t);
      } else {
        const bf16* xr = p.X + (size_t)t * p.ldx;
        for (int c = 8 * lane; c < K; c += 256) {
          float f[8];
          unpack8(ldcg16(xr + c), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) tot = fmaf(f[e], f[e], tot);
        }
      }
      tot = warp_sum_xor(tot);
      if (lane == 0) scale_s[t] = ref_mean_rsqrt(tot, K, a.eps);
    }
    __syncthreads();
    const bool write_xn = p.xn_out != nullptr && blockIdx.x == 0;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int t = nt * 8 + g;
      const float sc = t < T ? scale_s[t] : 0.f;
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
        for (int u = 0; u < JU; ++u) {
          const int xo = koff0 + c * (JU * 512) + u * 512;
          float f[8], w[8];
          unpack8(xv[c][nt][u], f);
          unpack8(*reinterpret_cast<const uint4*>(p.norm_w + xo), w);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = (sc * f[e]) * w[e];
          xv[c][nt][u] = pack8(f);
          if (write_xn && xrow[nt] != nullptr) *reinterpret_cast<uint4*>(p.xn_out + (size_t)t * K + xo) = xv[c][nt][u];
        }
    }
  }
  prof_stamp(a, 6);
#pragma unroll
  for (int tile = 0; tile < TILES; ++tile) {
    if (tile < n_tiles) {
      float acc[NM][NT][4];
#pragma unroll
      for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
        for (int u = 0; u < JU; ++u)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int m = 0; m < NM; ++m) {
              const uint4 x4 = xv[c][nt][u];
              mma_bf16_16816(acc[m][nt], wl[tile][c][m][u].x, wh[tile][c][m][u].x, wl[tile][c][m][u].y, wh[tile][c][m][u].y, x4.x, x4.y);
              mma_bf16_16816(acc[m][nt], wl[tile][c][m][u].z, wh[tile][c][m][u].z, wl[tile][c][m][u].w, wh[tile][c][m][u].w, x4.z, x4.w);
            }
#pragma unroll
      for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          float* r = red + (size_t)((nt * NM + m) * 8 + 2 * tg) * red_cs + warp * red_r + (tile << 4) + g;
          r[0] = acc[m][nt][0]; r[red_cs] = acc[m][nt][1]; r[8] = acc[m][nt][2]; r[red_cs + 8] = acc[m][nt][3];
        }
    }
  }
  mega_gemv_tail<DUAL, NT>(a, p, red, sqbuf, rraw, r0, r1, n_tiles, T, gb);
}

template <bool DUAL>
__device__ __forceinline__ void mega_gemv(const MegaArgs& a, const GemvP& p, unsigned char* smem, GridBar& gb) {
  const bool norm = p.xmode == X_NORM;
  const bool nt1 = p.T <= 8;
  // widest CTA of the phase (uniform over the grid): units per CTA -> tiles
  const int units = p.N >> 3, base = (int)(((unsigned)units * c_grid_magic) >> 24);
  const int max_tiles = (base + (units - base * (int)gridDim.x > 0 ? 1 : 0) + 1) >> 1;
  if (a.small_mode != 0) {
    if (DUAL) {
      if (norm && p.K == 1024 && max_tiles <= 2) {
        if (nt1) mega_gemv_small<DUAL, 1, true, 2, 1>(a, p, smem, gb);
        else mega_gemv_small<DUAL, 2, true, 2, 1>(a, p, smem, gb);
        return;
      }
    } else {
      if (norm && p.K == 1024 && max_tiles <= 2) {
        if (nt1) mega_gemv_small<false, 1, true, 2, 1>(a, p, smem, gb);
        else mega_gemv_small<false, 2, true, 2, 1>(a, p, smem, gb);
        return;
      }
      if (!norm && p.K == 2048 && max_tiles <= 1) {
        if (nt1) mega_gemv_small<false, 1, false, 1, 2>(a, p, smem, gb);
        else mega_gemv_small<false, 2, false, 1, 2>(a, p, smem, gb);
        return;
      }
      if (!norm && p.K == 3072 && max_tiles <= 1) {
        if (nt1) mega_gemv_small<false, 1, false, 1, 3>(a, p, smem, gb);
        else mega_gemv_small<false, 2, false, 1, 3>(a, p, smem, gb);
        return;
      }
    }
  }
  if (nt1) {
    if (norm) mega_gemv_t<DUAL, 1, true>(a, p, smem, gb);
    else mega_gemv_t<DUAL, 1, false>(a, p, smem, gb);
  } else {
    if (norm) mega_gemv_t<DUAL, 2, true>(a, p, smem, gb);
    else mega_gemv_t<DUAL, 2, false>(a, p, smem, gb);
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
  bf16* kcur = reinterpret_cast<bf16*>(red + 16 * 2 * 128);   // [128] this token's rotated K row
  bf16* vcur = kcur + 128;                                     // [128] this token's V row
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nh = p.heads + 2 * p.kv_heads;
  const float scale = rbf(0.08838834764831845f);
  for (int item = blockIdx.x; item < p.B * p.kv_heads; item += gridDim.x) {
    const int b = item / p.kv_heads, kvh = item - b * p.kv_heads;
    bf16* kbase = p.k_cache + ((size_t)b * p.kv_heads + kvh) * p.max_seq * 128;
    bf16* vbase = p.v_cache + ((size_t)b * p.kv_heads + kvh) * p.max_seq * 128;
    for (int s = 0; s < p.S; ++s) {
      const int t = b * p.S + s;
      const int pos = (p.pos_base ? __ldcg(p.pos_base + b) : 0) + p.pos_add + s;
      const int L = pos + 1;
      // Rows j < pos are already in the cache: start this warp's first K and V loads now, so they travel
      // together with the qkv loads below instead of one L2 round trip after another.
      uint2 kpre = make_uint2(0u, 0u), vpre = make_uint2(0u, 0u);
      if (warp < pos) {
        kpre = __ldcg(reinterpret_cast<const uint2*>(kbase + (size_t)warp * 128 + 4 * lane));
        vpre = __ldcg(reinterpret_cast<const uint2*>(vbase + (size_t)warp * 128 + 4 * lane));
      }
      // warps 0,1: q heads 2kvh, 2kvh+1; warp 2: k; warp 3: v
      if (warp < 4) {
        const int hh = warp < 2 ? 2 * kvh + warp : (warp == 2 ? p.heads + kvh : p.heads + p.kv_heads + kvh);
        const unsigned short* src = reinterpret_cast<const unsigned short*>(p.qkv + (size_t)t * nh * 128 + (size_t)hh * 128);
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(((uint32_t)__ldcg(src + lane + 32 * i)) << 16);
        if (warp == 3) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            vbase[(size_t)pos * 128 + lane + 32 * i] = f2bf(v[i]);
            vcur[lane + 32 * i] = f2bf(v[i]);
          }
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
            for (int i = 0; i < 4; ++i) {
              kbase[(size_t)pos * 128 + lane + 32 * i] = f2bf(o[i]);
              kcur[lane + 32 * i] = f2bf(o[i]);
            }
          }
        }
      }
      __syncthreads();      // q, kcur, vcur in smem
      float q0[4], q1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { q0[i] = qs[4 * lane + i]; q1[i] = qs[128 + 4 * lane + i]; }
      // row j of K or V: the prefetched register (first iteration), the cache (j < pos) or this token (j == pos)
      auto row = [&](const bf16* base, const bf16* cur, const uint2& pre, int j) -> uint2 {
        if (j == pos) return *reinterpret_cast<const uint2*>(cur + 4 * lane);
        if (j == warp) return pre;
        return __ldcg(reinterpret_cast<const uint2*>(base + (size_t)j * 128 + 4 * lane));
      };
      for (int j0 = warp; j0 < L; j0 += ATT_U * MEGA_WARPS) {
        uint2 ku[ATT_U];
#pragma unroll
        for (int q = 0; q < ATT_U; ++q) {
          const int j = j0 + q * MEGA_WARPS;
          ku[q] = j < L ? row(kbase, kcur, kpre, j) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < ATT_U; ++q) {
        const int j = j0 + q * MEGA_WARPS;
        if (j >= L) break;
        const uint2 u = ku[q];
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
      for (int j0 = warp; j0 < L; j0 += ATT_U * MEGA_WARPS) {
        uint2 vu[ATT_U];
#pragma unroll
        for (int q = 0; q < ATT_U; ++q) {
          const int j = j0 + q * MEGA_WARPS;
          vu[q] = j < L ? row(vbase, vcur, vpre, j) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < ATT_U; ++q) {
        const int j = j0 + q * MEGA_WARPS;
        if (j >= L) break;
        const uint2 u = vu[q];
        const float v0 = bf_lo(u.x), v1 = bf_hi(u.x), v2 = bf_lo(u.y), v3 = bf_hi(u.y);
        const float p0 = sc0[j], p1 = sc1[j];
        o0[0] = fmaf(p0, v0, o0[0]); o0[1] = fmaf(p0, v1, o0[1]); o0[2] = fmaf(p0, v2, o0[2]); o0[3] = fmaf(p0, v3, o0[3]);
        o1[0] = fmaf(p1, v0, o1[0]); o1[1] = fmaf(p1, v1, o1[1]); o1[2] = fmaf(p1, v2, o1[2]); o1[3] = fmaf(p1, v3, o1[3]);
        }
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
// Phase descriptors live in SHARED memory: thread 0 fills them, everybody reads them (broadcast).  Building
// them per thread on the stack would cost ~150 B of local-memory traffic per thread per phase (6 GB of DRAM
// writes per frame at 75 k threads x 546 phases -- measured with ncu before this change).
struct MegaShared {
  LayerW layers[MEGA_MAX_LAYERS];   // talker layers then code-predictor layers (pointer tables read by every descriptor fill)
  MegaArgs a;
  GemvP gp;
  AttnP ap;
  uint32_t codes[16];
};

#define MEGA_FILL_BEGIN(sh) if (threadIdx.x == 0) { GemvP& q = (sh).gp; q = GemvP{};
#define MEGA_FILL_END() } __syncthreads();

// One decoder stack over T = B*S tokens, in place on a.x  (DecoderLayer::forward x layers).
__device__ __noinline__ void mega_layers(MegaShared& sh, const MegaStack& st, int T, int S, const int* pos_base, int pos_add,
                                         bf16* kc, bf16* vc, int cache_seq, const bf16* cos_tab, const bf16* sin_tab,
                                         unsigned char* smem, const float* first_ss, const bf16* after_W, int after_N,
                                         int after_K, GridBar& gb) {
  const MegaArgs& a = sh.a;
  const int nh = st.heads + 2 * st.kv_heads;
  const size_t layer_stride = (size_t)a.B * st.kv_heads * cache_seq * 128;
  for (int l = 0; l < st.n_layers; ++l) {
    const LayerW* wp = st.layers + l;
    // P1: rms_norm(x) -> [q;k;v]
    MEGA_FILL_BEGIN(sh)
      q.W = wp->wqkv; q.N = nh * 128; q.K = st.H; q.T = T; q.xmode = X_NORM; q.X = a.x; q.ldx = st.H; q.norm_w = wp->in_ln;
      q.ss_in = l == 0 ? first_ss : a.ssB;
      q.epi = EPI_STORE; q.Y = a.qkv; q.ldy = nh * 128;
      q.next_W = wp->wo; q.next_N = st.H; q.next_K = st.heads * 128;
    MEGA_FILL_END()
    mega_gemv<false>(a, sh.gp, smem, gb);
    // P2: QK-norm, RoPE, KV append, attention
    if (threadIdx.x == 0) {
      AttnP& at = sh.ap;
      at.qkv = a.qkv; at.out = a.attn; at.k_cache = kc + l * layer_stride; at.v_cache = vc + l * layer_stride;
      at.q_norm_w = wp->q_norm; at.k_norm_w = wp->k_norm; at.cos_tab = cos_tab; at.sin_tab = sin_tab; at.pos_base = pos_base;
      at.pos_add = pos_add; at.S = S; at.B = a.B; at.heads = st.heads; at.kv_heads = st.kv_heads; at.max_seq = cache_seq;
    }
    grid_wait(gb);
    prof_stamp(a, 4);
    mega_attn(a, sh.ap, smem);
    __syncthreads();
    grid_arrive(gb);
    prof_stamp(a, 5);
    // P3: o_proj + residual: h1 = bf16(x + attn_out), partial sum of squares of the un-rounded sum
    MEGA_FILL_BEGIN(sh)
      q.W = wp->wo; q.N = st.H; q.K = st.heads * 128; q.T = T; q.xmode = X_PLAIN; q.X = a.attn; q.ldx = st.heads * 128;
      q.epi = EPI_O_H1; q.R = a.x; q.ldr = st.H; q.Y = a.h1; q.ldy = st.H; q.ss_out = a.ssA;
      q.next_W = wp->gate; q.next_W2 = wp->up; q.next_N = st.I; q.next_K = st.H;
    MEGA_FILL_END()
    mega_gemv<false>(a, sh.gp, smem, gb);
    // P4: post-attention RMSNorm (scale from P3's partials, applied to the rounded h1) -> SwiGLU(gate, up)
    MEGA_FILL_BEGIN(sh)
      q.W = wp->gate; q.W2 = wp->up; q.N = st.I; q.K = st.H; q.T = T; q.xmode = X_NORM; q.X = a.h1; q.ldx = st.H;
      q.norm_w = wp->post_ln; q.ss_in = a.ssA; q.epi = EPI_SWIGLU; q.Y = a.act; q.ldy = st.I;
      q.next_W = wp->down; q.next_N = st.H; q.next_K = st.I;
    MEGA_FILL_END()
    mega_gemv<true>(a, sh.gp, smem, gb);
    // P5: down_proj + residual -> x
    MEGA_FILL_BEGIN(sh)
      q.W = wp->down; q.N = st.H; q.K = st.I; q.T = T; q.xmode = X_PLAIN; q.X = a.act; q.ldx = st.I;
      q.epi = EPI_RESIDUAL; q.R = a.h1; q.ldr = st.H; q.Y = a.x; q.ldy = st.H; q.ss_out = a.ssB;
      if (l + 1 < st.n_layers) {
        q.next_W = st.layers[l + 1].wqkv; q.next_N = nh * 128; q.next_K = st.H;
      } else {
        q.next_W = after_W; q.next_N = after_N; q.next_K = after_K;
      }
    MEGA_FILL_END()
    mega_gemv<false>(a, sh.gp, smem, gb);
  }
}

__device__ __noinline__ void mega_sample(const SampleArgs& sa, int b, SampleSmem& sm) { sample_row_body(sa, b, sm); }
__device__ __noinline__ void mega_finish(const MegaArgs& a, uint32_t* s_codes) {
  EmbTable tab{};
  for (int i = 0; i < a.n_ac; ++i) tab.e[i] = a.cp_emb[i];
  for (int b = blockIdx.x; b < a.B; b += gridDim.x)
    frame_finish_row(a.fs, tab, a.codec_emb, a.step_input, a.H, a.B, a.n_ac, b, s_codes);
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_frames_mega_kernel(const MegaArgs args) {
  extern __shared__ __align__(128) unsigned char mega_smem[];
  __shared__ MegaShared sh;
  {
    // kernel parameters -> shared memory (taking their address would otherwise copy them to every thread's stack)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&args);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.a);
    for (int i = threadIdx.x; i < (int)(sizeof(MegaArgs) / 4); i += MEGA_THREADS) dst[i] = src[i];
    if (threadIdx.x == 0) s_prof_idx = g_prof_idx;
    const int nl = args.tk.n_layers + args.cp.n_layers;
    const uint32_t* l0 = reinterpret_cast<const uint32_t*>(args.tk.layers);
    const uint32_t* l1 = reinterpret_cast<const uint32_t*>(args.cp.layers);
    uint32_t* ld = reinterpret_cast<uint32_t*>(sh.layers);
    constexpr int W4 = (int)(sizeof(LayerW) / 4);
    for (int i = threadIdx.x; i < nl * W4; i += MEGA_THREADS)
      ld[i] = i < args.tk.n_layers * W4 ? l0[i] : l1[i - args.tk.n_layers * W4];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sh.a.tk.layers = sh.layers;
    sh.a.cp.layers = sh.layers + sh.a.tk.n_layers;
  }
  __syncthreads();
  const MegaArgs& a = sh.a;
  GridBar gb{a.bar, 0u, a.bar_mode};
  const int B = a.B;
  if (a.bench_barriers > 0) {
    for (int i = 0; i < a.bench_barriers; ++i) grid_sync(gb);
    return;
  }
  grid_arrive(gb);                 // every phase waits for its predecessor's arrive; this is the first phase's
  for (int frame = 0; frame < a.n_frames; ++frame) {
    {
      // frame prologue (its own phase): the previous frame's sampler is complete from here on
      grid_wait(gb);
      if (frame > 0 && a.do_sample) {
        // stop early once every row has sampled EOS (uniform decision: all CTAs read the same flags)
        int active = 0;
        for (int b = 0; b < B; ++b) active += __ldcg(a.fs.done + b) ? 0 : 1;
        if (active == 0) break;
      }
      if (a.do_cp && blockIdx.x == 0)
        for (int i = threadIdx.x; i < a.n_ac * B; i += MEGA_THREADS) a.fs.amax[i] = 0ull;
      __syncthreads();
      grid_arrive(gb);
    }
    if (a.do_cp) {
      // ---- code predictor: 15 dependent passes (code_predictor.rs:320-416) ----
      const int nh_cp = (a.cp.heads + 2 * a.cp.kv_heads) * 128;
      for (int g = 0; g < a.n_ac; ++g) {
        const int T = g == 0 ? 2 * B : B, S = g == 0 ? 2 : 1;
        MEGA_FILL_BEGIN(sh)
          q.N = a.C; q.K = a.H; q.T = T; q.xmode = g == 0 ? X_CP0 : X_CPG; q.g = g;
          q.emb = g == 0 ? a.codec_emb : a.cp_emb[g - 1];
          q.W = a.cp_proj_w; q.bias = a.cp_proj_b; q.epi = EPI_BIAS; q.Y = a.x; q.ldy = a.C; q.ss_out = a.ssB;
          q.next_W = a.cp.layers[0].wqkv; q.next_N = nh_cp; q.next_K = a.C;
        MEGA_FILL_END()
        if (a.cp_proj_w) {
          mega_gemv<false>(a, sh.gp, mega_smem, gb);
        } else {
          grid_wait(gb);
          if (blockIdx.x == 0) {
          // no projection (talker hidden == CP hidden): the gathered rows are the layer input
          const int K8 = a.H >> 3;
          for (int i = threadIdx.x; i < T * K8; i += MEGA_THREADS) {
            const int t = i / K8, qq = i - t * K8;
            const bf16* src;
            if (g == 0) {
              const int b = t >> 1;
              src = (t & 1) ? a.codec_emb + (size_t)__ldcg(a.fs.cur_tok + b) * a.H : a.fs.last_hidden + (size_t)b * a.H;
            } else {
              src = a.cp_emb[g - 1] + (size_t)argmax_key_index(__ldcg(a.fs.amax + (size_t)(g - 1) * B + t)) * a.H;
            }
            *reinterpret_cast<uint4*>(a.x + (size_t)t * a.C + qq * 8) = ldcg16(src + qq * 8);
          }
          if (g == 0) { if (threadIdx.x < B) a.fs.frame_codes[threadIdx.x * 16] = __ldcg(a.fs.cur_tok + threadIdx.x); }
          else if (threadIdx.x < B)
            a.fs.frame_codes[threadIdx.x * 16 + g] = argmax_key_index(__ldcg(a.fs.amax + (size_t)(g - 1) * B + threadIdx.x));
          }
          __syncthreads();
          grid_arrive(gb);
        }
        mega_layers(sh, a.cp, T, S, nullptr, g == 0 ? 0 : g + 1, a.cp_k, a.cp_v, a.cp_max_seq, a.cp_cos, a.cp_sin, mega_smem,
                    a.cp_proj_w ? a.ssB : nullptr, a.cp_head[g], a.cpV, a.C, gb);
        MEGA_FILL_BEGIN(sh)
          q.W = a.cp_head[g]; q.N = a.cpV; q.K = a.C; q.T = B; q.xmode = X_NORM; q.norm_w = a.cp_norm; q.ss_in = g == 0 ? nullptr : a.ssB;
          q.X = g == 0 ? a.x + a.C : a.x; q.ldx = g == 0 ? 2 * a.C : a.C;
          q.epi = EPI_LOGITS; q.amax = a.fs.amax + (size_t)g * B;
          q.Yf = a.cp_logits ? a.cp_logits + (size_t)g * B * a.cpV : nullptr;
          if (g + 1 < a.n_ac) {
            if (a.cp_proj_w) { q.next_W = a.cp_proj_w; q.next_N = a.C; q.next_K = a.H; }
            else { q.next_W = a.cp.layers[0].wqkv; q.next_N = nh_cp; q.next_K = a.C; }
          } else if (a.do_talker) {
            q.next_W = a.tk.layers[0].wqkv; q.next_N = (a.tk.heads + 2 * a.tk.kv_heads) * 128; q.next_K = a.H;
          }
        MEGA_FILL_END()
        mega_gemv<false>(a, sh.gp, mega_smem, gb);
      }
    }
    if (a.do_finish) {
      // ---- emit the frame, build the talker input (lib.rs:605-622) ----
      grid_wait(gb);
      mega_finish(a, sh.codes);
      __syncthreads();
      grid_arrive(gb);
    }
    if (a.do_talker) {
      // ---- talker step (talker.rs:716-736) ----
      const bf16* in = a.do_finish ? a.step_input : a.ext_step_input;
      grid_wait(gb);
      for (int i = blockIdx.x * MEGA_THREADS + threadIdx.x; i < B * (a.H >> 3); i += gridDim.x * MEGA_THREADS)
        reinterpret_cast<uint4*>(a.x)[i] = ldcg16(reinterpret_cast<const uint4*>(in) + i);
      __syncthreads();
      grid_arrive(gb);
      mega_layers(sh, a.tk, B, 1, a.fs.offset, 0, a.tk_k, a.tk_v, a.max_seq, a.t_cos, a.t_sin, mega_smem, nullptr,
                  a.codec_head, a.V, a.H, gb);
      MEGA_FILL_BEGIN(sh)
        q.W = a.codec_head; q.N = a.V; q.K = a.H; q.T = B; q.xmode = X_NORM; q.norm_w = a.t_norm; q.X = a.x; q.ldx = a.H;
        q.ss_in = a.ssB;
        q.xn_out = a.fs.last_hidden; q.epi = EPI_LOGITS; q.Yf = a.logits;
        if (a.do_cp && frame + 1 < a.n_frames) {       // the next frame starts with the CP projection (or its first layer)
          if (a.cp_proj_w) { q.next_W = a.cp_proj_w; q.next_N = a.C; q.next_K = a.H; }
          else { q.next_W = a.cp.layers[0].wqkv; q.next_N = (a.cp.heads + 2 * a.cp.kv_heads) * 128; q.next_K = a.C; }
        }
      MEGA_FILL_END()
      mega_gemv<false>(a, sh.gp, mega_smem, gb);
    }
    if (a.do_sample) {
      // ---- penalties + sampling + state update (lib.rs:639-651) ----
      SampleSmem& sm = *reinterpret_cast<SampleSmem*>(mega_smem);
      grid_wait(gb);
      for (int b = blockIdx.x; b < B; b += gridDim.x) mega_sample(a.smp, b, sm);
      __syncthreads();
      grid_arrive(gb);
    }
  }
}

#else
__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_frames_mega_kernel(const MegaArgs) {}
#endif

// Shared memory of the persistent kernel for a given model / batch / grid; returns 0 when a phase does not fit
// the kernel's static limits (then the multi-kernel path is used).
static size_t mega_smem_bytes(const q3_model_desc& d, int B, int max_seq, int grid) {
  size_t red_max = 0;
  bool ok = true;
  auto phase = [&](int N, int T, bool dual) {
    const int units = N / 8, per_cta = (units + grid - 1) / grid, tiles = (per_cta + 1) / 2;
    if (tiles > MEGA_MAX_TILES || T > MEGA_TMAX || N % 8 != 0) ok = false;
    const int NT = (T + 7) / 8;
    red_max = std::max(red_max, (size_t)NT * (dual ? 2 : 1) * 8 * (256 * tiles + 4) * 4);
  };
  const int nh = (d.heads + 2 * d.kv_heads) * 128, cnh = (d.cp_heads + 2 * d.cp_kv_heads) * 128;
  phase(nh, B, false); phase(d.hidden, B, false); phase(d.inter, B, true); phase(d.codec_vocab, B, false);
  phase(d.cp_hidden, 2 * B, false); phase(cnh, 2 * B, false); phase(d.cp_inter, 2 * B, true); phase(d.cp_vocab, B, false);
  if (!ok) return 0;
  size_t m = sizeof(SampleSmem);
  m = std::max(m, (size_t)(MEGA_TMAX + MEGA_TMAX * MEGA_SQ_STRIDE) * 4 + red_max);
  m = std::max(m, (size_t)(2 * std::max(max_seq, d.cp_max_seq) + 256 + 16 * 2 * 128) * 4 + 2 * 128 * 2);
  return m;
}
