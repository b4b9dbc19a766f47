// Persistent frame kernel, second generation: DATAFLOW phases.
//
// Same work as mega.cuh (one cooperative launch runs whole decode frames: 15 code-predictor passes, the
// 16-way embedding sum, the 28-layer talker step, the codec head and the sampler, generate_codes
// src/lib.rs:530-656), same skinny-GEMM formulation (weights loaded straight from global memory into
// mma.sync A fragments, split-K over the 16 warps of a CTA, fixed-order combine), but the ~550 dependent
// phases of a frame are no longer separated by fence-based grid barriers:
//
//   * every activation that crosses CTAs lives in a TAGGED buffer: 8-byte slots {payload, tag} where the
//     payload is two bf16 values (or one f32 for the post-attention sum h1, whose un-rounded value the
//     reference's fused residual+RMSNorm kernel needs, kernels/fused_residual_rmsnorm.cu:60-65) and the tag
//     is the number of the phase that wrote it.  A slot is written and read with single 64-bit relaxed
//     accesses, so a reader that sees the expected tag also sees the payload -- no release fence on the
//     producer side (a MEMBAR.GPU costs more than the arithmetic of a code-predictor phase) and no
//     separate "barrier, then load" round trip on the consumer side;
//   * the grid-wide counter survives as a fence-free HINT (relaxed add / relaxed poll by one thread) that
//     tells a CTA when polling the data is worth it; correctness rests on the tags alone.  Phases whose
//     outputs are not tagged (logits, arg-max keys, sampler state, the emitted codes) keep a real
//     release/acquire barrier: ~20 of the 552 phases of a frame;
//   * the consumer of an RMSNorm phase computes the row scale itself from the fragments it loaded (quad
//     shuffle + one shared-memory exchange), so the per-CTA partial-sum arrays and their extra
//     synchronisation are gone;
//   * the phase program (552 descriptors per frame) is built once on the host and copied into shared
//     memory at kernel start: no per-phase descriptor construction by a single thread;
//   * buffer reuse is safe without barriers because every skinny-GEMM phase consumes the FULL activation
//     vector of the phase before it: a CTA that starts writing in phase j has seen the tags of every
//     producer of phase j-1, each of which had finished all of its reads of phases < j-1 (program order),
//     and every buffer is rewritten at the earliest 5 phases after it was last read.
#pragma once
#include "mega.cuh"

typedef unsigned long long u64;

// Profiling hooks (phase stamps of block 0, per-CTA arrival stamps, tag re-read counters) are compiled in only with
// -DQ3_PROF=1 (the tools' library, libq3tts_b200_prof.so).  Every phase function of the persistent kernels runs ONCE per
// phase and the ~10 functions of a frame evict each other from the instruction caches (ncu: 111 instruction-cache misses
// per phase per SM = the whole function body), so code that is never executed in production still costs fetch bandwidth.
#ifdef Q3_PROF
#define M2_PROF_ENABLED true
#else
#define M2_PROF_ENABLED false
#endif

enum M2Kind { M2_GEMV = 0, M2_ATTN = 1, M2_PROLOGUE = 2, M2_GATHER = 3, M2_FINISH = 4, M2_COPYIN = 5, M2_SAMPLE = 6 };
enum M2Fmt { XF_BF16T = 0, XF_F32T = 1, XF_GATHER = 2, XF_NONE = 3 };
enum M2Flags { PF_WAIT_ACQ = 1, PF_ARRIVE_REL = 2, PF_DUAL = 4, PF_NORM = 8, PF_CP = 16, PF_CP0 = 32, PF_RING = 64 };

struct alignas(16) M2Phase {   // 160 bytes
  const bf16* W;        // GEMV: weights [N][K]; ATTN: K cache of the layer
  const bf16* W2;       // GEMV: dual partner; ATTN: V cache of the layer
  const void* X;        // tagged input (row of token t at X + t*ldx elements) ; COPYIN: plain bf16 [B][H]
  void* Y;              // tagged output
  const void* R;        // tagged residual (rows owned by this CTA)
  const bf16* aux;      // NORM: norm weight; EPI_BIAS: bias; ATTN: q_norm
  float* Yf;            // EPI_LOGITS: f32 logits or null
  u64* amax;            // EPI_LOGITS: packed arg-max keys [T] or null
  bf16* xn_out;         // NORM: normalised rows written by CTA 0 (talker last_hidden) or null
  const bf16* aux2;     // XF_GATHER: embedding table; ATTN: k_norm
  int N, K, T, ldx;     // ATTN: N = heads, K = kv_heads, ldx = (heads + 2 kv) * 128
  int ldy, ldr, g, kind;
  int flags, epi, xf, small;   // small: (tiles << 4) | chunks of the register-resident variant, 0 = streaming variant
  int pos_add, S, rf, yf;
  int next_gemv;        // index of the skinny-GEMM phase whose weight rows this phase prefetches into L2, or -1 (host-resolved)
  int pad_[3];
};
static_assert(sizeof(M2Phase) == 160, "M2Phase layout");

struct M2Args {
  const M2Phase* prog;
  int n_ph, n_frames;
  int B, H, n_ac;
  float eps;
  FrameState fs;
  const bf16 *cp_cos, *cp_sin, *t_cos, *t_sin;
  int max_seq, cp_max_seq;
  const bf16* codec_emb;
  const bf16* cp_emb[15];
  bf16* step_input;
  SampleArgs smp;
  unsigned* bar;
  unsigned* tag_ctr;    // device word: last tag used by earlier launches of this session
  int* err;             // mapped host word: non-zero when a watchdog fired
  unsigned long long* prof;
  int prof_cap;
  int do_sample;
  int prof_mode;        // 2: prof = [n_ph][grid] arrival stamps of the launch's second frame, prof[n_ph*grid] = tag re-reads
  int prefetch;         // 1: pull the next GEMV phase's weight rows into L2 at the end of every phase
  int bench_barriers;
  int ring_shift;       // ring kernel: log2 of the number of ring stages in use (<= 3)
  int pf_sleep;         // ring kernel: nanoseconds the producer warp sleeps between polls of a full ring
  int m4_slots;         // mega4.cuh: ring slots in use
  int m4_red2;          // mega4.cuh: the combine buffer is double-buffered over tiles
  u64* xchg;            // m2_attn_units: exchange slots of the splits of a long row, [rows x kv_heads x split_ns][M2_XCHG_SLOTS]
  int split_min_l;      // > 0: talker attention of this launch runs m2_attn_units, rows with >= split_min_l positions are split
  int split_ns;         // ... over this many CTAs (<= M2_SPLIT_NS_MAX)
};

// ---------------------------------------------------------------------------------------------------
// tagged slots
// Slot reads are plain L2 loads (ld.global.cg: no L1, so a re-read always goes back to L2).  A strong load
// (ld.relaxed.gpu -> LDG.E.STRONG.GPU) is served at the line's home L2 partition on this two-die part and measured
// 2-3x slower for the all-CTAs-read-the-same-lines pattern of the activation vectors; the tag makes a stale or
// torn-between-slots read harmless (it is simply repeated), and each 64-bit slot is written by ONE 64-bit store.
// HARDWARE-SPECIFIC ASSUMPTION: under the PTX memory model a weak load racing with another CTA's store is a data race;
// what this relies on is that sm_100 performs an aligned 8-byte (and each half of a 16-byte) global access as a single
// copy, so a reader sees a slot's {payload, tag} pair entirely old or entirely new.  -DM2_STRONG_LOADS builds the
// model-conformant variant (ld.relaxed.gpu); the repeat test of tests/test_gpu_parity.py (150 identical frames, bit-equal
// logits) and the per-frame follow-mode comparison are the stress tests of the default.
#ifndef M2_STRONG_LOADS
__device__ __forceinline__ void ld_slot2(const void* p, u64& a, u64& b) {
  asm volatile("ld.global.cg.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ u64 ld_slot(const void* p) {
  u64 a;
  asm volatile("ld.global.cg.b64 %0, [%1];" : "=l"(a) : "l"(p));
  return a;
}
#else
__device__ __forceinline__ void ld_slot2(const void* p, u64& a, u64& b) {
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ u64 ld_slot(const void* p) {
  u64 a;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(a) : "l"(p));
  return a;
}
#endif
__device__ __forceinline__ void st_slot(void* p, uint32_t payload, uint32_t tag) {
  const u64 v = ((u64)tag << 32) | (u64)payload;
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_slot2(void* p, uint32_t p0, uint32_t p1, uint32_t tag) {
  const u64 v0 = ((u64)tag << 32) | (u64)p0, v1 = ((u64)tag << 32) | (u64)p1;
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(v0), "l"(v1) : "memory");
}
__device__ __forceinline__ uint32_t slot_tag(u64 s) { return (uint32_t)(s >> 32); }
__device__ __forceinline__ uint32_t slot_val(u64 s) { return (uint32_t)s; }

// Barrier of the 512 compute threads (named barrier 1).  In the dataflow kernel these are all the threads of the
// CTA; in the TMA-ring kernel (mega3.cuh) a 17th warp streams weights and does not take part.
__device__ __forceinline__ void m2_csync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

constexpr unsigned M2_SPIN_LIMIT = 1u << 22;     // polls before the watchdog gives up (~1-2 s)
constexpr unsigned M2_RETRY_LIMIT = 1u << 20;    // tag re-reads before the watchdog gives up

struct M2Sync {
  unsigned* ctr;
  int* err;
  unsigned epoch, G;
  bool dead;
  unsigned long long* arr;   // profiling (prof_mode 2): arrival time of every CTA for every phase of one frame, or null
  unsigned* retries;         // profiling: tag re-read counter, or null
};
__device__ __forceinline__ void m2_fail(M2Sync& gs, int code) {
  if (gs.err != nullptr) atomicCAS(gs.err, 0, code);
  gs.dead = true;
}
// The phase functions take the sync state BY VALUE and return (epoch | dead << 32): passed by reference to a
// non-inlined function it lived in local memory, and thread 0 paid a chain of LDL/STL round trips per phase.
__device__ __forceinline__ unsigned long long m2_pack(const M2Sync& gs) {
  return (unsigned long long)gs.epoch | ((unsigned long long)(gs.dead ? 1u : 0u) << 32);
}
__device__ __forceinline__ void m2_unpack(M2Sync& gs, unsigned long long r) {
  gs.epoch = (unsigned)r;
  gs.dead = (r >> 32) != 0ull;
}
// Every CTA executes one wait and one arrive per phase.  Relaxed form = hint only (tags carry the data
// dependence); acquire/release form = real grid barrier for phases with untagged inputs/outputs.
__device__ __forceinline__ void m2_wait(M2Sync& gs, int flags) {
  if (threadIdx.x == 0 && !gs.dead) {
    const unsigned target = (gs.epoch + 1u) * gs.G;
    unsigned v, it = 0;
    if (flags & PF_WAIT_ACQ) {
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gs.ctr) : "memory");
        if (++it > M2_SPIN_LIMIT) { m2_fail(gs, 1000000 + (int)gs.epoch); break; }
      } while (v < target);
    } else {
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gs.ctr));
        if (++it > M2_SPIN_LIMIT) { m2_fail(gs, 2000000 + (int)gs.epoch); break; }
      } while (v < target);
    }
  }
  gs.epoch += 1u;
  m2_csync();
}
// callers make sure every thread's stores of the phase were issued (a __syncthreads) before this
__device__ __forceinline__ void m2_arrive(M2Sync& gs, int flags) {
  if (threadIdx.x == 0) {
    if (M2_PROF_ENABLED && gs.arr != nullptr) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      gs.arr[3 * gs.G + blockIdx.x] = t;
    }
    if (flags & PF_ARRIVE_REL) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gs.ctr) : "memory");
    else asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(gs.ctr) : "memory");
  }
}

// profiling (prof_mode 2): stamp k (0 wait passed, 1 activations ready, 2 MMA loop done, 3 arrive) of this CTA
__device__ __forceinline__ void m2_stamp(M2Sync& gs, int k) {
  if (M2_PROF_ENABLED && gs.arr != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    gs.arr[k * gs.G + blockIdx.x] = t;
  }
}
__device__ unsigned int g_prof2_idx;
__shared__ unsigned int s_prof2_idx;
__device__ __forceinline__ void prof2(const M2Args& a, int tag) {
  if (M2_PROF_ENABLED && a.prof != nullptr && a.prof_mode != 2 && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned i = s_prof2_idx++;
    if ((int)i < a.prof_cap) a.prof[i] = (t << 8) | (unsigned long long)(tag & 0xff);
    g_prof2_idx = i + 1;
  }
}

// next GEMV phase's weight rows of this CTA -> L2 (they do not depend on activations)
// a.prefetch 1 (default): one thread (lane 0 of warp 1) requests the whole range; 2 (Q3_PF_SPLIT=1): lane 0 of each of the
// 16 warps requests one sixteenth -- an experiment: a single cp.async.bulk INTO SHARED MEMORY of 50-130 KB holds its issuing
// thread for more than a microsecond (mega5.cuh), the L2 prefetch form does not (the split is 3 % slower).
__device__ __forceinline__ void m2_prefetch(const M2Args& a, const M2Phase* nx) {
  if (a.prefetch == 0 || nx == nullptr) return;
  if (a.prefetch == 1 ? threadIdx.x != 32 : (threadIdx.x & 31) != 0) return;
  int r0, r1;
  mega_row_range(nx->N, r0, r1);
  if (r1 <= r0) return;
  size_t bytes = (size_t)(r1 - r0) * nx->K * 2, off = 0;
  if (a.prefetch != 1) { bytes /= MEGA_WARPS; off = (size_t)(threadIdx.x >> 5) * bytes; }     // multiples of 64 bytes
  l2_prefetch_bulk(reinterpret_cast<const char*>(nx->W + (size_t)r0 * nx->K) + off, bytes);
  if (nx->W2 != nullptr) l2_prefetch_bulk(reinterpret_cast<const char*>(nx->W2 + (size_t)r0 * nx->K) + off, bytes);
}

// ---------------------------------------------------------------------------------------------------
// 8 consecutive elements [k, k+8) of a token row -> packed bf16x8 (the B fragments of two MMAs).
// sq accumulates the squares the RMSNorm of the consuming phase is defined on: the bf16 values for
// XF_BF16T, the UN-rounded f32 sums for XF_F32T.  bad collects tag mismatches.
template <int XF>
__device__ __forceinline__ uint4 m2_load_x8(const char* row, int k, uint32_t xtag, uint32_t& bad, float& sq) {
  uint4 r;
  if constexpr (XF == XF_BF16T) {
    u64 s0, s1, s2, s3;
    ld_slot2(row + (size_t)k * 4, s0, s1);
    ld_slot2(row + (size_t)k * 4 + 16, s2, s3);
    bad |= (slot_tag(s0) ^ xtag) | (slot_tag(s1) ^ xtag) | (slot_tag(s2) ^ xtag) | (slot_tag(s3) ^ xtag);
    r = make_uint4(slot_val(s0), slot_val(s1), slot_val(s2), slot_val(s3));
    float f[8];
    unpack8(r, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) sq = fmaf(f[e], f[e], sq);
  } else if constexpr (XF == XF_F32T) {
    u64 s[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld_slot2(row + (size_t)k * 8 + 16 * i, s[2 * i], s[2 * i + 1]);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      bad |= slot_tag(s[e]) ^ xtag;
      f[e] = __uint_as_float(slot_val(s[e]));
      sq = fmaf(f[e], f[e], sq);
    }
    r = pack8(f);
  } else {
    r = ldcg16(row + (size_t)k * 2);
  }
  return r;
}
template <int XF>
__device__ __forceinline__ size_t m2_row_bytes(int ldx) {
  return XF == XF_BF16T ? (size_t)ldx * 4 : (XF == XF_F32T ? (size_t)ldx * 8 : (size_t)ldx * 2);
}

// token rows of this lane (token nt*8 + g of each n-tile) + the code bookkeeping of the gather modes
template <int NT, int XF>
__device__ __forceinline__ void m2_token_rows(const M2Args& a, const M2Phase& p, int g, const char* (&xrow)[NT]) {
  const int T = p.T, K = p.K;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int t = nt * 8 + g;
    xrow[nt] = nullptr;
    if (t < T) {
      if constexpr (XF == XF_GATHER) {
        if (p.flags & PF_CP0) {
          const int b = (t >> 1) + p.pos_add;          // pos_add: first batch row of this pass-0 group (batch > 8)
          xrow[nt] = reinterpret_cast<const char*>((t & 1) ? p.aux2 + (size_t)__ldcg(a.fs.cur_tok + b) * K
                                                           : a.fs.last_hidden + (size_t)b * K);
        } else {
          xrow[nt] = reinterpret_cast<const char*>(
              p.aux2 + (size_t)argmax_key_index(__ldcg(a.fs.amax + (size_t)(p.g - 1) * a.B + t)) * K);
        }
      } else {
        xrow[nt] = reinterpret_cast<const char*>(p.X) + (size_t)t * m2_row_bytes<XF>(p.ldx);
      }
    }
  }
  if constexpr (XF == XF_GATHER) {
    if (blockIdx.x == 0) {
      const int tid = threadIdx.x;
      if (!(p.flags & PF_CP0) && tid < T)
        a.fs.frame_codes[tid * 16 + p.g] = argmax_key_index(__ldcg(a.fs.amax + (size_t)(p.g - 1) * a.B + tid));
      if ((p.flags & PF_CP0) && tid < (T >> 1)) a.fs.frame_codes[(tid + p.pos_add) * 16] = __ldcg(a.fs.cur_tok + tid + p.pos_add);
    }
  }
}

// residual inputs of the epilogue: rows owned by this CTA, written by this CTA two or three phases ago.
// Combine mapping: output idx = tid + it*512 -> token t = idx >> 6, CTA-local row = idx & 63.
template <int NT>
__device__ __forceinline__ void m2_load_residual(const M2Phase& p, int r0, int r1, float (&rres)[MEGA_MAX_OUT]) {
  const int tid = threadIdx.x;
  const int epi = p.epi, T = p.T, rf = p.rf, ldr = p.ldr;
  const u64* const R64 = reinterpret_cast<const u64*>(p.R);
  const bool has_r = epi == EPI_RESIDUAL || epi == EPI_O_H1;
#pragma unroll
  for (int it = 0; it < MEGA_MAX_OUT; ++it) {
    rres[it] = 0.f;
    if (NT == 1 && it >= 1) continue;          // one token tile: the second output iteration does not exist (m2_tail)
    const int idx = tid + it * MEGA_THREADS;
    const int t = idx >> 6, n = r0 + (idx & 63);
    if (has_r && t < T && n < r1) {
      if (rf == XF_F32T) {
        const u64 s = ld_slot(R64 + (size_t)t * ldr + n);
        rres[it] = rbf(__uint_as_float(slot_val(s)));            // h1 as the reference stores it: bf16(x + attn)
      } else {
        const u64 s = ld_slot(R64 + (((size_t)t * ldr + n) >> 1));
        rres[it] = (n & 1) ? bf_hi(slot_val(s)) : bf_lo(slot_val(s));
      }
    }
  }
}

// Cross-warp combine (fixed order), fused epilogue, tagged stores, arrive: the common tail of the GEMV phases.
// red: [nt][m][token col 0..7][warp][CTA-local row] f32, column stride 16*R + 4 floats (see mega.cuh).
template <bool DUAL, int NT>
__device__ __forceinline__ void m2_tail(const M2Args& a, const M2Phase& p, const M2Phase* nx, float* red,
                                        const float (&rres)[MEGA_MAX_OUT], int r0, int r1, int n_tiles, M2Sync& gs,
                                        const uint32_t tag) {
  constexpr int NM = DUAL ? 2 : 1;
  const int tid = threadIdx.x, lane = tid & 31, T = p.T;
  const bool prof_on = M2_PROF_ENABLED && a.prof != nullptr;
  const int red_r = n_tiles << 4, red_cs = 16 * red_r + 4;
  if (prof_on) m2_stamp(gs, 2);
  m2_csync();
  if (prof_on) prof2(a, 4);
  // descriptor fields used below, read once (the slot stores are asm volatile with a memory clobber, so the compiler
  // would re-read them from shared memory in every iteration)
  const int epi = p.epi, yf = p.yf, ldy = p.ldy, pN = p.N;
  u64* const Y64 = reinterpret_cast<u64*>(p.Y);
  float* const Yf = p.Yf;
  u64* const amax = p.amax;
  const bf16* const bias = p.aux;
  // idx = tid + it * 512 -> token idx >> 6: with one token tile (T <= 8) the second iteration has no valid output
  constexpr int N_IT = NT == 1 ? 1 : MEGA_MAX_OUT;
#pragma unroll
  for (int it = 0; it < N_IT; ++it) {
    const int idx = tid + it * MEGA_THREADS;
    const int t = idx >> 6, rem = idx & 63;
    const int n = r0 + rem, nt = t >> 3, col = t & 7;
    const bool valid = t < T && n < r1 && nt < NT;
    float outv = 0.f;
    u64 key = 0ull;
    if (valid) {
      float v0 = 0.f, v1 = 0.f;
      const float* rb = red + (size_t)((nt * NM) * 8 + col) * red_cs + rem;
#pragma unroll
      for (int w = 0; w < MEGA_WARPS; ++w) {
        v0 += rb[w * red_r];
        if (DUAL) v1 += rb[8 * red_cs + w * red_r];
      }
      const float v = rbf(v0);
      switch (epi) {
        case EPI_STORE: outv = v; break;
        case EPI_BIAS: outv = rbf(v + bf2f(bias[n])); break;
        case EPI_RESIDUAL: outv = rbf(rres[it] + v); break;
        case EPI_O_H1: outv = rres[it] + v; break;       // x + attn_out, un-rounded (fused_residual_rmsnorm.cu:60-65)
        case EPI_SWIGLU: outv = rbf(rbf(silu_f(v)) * rbf(v1)); break;
        case EPI_LOGITS: {
          if (Yf != nullptr) Yf[(size_t)t * pN + n] = v;
          key = argmax_key(v, n);
        } break;
        default: break;
      }
    }
    if (yf == XF_BF16T) {
      const float other = __shfl_xor_sync(0xffffffffu, outv, 1);
      if (valid && !(lane & 1))
        st_slot(Y64 + (((size_t)t * ldy + n) >> 1), pack2(outv, other), tag);
    } else if (yf == XF_F32T) {
      if (valid) st_slot(Y64 + (size_t)t * ldy + n, __float_as_uint(outv), tag);
    }
    if (epi == EPI_LOGITS && amax != nullptr) {
      // a warp covers 32 consecutive rows of ONE token: one atomic per warp instead of 32 on the same address
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const u64 other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
      }
      if (lane == 0 && key != 0ull) atomicMax(amax + t, key);
    }
  }
  m2_csync();
  m2_arrive(gs, p.flags);
  if (prof_on) prof2(a, 5);
  m2_prefetch(a, nx);
}

// RMSNorm row scale from the squares every lane accumulated over its own k-slices: quad shuffle, one exchange
// through shared memory, 16 warp partials summed in warp order (fixed order: independent of the batch size).
template <int NT>
__device__ __forceinline__ void m2_row_scales(const M2Args& a, float* part_s, float (&sq)[NT], int K, float (&xsc)[NT]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    sq[nt] += __shfl_xor_sync(0xffffffffu, sq[nt], 1);
    sq[nt] += __shfl_xor_sync(0xffffffffu, sq[nt], 2);
    if (tg == 0) part_s[(nt * 8 + g) * 16 + warp] = sq[nt];
  }
  m2_csync();
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const float4* pp = reinterpret_cast<const float4*>(part_s + (nt * 8 + g) * 16);
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = pp[i];
      tot += v.x; tot += v.y; tot += v.z; tot += v.w;
    }
    xsc[nt] = ref_mean_rsqrt(tot, K, a.eps);
  }
}

__device__ __forceinline__ uint4 m2_apply_norm(const uint4& x4, const uint4& w4, float sc) {
  float f[8], w[8];
  unpack8(x4, f);
  unpack8(w4, w);
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = (sc * f[e]) * w[e];
  return pack8(f);
}

// smem: scale[16] | part[16 tokens][16 warps] | red
constexpr int M2_RED_OFF = (16 + 256) * 4;

// ---------------------------------------------------------------------------------------------------
// Streaming variant: any K (multiple of 32), up to MEGA_MAX_TILES 16-row tiles per CTA; weight chunks are
// streamed through registers (two k-steps per warp in flight), the first chunk requested before the wait.
// XRES (NORM only, K a multiple of 1024 and <= 2048): the activations stay in registers after the scale pass.
template <bool DUAL, int NT, int XF, bool NORM, bool XRES>
__device__ __noinline__ unsigned long long m2_gemv(const M2Args& a, const M2Phase& p, const M2Phase* nx, unsigned char* smem,
                                                   M2Sync gs, const uint32_t tag) {
  float* part_s = reinterpret_cast<float*>(smem) + 16;
  float* red = reinterpret_cast<float*>(smem + M2_RED_OFF);
  const int K = p.K, T = p.T;
  const bool prof_on = M2_PROF_ENABLED && a.prof != nullptr;      // one shared-memory read instead of one per stamp
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const uint32_t xtag = tag - 1u;
  int r0, r1;
  mega_row_range(p.N, r0, r1);
  if (r1 <= r0) {
    m2_wait(gs, p.flags);
    m2_arrive(gs, p.flags);
    m2_prefetch(a, nx);
    return m2_pack(gs);
  }
  if (prof_on) prof2(a, 1);
  constexpr int NM = DUAL ? 2 : 1;
  constexpr int JU = 2;
  constexpr int XC = XRES ? 2 : 1;
  const int ksteps = K >> 5;
  const int jn = ksteps > warp ? (ksteps - warp + MEGA_WARPS - 1) / MEGA_WARPS : 0;
  const int n_chunks = max(1, ((ksteps + MEGA_WARPS - 1) / MEGA_WARPS + JU - 1) / JU);
  const int n_tiles = (r1 - r0 + 15) >> 4;
  const int red_r = n_tiles << 4, red_cs = 16 * red_r + 4;
  const int koff0 = warp * 32 + 8 * tg;
  const bool k_full = (ksteps % (MEGA_WARPS * JU)) == 0;
  const bool write_xn = NORM && p.xn_out != nullptr && blockIdx.x == 0;

  float rres[MEGA_MAX_OUT];

  uint4 wl[NM][JU], wh[NM][JU];
  auto load_w = [&](int tile, int c) {
    const int n0 = r0 + (tile << 4);
    const bool hi_ok = (n0 + 8) < r1;
    const size_t woff = (size_t)(n0 + g) * K + koff0 + (size_t)c * (JU * 512);
    if (k_full && hi_ok) {
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
  // ---- before the wait: first weights, (XRES) the norm weights, then the residual rows (their conversion waits for
  // the data, so they come last) ----
  load_w(0, 0);
  uint4 wnr[XC][JU];
  if constexpr (XRES) {
#pragma unroll
    for (int c = 0; c < XC; ++c)
#pragma unroll
      for (int u = 0; u < JU; ++u) {
        wnr[c][u] = make_uint4(0, 0, 0, 0);
        if (c < n_chunks) wnr[c][u] = *reinterpret_cast<const uint4*>(p.aux + koff0 + (c * JU + u) * 512);
      }
  }
  m2_load_residual<NT>(p, r0, r1, rres);
  m2_wait(gs, p.flags);
  if (prof_on) prof2(a, 2);
  if (prof_on) m2_stamp(gs, 0);
  const char* xrow[NT];
  m2_token_rows<NT, XF>(a, p, g, xrow);

  // synchronous, tag-verified load of one chunk of this lane's activations
  uint4 xv[NT][JU];
  float sqc[NT];
  auto load_x_sync = [&](int c) {
    unsigned tries = 0;
    for (;;) {
      uint32_t bad = 0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) sqc[nt] = 0.f;
#pragma unroll
      for (int u = 0; u < JU; ++u) {
        const bool ok = k_full || (c * JU + u) < jn;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          xv[nt][u] = make_uint4(0, 0, 0, 0);
          if (ok && xrow[nt] != nullptr) xv[nt][u] = m2_load_x8<XF>(xrow[nt], koff0 + (c * JU + u) * 512, xtag, bad, sqc[nt]);
        }
      }
      if (XF == XF_GATHER || !__any_sync(0xffffffffu, bad != 0)) break;
      if (M2_PROF_ENABLED && gs.retries != nullptr && (threadIdx.x & 31) == 0) atomicAdd(gs.retries, 1u);
      if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 3000000 + (int)gs.epoch); break; }
    }
  };

  float xsc[NT];
  uint4 xres[XC][JU][NT];
  if constexpr (NORM) {
    float sq[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) sq[nt] = 0.f;
    if constexpr (XRES) {
      // both chunks' slots in one batch of loads (one round trip instead of two)
      unsigned tries = 0;
      for (;;) {
        uint32_t bad = 0;
        float sqa[XC][NT];
#pragma unroll
        for (int c = 0; c < XC; ++c)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            sqa[c][nt] = 0.f;
#pragma unroll
            for (int u = 0; u < JU; ++u) {
              xres[c][u][nt] = make_uint4(0, 0, 0, 0);
              if (c < n_chunks && xrow[nt] != nullptr)
                xres[c][u][nt] = m2_load_x8<XF>(xrow[nt], koff0 + (c * JU + u) * 512, xtag, bad, sqa[c][nt]);
            }
          }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          sq[nt] = sqa[0][nt];
          if (XC > 1) sq[nt] += sqa[XC - 1][nt];
        }
        if (!__any_sync(0xffffffffu, bad != 0)) break;
        if (M2_PROF_ENABLED && gs.retries != nullptr && (threadIdx.x & 31) == 0) atomicAdd(gs.retries, 1u);
        if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 3500000 + (int)gs.epoch); break; }
      }
    } else {
      for (int c = 0; c < n_chunks; ++c) {
        load_x_sync(c);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) sq[nt] += sqc[nt];
      }
    }
    m2_row_scales<NT>(a, part_s, sq, K, xsc);
    if constexpr (XRES) {
#pragma unroll
      for (int c = 0; c < XC; ++c)
        if (c < n_chunks) {
#pragma unroll
          for (int u = 0; u < JU; ++u)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              xres[c][u][nt] = m2_apply_norm(xres[c][u][nt], wnr[c][u], xsc[nt]);
              if (write_xn && xrow[nt] != nullptr)
                *reinterpret_cast<uint4*>(p.xn_out + (size_t)(nt * 8 + g) * K + koff0 + (c * JU + u) * 512) = xres[c][u][nt];
            }
        }
    }
  }
  // raw tagged slots of the next chunk, requested right after the previous chunk was consumed (non-NORM bf16 input)
  u64 xr[NT][JU][4];
  auto load_x_raw = [&](int c) {
#pragma unroll
    for (int u = 0; u < JU; ++u) {
      const bool ok = k_full || (c * JU + u) < jn;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (ok && xrow[nt] != nullptr) {
          const char* q = xrow[nt] + (size_t)(koff0 + (c * JU + u) * 512) * 4;
          ld_slot2(q, xr[nt][u][0], xr[nt][u][1]);
          ld_slot2(q + 16, xr[nt][u][2], xr[nt][u][3]);
        } else {
          // slots of tokens / k-steps that do not exist read as zero payload with the expected tag
          xr[nt][u][0] = xr[nt][u][1] = xr[nt][u][2] = xr[nt][u][3] = (u64)xtag << 32;
        }
      }
    }
  };
  auto load_x_plain = [&](int c) {
#pragma unroll
    for (int u = 0; u < JU; ++u) {
      const bool ok = k_full || (c * JU + u) < jn;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        xv[nt][u] = make_uint4(0, 0, 0, 0);
        if (ok && xrow[nt] != nullptr) xv[nt][u] = ldcg16(xrow[nt] + (size_t)(koff0 + (c * JU + u) * 512) * 2);
      }
    }
  };
  constexpr bool RAW = !NORM && XF == XF_BF16T;
  constexpr bool PLAIN = !NORM && XF == XF_GATHER;
  if constexpr (RAW) load_x_raw(0);
  if constexpr (PLAIN) load_x_plain(0);
  if (prof_on) prof2(a, 3);
  if (prof_on) m2_stamp(gs, 1);
  float acc[NM][NT][4];
  for (int tile = 0; tile < n_tiles; ++tile) {
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
      if constexpr (RAW) {
        // verify the tags of the chunk that is about to be consumed; re-read until the producers' stores landed
        unsigned tries = 0;
        for (;;) {
          uint32_t bad = 0;
#pragma unroll
          for (int u = 0; u < JU; ++u)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int i = 0; i < 4; ++i) bad |= slot_tag(xr[nt][u][i]) ^ xtag;
          if (!__any_sync(0xffffffffu, bad != 0)) break;
          if (M2_PROF_ENABLED && gs.retries != nullptr && (threadIdx.x & 31) == 0) atomicAdd(gs.retries, 1u);
      if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 4000000 + (int)gs.epoch); break; }
          load_x_raw(c);
        }
#pragma unroll
        for (int u = 0; u < JU; ++u)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
            xv[nt][u] = make_uint4(slot_val(xr[nt][u][0]), slot_val(xr[nt][u][1]), slot_val(xr[nt][u][2]), slot_val(xr[nt][u][3]));
      }
      if constexpr (NORM && !XRES) {
        load_x_sync(c);
#pragma unroll
        for (int u = 0; u < JU; ++u) {
          const bool ok = k_full || (c * JU + u) < jn;
          uint4 w4 = make_uint4(0, 0, 0, 0);
          if (ok) w4 = *reinterpret_cast<const uint4*>(p.aux + koff0 + (c * JU + u) * 512);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            xv[nt][u] = m2_apply_norm(xv[nt][u], w4, xsc[nt]);
            if (write_xn && tile == 0 && ok && xrow[nt] != nullptr)
              *reinterpret_cast<uint4*>(p.xn_out + (size_t)(nt * 8 + g) * K + koff0 + (c * JU + u) * 512) = xv[nt][u];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < JU; ++u) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint4 x4;
          if constexpr (XRES) x4 = (c == 0) ? xres[0][u][nt] : xres[XC - 1][u][nt];
          else x4 = xv[nt][u];
#pragma unroll
          for (int m = 0; m < NM; ++m) {
            mma_bf16_16816(acc[m][nt], wl[m][u].x, wh[m][u].x, wl[m][u].y, wh[m][u].y, x4.x, x4.y);
            mma_bf16_16816(acc[m][nt], wl[m][u].z, wh[m][u].z, wl[m][u].w, wh[m][u].w, x4.z, x4.w);
          }
        }
      }
      if (c + 1 < n_chunks) {
        load_w(tile, c + 1);
        if constexpr (RAW) load_x_raw(c + 1);
        if constexpr (PLAIN) load_x_plain(c + 1);
      } else if (tile + 1 < n_tiles) {
        load_w(tile + 1, 0);
        if constexpr (RAW) load_x_raw(0);
        if constexpr (PLAIN) load_x_plain(0);
      }
    }
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float* r = red + (size_t)((nt * NM + m) * 8 + 2 * tg) * red_cs + warp * red_r + (tile << 4) + g;
        r[0] = acc[m][nt][0]; r[red_cs] = acc[m][nt][1]; r[8] = acc[m][nt][2]; r[red_cs + 8] = acc[m][nt][3];
      }
  }
  m2_tail<DUAL, NT>(a, p, nx, red, rres, r0, r1, n_tiles, gs, tag);
  return m2_pack(gs);
}

// ---------------------------------------------------------------------------------------------------
// Register-resident variant: K == CHUNKS * 1024 and at most TILES 16-row tiles per CTA, so ALL of a lane's
// weight fragments (<= 16 x 128 bit) are requested before the wait; after it the activations are fetched once.
template <bool DUAL, int NT, int XF, bool NORM, int TILES, int CHUNKS>
__device__ __noinline__ unsigned long long m2_gemv_small(const M2Args& a, const M2Phase& p, const M2Phase* nx,
                                                         unsigned char* smem, M2Sync gs, const uint32_t tag) {
  float* part_s = reinterpret_cast<float*>(smem) + 16;
  float* red = reinterpret_cast<float*>(smem + M2_RED_OFF);
  const int K = p.K;
  const bool prof_on = M2_PROF_ENABLED && a.prof != nullptr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const uint32_t xtag = tag - 1u;
  constexpr int NM = DUAL ? 2 : 1;
  constexpr int JU = 2;
  int r0, r1;
  mega_row_range(p.N, r0, r1);
  if (r1 <= r0) {
    m2_wait(gs, p.flags);
    m2_arrive(gs, p.flags);
    m2_prefetch(a, nx);
    return m2_pack(gs);
  }
  if (prof_on) prof2(a, 1);
  const int n_tiles = (r1 - r0 + 15) >> 4;
  const int red_r = n_tiles << 4, red_cs = 16 * red_r + 4;
  const int koff0 = warp * 32 + 8 * tg;
  float rres[MEGA_MAX_OUT];
  // ---- before the wait: every weight fragment of the phase, the norm weights, then the residual rows ----
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
  uint4 wn[CHUNKS][JU];
  if constexpr (NORM) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
      for (int u = 0; u < JU; ++u) wn[c][u] = *reinterpret_cast<const uint4*>(p.aux + koff0 + (c * JU + u) * 512);
  }
  m2_load_residual<NT>(p, r0, r1, rres);
  m2_wait(gs, p.flags);
  if (prof_on) prof2(a, 2);
  if (prof_on) m2_stamp(gs, 0);
  // ---- after the wait: activations (once, tag-verified), scales ----
  const char* xrow[NT];
  m2_token_rows<NT, XF>(a, p, g, xrow);
  uint4 xv[CHUNKS][JU][NT];
  float sq[NT];
  {
    unsigned tries = 0;
    for (;;) {
      uint32_t bad = 0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) sq[nt] = 0.f;
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
        for (int u = 0; u < JU; ++u)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            xv[c][u][nt] = make_uint4(0, 0, 0, 0);
            if (xrow[nt] != nullptr) xv[c][u][nt] = m2_load_x8<XF>(xrow[nt], koff0 + (c * JU + u) * 512, xtag, bad, sq[nt]);
          }
      if (XF == XF_GATHER || !__any_sync(0xffffffffu, bad != 0)) break;
      if (M2_PROF_ENABLED && gs.retries != nullptr && (threadIdx.x & 31) == 0) atomicAdd(gs.retries, 1u);
      if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 5000000 + (int)gs.epoch); break; }
    }
  }
  if constexpr (NORM) {
    float xsc[NT];
    m2_row_scales<NT>(a, part_s, sq, K, xsc);
    const bool write_xn = p.xn_out != nullptr && blockIdx.x == 0;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
      for (int u = 0; u < JU; ++u)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          xv[c][u][nt] = m2_apply_norm(xv[c][u][nt], wn[c][u], xsc[nt]);
          if (write_xn && xrow[nt] != nullptr)
            *reinterpret_cast<uint4*>(p.xn_out + (size_t)(nt * 8 + g) * K + koff0 + (c * JU + u) * 512) = xv[c][u][nt];
        }
  }
  if (prof_on) prof2(a, 3);
  if (prof_on) m2_stamp(gs, 1);
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
              const uint4 x4 = xv[c][u][nt];
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
  m2_tail<DUAL, NT>(a, p, nx, red, rres, r0, r1, n_tiles, gs, tag);
  return m2_pack(gs);
}

// variant dispatch (uniform over the grid: everything comes from the descriptor)
__device__ __forceinline__ unsigned long long m2_gemv_dispatch(const M2Args& a, const M2Phase& p, const M2Phase* nx,
                                                               unsigned char* smem, const M2Sync& gs, const uint32_t tag) {
  const bool nt1 = p.T <= 8;
  const bool dual = (p.flags & PF_DUAL) != 0, norm = (p.flags & PF_NORM) != 0;
  const int sm = p.small;
  if (sm != 0) {
    if (dual) {             // gate/up: NORM, f32 input (h1)
      if (nt1) return m2_gemv_small<true, 1, XF_F32T, true, 2, 1>(a, p, nx, smem, gs, tag);
      else return m2_gemv_small<true, 2, XF_F32T, true, 2, 1>(a, p, nx, smem, gs, tag);
    } else if (norm) {
      if (sm == 0x21) {
        if (nt1) return m2_gemv_small<false, 1, XF_BF16T, true, 2, 1>(a, p, nx, smem, gs, tag);
        else return m2_gemv_small<false, 2, XF_BF16T, true, 2, 1>(a, p, nx, smem, gs, tag);
      } else {              // 0x22, T <= 8 only
        return m2_gemv_small<false, 1, XF_BF16T, true, 2, 2>(a, p, nx, smem, gs, tag);
      }
    } else if (p.xf == XF_GATHER) {
      if (nt1) return m2_gemv_small<false, 1, XF_GATHER, false, 1, 2>(a, p, nx, smem, gs, tag);
      else return m2_gemv_small<false, 2, XF_GATHER, false, 1, 2>(a, p, nx, smem, gs, tag);
    } else if (sm == 0x12) {
      if (nt1) return m2_gemv_small<false, 1, XF_BF16T, false, 1, 2>(a, p, nx, smem, gs, tag);
      else return m2_gemv_small<false, 2, XF_BF16T, false, 1, 2>(a, p, nx, smem, gs, tag);
    } else {                // 0x13
      if (nt1) return m2_gemv_small<false, 1, XF_BF16T, false, 1, 3>(a, p, nx, smem, gs, tag);
      else return m2_gemv_small<false, 2, XF_BF16T, false, 1, 3>(a, p, nx, smem, gs, tag);
    }
  }
  const bool xres = norm && p.K <= 2048 && (p.K & 1023) == 0;
  if (dual) {               // NORM, f32 input
    if (xres) {
      if (nt1) return m2_gemv<true, 1, XF_F32T, true, true>(a, p, nx, smem, gs, tag);
      else return m2_gemv<true, 2, XF_F32T, true, true>(a, p, nx, smem, gs, tag);
    } else {
      if (nt1) return m2_gemv<true, 1, XF_F32T, true, false>(a, p, nx, smem, gs, tag);
      else return m2_gemv<true, 2, XF_F32T, true, false>(a, p, nx, smem, gs, tag);
    }
  } else if (norm) {
    if (xres) {
      if (nt1) return m2_gemv<false, 1, XF_BF16T, true, true>(a, p, nx, smem, gs, tag);
      else return m2_gemv<false, 2, XF_BF16T, true, true>(a, p, nx, smem, gs, tag);
    } else {
      if (nt1) return m2_gemv<false, 1, XF_BF16T, true, false>(a, p, nx, smem, gs, tag);
      else return m2_gemv<false, 2, XF_BF16T, true, false>(a, p, nx, smem, gs, tag);
    }
  } else if (p.xf == XF_GATHER) {
    if (nt1) return m2_gemv<false, 1, XF_GATHER, false, false>(a, p, nx, smem, gs, tag);
    else return m2_gemv<false, 2, XF_GATHER, false, false>(a, p, nx, smem, gs, tag);
  } else {
    if (nt1) return m2_gemv<false, 1, XF_BF16T, false, false>(a, p, nx, smem, gs, tag);
    else return m2_gemv<false, 2, XF_BF16T, false, false>(a, p, nx, smem, gs, tag);
  }
}

// ---------------------------------------------------------------------------------------------------
// QK-norm + RoPE + KV append + attention for one (row b, kv head) item per CTA; S tokens per row in order
// (Attention::forward, transformer.rs:294-369, matmul path: scores rounded to bf16, scale applied as a separate
// bf16 multiply, f32 softmax rounded to bf16, f32-accumulated PV rounded to bf16).
// Input: tagged qkv rows; output: tagged attention rows.  The K/V rows of earlier positions were written by THIS
// CTA (same item -> same CTA in every frame) or by the prefill kernels, so they need no tags.
constexpr int M2_ATT_FAST_L = 16;
// cache rows per warp whose loads are in flight together in the general path: the loop is latency-bound (one L2 / HBM round trip
// per batch of rows), so at 2000 cached positions 16 rows per batch instead of 8 nearly halves the phase (same row order per
// warp, hence bit-identical results)
#ifndef M2_ATT_U
#define M2_ATT_U 16
#endif
__device__ __noinline__ unsigned long long m2_attn_units(const M2Args& a, const M2Phase& p, unsigned char* smem, M2Sync gs,
                                                         const uint32_t tag);
__device__ __noinline__ unsigned long long m2_attn(const M2Args& a, const M2Phase& p, unsigned char* smem, M2Sync gs,
                                                   const uint32_t tag) {
  if (a.split_min_l > 0 && !(p.flags & PF_CP) && p.S == 1) return m2_attn_units(a, p, smem, gs, tag);
  const bool cp = (p.flags & PF_CP) != 0;
  const bool prof_on = M2_PROF_ENABLED && a.prof != nullptr;
  const int max_seq = cp ? a.cp_max_seq : a.max_seq;
  const bf16* cos_tab = cp ? a.cp_cos : a.t_cos;
  const bf16* sin_tab = cp ? a.cp_sin : a.t_sin;
  const int* pos_base = cp ? nullptr : a.fs.offset;
  const int heads = p.N, kv_heads = p.K, S = p.S, nh = heads + 2 * kv_heads;
  const int sc_n = (max_seq + 3) & ~3;                   // keeps everything behind the score arrays 16-byte aligned
  float* sc0 = reinterpret_cast<float*>(smem);          // [max_seq]
  float* sc1 = sc0 + sc_n;
  float* qs = sc1 + sc_n;                                // [2][128] rotated queries (bf16 values)
  float* red = qs + 256;                                 // [16][2][128]
  bf16* Ks = reinterpret_cast<bf16*>(red + 16 * 2 * 128);   // [16][128] fast path: K rows (row pos = this token)
  bf16* Vs = Ks + M2_ATT_FAST_L * 128;                        // [16][128]
  bf16* kcur = Vs + M2_ATT_FAST_L * 128;                      // [128] general path: this token's rotated K row
  bf16* vcur = kcur + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t xtag = tag - 1u;
  const float scale = rbf(0.08838834764831845f);
  const int n_items = (p.T / p.S) * kv_heads;       // rows of this phase (a pass-0 group of a batch > 8 has fewer than a.B)
  const bool worker = (int)blockIdx.x < n_items;
  // rows that can be requested before the wait (fast path, first token of the item)
  int b0 = 0, pos0 = 0;
  if (worker) {
    b0 = blockIdx.x / kv_heads;
    pos0 = (pos_base ? __ldcg(pos_base + b0) : 0) + p.pos_add;
  }
  m2_wait(gs, p.flags);
  if (prof_on) prof2(a, 6);
  if (prof_on) m2_stamp(gs, 0);
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / kv_heads, kvh = item - b * kv_heads;
    bf16* kbase = const_cast<bf16*>(p.W) + ((size_t)b * kv_heads + kvh) * max_seq * 128;
    bf16* vbase = const_cast<bf16*>(p.W2) + ((size_t)b * kv_heads + kvh) * max_seq * 128;
    const int posb = item == (int)blockIdx.x ? pos0 : (pos_base ? __ldcg(pos_base + b) : 0) + p.pos_add;
    for (int s = 0; s < S; ++s) {
      const int t = b * S + s;
      const int pos = posb + s;
      const int L = pos + 1;
      const bool fast = L <= M2_ATT_FAST_L;
      uint2 kpre = make_uint2(0u, 0u), vpre = make_uint2(0u, 0u);
      if (fast) {
        // warps 4..15: rows j < pos of K and V -> shared memory (16 lanes x 16 B per row)
        if (warp >= 4) {
          for (int i = tid - 128; i < 2 * pos * 16; i += MEGA_THREADS - 128) {
            const int r = i >> 4, c16 = i & 15;
            const bool isv = r >= pos;
            const int j = isv ? r - pos : r;
            const uint4 v = ldcg16((isv ? vbase : kbase) + (size_t)j * 128 + c16 * 8);
            *reinterpret_cast<uint4*>((isv ? Vs : Ks) + j * 128 + c16 * 8) = v;
          }
        }
      } else if (warp < pos) {
        kpre = __ldcg(reinterpret_cast<const uint2*>(kbase + (size_t)warp * 128 + 4 * lane));
        vpre = __ldcg(reinterpret_cast<const uint2*>(vbase + (size_t)warp * 128 + 4 * lane));
      }
      // warps 0,1: q heads 2kvh, 2kvh+1; warp 2: k; warp 3: v.  Lane l holds elements 2l, 2l+1, 64+2l, 65+2l.
      if (warp < 4) {
        const int hh = warp < 2 ? 2 * kvh + warp : (warp == 2 ? heads + kvh : heads + kv_heads + kvh);
        const u64* src = reinterpret_cast<const u64*>(p.X) + (((size_t)t * nh + hh) * 128 >> 1);
        u64 s0, s1;
        unsigned tries = 0;
        for (;;) {
          s0 = ld_slot(src + lane);
          s1 = ld_slot(src + 32 + lane);
          const uint32_t bad = (slot_tag(s0) ^ xtag) | (slot_tag(s1) ^ xtag);
          if (!__any_sync(0xffffffffu, bad != 0)) break;
          if (M2_PROF_ENABLED && gs.retries != nullptr && (threadIdx.x & 31) == 0) atomicAdd(gs.retries, 1u);
      if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 6000000 + (int)gs.epoch); break; }
        }
        float v[4] = {bf_lo(slot_val(s0)), bf_hi(slot_val(s0)), bf_lo(slot_val(s1)), bf_hi(slot_val(s1))};
        const int d0 = 2 * lane;           // v[0],v[1] -> d0, d0+1 ; v[2],v[3] -> 64+d0, 65+d0
        if (warp == 3) {
          bf16* dst = fast ? Vs + pos * 128 : vcur;
          const uint32_t lo = slot_val(s0), hi = slot_val(s1);
          *reinterpret_cast<uint32_t*>(vbase + (size_t)pos * 128 + d0) = lo;
          *reinterpret_cast<uint32_t*>(vbase + (size_t)pos * 128 + 64 + d0) = hi;
          *reinterpret_cast<uint32_t*>(dst + d0) = lo;
          *reinterpret_cast<uint32_t*>(dst + 64 + d0) = hi;
        } else {
          const bf16* nw = warp < 2 ? p.aux : p.aux2;
          float tmp = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) tmp = fmaf(v[i], v[i], tmp);
          tmp = warp_sum_xor(tmp);
          const float sc = ref_mean_rsqrt(tmp, 128, a.eps);
          const uint32_t w0 = *reinterpret_cast<const uint32_t*>(nw + d0), w1 = *reinterpret_cast<const uint32_t*>(nw + 64 + d0);
          const float n0 = rbf((sc * v[0]) * bf_lo(w0)), n1 = rbf((sc * v[1]) * bf_hi(w0));
          const float n2 = rbf((sc * v[2]) * bf_lo(w1)), n3 = rbf((sc * v[3]) * bf_hi(w1));
          const uint32_t cw = *reinterpret_cast<const uint32_t*>(cos_tab + (size_t)pos * 64 + d0);
          const uint32_t sw = *reinterpret_cast<const uint32_t*>(sin_tab + (size_t)pos * 64 + d0);
          const float c0 = bf_lo(cw), c1 = bf_hi(cw), sn0 = bf_lo(sw), sn1 = bf_hi(sw);
          const float o0 = rbf(rbf(n0 * c0) - rbf(n2 * sn0)), o1 = rbf(rbf(n1 * c1) - rbf(n3 * sn1));
          const float o2 = rbf(rbf(n2 * c0) + rbf(n0 * sn0)), o3 = rbf(rbf(n3 * c1) + rbf(n1 * sn1));
          if (warp < 2) {
            *reinterpret_cast<float2*>(qs + warp * 128 + d0) = make_float2(o0, o1);
            *reinterpret_cast<float2*>(qs + warp * 128 + 64 + d0) = make_float2(o2, o3);
          } else {
            bf16* dst = fast ? Ks + pos * 128 : kcur;
            const uint32_t lo = pack2(o0, o1), hi = pack2(o2, o3);
            *reinterpret_cast<uint32_t*>(kbase + (size_t)pos * 128 + d0) = lo;
            *reinterpret_cast<uint32_t*>(kbase + (size_t)pos * 128 + 64 + d0) = hi;
            *reinterpret_cast<uint32_t*>(dst + d0) = lo;
            *reinterpret_cast<uint32_t*>(dst + 64 + d0) = hi;
          }
        }
      }
      m2_csync();      // q, this token's k/v (and, fast path, all earlier rows) in smem
      float q0[4], q1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { q0[i] = qs[4 * lane + i]; q1[i] = qs[128 + 4 * lane + i]; }
      if (fast) {
        // one position per warp; both heads
        if (warp < L) {
          const uint2 u = *reinterpret_cast<const uint2*>(Ks + warp * 128 + 4 * lane);
          const float k0 = bf_lo(u.x), k1 = bf_hi(u.x), k2 = bf_lo(u.y), k3 = bf_hi(u.y);
          float d0 = q0[0] * k0 + q0[1] * k1 + q0[2] * k2 + q0[3] * k3;
          float d1 = q1[0] * k0 + q1[1] * k1 + q1[2] * k2 + q1[3] * k3;
          d0 = warp_sum_xor(d0);
          d1 = warp_sum_xor(d1);
          if (lane == 0) {
            sc0[warp] = rbf(rbf(d0) * scale);
            sc1[warp] = rbf(rbf(d1) * scale);
          }
        }
        m2_csync();
        // softmax of the (<= 16) scores of head h by warp h, one position per lane; probabilities (rounded to bf16 as the
        // reference's softmax output is) go back into the score array
        if (warp < 2) {
          float* sc = warp == 0 ? sc0 : sc1;
          const float x = lane < L ? sc[lane] : -INFINITY;
          const float m = warp_max(x);
          const float e = lane < L ? expf(x - m) : 0.f;
          const float sum = warp_sum_xor(e);
          if (lane < L) sc[lane] = rbf(e / sum);
        }
        m2_csync();
        float outv = 0.f;
        const int h = (tid >> 7) & 1, d = tid & 127;
        if (tid < 256) {
          const float* sc = h == 0 ? sc0 : sc1;
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < M2_ATT_FAST_L; ++j)
            if (j < L) acc = fmaf(sc[j], bf2f(Vs[j * 128 + d]), acc);
          outv = rbf(acc);
        }
        const float other = __shfl_xor_sync(0xffffffffu, outv, 1);
        if (tid < 256 && !(lane & 1))
          st_slot(reinterpret_cast<u64*>(p.Y) + (((size_t)t * heads + (2 * kvh + h)) * 128 + d >> 1), pack2(outv, other), tag);
        if (s + 1 < S || item + (int)gridDim.x < n_items) m2_csync();
        continue;
      }
      // ---- general path ----
      auto row = [&](const bf16* base, const bf16* cur, const uint2& pre, int j) -> uint2 {
        if (j == pos) return *reinterpret_cast<const uint2*>(cur + 4 * lane);
        if (j == warp) return pre;
        return __ldcg(reinterpret_cast<const uint2*>(base + (size_t)j * 128 + 4 * lane));
      };
      for (int j0 = warp; j0 < L; j0 += M2_ATT_U * MEGA_WARPS) {
        uint2 ku[M2_ATT_U];
#pragma unroll
        for (int q = 0; q < M2_ATT_U; ++q) {
          const int j = j0 + q * MEGA_WARPS;
          ku[q] = j < L ? row(kbase, kcur, kpre, j) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < M2_ATT_U; ++q) {
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
      m2_csync();
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
      m2_csync();
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j0 = warp; j0 < L; j0 += M2_ATT_U * MEGA_WARPS) {
        uint2 vu[M2_ATT_U];
#pragma unroll
        for (int q = 0; q < M2_ATT_U; ++q) {
          const int j = j0 + q * MEGA_WARPS;
          vu[q] = j < L ? row(vbase, vcur, vpre, j) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < M2_ATT_U; ++q) {
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
      m2_csync();
      {
        float outv = 0.f;
        const int h = (tid >> 7) & 1, d = tid & 127;
        if (tid < 256) {
          float acc = 0.f;
#pragma unroll
          for (int w = 0; w < MEGA_WARPS; ++w) acc += red[(w * 2 + h) * 128 + d];
          outv = rbf(acc);
        }
        const float other = __shfl_xor_sync(0xffffffffu, outv, 1);
        if (tid < 256 && !(lane & 1))
          st_slot(reinterpret_cast<u64*>(p.Y) + (((size_t)t * heads + (2 * kvh + h)) * 128 + d >> 1), pack2(outv, other), tag);
      }
      m2_csync();
    }
  }
  m2_csync();
  m2_arrive(gs, p.flags);
  if (prof_on) prof2(a, 7);
  return m2_pack(gs);
}

// ---------------------------------------------------------------------------------------------------
// Split-KV form of the talker's decode attention (one token per row), used for the launches in which some row's context can
// reach a.split_min_l positions (host decision, q3tts.cu mega2_launch).  A row whose context is shorter takes exactly the
// code path of m2_attn (same row order per warp, same reduction trees: bit-identical results); a longer row is cut into
// a.split_ns (4) ranges of cache rows handled by as many CTAs, which exchange the softmax maximum, the normaliser and
// their partial P*V sums through tagged slots in global memory (a.xchg).  Whether a row is split depends on ITS context
// length only, never on the other rows of the batch, so a row of a batch still equals its batch-1 run bit for bit.
// At 2000 positions m2_attn is bound by the latency of its batches of cache rows on the 8 x B CTAs that have work
// (~40 us per layer); here up to 4 x 8 x B CTAs share them.  The current position's K / V rows are appended by the last
// split only; every other cache row was written at least one frame (549 phases, several of them release / acquire
// barriers) earlier, by whichever CTA owned the position then.
constexpr int M2_SPLIT_NS_MAX = 8;            // a.split_ns: 4 by default, Q3_SPLIT_NS=8 for single-stream long-form (host, q3tts.cu)
constexpr int M2_XCHG_SLOTS = 4 + 256;          // per split: max[2], sum[2], partial P*V [2][128]
__device__ __noinline__ unsigned long long m2_attn_units(const M2Args& a, const M2Phase& p, unsigned char* smem, M2Sync gs,
                                                         const uint32_t tag) {
  const bool cp = (p.flags & PF_CP) != 0;
  const bool prof_on = M2_PROF_ENABLED && a.prof != nullptr;
  const int max_seq = cp ? a.cp_max_seq : a.max_seq;
  const bf16* cos_tab = cp ? a.cp_cos : a.t_cos;
  const bf16* sin_tab = cp ? a.cp_sin : a.t_sin;
  const int* pos_base = cp ? nullptr : a.fs.offset;
  const int heads = p.N, kv_heads = p.K, S = p.S, nh = heads + 2 * kv_heads;
  const int sc_n = (max_seq + 3) & ~3;                   // keeps everything behind the score arrays 16-byte aligned
  float* sc0 = reinterpret_cast<float*>(smem);          // [max_seq]
  float* sc1 = sc0 + sc_n;
  float* qs = sc1 + sc_n;                                // [2][128] rotated queries (bf16 values)
  float* red = qs + 256;                                 // [16][2][128]
  bf16* Ks = reinterpret_cast<bf16*>(red + 16 * 2 * 128);   // [16][128] fast path: K rows (row pos = this token)
  bf16* Vs = Ks + M2_ATT_FAST_L * 128;                        // [16][128]
  bf16* kcur = Vs + M2_ATT_FAST_L * 128;                      // [128] general path: this token's rotated K row
  bf16* vcur = kcur + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t xtag = tag - 1u;
  const float scale = rbf(0.08838834764831845f);
  const int n_rows = p.T;                           // S == 1: one token per row
  int* s_L = reinterpret_cast<int*>(vcur + 128);     // [16] context length of every row of the phase
  int* s_ub = s_L + 16;                              // [17] first work unit of every row
  m2_wait(gs, p.flags);
  if (prof_on) prof2(a, 6);
  if (prof_on) m2_stamp(gs, 0);
  if (tid < n_rows) s_L[tid] = __ldcg(pos_base + tid) + p.pos_add + 1;
  m2_csync();
  if (tid == 0) {
    int u = 0;
    for (int r = 0; r < n_rows; ++r) { s_ub[r] = u; u += kv_heads * (s_L[r] >= a.split_min_l ? a.split_ns : 1); }
    s_ub[n_rows] = u;
  }
  m2_csync();
  const int n_units = s_ub[n_rows];
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    int b = 0;
    while (b + 1 < n_rows && s_ub[b + 1] <= unit) ++b;
    const int L = s_L[b], pos = L - 1;
    const int ns = L >= a.split_min_l ? a.split_ns : 1;
    const int ru = unit - s_ub[b], kvh = ru / ns, split = ru - kvh * ns;
    bf16* kbase = const_cast<bf16*>(p.W) + ((size_t)b * kv_heads + kvh) * max_seq * 128;
    bf16* vbase = const_cast<bf16*>(p.W2) + ((size_t)b * kv_heads + kvh) * max_seq * 128;
    // rows [j_lo, j_hi) of the cache belong to this split; the last split owns the current position (it appends K / V)
    const int chunk = (L + ns - 1) / ns;
    const int j_lo = split * chunk, j_hi = min(L, j_lo + chunk);
    const bool owner = split == ns - 1;
    u64* xg = a.xchg + (size_t)(unit - split) * M2_XCHG_SLOTS;      // exchange slots of the item's splits: [ns][M2_XCHG_SLOTS]
    {
      const int s = 0;
      const int t = b;
      const bool fast = L <= M2_ATT_FAST_L;      // (ns == 1 there: split_min_l > 16)
      uint2 kpre = make_uint2(0u, 0u), vpre = make_uint2(0u, 0u);
      if (fast) {
        // warps 4..15: rows j < pos of K and V -> shared memory (16 lanes x 16 B per row)
        if (warp >= 4) {
          for (int i = tid - 128; i < 2 * pos * 16; i += MEGA_THREADS - 128) {
            const int r = i >> 4, c16 = i & 15;
            const bool isv = r >= pos;
            const int j = isv ? r - pos : r;
            const uint4 v = ldcg16((isv ? vbase : kbase) + (size_t)j * 128 + c16 * 8);
            *reinterpret_cast<uint4*>((isv ? Vs : Ks) + j * 128 + c16 * 8) = v;
          }
        }
      } else if (ns == 1 && warp < pos) {
        kpre = __ldcg(reinterpret_cast<const uint2*>(kbase + (size_t)warp * 128 + 4 * lane));
        vpre = __ldcg(reinterpret_cast<const uint2*>(vbase + (size_t)warp * 128 + 4 * lane));
      }
      // warps 0,1: q heads 2kvh, 2kvh+1; warp 2: k; warp 3: v.  Lane l holds elements 2l, 2l+1, 64+2l, 65+2l.
      if (warp < 2 || (warp < 4 && owner)) {
        const int hh = warp < 2 ? 2 * kvh + warp : (warp == 2 ? heads + kvh : heads + kv_heads + kvh);
        const u64* src = reinterpret_cast<const u64*>(p.X) + (((size_t)t * nh + hh) * 128 >> 1);
        u64 s0, s1;
        unsigned tries = 0;
        for (;;) {
          s0 = ld_slot(src + lane);
          s1 = ld_slot(src + 32 + lane);
          const uint32_t bad = (slot_tag(s0) ^ xtag) | (slot_tag(s1) ^ xtag);
          if (!__any_sync(0xffffffffu, bad != 0)) break;
          if (M2_PROF_ENABLED && gs.retries != nullptr && (threadIdx.x & 31) == 0) atomicAdd(gs.retries, 1u);
      if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 6000000 + (int)gs.epoch); break; }
        }
        float v[4] = {bf_lo(slot_val(s0)), bf_hi(slot_val(s0)), bf_lo(slot_val(s1)), bf_hi(slot_val(s1))};
        const int d0 = 2 * lane;           // v[0],v[1] -> d0, d0+1 ; v[2],v[3] -> 64+d0, 65+d0
        if (warp == 3) {
          bf16* dst = fast ? Vs + pos * 128 : vcur;
          const uint32_t lo = slot_val(s0), hi = slot_val(s1);
          *reinterpret_cast<uint32_t*>(vbase + (size_t)pos * 128 + d0) = lo;
          *reinterpret_cast<uint32_t*>(vbase + (size_t)pos * 128 + 64 + d0) = hi;
          *reinterpret_cast<uint32_t*>(dst + d0) = lo;
          *reinterpret_cast<uint32_t*>(dst + 64 + d0) = hi;
        } else {
          const bf16* nw = warp < 2 ? p.aux : p.aux2;
          float tmp = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) tmp = fmaf(v[i], v[i], tmp);
          tmp = warp_sum_xor(tmp);
          const float sc = ref_mean_rsqrt(tmp, 128, a.eps);
          const uint32_t w0 = *reinterpret_cast<const uint32_t*>(nw + d0), w1 = *reinterpret_cast<const uint32_t*>(nw + 64 + d0);
          const float n0 = rbf((sc * v[0]) * bf_lo(w0)), n1 = rbf((sc * v[1]) * bf_hi(w0));
          const float n2 = rbf((sc * v[2]) * bf_lo(w1)), n3 = rbf((sc * v[3]) * bf_hi(w1));
          const uint32_t cw = *reinterpret_cast<const uint32_t*>(cos_tab + (size_t)pos * 64 + d0);
          const uint32_t sw = *reinterpret_cast<const uint32_t*>(sin_tab + (size_t)pos * 64 + d0);
          const float c0 = bf_lo(cw), c1 = bf_hi(cw), sn0 = bf_lo(sw), sn1 = bf_hi(sw);
          const float o0 = rbf(rbf(n0 * c0) - rbf(n2 * sn0)), o1 = rbf(rbf(n1 * c1) - rbf(n3 * sn1));
          const float o2 = rbf(rbf(n2 * c0) + rbf(n0 * sn0)), o3 = rbf(rbf(n3 * c1) + rbf(n1 * sn1));
          if (warp < 2) {
            *reinterpret_cast<float2*>(qs + warp * 128 + d0) = make_float2(o0, o1);
            *reinterpret_cast<float2*>(qs + warp * 128 + 64 + d0) = make_float2(o2, o3);
          } else {
            bf16* dst = fast ? Ks + pos * 128 : kcur;
            const uint32_t lo = pack2(o0, o1), hi = pack2(o2, o3);
            *reinterpret_cast<uint32_t*>(kbase + (size_t)pos * 128 + d0) = lo;
            *reinterpret_cast<uint32_t*>(kbase + (size_t)pos * 128 + 64 + d0) = hi;
            *reinterpret_cast<uint32_t*>(dst + d0) = lo;
            *reinterpret_cast<uint32_t*>(dst + 64 + d0) = hi;
          }
        }
      }
      m2_csync();      // q, this token's k/v (and, fast path, all earlier rows) in smem
      float q0[4], q1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { q0[i] = qs[4 * lane + i]; q1[i] = qs[128 + 4 * lane + i]; }
      if (fast) {
        // one position per warp; both heads
        if (warp < L) {
          const uint2 u = *reinterpret_cast<const uint2*>(Ks + warp * 128 + 4 * lane);
          const float k0 = bf_lo(u.x), k1 = bf_hi(u.x), k2 = bf_lo(u.y), k3 = bf_hi(u.y);
          float d0 = q0[0] * k0 + q0[1] * k1 + q0[2] * k2 + q0[3] * k3;
          float d1 = q1[0] * k0 + q1[1] * k1 + q1[2] * k2 + q1[3] * k3;
          d0 = warp_sum_xor(d0);
          d1 = warp_sum_xor(d1);
          if (lane == 0) {
            sc0[warp] = rbf(rbf(d0) * scale);
            sc1[warp] = rbf(rbf(d1) * scale);
          }
        }
        m2_csync();
        // softmax of the (<= 16) scores of head h by warp h, one position per lane; probabilities (rounded to bf16 as the
        // reference's softmax output is) go back into the score array
        if (warp < 2) {
          float* sc = warp == 0 ? sc0 : sc1;
          const float x = lane < L ? sc[lane] : -INFINITY;
          const float m = warp_max(x);
          const float e = lane < L ? expf(x - m) : 0.f;
          const float sum = warp_sum_xor(e);
          if (lane < L) sc[lane] = rbf(e / sum);
        }
        m2_csync();
        float outv = 0.f;
        const int h = (tid >> 7) & 1, d = tid & 127;
        if (tid < 256) {
          const float* sc = h == 0 ? sc0 : sc1;
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < M2_ATT_FAST_L; ++j)
            if (j < L) acc = fmaf(sc[j], bf2f(Vs[j * 128 + d]), acc);
          outv = rbf(acc);
        }
        const float other = __shfl_xor_sync(0xffffffffu, outv, 1);
        if (tid < 256 && !(lane & 1))
          st_slot(reinterpret_cast<u64*>(p.Y) + (((size_t)t * heads + (2 * kvh + h)) * 128 + d >> 1), pack2(outv, other), tag);
        if (unit + (int)gridDim.x < n_units) m2_csync();
        continue;
      }
      // ---- general path: rows [j_lo, j_hi) (the whole context when the row is not split) ----
      auto row = [&](const bf16* base, const bf16* cur, const uint2& pre, int j) -> uint2 {
        if (j == pos) return *reinterpret_cast<const uint2*>(cur + 4 * lane);
        if (ns == 1 && j == warp) return pre;
        return __ldcg(reinterpret_cast<const uint2*>(base + (size_t)j * 128 + 4 * lane));
      };
      for (int j0 = j_lo + warp; j0 < j_hi; j0 += M2_ATT_U * MEGA_WARPS) {
        uint2 ku[M2_ATT_U];
#pragma unroll
        for (int q = 0; q < M2_ATT_U; ++q) {
          const int j = j0 + q * MEGA_WARPS;
          ku[q] = j < j_hi ? row(kbase, kcur, kpre, j) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < M2_ATT_U; ++q) {
          const int j = j0 + q * MEGA_WARPS;
          if (j >= j_hi) break;
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
      m2_csync();
      if (warp < 2) {
        // softmax of head `warp` over the WHOLE context: the maximum and the normaliser are exchanged between the splits of
        // the item (tagged 8-byte slots, as the activations), so every probability is exp(s - global max) / global sum rounded
        // to bf16 exactly as in the unsplit kernel; the sum adds the splits' partial sums in split order
        float* sc = warp == 0 ? sc0 : sc1;
        float m = -INFINITY;
        for (int j = j_lo + lane; j < j_hi; j += 32) m = fmaxf(m, sc[j]);
        m = warp_max(m);
        if (ns > 1) {
          if (lane == 0) st_slot(xg + (size_t)split * M2_XCHG_SLOTS + warp, __float_as_uint(m), tag);
          float mo = -INFINITY;
          if (lane < ns) {
            unsigned tries = 0;
            for (;;) {
              const u64 sl = ld_slot(xg + (size_t)lane * M2_XCHG_SLOTS + warp);
              if (slot_tag(sl) == tag) { mo = __uint_as_float(slot_val(sl)); break; }
              if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 6100000 + (int)gs.epoch); break; }
            }
          }
          m = warp_max(mo);
        }
        float sum = 0.f;
        for (int j = j_lo + lane; j < j_hi; j += 32) {
          const float e = expf(sc[j] - m);
          sc[j] = e;
          sum += e;
        }
        sum = warp_sum_xor(sum);
        if (ns > 1) {
          if (lane == 0) st_slot(xg + (size_t)split * M2_XCHG_SLOTS + 2 + warp, __float_as_uint(sum), tag);
          float so = 0.f;
          if (lane < ns) {
            unsigned tries = 0;
            for (;;) {
              const u64 sl = ld_slot(xg + (size_t)lane * M2_XCHG_SLOTS + 2 + warp);
              if (slot_tag(sl) == tag) { so = __uint_as_float(slot_val(sl)); break; }
              if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 6200000 + (int)gs.epoch); break; }
            }
          }
          sum = 0.f;
          for (int q = 0; q < ns; ++q) sum += __shfl_sync(0xffffffffu, so, q);      // split order
        }
        for (int j = j_lo + lane; j < j_hi; j += 32) sc[j] = rbf(sc[j] / sum);
      }
      m2_csync();
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j0 = j_lo + warp; j0 < j_hi; j0 += M2_ATT_U * MEGA_WARPS) {
        uint2 vu[M2_ATT_U];
#pragma unroll
        for (int q = 0; q < M2_ATT_U; ++q) {
          const int j = j0 + q * MEGA_WARPS;
          vu[q] = j < j_hi ? row(vbase, vcur, vpre, j) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < M2_ATT_U; ++q) {
          const int j = j0 + q * MEGA_WARPS;
          if (j >= j_hi) break;
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
      m2_csync();
      {
        float outv = 0.f;
        const int h = (tid >> 7) & 1, d = tid & 127;
        if (tid < 256) {
          float acc = 0.f;
#pragma unroll
          for (int w = 0; w < MEGA_WARPS; ++w) acc += red[(w * 2 + h) * 128 + d];
          if (ns > 1) {
            // partial sums of the splits: split s > 0 publishes its 2 x 128 values, split 0 adds them in split order
            if (split != 0) {
              st_slot(xg + (size_t)split * M2_XCHG_SLOTS + 4 + tid, __float_as_uint(acc), tag);
            } else {
              for (int q = 1; q < ns; ++q) {
                unsigned tries = 0;
                for (;;) {
                  const u64 sl = ld_slot(xg + (size_t)q * M2_XCHG_SLOTS + 4 + tid);
                  if (slot_tag(sl) == tag) { acc += __uint_as_float(slot_val(sl)); break; }
                  if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 6300000 + (int)gs.epoch); break; }
                }
              }
            }
          }
          outv = rbf(acc);
        }
        const float other = __shfl_xor_sync(0xffffffffu, outv, 1);
        if (tid < 256 && !(lane & 1) && split == 0)
          st_slot(reinterpret_cast<u64*>(p.Y) + (((size_t)t * heads + (2 * kvh + h)) * 128 + d >> 1), pack2(outv, other), tag);
      }
      m2_csync();
    }
  }
  m2_csync();
  m2_arrive(gs, p.flags);
  if (prof_on) prof2(a, 7);
  return m2_pack(gs);
}

// ---------------------------------------------------------------------------------------------------
// 8 consecutive bf16 values -> 4 tagged slots (32 bytes)
__device__ __forceinline__ void m2_store_row8(u64* dst_slots, const uint4& v, uint32_t tag) {
  st_slot2(dst_slots, v.x, v.y, tag);
  st_slot2(dst_slots + 2, v.z, v.w, tag);
}

__device__ __noinline__ void m2_sample(const SampleArgs& sa, int b, SampleSmem& sm) { sample_row_body(sa, b, sm); }
__device__ __noinline__ void m2_finish(const M2Args& a, const M2Phase& p, uint32_t* s_codes, const uint32_t tag) {
  EmbTable tab{};
  for (int i = 0; i < a.n_ac; ++i) tab.e[i] = a.cp_emb[i];
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    frame_finish_row(a.fs, tab, a.codec_emb, a.step_input, a.H, a.B, a.n_ac, b, s_codes);
    // the talker input of this row, re-read by the threads that wrote it, as tagged slots
    for (int c = threadIdx.x * 8; c < a.H; c += blockDim.x * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(a.step_input + (size_t)b * a.H + c);
      m2_store_row8(reinterpret_cast<u64*>(p.Y) + (((size_t)b * a.H + c) >> 1), v, tag);
    }
  }
}

constexpr int M2_MAX_PHASES = 600;

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_frames_mega2_kernel(const M2Args args) {
  extern __shared__ __align__(128) unsigned char m2_smem[];
  __shared__ M2Args sa;
  __shared__ uint32_t s_codes[16];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&args);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sa);
    for (int i = threadIdx.x; i < (int)(sizeof(M2Args) / 4); i += MEGA_THREADS) dst[i] = src[i];
    if (threadIdx.x == 0) s_prof2_idx = g_prof2_idx;
    // the phase program -> shared memory
    const uint4* ps = reinterpret_cast<const uint4*>(args.prog);
    uint4* pd = reinterpret_cast<uint4*>(m2_smem);
    for (int i = threadIdx.x; i < args.n_ph * (int)(sizeof(M2Phase) / 16); i += MEGA_THREADS) pd[i] = ps[i];
  }
  __syncthreads();
  const M2Args& a = sa;
  const M2Phase* prog = reinterpret_cast<const M2Phase*>(m2_smem);
  unsigned char* work = m2_smem + (((size_t)a.n_ph * sizeof(M2Phase) + 127) & ~(size_t)127);
  M2Sync gs{a.bar, a.err, 0u, gridDim.x, false, nullptr, nullptr};
  const int B = a.B;
  if (a.bench_barriers > 0) {
    m2_arrive(gs, PF_ARRIVE_REL);
    for (int i = 0; i < a.bench_barriers; ++i) {
      m2_wait(gs, a.bench_barriers & 1 ? PF_WAIT_ACQ : 0);
      m2_arrive(gs, a.bench_barriers & 1 ? PF_ARRIVE_REL : 0);
    }
    return;
  }
  const uint32_t tag0 = __ldcg(a.tag_ctr);
  const bool prof_any = M2_PROF_ENABLED && a.prof != nullptr;
  uint32_t seq = 0;
  bool stop = false;
  m2_arrive(gs, PF_ARRIVE_REL);       // every phase waits for its predecessor's arrive; this is the first phase's
  for (int frame = 0; frame < a.n_frames && !stop; ++frame) {
    for (int i = 0; i < a.n_ph; ++i) {
      const M2Phase& p = prog[i];
      seq += 1u;
      if (prof_any && a.prof_mode == 2) {
        gs.arr = frame == 1 ? a.prof + (size_t)i * 4 * gridDim.x : nullptr;
        gs.retries = frame == 1 ? reinterpret_cast<unsigned*>(a.prof + (size_t)a.n_ph * 4 * gridDim.x) + i : nullptr;
        if (frame == 1 && i == 0 && threadIdx.x == 0) {
          unsigned smid;
          asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          reinterpret_cast<unsigned*>(a.prof + (size_t)a.n_ph * 4 * gridDim.x)[1024 + blockIdx.x] = smid;
        }
      }
      const uint32_t tag = tag0 + seq;
      // next skinny-GEMM phase (for the L2 prefetch of its weights)
      // HBM is idle while the 64 attention CTAs work and during the short o_proj phase that follows, so the gate/up
      // rows (the largest matrix of a layer) are requested at the START of the attention phase, two phases ahead;
      // o_proj itself then prefetches nothing.  (Measured before this change: the talker gate/up phase spent 5.7 us
      // in its wait while the 50 MB prefetch issued at the end of o_proj drained, then streamed from L2.)
      // (the prefetch plan is resolved on the host: M2Phase::next_gemv)
      const int nxi = p.next_gemv;
      const M2Phase* nx = (nxi >= 0 && (nxi > i || frame + 1 < a.n_frames)) ? &prog[nxi] : nullptr;
      switch (p.kind) {
        case M2_GEMV: m2_unpack(gs, m2_gemv_dispatch(a, p, nx, work, gs, tag)); break;
        case M2_ATTN:
          m2_prefetch(a, nx);
          m2_unpack(gs, m2_attn(a, p, work, gs, tag));
          break;
        case M2_PROLOGUE: {
          m2_wait(gs, p.flags);
          if (frame > 0 && a.do_sample) {
            // stop early once every row has sampled EOS (uniform decision: all CTAs read the same flags)
            int active = 0;
            for (int b = 0; b < B; ++b) active += __ldcg(a.fs.done + b) ? 0 : 1;
            if (active == 0) stop = true;
          }
          if (!stop && blockIdx.x == 0)
            for (int k = threadIdx.x; k < a.n_ac * B; k += MEGA_THREADS) a.fs.amax[k] = 0ull;
          __syncthreads();
          m2_arrive(gs, p.flags);
        } break;
        case M2_GATHER: {
          // no projection (talker hidden == CP hidden): the gathered rows are the layer input
          m2_wait(gs, p.flags);
          if (blockIdx.x == 0) {
            const int K8 = p.K >> 3, T = p.T, g = p.g;
            for (int k = threadIdx.x; k < T * K8; k += MEGA_THREADS) {
              const int t = k / K8, qq = k - t * K8;
              const bf16* src;
              if (g == 0) {
                const int b = (t >> 1) + p.pos_add;
                src = (t & 1) ? p.aux2 + (size_t)__ldcg(a.fs.cur_tok + b) * p.K : a.fs.last_hidden + (size_t)b * p.K;
              } else {
                src = p.aux2 + (size_t)argmax_key_index(__ldcg(a.fs.amax + (size_t)(g - 1) * B + t)) * p.K;
              }
              m2_store_row8(reinterpret_cast<u64*>(p.Y) + (((size_t)t * p.ldy + qq * 8) >> 1), ldcg16(src + qq * 8), tag);
            }
            if (g == 0) { if ((int)threadIdx.x < (T >> 1)) a.fs.frame_codes[(threadIdx.x + p.pos_add) * 16] = __ldcg(a.fs.cur_tok + threadIdx.x + p.pos_add); }
            else if (threadIdx.x < B)
              a.fs.frame_codes[threadIdx.x * 16 + g] = argmax_key_index(__ldcg(a.fs.amax + (size_t)(g - 1) * B + threadIdx.x));
          }
          __syncthreads();
          m2_arrive(gs, p.flags);
        } break;
        case M2_FINISH: {
          m2_wait(gs, p.flags);
          m2_finish(a, p, s_codes, tag);
          __syncthreads();
          m2_arrive(gs, p.flags);
        } break;
        case M2_COPYIN: {
          m2_wait(gs, p.flags);
          for (int k = blockIdx.x * MEGA_THREADS + threadIdx.x; k < B * (a.H >> 3); k += gridDim.x * MEGA_THREADS)
            m2_store_row8(reinterpret_cast<u64*>(p.Y) + (size_t)k * 4, ldcg16(reinterpret_cast<const uint4*>(p.X) + k), tag);
          __syncthreads();
          m2_arrive(gs, p.flags);
        } break;
        case M2_SAMPLE: {
          SampleSmem& sm = *reinterpret_cast<SampleSmem*>(work);
          m2_wait(gs, p.flags);
          for (int b = blockIdx.x; b < B; b += gridDim.x) m2_sample(a.smp, b, sm);
          __syncthreads();
          m2_arrive(gs, p.flags);
        } break;
        default: break;
      }
      if (stop) break;
    }
  }
  // the tag counter of the session: read by every CTA at the start of the NEXT launch
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.tag_ctr = tag0 + seq;
}

// shared-memory size: program + the largest work area; 0 when a phase does not fit the static limits
static size_t mega2_smem_bytes(const q3_model_desc& d, int B, int max_seq, int grid, int n_ph) {
  size_t red_max = 0;
  bool ok = true;
  auto phase = [&](int N, int T, bool dual) {
    const int units = N / 8, per_cta = (units + grid - 1) / grid, tiles = (per_cta + 1) / 2;
    if (tiles > MEGA_MAX_TILES || T > MEGA_TMAX || N % 8 != 0) ok = false;
    const int NT = (T + 7) / 8;
    red_max = std::max(red_max, (size_t)NT * (dual ? 2 : 1) * 8 * (256 * tiles + 4) * 4);
  };
  const int nh = (d.heads + 2 * d.kv_heads) * 128, cnh = (d.cp_heads + 2 * d.cp_kv_heads) * 128;
  // code-predictor pass 0 has two tokens per row: batches above 8 run it in row groups of 8 (m2_build_program)
  const int T0 = 2 * std::min(B, 8), T1 = B;
  phase(nh, B, false); phase(d.hidden, B, false); phase(d.inter, B, true); phase(d.codec_vocab, B, false);
  phase(d.cp_hidden, std::max(T0, T1), false); phase(cnh, std::max(T0, T1), false); phase(d.cp_inter, std::max(T0, T1), true);
  phase(d.cp_vocab, B, false);
  if (!ok || n_ph > M2_MAX_PHASES) return 0;
  size_t m = sizeof(SampleSmem);
  m = std::max(m, (size_t)M2_RED_OFF + red_max);
  m = std::max(m, (size_t)(2 * (std::max(max_seq, d.cp_max_seq) + 3) + 256 + 16 * 2 * 128) * 4 + (2 * M2_ATT_FAST_L * 128 + 256) * 2 + 256);
  return (((size_t)n_ph * sizeof(M2Phase) + 127) & ~(size_t)127) + m;
}
