// Persistent frame kernel, fourth generation: the dataflow phases of mega2.cuh with the WEIGHTS STREAMED AHEAD OF THE
// DEPENDENCY CHAIN by a warp-specialised TMA producer (generate_codes src/lib.rs:530-656; talker step
// src/models/talker.rs:716-736; code-predictor frame src/models/code_predictor.rs:320-416).
//
// Why (measured on B200, profiles/r1_mega2_skew_segments.log): a code-predictor phase of the dataflow kernel took
// ~4.5 us against ~0.9 us of weight streaming.  Two of those microseconds sat between the last CTA's arrive and a CTA
// getting past its wait, because every warp first requested its 55-85 KB of weights into registers and the poll, the
// activation loads and everything after them queued behind that data on the SM's ingress.  Weights do not depend on
// activations, so they need not be on the dependent path at all:
//   * a 17th warp (the producer) walks the phase program ahead of the 16 compute warps and copies each 16-row x
//     1024-column weight tile into a shared-memory ring with one cp.async.bulk per row (2 KB each, rows padded to
//     2112 bytes so the fragment loads are bank-conflict free), one full/empty mbarrier pair per ring slot.  The ring
//     (5 slots x 33 KB) holds about two code-predictor phases, i.e. the weights of phase j+1 and j+2 land while phase j
//     waits on its barrier, loads its activations and combines;
//   * the compute warps read mma.sync A fragments with two LDS.128 per 16x32 sub-tile (the k-permutation trick of
//     mega.cuh carries over: a lane's 16 contiguous bytes ARE its fragments when the activations use the same
//     permutation), one mbarrier wait per 33 KB slot (the round-1 ring kernel, mega3.cuh, paid one per 17 KB and a
//     sleeping poll loop, and lost), and K is split over the 16 warps inside a slot;
//   * partial sums are combined PER 16-ROW TILE in a fixed 4 x 4 tree (four threads per output, each adding four warps
//     in order, then a butterfly): the combine buffer shrinks from 49-98 KB to 19-37 KB, which is what makes room for
//     the ring, and the tree does not depend on the batch size (rows of a batch equal their batch-1 runs bit for bit);
//   * phase descriptors are staged by the producer as well (8-slot ring), so no shared memory is spent on the 88 KB
//     phase program;
//   * everything else -- tagged activation slots, the fence-free hint barrier, consumer-side RMSNorm, the attention
//     phase, frame finish and the sampler -- is mega2.cuh's, bit-for-bit the same arithmetic.
// Models with a skinny-GEMM K that is not a multiple of 1024 stay on the mega2.cuh kernel.
#pragma once
#include "mega2.cuh"

constexpr int M4_THREADS = MEGA_THREADS + 64;        // 16 compute warps + the weight producer warp + the descriptor stager warp
                                                     // (17 to 20 warps cost the same: 5 warps per scheduler -> 96 registers per thread)
constexpr int M4_KC = 1024;                          // columns per ring slot
constexpr int M4_ROW_BYTES = M4_KC * 2 + 64;         // 2112: rows g and g+1 start 64 bytes apart modulo 128
constexpr int M4_SLOT_BYTES = 16 * M4_ROW_BYTES;     // 33792
constexpr int M4_MAX_SLOTS = 6;
constexpr int M4_DESCS = 16;                         // descriptor ring slots
constexpr int M4_RED_WS = 18;                        // floats per (column, warp): 16 rows + 2 (conflict-free 4-way split reads)
constexpr int M4_RED_CS = 16 * M4_RED_WS + 4;        // 292 floats per token column (conflict-free fragment writes)
constexpr unsigned M4_MBAR_SPIN = 1u << 20;          // mbarrier polls before the watchdog gives up

__device__ __forceinline__ uint32_t m4_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m4_mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ bool m4_mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void m4_mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
// Release of a ring slot by lane 0 on behalf of its warp.  `z` MUST be computed from the registers the slot's fragment loads
// wrote (see the call site): ptxas schedules an mbarrier.arrive as soon as its operands are ready and is free to hoist it
// above the MMAs that consume the fragments -- with the LDS still in flight -- so the producer could refill the slot under
// them.  That race showed as run-to-run differences once the ring was deep enough for the producer to be waiting on that
// very slot (5 slots, 16 tokens); tools/cp_repeat.py and tools/cp_bisect.py are the hunt.  z is always 0 at run time (it is
// masked with a kernel argument that is 0 in every decode launch), which the compiler cannot know.
__device__ __forceinline__ void m4_mbar_release_slot(uint32_t addr, uint32_t z) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr + z) : "memory");
}
__device__ __forceinline__ void m4_mbar_expect(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m4_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ uint4 m4_lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// CTA-wide bookkeeping in (static) shared memory
struct M4Shared {
  M2Args a;
  M2Phase desc[M4_DESCS];
  unsigned long long full[M4_MAX_SLOTS], empty[M4_MAX_SLOTS], desc_full[M4_DESCS];
  volatile unsigned done;       // the compute warps have finished every phase < done
  volatile unsigned decided;    // q + 1 of the last PROLOGUE phase whose stop decision has been taken
  volatile int stop;            // every row has sampled EOS: the producer leaves its loop
  volatile int dead;            // a watchdog fired somewhere in this CTA
  unsigned long long* prof_arr;   // prof_mode 2: arrival stamps of the current phase, or null
  unsigned* prof_retries;
  uint32_t codes[16];
};

// per-thread mutable state of a compute thread.  Phase functions take and return it PACKED in 64 bits (epoch | slot << 32 |
// par << 40 | dead << 41): a struct passed to or returned from a non-inlined function lives in local memory, and thread 0
// pays a chain of LDL/STL round trips per phase (measured on the dataflow kernel, mega2.cuh).
struct M4State {
  unsigned epoch;      // grid barrier epoch
  unsigned slot, par;  // ring position of the next slot to consume and its full-barrier parity
  bool dead;
};
__device__ __forceinline__ unsigned long long m4_pack(const M4State& st) {
  return (unsigned long long)st.epoch | ((unsigned long long)(st.slot & 0xffu) << 32) | ((unsigned long long)(st.par & 1u) << 40) |
         ((unsigned long long)(st.dead ? 1u : 0u) << 41);
}
__device__ __forceinline__ M4State m4_unpack(unsigned long long r) {
  M4State st;
  st.epoch = (unsigned)r;
  st.slot = (unsigned)(r >> 32) & 0xffu;
  st.par = (unsigned)(r >> 40) & 1u;
  st.dead = ((r >> 41) & 1ull) != 0ull;
  return st;
}

// One lane per warp polls the barrier, the warp then re-converges: 512 threads polling one mbarrier serialise on the
// shared-memory atomic unit (measured: ~0.5 us per slot with every thread polling, against ~0.1 us of loads and MMAs).
// The __syncwarp orders the other lanes' shared-memory reads of the slot behind lane 0's acquire of the barrier.
__device__ __forceinline__ bool m4_wait_full(M4Shared& sh, uint32_t bar, uint32_t parity) {
  bool ok = true;
  if ((threadIdx.x & 31) == 0) {
    unsigned it = 0;
    while (!m4_mbar_try(bar, parity)) {
      if (++it > M4_MBAR_SPIN) {
        if (sh.a.err != nullptr) atomicCAS(sh.a.err, 0, 7000000);
        sh.dead = 1;
        ok = false;
        break;
      }
    }
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
}

// profiling: a DURATION (ns) recorded under tag 10 -- time thread 0 of block 0 spent waiting for ring slots in a phase
__device__ __forceinline__ void prof2_dur(const M2Args& a, unsigned long long ns) {
  if (M2_PROF_ENABLED && a.prof != nullptr && a.prof_mode != 2 && blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned i = s_prof2_idx++;
    if ((int)i < a.prof_cap) a.prof[i] = (ns << 8) | 10ull;
    g_prof2_idx = i + 1;
  }
}
__device__ __forceinline__ unsigned long long m4_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// hint wait of a phase + the "phases < q are finished" mark the producer recycles descriptor slots by: after the
// barrier inside m2_wait every compute thread has left phase q-1
__device__ __forceinline__ void m4_wait(M4Shared& sh, M2Sync& gs, int flags, unsigned q) {
  m2_wait(gs, flags);
  if (threadIdx.x == 0) sh.done = q;
}

// ---------------------------------------------------------------------------------------------------
// Combine of up to TP 16-row tiles in one pass: red holds, per tile, the 16 warps' partial sums as
// [nt][m][token col][warp][row] f32 (M4_RED_CS floats per column, M4_RED_WS per warp).  Work item idx = tid + j*512 of tile
// tt: output o = idx >> 2 (row = o & 15, token = o >> 4), part = idx & 3 adds warps 4*part .. 4*part+3 in order, then a
// two-step butterfly over the four parts: the same tree for every batch size.  Fused epilogue and tagged stores as m2_tail.
template <bool DUAL, int NT, int TP>
__device__ __forceinline__ void m4_tile_tail(const M2Phase& p, const float* red0, const float (&rres)[TP][NT], const int n00,
                                             const int tp, const int r1, const uint32_t tag) {
  constexpr int NM = DUAL ? 2 : 1;
  constexpr int RED_FLOATS = NT * NM * 8 * M4_RED_CS;
  const int tid = threadIdx.x, lane = tid & 31, T = p.T;
  const int epi = p.epi, yf = p.yf, ldy = p.ldy, pN = p.N;
  u64* const Y64 = reinterpret_cast<u64*>(p.Y);
  float* const Yf = p.Yf;
  u64* const amax = p.amax;
  const bf16* const bias = p.aux;
#pragma unroll
  for (int tt = 0; tt < TP; ++tt) {
    if (tt >= tp) break;
    const float* red = red0 + tt * RED_FLOATS;
    const int n0 = n00 + (tt << 4);
#pragma unroll
    for (int it = 0; it < NT; ++it) {
      const int idx = tid + it * MEGA_THREADS;
      const int o = idx >> 2, part = idx & 3;
      const int row = o & 15, t = o >> 4;              // t < 8 * NT
      const int nt = t >> 3, col = t & 7;
      const int n = n0 + row;
      const bool valid = t < T && n < r1;
      const float* rb = red + (size_t)((nt * NM) * 8 + col) * M4_RED_CS + (part * 4) * M4_RED_WS + row;
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        v0 += rb[w * M4_RED_WS];
        if (DUAL) v1 += rb[8 * M4_RED_CS + w * M4_RED_WS];
      }
      v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
      v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
      if (DUAL) {
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
      }
      float outv = 0.f;
      u64 key = 0ull;
      if (valid) {
        const float v = rbf(v0);
        switch (epi) {
          case EPI_STORE: outv = v; break;
          case EPI_BIAS: outv = rbf(v + bf2f(bias[n])); break;
          case EPI_RESIDUAL: outv = rbf(rres[tt][it] + v); break;
          case EPI_O_H1: outv = rres[tt][it] + v; break;    // x + attn_out, un-rounded (fused_residual_rmsnorm.cu:60-65)
          case EPI_SWIGLU: outv = rbf(rbf(silu_f(v)) * rbf(v1)); break;
          case EPI_LOGITS: {
            if (Yf != nullptr && part == 0) Yf[(size_t)t * pN + n] = v;
            key = argmax_key(v, n);
          } break;
          default: break;
        }
      }
      if (yf == XF_BF16T) {
        const float other = __shfl_down_sync(0xffffffffu, outv, 4);       // row + 1 of the same token
        if (valid && part == 0 && !(row & 1)) st_slot(Y64 + (((size_t)t * ldy + n) >> 1), pack2(outv, other), tag);
      } else if (yf == XF_F32T) {
        if (valid && part == 0) st_slot(Y64 + (size_t)t * ldy + n, __float_as_uint(outv), tag);
      }
      if (epi == EPI_LOGITS && amax != nullptr) {
        // a warp covers 8 consecutive rows of ONE token: one atomic per warp
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
          const u64 other = __shfl_xor_sync(0xffffffffu, key, sft);
          key = other > key ? other : key;
        }
        if (lane == 0 && key != 0ull) atomicMax(amax + t, key);
      }
    }
  }
}

// residual inputs of this thread's work items for the tiles that start at row n00, n00 + 16, .. (rows owned by this CTA,
// written by this CTA two or three phases ago: no tag check, as in m2_load_residual)
template <int NT, int TP>
__device__ __forceinline__ void m4_load_residual(const M2Phase& p, int n00, int r1, float (&rres)[TP][NT]) {
  const int tid = threadIdx.x;
  const int epi = p.epi, T = p.T, rf = p.rf, ldr = p.ldr;
  const u64* const R64 = reinterpret_cast<const u64*>(p.R);
  const bool has_r = epi == EPI_RESIDUAL || epi == EPI_O_H1;
#pragma unroll
  for (int tt = 0; tt < TP; ++tt)
#pragma unroll
    for (int it = 0; it < NT; ++it) {
      rres[tt][it] = 0.f;
      const int o = (tid + it * MEGA_THREADS) >> 2;
      const int row = o & 15, t = o >> 4;
      const int n = n00 + (tt << 4) + row;
      if (has_r && t < T && n < r1) {
        if (rf == XF_F32T) {
          const u64 sl = ld_slot(R64 + (size_t)t * ldr + n);
          rres[tt][it] = rbf(__uint_as_float(slot_val(sl)));            // h1 as the reference stores it: bf16(x + attn)
        } else {
          const u64 sl = ld_slot(R64 + (((size_t)t * ldr + n) >> 1));
          rres[tt][it] = (n & 1) ? bf_hi(slot_val(sl)) : bf_lo(slot_val(sl));
        }
      }
    }
}

// ---------------------------------------------------------------------------------------------------
// One skinny-GEMM phase on the compute warps.  Weights come from the ring in the order the producer streams them:
// tile-major, then matrix (gate, up), then k-chunk (1024 columns).  The activations of KCH chunks (one GROUP) are fetched
// in one batch of tag-verified loads and stay in registers; K == groups * KCH * 1024.  NORM / DUAL phases have one group
// (K <= 2048: the scale needs the whole row); the down projections (K = 3072 / 6144) use KCH = 3 or 6 and, where the
// registers do not reach (16 tokens x 6144), two groups.  (The first version streamed the activation chunks of K > 2048
// one at a time behind the MMAs: each chunk exposed an L2 round trip, 3.2 us for the three chunks of a code-predictor
// down projection against 1.3 us for one batch of loads.)
// smem work area: scale[16] | part[16][16] | red[1 or 2] (double-buffered over tiles when it fits: one barrier per tile).
template <bool DUAL, int NT, int XF, bool NORM, int KCH>
__device__ __noinline__ unsigned long long m4_gemv(M4Shared& sh, const M2Phase& p, unsigned char* ring, unsigned char* work,
                                                   const unsigned long long st_packed, const uint32_t tag, const unsigned q) {
  const M2Args& a = sh.a;
  M4State st = m4_unpack(st_packed);
  constexpr int NM = DUAL ? 2 : 1;
  float* part_s = reinterpret_cast<float*>(work) + 16;
  float* red0 = reinterpret_cast<float*>(work + M2_RED_OFF);
  constexpr int RED_FLOATS = NT * NM * 8 * M4_RED_CS;
  const int K = p.K;
  const bool prof_on = M2_PROF_ENABLED && a.prof != nullptr;
  M2Sync gs{a.bar, a.err, st.epoch, gridDim.x, st.dead, prof_on ? sh.prof_arr : nullptr, prof_on ? sh.prof_retries : nullptr};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const uint32_t xtag = tag - 1u;
  int r0, r1;
  mega_row_range(p.N, r0, r1);
  if (r1 <= r0) {
    m4_wait(sh, gs, p.flags, q);
    m2_arrive(gs, p.flags);
    st.epoch = gs.epoch; st.dead = gs.dead;
    return m4_pack(st);
  }
  if (prof_on) prof2(a, 1);
  const int n_tiles = (r1 - r0 + 15) >> 4;
  const int groups = (NORM || DUAL) ? 1 : (K >> 10) / KCH;
  const int koff0 = warp * 32 + 8 * tg;              // element offset of this lane inside a 512-column half chunk
  // ---- before the wait: residual rows of the first pass, norm weights ----
  constexpr int TP = NT == 1 ? 2 : 1;                // tiles per combine pass (two when the work area holds two tiles' partial sums)
  float rres[TP][NT];
  m4_load_residual<NT, TP>(p, r0, r1, rres);
  uint4 wn[NORM ? KCH : 1][2];
  if constexpr (NORM) {
#pragma unroll
    for (int c = 0; c < KCH; ++c)
#pragma unroll
      for (int u = 0; u < 2; ++u) wn[c][u] = *reinterpret_cast<const uint4*>(p.aux + koff0 + (c * 2 + u) * 512);
  }
  m4_wait(sh, gs, p.flags, q);
  if (prof_on) prof2(a, 2);
  if (prof_on) m2_stamp(gs, 0);
  const char* xrow[NT];
  m2_token_rows<NT, XF>(a, p, g, xrow);
  // ---- activations of one group: one batch of tag-verified loads ----
  uint4 xv[KCH][2][NT];
  float sq[NT];
  auto load_group = [&](int grp) {
    unsigned tries = 0;
    const int kbase = grp * (KCH * 1024) + koff0;
    for (;;) {
      uint32_t bad = 0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) sq[nt] = 0.f;
#pragma unroll
      for (int c = 0; c < KCH; ++c)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            xv[c][u][nt] = make_uint4(0, 0, 0, 0);
            if (xrow[nt] != nullptr) xv[c][u][nt] = m2_load_x8<XF>(xrow[nt], kbase + (c * 2 + u) * 512, xtag, bad, sq[nt]);
          }
      if (XF == XF_GATHER || !__any_sync(0xffffffffu, bad != 0)) break;
      if (M2_PROF_ENABLED && gs.retries != nullptr && lane == 0) atomicAdd(gs.retries, 1u);
      if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 5000000 + (int)gs.epoch); break; }
    }
  };
  load_group(0);
  if constexpr (NORM) {
    float xsc[NT];
    m2_row_scales<NT>(a, part_s, sq, K, xsc);
    const bool write_xn = p.xn_out != nullptr && blockIdx.x == 0;
#pragma unroll
    for (int c = 0; c < KCH; ++c)
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          xv[c][u][nt] = m2_apply_norm(xv[c][u][nt], wn[c][u], xsc[nt]);
          if (write_xn && xrow[nt] != nullptr)
            *reinterpret_cast<uint4*>(p.xn_out + (size_t)(nt * 8 + g) * K + koff0 + (c * 2 + u) * 512) = xv[c][u][nt];
        }
  }
  if (prof_on) prof2(a, 3);
  if (prof_on) m2_stamp(gs, 1);
  // ---- the slot stream ----
  const uint32_t ring_s = m4_smem(ring), full_s = m4_smem(sh.full), empty_s = m4_smem(sh.empty);
  const uint32_t frag_off = (uint32_t)(g * M4_ROW_BYTES + warp * 64 + tg * 16);
  const unsigned n_slots = (unsigned)a.m4_slots;
  unsigned slot = st.slot, par = st.par;
  bool alive = !sh.dead;
  unsigned long long t_slots = 0ull;
  const uint32_t zmask = (uint32_t)a.bench_barriers;      // 0 in every decode launch; opaque to the compiler (m4_mbar_release_slot)
#pragma unroll 1
  for (int t0 = 0; t0 < n_tiles; t0 += TP) {
    const int tp = min(TP, n_tiles - t0);
    float acc[TP][NM][NT][4];
#pragma unroll
    for (int tt = 0; tt < TP; ++tt)
#pragma unroll
      for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[tt][m][nt][i] = 0.f;
#pragma unroll
    for (int tt = 0; tt < TP; ++tt) {
      if (tt < tp) {
#pragma unroll
        for (int m = 0; m < NM; ++m) {
#pragma unroll 1
          for (int grp = 0; grp < groups; ++grp) {
            if (grp > 0 || (t0 + tt > 0 && groups > 1)) load_group(grp);      // (only K = 6144 has more than one group)
#pragma unroll
            for (int c = 0; c < KCH; ++c) {
              if (prof_on) {
                const unsigned long long tw0 = m4_now();
                if (alive) alive = m4_wait_full(sh, full_s + slot * 8, par);
                t_slots += m4_now() - tw0;
              } else {
                if (alive) alive = m4_wait_full(sh, full_s + slot * 8, par);
              }
              const uint32_t fa = ring_s + slot * M4_SLOT_BYTES + frag_off;
              uint4 wl[2], wh[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                wl[u] = m4_lds128(fa + u * 1024);
                wh[u] = m4_lds128(fa + u * 1024 + 8 * M4_ROW_BYTES);
              }
#pragma unroll
              for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                  const uint4 x4 = xv[c][u][nt];
                  mma_bf16_16816(acc[tt][m][nt], wl[u].x, wh[u].x, wl[u].y, wh[u].y, x4.x, x4.y);
                  mma_bf16_16816(acc[tt][m][nt], wl[u].z, wh[u].z, wl[u].w, wh[u].w, x4.z, x4.w);
                }
              {
                // the release waits for every lane's fragment loads: one warp-wide OR over a word of each LDS.128 result
                const uint32_t z = __reduce_or_sync(0xffffffffu, (wl[0].x ^ wh[0].x ^ wl[1].x ^ wh[1].x) & zmask);
                if (lane == 0) m4_mbar_release_slot(empty_s + slot * 8, z);
              }
              if (++slot == n_slots) { slot = 0; par ^= 1u; }
            }
          }
        }
      }
    }
    // partial sums of this pass -> red (one buffer per tile of the pass)
#pragma unroll
    for (int tt = 0; tt < TP; ++tt)
      if (tt < tp) {
#pragma unroll
        for (int m = 0; m < NM; ++m)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float* r = red0 + tt * RED_FLOATS + (size_t)((nt * NM + m) * 8 + 2 * tg) * M4_RED_CS + warp * M4_RED_WS + g;
            r[0] = acc[tt][m][nt][0]; r[M4_RED_CS] = acc[tt][m][nt][1]; r[8] = acc[tt][m][nt][2]; r[M4_RED_CS + 8] = acc[tt][m][nt][3];
          }
      }
    const bool last = t0 + tp >= n_tiles;
    if (prof_on && last) m2_stamp(gs, 2);
    m2_csync();
    if (prof_on && last) { prof2_dur(a, t_slots); prof2(a, 4); }
    m4_tile_tail<DUAL, NT, TP>(p, red0, rres, r0 + (t0 << 4), tp, r1, tag);
    if (!last) {
      m4_load_residual<NT, TP>(p, r0 + ((t0 + tp) << 4), r1, rres);
      m2_csync();                 // the next pass's partial sums wait for this pass's combine reads
    }
  }
  m2_csync();
  m2_arrive(gs, p.flags);
  if (prof_on) prof2(a, 5);
  st.epoch = gs.epoch; st.dead = gs.dead; st.slot = slot; st.par = par;
  return m4_pack(st);
}

// which (K, dual, norm, input format) combinations the ring kernel implements
__host__ __device__ inline bool m4_gemv_supported(int N, int K, bool dual, bool norm, int xf) {
  if (K < 1024 || (K & 1023) != 0 || (N & 7) != 0) return false;
  if (dual) return norm && xf == XF_F32T && K <= 2048;
  if (norm) return xf == XF_BF16T && K <= 2048;
  if (xf == XF_GATHER) return K <= 2048;
  return xf == XF_BF16T;
}

template <int NT>
__device__ __forceinline__ unsigned long long m4_gemv_nt(M4Shared& sh, const M2Phase& p, unsigned char* ring, unsigned char* work,
                                                         const unsigned long long st, const uint32_t tag, const unsigned q) {
  const bool dual = (p.flags & PF_DUAL) != 0, norm = (p.flags & PF_NORM) != 0;
  const int kch = p.K >> 10;
  if (dual) {
    if (kch == 1) return m4_gemv<true, NT, XF_F32T, true, 1>(sh, p, ring, work, st, tag, q);
    return m4_gemv<true, NT, XF_F32T, true, 2>(sh, p, ring, work, st, tag, q);
  }
  if (norm) {
    if (kch == 1) return m4_gemv<false, NT, XF_BF16T, true, 1>(sh, p, ring, work, st, tag, q);
    return m4_gemv<false, NT, XF_BF16T, true, 2>(sh, p, ring, work, st, tag, q);
  }
  if (p.xf == XF_GATHER) {
    if (kch == 1) return m4_gemv<false, NT, XF_GATHER, false, 1>(sh, p, ring, work, st, tag, q);
    return m4_gemv<false, NT, XF_GATHER, false, 2>(sh, p, ring, work, st, tag, q);
  }
  // plain bf16-slot input: the largest group the 96 registers of a thread hold without spilling (the raw 8-byte slots of a
  // group are in flight together, twice the payload): 3 chunks of 8 tokens, 2 chunks of 16
  if (NT == 1 && kch % 3 == 0) return m4_gemv<false, NT, XF_BF16T, false, NT == 1 ? 3 : 1>(sh, p, ring, work, st, tag, q);
  if (kch % 2 == 0) return m4_gemv<false, NT, XF_BF16T, false, 2>(sh, p, ring, work, st, tag, q);
  return m4_gemv<false, NT, XF_BF16T, false, 1>(sh, p, ring, work, st, tag, q);
}
__device__ __forceinline__ unsigned long long m4_gemv_dispatch(M4Shared& sh, const M2Phase& p, unsigned char* ring,
                                                               unsigned char* work, const unsigned long long st, const uint32_t tag,
                                                               const unsigned q) {
  if (p.T <= 8) return m4_gemv_nt<1>(sh, p, ring, work, st, tag, q);
  return m4_gemv_nt<2>(sh, p, ring, work, st, tag, q);
}

// ---------------------------------------------------------------------------------------------------
// frame finish on the 512 compute threads (m2_finish with the compute-warp barrier)
__device__ __noinline__ void m4_finish(const M2Args& a, const M2Phase& p, uint32_t* s_codes, const uint32_t tag, int frame) {
  EmbTable tab{};
  for (int i = 0; i < a.n_ac; ++i) tab.e[i] = a.cp_emb[i];
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    frame_finish_row<SyncCompute512>(a.fs, tab, a.codec_emb, a.step_input, a.H, a.B, a.n_ac, b, s_codes);
    // the talker input of this row, re-read by the threads that wrote it, as tagged slots
    for (int c = threadIdx.x * 8; c < a.H; c += MEGA_THREADS * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(a.step_input + (size_t)b * a.H + c);
      m2_store_row8(reinterpret_cast<u64*>(p.Y) + (((size_t)b * a.H + c) >> 1), v, tag);
    }
  }
}
__device__ __noinline__ void m4_sample(const SampleArgs& sa, int b, SampleSmem& sm) { sample_row_body<SyncCompute512>(sa, b, sm); }

// ---------------------------------------------------------------------------------------------------
// The descriptor stager warp: copies the phase descriptors of the program into the shared-memory descriptor ring, up to
// M4_DESCS phases ahead of the compute warps.  Decoupled from the weight producer: in the first version one warp did
// both, a descriptor was fetched (a global-memory round trip) only after all weight copies of the previous phase had
// been issued, and the compute warps waited ~0.5 us for it at the top of every phase.
__device__ __noinline__ void m4_stager(M4Shared& sh) {
  const M2Args& a = sh.a;
  const int lane = threadIdx.x & 31;
  const uint32_t dfull_s = m4_smem(sh.desc_full);
  unsigned q = 0;
  for (int frame = 0; frame < a.n_frames; ++frame) {
    for (int i = 0; i < a.n_ph; ++i, ++q) {
      // descriptor slot q % M4_DESCS held phase q - M4_DESCS: free once the compute warps have finished it
      unsigned it = 0;
      while ((int)(q - sh.done) >= M4_DESCS) {
        if (sh.stop || sh.dead) return;
        __nanosleep(32);
        if (++it > (1u << 24)) { sh.dead = 1; return; }
      }
      const uint4* src = reinterpret_cast<const uint4*>(a.prog + i);
      uint4* dst = reinterpret_cast<uint4*>(&sh.desc[q % M4_DESCS]);
      if (lane < (int)(sizeof(M2Phase) / 16)) dst[lane] = __ldg(src + lane);
      __syncwarp();
      if (lane == 0) m4_mbar_arrive(dfull_s + (q % M4_DESCS) * 8);
    }
  }
}

// The weight producer warp: weight tiles in program order, as far ahead as the ring allows.
__device__ __noinline__ void m4_producer(M4Shared& sh, unsigned char* ring) {
  const M2Args& a = sh.a;
  const int lane = threadIdx.x & 31;
  const uint32_t ring_s = m4_smem(ring), full_s = m4_smem(sh.full), empty_s = m4_smem(sh.empty),
                 dfull_s = m4_smem(sh.desc_full);
  const unsigned n_slots = (unsigned)a.m4_slots;
  unsigned slot = 0, par = 1;          // empty-barrier parity: a fresh barrier passes a wait on parity 1
  unsigned q = 0;
  for (int frame = 0; frame < a.n_frames; ++frame) {
    for (int i = 0; i < a.n_ph; ++i, ++q) {
      // this phase's descriptor (the producer runs ahead of the compute warps, so the slot is not recycled under it)
      {
        unsigned it = 0;
        while (!m4_mbar_try(dfull_s + (q % M4_DESCS) * 8, (q / M4_DESCS) & 1u)) {
          if (sh.stop || sh.dead) return;
          if (++it > M4_MBAR_SPIN) { sh.dead = 1; return; }
        }
      }
      const M2Phase* dsl = &sh.desc[q % M4_DESCS];
      const int kind = dsl->kind;
      if (kind == M2_GEMV) {
        int r0, r1;
        mega_row_range(dsl->N, r0, r1);
        if (r1 <= r0) continue;
        const int K = dsl->K, kch = K >> 10, nm = (dsl->flags & PF_DUAL) ? 2 : 1;
        const bf16* W0 = dsl->W;
        const bf16* W1 = dsl->W2;
        const int n_tiles = (r1 - r0 + 15) >> 4;
        for (int tile = 0; tile < n_tiles; ++tile) {
          const int n0 = r0 + (tile << 4);
          const int rows = min(16, r1 - n0);
          for (int m = 0; m < nm; ++m) {
            const bf16* Wm = (m == 0 ? W0 : W1) + (size_t)(n0 + (lane & 15)) * K;
            for (int c = 0; c < kch; ++c) {
              unsigned it = 0;
              while (!m4_mbar_try(empty_s + slot * 8, par)) {
                if (sh.stop || sh.dead) return;
                if (++it > M4_MBAR_SPIN) { sh.dead = 1; return; }
              }
              if (lane == 0) m4_mbar_expect(full_s + slot * 8, (uint32_t)rows * (M4_KC * 2));
              __syncwarp();
              if (lane < rows)
                m4_bulk_g2s(ring_s + slot * M4_SLOT_BYTES + lane * M4_ROW_BYTES, Wm + (size_t)c * M4_KC, M4_KC * 2, full_s + slot * 8);
              if (++slot == n_slots) { slot = 0; par ^= 1u; }
            }
          }
        }
      } else if (kind == M2_PROLOGUE) {
        if (frame > 0 && a.do_sample) {
          // the compute warps decide here whether the loop ends; do not stream past that decision
          unsigned it = 0;
          while ((int)(sh.decided - q) <= 0) {
            if (sh.stop || sh.dead) return;
            __nanosleep(32);
            if (++it > (1u << 24)) { sh.dead = 1; return; }
          }
          if (sh.stop) return;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(M4_THREADS, 1) decode_frames_mega4_kernel(const M2Args args) {
  extern __shared__ __align__(128) unsigned char m4_dyn[];
  __shared__ M4Shared sh;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&args);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.a);
    for (int i = threadIdx.x; i < (int)(sizeof(M2Args) / 4); i += M4_THREADS) dst[i] = src[i];
    if (threadIdx.x == 0) {
      s_prof2_idx = g_prof2_idx;
      sh.done = 0u; sh.decided = 0u; sh.stop = 0; sh.dead = 0;
      for (int s = 0; s < M4_MAX_SLOTS; ++s) {
        m4_mbar_init(m4_smem(&sh.full[s]), 1u);
        m4_mbar_init(m4_smem(&sh.empty[s]), MEGA_WARPS);
      }
      for (int s = 0; s < M4_DESCS; ++s) m4_mbar_init(m4_smem(&sh.desc_full[s]), 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  unsigned char* ring = m4_dyn;
  unsigned char* work = m4_dyn + (size_t)args.m4_slots * M4_SLOT_BYTES;
  if (threadIdx.x >= MEGA_THREADS) {
    if (threadIdx.x < MEGA_THREADS + 32) m4_producer(sh, ring);
    else m4_stager(sh);
    return;
  }
  const M2Args& a = sh.a;
  const int B = a.B;
  const uint32_t tag0 = __ldcg(a.tag_ctr);
  const uint32_t dfull_s = m4_smem(sh.desc_full);
  M4State st{0u, 0u, 0u, false};
  if (threadIdx.x == 0) { sh.prof_arr = nullptr; sh.prof_retries = nullptr; }
  {
    M2Sync gs{a.bar, a.err, 0u, gridDim.x, false, nullptr, nullptr};
    m2_arrive(gs, PF_ARRIVE_REL);       // every phase waits for its predecessor's arrive; this is the first phase's
  }
  uint32_t q = 0;
  bool stop = false;
  for (int frame = 0; frame < a.n_frames && !stop; ++frame) {
    for (int i = 0; i < a.n_ph; ++i, ++q) {
      // this phase's descriptor (staged by the producer warp)
      if (!m4_wait_full(sh, dfull_s + (q % M4_DESCS) * 8, (q / M4_DESCS) & 1u)) st.dead = true;
      const M2Phase& p = sh.desc[q % M4_DESCS];
      const uint32_t tag = tag0 + q + 1u;
      if (M2_PROF_ENABLED && a.prof != nullptr) prof2(a, 8);
      unsigned long long* arr = nullptr;
      unsigned* retries = nullptr;
      if (M2_PROF_ENABLED && a.prof_mode == 2) {
        arr = frame == 1 ? a.prof + (size_t)i * 4 * gridDim.x : nullptr;
        retries = frame == 1 ? reinterpret_cast<unsigned*>(a.prof + (size_t)a.n_ph * 4 * gridDim.x) + i : nullptr;
        if (threadIdx.x == 0) {
          // read by the phase functions after the barrier inside their wait
          sh.prof_arr = arr; sh.prof_retries = retries;
          if (frame == 1 && i == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            reinterpret_cast<unsigned*>(a.prof + (size_t)a.n_ph * 4 * gridDim.x)[1024 + blockIdx.x] = smid;
          }
        }
        m2_csync();
      }
      if (p.kind == M2_GEMV) {
        st = m4_unpack(m4_gemv_dispatch(sh, p, ring, work, m4_pack(st), tag, q));
      } else {
        M2Sync gs{a.bar, a.err, st.epoch, gridDim.x, st.dead, arr, retries};
        switch (p.kind) {
          case M2_ATTN:
            // (no "finished" mark here: m2_attn waits inside; sh.done lags by this one phase, which only shortens the
            // producer's descriptor lookahead from 8 to 7 phases)
            m2_unpack(gs, m2_attn(a, p, work, gs, tag));
            break;
          case M2_PROLOGUE: {
            m4_wait(sh, gs, p.flags, q);
            if (frame > 0 && a.do_sample) {
              int active = 0;
              for (int b = 0; b < B; ++b) active += __ldcg(a.fs.done + b) ? 0 : 1;
              if (active == 0) stop = true;
            }
            if (!stop && blockIdx.x == 0)
              for (int k = threadIdx.x; k < a.n_ac * B; k += MEGA_THREADS) a.fs.amax[k] = 0ull;
            if (threadIdx.x == 0) {
              if (stop) sh.stop = 1;
              __threadfence_block();
              sh.decided = q + 1u;               // the producer may stream the next frame's weights
            }
            m2_csync();
            m2_arrive(gs, p.flags);
          } break;
          case M2_GATHER: {
            m4_wait(sh, gs, p.flags, q);
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
            m2_csync();
            m2_arrive(gs, p.flags);
          } break;
          case M2_FINISH: {
            m4_wait(sh, gs, p.flags, q);
            m4_finish(a, p, sh.codes, tag, frame);
            m2_csync();
            m2_arrive(gs, p.flags);
          } break;
          case M2_SAMPLE: {
            SampleSmem& sm = *reinterpret_cast<SampleSmem*>(work);
            m4_wait(sh, gs, p.flags, q);
            for (int b = blockIdx.x; b < B; b += gridDim.x) m4_sample(a.smp, b, sm);
            m2_csync();
            m2_arrive(gs, p.flags);
          } break;
          case M2_COPYIN: {
            m4_wait(sh, gs, p.flags, q);
            for (int k = blockIdx.x * MEGA_THREADS + threadIdx.x; k < B * (a.H >> 3); k += gridDim.x * MEGA_THREADS)
              m2_store_row8(reinterpret_cast<u64*>(p.Y) + (size_t)k * 4, ldcg16(reinterpret_cast<const uint4*>(p.X) + k), tag);
            m2_csync();
            m2_arrive(gs, p.flags);
          } break;
          default: break;
        }
        st.epoch = gs.epoch; st.dead = gs.dead;
      }
      if (stop) break;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.tag_ctr = tag0 + q + (stop ? 1u : 0u);
}

// dynamic shared memory: ring + work area; returns 0 when the model does not fit.  n_slots is an output.
// Combine buffer: two tiles of 8 tokens (TP = 2, NT = 1) or one tile of 16 tokens (NT = 2), dual -- the same 37 KB.
static size_t mega4_smem_bytes(const q3_model_desc& d, int B, int max_seq, int* n_slots, int* red2) {
  (void)B;
  const size_t red = (size_t)2 * 2 * 8 * M4_RED_CS * 4;
  const size_t attn = (size_t)(2 * (std::max(max_seq, d.cp_max_seq) + 3) + 256 + 16 * 2 * 128) * 4 + (2 * M2_ATT_FAST_L * 128 + 256) * 2 + 256;      // + the row tables of m2_attn_units
  const size_t avail = 227 * 1024 - 5120;          // static shared memory of the kernel: M4Shared (4 KB) + profiling index
  size_t work = std::max(std::max((size_t)M2_RED_OFF + red, attn), sizeof(SampleSmem));
  *red2 = 1;
  work = (work + 127) & ~(size_t)127;
  if (work + 3 * M4_SLOT_BYTES > avail) return 0;
  *n_slots = (int)std::min<size_t>(M4_MAX_SLOTS, (avail - work) / M4_SLOT_BYTES);
  return (size_t)(*n_slots) * M4_SLOT_BYTES + work;
}
