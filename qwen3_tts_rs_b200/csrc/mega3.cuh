// Persistent frame kernel, third generation: the dataflow phases of mega2.cuh with the WEIGHTS STREAMED BY TMA
// into a shared-memory ring by a dedicated producer warp (warp-specialised producer / consumer).
//
// Why: measured on B200 (tools/profile_skew.py), a code-predictor phase of the dataflow kernel spent ~2 us between
// the last CTA's arrive and the moment a CTA got past its wait -- three times the 0.66 us of a bare hint barrier --
// because every warp first issues its weight loads into registers (55-85 KB per SM per phase) and the poll, the
// activation loads and everything after them queue behind that data on the SM's ingress path.  Registers cannot
// hold weights across the previous phase's arithmetic, shared memory can: here one extra warp issues
// cp.async.bulk (TMA, 1 KB per weight row per 512-column chunk) into an 8-stage / 136 KB ring as soon as a stage
// is free, running up to several phases ahead of the compute warps, so a phase's weights arrive while the
// previous phases wait on their barriers, and the compute warps' own loads are only the activations.
//   * ring stage = one "tile chunk": 16 weight rows x 512 columns (bf16), rows padded to 1088 bytes so that the
//     128-bit fragment loads (lane = row g, 16-byte column tg) are bank-conflict free without a swizzle;
//   * full/empty mbarriers per stage: the producer arms full[s] with expect_tx and the 16 row copies complete it;
//     each of the 16 compute warps arrives on empty[s] after its two LDS.128 + MMAs of the chunk;
//   * within a chunk the 16 warps split K (warp w owns k-step w = 32 columns), partial sums are combined in a
//     fixed order exactly as in mega2.cuh; the k-permutation trick (a lane's 16 contiguous bytes of a weight row
//     ARE its mma.m16n8k16 A fragments when the activations use the same permutation) carries over unchanged;
//   * phase descriptors are streamed by the producer warp as well (8-slot ring in shared memory), so the whole
//     227 KB minus the 66 KB combine buffer is available to the weight ring.
// Models whose K dimensions are not multiples of 512 stay on the mega2.cuh kernel.
#pragma once
#include "mega2.cuh"

constexpr int M3_STAGES = 8;
constexpr int M3_ROW_BYTES = 1024 + 64;
constexpr int M3_STAGE_BYTES = 16 * M3_ROW_BYTES;   // 17408
constexpr int M3_KC = 512;                          // columns per chunk
constexpr int M3_DESCS = 8;                         // descriptor ring slots
constexpr int M3_THREADS = MEGA_THREADS + 32;       // 16 compute warps + the producer warp

__device__ __forceinline__ uint32_t m3_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m3_mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ bool m3_mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void m3_mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void m3_mbar_expect(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m3_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}

// The round-1 ring kernel is history (slower than the dataflow kernel, superseded by mega4.cuh): compiled only into the
// development library (-DQ3_ALL_GENERATIONS); the product library carries a stub and refuses Q3_MEGA=3.
// which (K, dual, norm, input format) combinations the ring kernel implements
__host__ __device__ inline bool m3_gemv_supported(int K, bool dual, bool norm, int xf, int T) {
  if (dual) return norm && xf == XF_F32T && (K == 1024 || K == 2048);
  if (norm) return xf == XF_BF16T && (K == 1024 || K == 2048);
  if (xf == XF_GATHER) return K == 2048;
  if (xf != XF_BF16T) return false;
  return K == 1024 || K == 2048 || K == 3072 || (K == 6144 && T <= 8);
}

#ifdef Q3_ALL_GENERATIONS
// CTA-wide bookkeeping in shared memory
struct M3Shared {
  M2Args a;
  M2Phase desc[M3_DESCS];
  unsigned long long full[M3_STAGES], empty[M3_STAGES], desc_full[M3_DESCS];
  volatile unsigned done;       // phases completed by the compute warps
  volatile int stop;            // every row has sampled EOS: the producer leaves its loop
  volatile int dead;            // a watchdog fired somewhere in this CTA
  uint32_t codes[16];
};

// per-thread mutable state of a compute thread, passed to and returned from the phase functions by value
struct M3State {
  unsigned epoch;      // grid barrier epoch
  unsigned cq;         // tile chunks consumed so far (ring position)
  bool dead;
  unsigned long long* arr;
  unsigned* retries;
};

__device__ __forceinline__ bool m3_wait_full(M3Shared& sh, uint32_t bar, uint32_t parity) {
  unsigned it = 0;
  while (!m3_mbar_try(bar, parity)) {
    if (++it > (1u << 22)) {
      if (sh.a.err != nullptr) atomicCAS(sh.a.err, 0, 7000000);
      sh.dead = 1;
      return false;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------
// One skinny-GEMM phase on the compute warps: weights from the ring, activations (tag-verified) in registers.
// K == KCH * 512.  smem work area: scale[16] | part[16][16] | red (as mega2.cuh).
template <bool DUAL, int NT, int XF, bool NORM, int KCH>
__device__ __noinline__ M3State m3_gemv(M3Shared& sh, const M2Phase& p, unsigned char* ring, unsigned char* work, M3State st,
                                        const uint32_t tag) {
  const M2Args& a = sh.a;
  float* part_s = reinterpret_cast<float*>(work) + 16;
  float* red = reinterpret_cast<float*>(work + M2_RED_OFF);
  M2Sync gs{a.bar, a.err, st.epoch, gridDim.x, st.dead, st.arr, st.retries};
  const int K = p.K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const uint32_t xtag = tag - 1u;
  constexpr int NM = DUAL ? 2 : 1;
  int r0, r1;
  mega_row_range(p.N, r0, r1);
  if (r1 <= r0) {
    m2_wait(gs, p.flags);
    m2_arrive(gs, p.flags);
    st.epoch = gs.epoch; st.dead = gs.dead;
    return st;
  }
  prof2(a, 1);
  const int n_tiles = (r1 - r0 + 15) >> 4;
  const int red_r = n_tiles << 4, red_cs = 16 * red_r + 4;
  const int koff0 = warp * 32 + 8 * tg;              // element offset of this lane inside a chunk's 512 columns
  float rres[MEGA_MAX_OUT];
  m2_load_residual<NT>(p, r0, r1, rres);
  uint4 wn[KCH];
  if constexpr (NORM) {
#pragma unroll
    for (int c = 0; c < KCH; ++c) wn[c] = *reinterpret_cast<const uint4*>(p.aux + koff0 + c * M3_KC);
  }
  m2_wait(gs, p.flags);
  prof2(a, 2);
  m2_stamp(gs, 0);
  // ---- activations (once, tag-verified), scales ----
  const char* xrow[NT];
  m2_token_rows<NT, XF>(a, p, g, xrow);
  uint4 xv[KCH][NT];
  float sq[NT];
  {
    unsigned tries = 0;
    for (;;) {
      uint32_t bad = 0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) sq[nt] = 0.f;
#pragma unroll
      for (int c = 0; c < KCH; ++c)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          xv[c][nt] = make_uint4(0, 0, 0, 0);
          if (xrow[nt] != nullptr) xv[c][nt] = m2_load_x8<XF>(xrow[nt], koff0 + c * M3_KC, xtag, bad, sq[nt]);
        }
      if (XF == XF_GATHER || !__any_sync(0xffffffffu, bad != 0)) break;
      if (gs.retries != nullptr && lane == 0) atomicAdd(gs.retries, 1u);
      if (++tries > M2_RETRY_LIMIT) { m2_fail(gs, 5000000 + (int)gs.epoch); break; }
    }
  }
  if constexpr (NORM) {
    float xsc[NT];
    m2_row_scales<NT>(a, part_s, sq, K, xsc);
    const bool write_xn = p.xn_out != nullptr && blockIdx.x == 0;
#pragma unroll
    for (int c = 0; c < KCH; ++c)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        xv[c][nt] = m2_apply_norm(xv[c][nt], wn[c], xsc[nt]);
        if (write_xn && xrow[nt] != nullptr)
          *reinterpret_cast<uint4*>(p.xn_out + (size_t)(nt * 8 + g) * K + koff0 + c * M3_KC) = xv[c][nt];
      }
  }
  prof2(a, 3);
  m2_stamp(gs, 1);
  // ---- the chunk stream: tile-major, then k-chunk, then (gate, up) ----
  const uint32_t ring_s = m3_smem(ring), full_s = m3_smem(sh.full), empty_s = m3_smem(sh.empty);
  const uint32_t frag_off = (uint32_t)(g * M3_ROW_BYTES + warp * 64 + tg * 16);
  unsigned cq = st.cq;
  const unsigned rshift = (unsigned)a.ring_shift, rmask = (1u << rshift) - 1u;
  bool alive = !sh.dead;
  for (int tile = 0; tile < n_tiles; ++tile) {
    float acc[NM][NT][4];
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;
#pragma unroll
    for (int c = 0; c < KCH; ++c) {
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const unsigned s = cq & rmask, par = (cq >> rshift) & 1u;
        if (alive) alive = m3_wait_full(sh, full_s + s * 8, par);
        uint4 wl, wh;
        const uint32_t fa = ring_s + s * M3_STAGE_BYTES + frag_off;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wl.x), "=r"(wl.y), "=r"(wl.z), "=r"(wl.w) : "r"(fa));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(wh.x), "=r"(wh.y), "=r"(wh.z), "=r"(wh.w)
                     : "r"(fa + 8 * M3_ROW_BYTES));
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint4 x4 = xv[c][nt];
          mma_bf16_16816(acc[m][nt], wl.x, wh.x, wl.y, wh.y, x4.x, x4.y);
          mma_bf16_16816(acc[m][nt], wl.z, wh.z, wl.w, wh.w, x4.z, x4.w);
        }
        __syncwarp();
        if (lane == 0) m3_mbar_arrive(empty_s + s * 8);
        ++cq;
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
  m2_tail<DUAL, NT>(a, p, nullptr, red, rres, r0, r1, n_tiles, gs, tag);
  st.epoch = gs.epoch; st.dead = gs.dead; st.cq = cq;
  return st;
}

__device__ __forceinline__ M3State m3_gemv_dispatch(M3Shared& sh, const M2Phase& p, unsigned char* ring, unsigned char* work,
                                                    M3State st, const uint32_t tag) {
  const bool nt1 = p.T <= 8;
  const bool dual = (p.flags & PF_DUAL) != 0, norm = (p.flags & PF_NORM) != 0;
  const int K = p.K;
#define M3_CALL(DUAL_, XF_, NORM_, KCH_)                                                        \
  return nt1 ? m3_gemv<DUAL_, 1, XF_, NORM_, KCH_>(sh, p, ring, work, st, tag)                  \
             : m3_gemv<DUAL_, 2, XF_, NORM_, KCH_>(sh, p, ring, work, st, tag)
  if (dual) {
    if (K == 1024) { M3_CALL(true, XF_F32T, true, 2); }
    M3_CALL(true, XF_F32T, true, 4);
  }
  if (norm) {
    if (K == 1024) { M3_CALL(false, XF_BF16T, true, 2); }
    M3_CALL(false, XF_BF16T, true, 4);
  }
  if (p.xf == XF_GATHER) { M3_CALL(false, XF_GATHER, false, 4); }
  if (K == 1024) { M3_CALL(false, XF_BF16T, false, 2); }
  if (K == 2048) { M3_CALL(false, XF_BF16T, false, 4); }
  if (K == 3072) { M3_CALL(false, XF_BF16T, false, 6); }
  return m3_gemv<false, 1, XF_BF16T, false, 12>(sh, p, ring, work, st, tag);
#undef M3_CALL
}

// ---------------------------------------------------------------------------------------------------
// FINISH / SAMPLE bodies run on all 17 warps (frame_finish_row and sample_row_body use __syncthreads and blockDim).
__device__ __forceinline__ void m3_full_body(M3Shared& sh, const M2Phase& p, unsigned char* work, const uint32_t tag) {
  __syncthreads();
  if (p.kind == M2_FINISH) {
    m2_finish(sh.a, p, sh.codes, tag);
  } else {
    SampleSmem& sm = *reinterpret_cast<SampleSmem*>(work);
    for (int b = blockIdx.x; b < sh.a.B; b += gridDim.x) m2_sample(sh.a.smp, b, sm);
  }
  __syncthreads();
}

// The producer warp: phase descriptors and weight chunks, in program order, as far ahead as the rings allow.
__device__ __noinline__ void m3_producer(M3Shared& sh, unsigned char* ring, unsigned char* work) {
  const M2Args& a = sh.a;
  const int lane = threadIdx.x & 31;
  const uint32_t ring_s = m3_smem(ring), full_s = m3_smem(sh.full), empty_s = m3_smem(sh.empty),
                 dfull_s = m3_smem(sh.desc_full);
  unsigned pq = 0, q = 0;
  const unsigned rshift = (unsigned)a.ring_shift, rmask = (1u << rshift) - 1u;
  const uint32_t tag0 = __ldcg(a.tag_ctr);
  for (int frame = 0; frame < a.n_frames; ++frame) {
    for (int i = 0; i < a.n_ph; ++i, ++q) {
      // descriptor slot q % M3_DESCS is free once the compute warps finished phase q - M3_DESCS
      {
        unsigned it = 0;
        while ((int)(q - sh.done) >= M3_DESCS) {
          if (sh.stop || sh.dead) return;
          __nanosleep(64);
          if (++it > (1u << 24)) { sh.dead = 1; return; }
        }
      }
      M2Phase* slot = &sh.desc[q % M3_DESCS];
      {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.prog + i);
        uint32_t* dst = reinterpret_cast<uint32_t*>(slot);
        for (int w = lane; w < (int)(sizeof(M2Phase) / 4); w += 32) dst[w] = __ldg(src + w);
      }
      __syncwarp();
      if (lane == 0) m3_mbar_arrive(dfull_s + (q % M3_DESCS) * 8);
      const int kind = slot->kind;
      if (kind == M2_GEMV) {
        int r0, r1;
        mega_row_range(slot->N, r0, r1);
        if (r1 <= r0) continue;
        const int K = slot->K, kch = K / M3_KC, nm = (slot->flags & PF_DUAL) ? 2 : 1;
        const bf16* W0 = slot->W;
        const bf16* W1 = slot->W2;
        const int n_tiles = (r1 - r0 + 15) >> 4;
        for (int tile = 0; tile < n_tiles; ++tile) {
          const int n0 = r0 + (tile << 4);
          const int rows = min(16, r1 - n0);
          for (int c = 0; c < kch; ++c)
            for (int m = 0; m < nm; ++m) {
              const unsigned s = pq & rmask, par = ((pq >> rshift) & 1u) ^ 1u;
              unsigned it = 0;
              while (!m3_mbar_try(empty_s + s * 8, par)) {
                if (sh.stop || sh.dead) return;
                __nanosleep(a.pf_sleep);      // a spinning 17th warp steals issue slots from four compute warps
                if (++it > (1u << 22)) { sh.dead = 1; return; }
              }
              if (lane == 0) m3_mbar_expect(full_s + s * 8, (uint32_t)rows * 1024u);
              __syncwarp();
              if (lane < rows) {
                const bf16* src = (m == 0 ? W0 : W1) + (size_t)(n0 + lane) * K + (size_t)c * M3_KC;
                m3_bulk_g2s(ring_s + s * M3_STAGE_BYTES + lane * M3_ROW_BYTES, src, 1024u, full_s + s * 8);
              }
              ++pq;
            }
        }
      } else if (kind == M2_PROLOGUE) {
        if (frame > 0 && a.do_sample) {
          // the compute warps decide here whether the loop ends; do not stream past that decision
          unsigned it = 0;
          while ((int)(sh.done - q) <= 0) {
            if (sh.dead) return;
            __nanosleep(64);
            if (++it > (1u << 24)) { sh.dead = 1; return; }
          }
          if (sh.stop) return;
        }
      } else if (kind == M2_FINISH || kind == M2_SAMPLE) {
        m3_full_body(sh, *slot, work, tag0 + q + 1u);
      }
    }
  }
}

__global__ void __launch_bounds__(M3_THREADS, 1) decode_frames_mega3_kernel(const M2Args args) {
  extern __shared__ __align__(128) unsigned char m3_dyn[];
  __shared__ M3Shared sh;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&args);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.a);
    for (int i = threadIdx.x; i < (int)(sizeof(M2Args) / 4); i += M3_THREADS) dst[i] = src[i];
    if (threadIdx.x == 0) {
      s_prof2_idx = g_prof2_idx;
      sh.done = 0u; sh.stop = 0; sh.dead = 0;
      for (int s = 0; s < M3_STAGES; ++s) {
        m3_mbar_init(m3_smem(&sh.full[s]), 1u);
        m3_mbar_init(m3_smem(&sh.empty[s]), MEGA_WARPS);
      }
      for (int s = 0; s < M3_DESCS; ++s) m3_mbar_init(m3_smem(&sh.desc_full[s]), 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  unsigned char* ring = m3_dyn;
  unsigned char* work = m3_dyn + M3_STAGES * M3_STAGE_BYTES;
  if (threadIdx.x >= MEGA_THREADS) {
    m3_producer(sh, ring, work);
    return;
  }
  const M2Args& a = sh.a;
  const int B = a.B;
  const uint32_t tag0 = __ldcg(a.tag_ctr);
  const uint32_t dfull_s = m3_smem(sh.desc_full);
  M3State st{0u, 0u, false, nullptr, nullptr};
  {
    M2Sync gs{a.bar, a.err, 0u, gridDim.x, false, nullptr, nullptr};
    m2_arrive(gs, PF_ARRIVE_REL);       // every phase waits for its predecessor's arrive; this is the first phase's
  }
  uint32_t q = 0;
  bool stop = false;
  for (int frame = 0; frame < a.n_frames && !stop; ++frame) {
    for (int i = 0; i < a.n_ph; ++i, ++q) {
      // this phase's descriptor (streamed by the producer warp)
      {
        unsigned it = 0;
        while (!m3_mbar_try(dfull_s + (q % M3_DESCS) * 8, (q / M3_DESCS) & 1u)) {
          if (++it > (1u << 22)) { sh.dead = 1; break; }
        }
      }
      const M2Phase& p = sh.desc[q % M3_DESCS];
      const uint32_t tag = tag0 + q + 1u;
      if (a.prof_mode == 2) {
        st.arr = frame == 1 ? a.prof + (size_t)i * 4 * gridDim.x : nullptr;
        st.retries = frame == 1 ? reinterpret_cast<unsigned*>(a.prof + (size_t)a.n_ph * 4 * gridDim.x) + i : nullptr;
        if (frame == 1 && i == 0 && threadIdx.x == 0) {
          unsigned smid;
          asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          reinterpret_cast<unsigned*>(a.prof + (size_t)a.n_ph * 4 * gridDim.x)[1024 + blockIdx.x] = smid;
        }
      }
      if (p.kind == M2_GEMV) {
        st = m3_gemv_dispatch(sh, p, ring, work, st, tag);
      } else {
        M2Sync gs{a.bar, a.err, st.epoch, gridDim.x, st.dead, st.arr, st.retries};
        switch (p.kind) {
          case M2_ATTN: m2_unpack(gs, m2_attn(a, p, work, gs, tag)); break;
          case M2_PROLOGUE: {
            m2_wait(gs, p.flags);
            if (frame > 0 && a.do_sample) {
              int active = 0;
              for (int b = 0; b < B; ++b) active += __ldcg(a.fs.done + b) ? 0 : 1;
              if (active == 0) stop = true;
            }
            if (!stop && blockIdx.x == 0)
              for (int k = threadIdx.x; k < a.n_ac * B; k += MEGA_THREADS) a.fs.amax[k] = 0ull;
            if (stop && threadIdx.x == 0) sh.stop = 1;
            m2_csync();
            m2_arrive(gs, p.flags);
          } break;
          case M2_GATHER: {
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
            m2_csync();
            m2_arrive(gs, p.flags);
          } break;
          case M2_FINISH:
          case M2_SAMPLE: {
            m2_wait(gs, p.flags);
            m3_full_body(sh, p, work, tag);
            m2_arrive(gs, p.flags);
          } break;
          case M2_COPYIN: {
            m2_wait(gs, p.flags);
            for (int k = blockIdx.x * MEGA_THREADS + threadIdx.x; k < B * (a.H >> 3); k += gridDim.x * MEGA_THREADS)
              m2_store_row8(reinterpret_cast<u64*>(p.Y) + (size_t)k * 4, ldcg16(reinterpret_cast<const uint4*>(p.X) + k), tag);
            m2_csync();
            m2_arrive(gs, p.flags);
          } break;
          default: break;
        }
        st.epoch = gs.epoch; st.dead = gs.dead;
      }
      // the descriptor slot may be recycled: every compute thread is past its last read of it
      m2_csync();
      if (threadIdx.x == 0) sh.done = q + 1u;
      if (stop) break;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.tag_ctr = tag0 + q + (stop ? 1u : 0u);
}

#else
__global__ void __launch_bounds__(M3_THREADS, 1) decode_frames_mega3_kernel(const M2Args) {}
#endif

static size_t mega3_smem_bytes(const q3_model_desc& d, int B, int max_seq, int grid) {
  const size_t work = mega2_smem_bytes(d, B, max_seq, grid, 0);   // work area only (no program in shared memory)
  if (work == 0) return 0;
  return (size_t)M3_STAGES * M3_STAGE_BYTES + work;
}
