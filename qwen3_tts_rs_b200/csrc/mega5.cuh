// Persistent frame kernel, fifth generation: the dataflow kernel of mega2.cuh, with the weights of its REGISTER-RESIDENT
// phases delivered by TMA through a shared-memory ring instead of by 128-bit global loads issued in front of the wait
// (generate_codes src/lib.rs:530-656; code-predictor frame src/models/code_predictor.rs:320-416; talker step
// src/models/talker.rs:716-736).
//
// What round 2 measured (profiles/r2_mega4_vs_mega2.md):
//   * in the dataflow kernel a code-predictor phase waits 1.1-2.3 us at its barrier although a bare hint barrier costs
//     0.65 us: the 55-85 KB of weight loads issued before the wait occupy the SM's load path, and the poll of the barrier
//     word queues behind them;
//   * the warp-specialised ring kernel (mega4.cuh) brings that wait down to 0.6-0.7 us, but pays for it twice: its 17th
//     warp caps every thread at 96 registers (5 warps per scheduler), and its weights are consumed slot by slot behind
//     mbarrier waits inside the dependent part of the phase -- 4.30 ms per frame against 2.93.
// This generation keeps what worked on each side:
//   * NO extra warp (16 warps, 128 registers);
//   * ONE prefetch buffer per CTA (all the shared memory mega2's work area leaves free, ~156 KB) holds the rows of the NEXT
//     register-resident phase: a CTA's rows [r0, r1) x K are contiguous in global memory, so the fill is one bulk copy per
//     matrix, split into sixteen so that every warp issues its own small piece (a single 50-130 KB cp.async.bulk holds the
//     issuing thread for ~1.9 us);
//   * a phase copies its fragments buffer -> registers at its ENTRY, before the wait (16 x LDS.128 per lane at most); the
//     block barrier that ends the wait proves every warp has done so, and right after it the warps issue the next ring
//     phase's rows into the same buffer (found by thread 0 scanning the descriptor ring, never past an undecided
//     PROLOGUE).  After the wait the phase is mega2's register-resident phase, instruction for instruction;
//   * the two big talker phases (gate/up, down: 340 / 170 KB per SM, bandwidth-bound) keep streaming global -> registers
//     as in mega2.cuh; the 88 KB phase program moves out of shared memory (16-entry descriptor ring staged 8 phases ahead
//     by warp 1), which is what makes room for the buffer.
// Result: bit-identical codes to mega2, deterministic, and SLOWER -- 4.22 ms per frame against 2.89 (1.7B, batch 8): the
// buffer is full 0.3-0.4 us after a phase asks for it, but the extra shared-memory hop and the bulk traffic in flight
// during the activation loads cost more than the shorter wait returns.  Opt-in (Q3_MEGA=5); kept as a measured result.
#pragma once
#include "mega4.cuh"

constexpr int M5_DESCS = 16;
constexpr int M5_DESC_AHEAD = 8;

// One prefetch buffer per CTA: the weight rows [r0, r1) x K of the NEXT register-resident phase (gate rows then up rows for
// the SwiGLU pair) -- contiguous in global memory.  It is filled while the current phase runs: the phase copies its
// fragments out of the buffer at its entry (before its barrier wait); once every warp has done so -- the block barrier that
// ends the wait -- the 16 warps issue one sixteenth of the next phase's rows each (one cp.async.bulk per warp and matrix: a
// single 50-130 KB copy holds its issuing thread for more than a microsecond, sixteen small ones issue in parallel).
struct M5Ring {
  unsigned long long full;         // mbarrier: all bytes of the region `issued_q` have landed
  M2Phase desc[M5_DESCS];
  volatile unsigned q_prod;        // absolute phase index (frame * n_ph + i) the scan looks at next
  volatile int issued_q;           // latest phase whose rows were issued into the buffer (-1: none yet)
  volatile int pend_q;             // phase the warps issue at the next issue point, or -1
  volatile int ready_q;            // = issued_q, but written only AFTER the block barrier that follows the scan: what a phase
                                   // tests at its entry (thread 0 may run a scan ahead of a slow warp's entry; it cannot pass
                                   // the barrier ahead of it)
  const char* volatile pend_w;     // its rows (first matrix / second matrix), bytes per matrix
  const char* volatile pend_w2;
  volatile unsigned pend_bytes;
  volatile unsigned staged;        // descriptors of all phases < staged are in desc[]
  unsigned q_total;
  int n_ph, do_sample;
  uint32_t buf_s;                  // shared-window address of the buffer
  int* err;
};

// Thread 0, buffer free (every warp of the CTA has copied its fragments of phase <= q_cur out of it, or is about to pass the
// barrier that proves it): finds the next ring phase at or after `from` in which this CTA has rows, arms the barrier with its
// byte count and publishes it as pending; the warps issue it after their next block barrier (m5_issue).
__device__ __noinline__ void m5_scan(M5Ring* ring_p, const unsigned q_cur, const unsigned from) {
  M5Ring& r = *ring_p;
  r.pend_q = -1;
  if (r.issued_q > (int)q_cur) return;            // a later phase's rows are already in the buffer (this CTA had no rows in q_cur)
  unsigned q = r.q_prod > from ? r.q_prod : from;
  for (int guard = 0; guard < 12; ++guard) {
    if (q >= r.q_total || q >= r.staged) break;
    const M2Phase& d = r.desc[q % M5_DESCS];
    const int flags = d.flags;
    if (!(flags & PF_RING)) {
      // the PROLOGUE of every frame but the first decides whether the loop ends: never stream past one that is still ahead
      if (d.kind == M2_PROLOGUE && q >= (unsigned)r.n_ph && r.do_sample && q > q_cur) break;
      ++q;
      continue;
    }
    int r0, r1;
    mega_row_range(d.N, r0, r1);
    if (r1 <= r0) { ++q; continue; }
    const uint32_t bytes = (uint32_t)(r1 - r0) * (uint32_t)d.K * 2u;
    const bool dual = (flags & PF_DUAL) != 0;
    r.pend_w = reinterpret_cast<const char*>(d.W + (size_t)r0 * d.K);
    r.pend_w2 = dual ? reinterpret_cast<const char*>(d.W2 + (size_t)r0 * d.K) : nullptr;
    r.pend_bytes = bytes;
    m4_mbar_expect(m4_smem(&r.full), dual ? 2u * bytes : bytes);
    r.issued_q = (int)q;
    r.pend_q = (int)q;
    r.q_prod = q + 1u;
    return;
  }
  r.q_prod = q;
}
// every warp, after a block barrier that follows m5_scan: its sixteenth of the pending rows
__device__ __forceinline__ void m5_issue(M5Ring& r) {
  if ((threadIdx.x & 31) == 0 && r.pend_q >= 0) {
    if (threadIdx.x == 0) r.ready_q = r.pend_q;
    const uint32_t bytes = r.pend_bytes, slice = bytes / MEGA_WARPS, off = (threadIdx.x >> 5) * slice;
    const uint32_t full_a = m4_smem(&r.full);
    m4_bulk_g2s(r.buf_s + off, r.pend_w + off, slice, full_a);
    const char* w2 = r.pend_w2;
    if (w2 != nullptr) m4_bulk_g2s(r.buf_s + bytes + off, w2 + off, slice, full_a);
  }
}

// ---------------------------------------------------------------------------------------------------
// Register-resident skinny-GEMM phase (m2_gemv_small of mega2.cuh) with its weight fragments taken from the prefetch buffer at
// the phase's entry.  Returns the packed state (epoch | par << 40 | dead << 41) of mega4.cuh; par = parity of the buffer's
// full barrier for the next ring phase.
template <bool DUAL, int NT, int XF, bool NORM, int TILES, int CHUNKS>
__device__ __noinline__ unsigned long long m5_gemv_small(const M2Args& a, M5Ring& ring, const M2Phase& p, unsigned char* smem,
                                                         M2Sync gs, const unsigned long long st_packed, const uint32_t tag,
                                                         const unsigned q_abs) {
  M4State st = m4_unpack(st_packed);
  float* part_s = reinterpret_cast<float*>(smem) + 16;
  float* red = reinterpret_cast<float*>(smem + M2_RED_OFF);
  const int K = p.K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const uint32_t xtag = tag - 1u;
  constexpr int NM = DUAL ? 2 : 1;
  constexpr int JU = 2;
  int r0, r1;
  mega_row_range(p.N, r0, r1);
  if (r1 <= r0) {
    if (tid == 0) m5_scan(&ring, q_abs, q_abs + 1u);
    m2_wait(gs, p.flags);
    m5_issue(ring);
    m2_arrive(gs, p.flags);
    st.epoch = gs.epoch; st.dead = gs.dead;
    return m4_pack(st);
  }
  const int n_tiles = (r1 - r0 + 15) >> 4;
  const int koff0 = warp * 32 + 8 * tg;
  float rres[MEGA_MAX_OUT];
  // ---- before the wait: every weight fragment of the phase, buffer -> registers; then the buffer is handed back ----
  uint4 wl[TILES][CHUNKS][NM][JU], wh[TILES][CHUNKS][NM][JU];
  const bool prof_on = M2_PROF_ENABLED && a.prof != nullptr;
  if (prof_on) prof2(a, 1);
  if (ring.ready_q != (int)q_abs) {
    // this phase's rows were not issued ahead (first ring phase of a frame, or of the launch): issue them now
    __syncthreads();
    if (tid == 0) m5_scan(&ring, q_abs, q_abs);
    __syncthreads();
    m5_issue(ring);
  }
  {
    const uint32_t zmask = (uint32_t)a.bench_barriers;      // 0 in every decode launch; opaque to the compiler (m4_mbar_release_slot)
    const uint32_t full_a = m4_smem(&ring.full);
    if (!gs.dead) {
      // one lane polls, the warp re-converges (m4_wait_full)
      bool ok = true;
      if (lane == 0) {
        unsigned it = 0;
        while (!m4_mbar_try(full_a, st.par)) {
          if (++it > M4_MBAR_SPIN) { if (ring.err != nullptr) atomicCAS(ring.err, 0, 7500000); ok = false; break; }
        }
      }
      __syncwarp();
      if (__shfl_sync(0xffffffffu, ok ? 1 : 0, 0) == 0) gs.dead = true;
    }
    st.par ^= 1u;
    if (prof_on) prof2(a, 9);            // buffer full
    const int rows_total = r1 - r0;
    const uint32_t row_b = (uint32_t)K * 2u;
    uint32_t chk = 0u;
#pragma unroll
    for (int tile = 0; tile < TILES; ++tile) {
      const bool lo_ok = tile < n_tiles, hi_ok = lo_ok && (tile << 4) + 8 < rows_total;
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const uint32_t base = ring.buf_s + (uint32_t)(m * rows_total + (tile << 4) + g) * row_b + (uint32_t)koff0 * 2u;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
          for (int u = 0; u < JU; ++u) {
            wl[tile][c][m][u] = make_uint4(0, 0, 0, 0);
            wh[tile][c][m][u] = make_uint4(0, 0, 0, 0);
            if (lo_ok) { wl[tile][c][m][u] = m4_lds128(base + (c * JU + u) * 1024); chk ^= wl[tile][c][m][u].x; }
            if (hi_ok) { wh[tile][c][m][u] = m4_lds128(base + 8 * row_b + (c * JU + u) * 1024); chk ^= wh[tile][c][m][u].x; }
          }
      }
    }
    // the buffer is refilled after the block barrier that ends the wait below: every fragment load must have COMPLETED before
    // this warp arrives there, so the arrival is made to depend on a word of every load (z = 0 at run time; opaque to the
    // compiler -- m4_mbar_release_slot explains what happened without it)
    const uint32_t z = __reduce_or_sync(0xffffffffu, chk & zmask);
    if (z != 0u) asm volatile("trap;");
  }
  if (tid == 0) m5_scan(&ring, q_abs, q_abs + 1u);
  if (prof_on) prof2(a, 11);             // fragments in registers, buffer handed back
  uint4 wn[CHUNKS][JU];
  if constexpr (NORM) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
      for (int u = 0; u < JU; ++u) wn[c][u] = *reinterpret_cast<const uint4*>(p.aux + koff0 + (c * JU + u) * 512);
  }
  m2_load_residual<NT>(p, r0, r1, rres);
  m2_wait(gs, p.flags);
  m5_issue(ring);                        // every warp's fragments are in registers: the next ring phase's rows -> the buffer
  if (prof_on) prof2(a, 2);
  // ---- after the wait: activations (once, tag-verified), scales -- m2_gemv_small from here on ----
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
  const int red_r = n_tiles << 4, red_cs = 16 * red_r + 4;
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
  m2_tail<DUAL, NT>(a, p, nullptr, red, rres, r0, r1, n_tiles, gs, tag);
  st.epoch = gs.epoch; st.dead = gs.dead;
  return m4_pack(st);
}

// which register-resident variant serves a ring phase: the same table as m2_gemv_dispatch (p.small chosen by the host)
__device__ __forceinline__ unsigned long long m5_gemv_dispatch(const M2Args& a, M5Ring& ring, const M2Phase& p, unsigned char* smem,
                                                               const M2Sync& gs, const unsigned long long st, const uint32_t tag,
                                                               const unsigned q) {
  const bool nt1 = p.T <= 8;
  const bool dual = (p.flags & PF_DUAL) != 0, norm = (p.flags & PF_NORM) != 0;
  const int sm = p.small;
  if (dual) {             // gate/up: NORM, f32 input (h1)
    if (nt1) return m5_gemv_small<true, 1, XF_F32T, true, 2, 1>(a, ring, p, smem, gs, st, tag, q);
    return m5_gemv_small<true, 2, XF_F32T, true, 2, 1>(a, ring, p, smem, gs, st, tag, q);
  } else if (norm) {
    if (sm == 0x21) {
      if (nt1) return m5_gemv_small<false, 1, XF_BF16T, true, 2, 1>(a, ring, p, smem, gs, st, tag, q);
      return m5_gemv_small<false, 2, XF_BF16T, true, 2, 1>(a, ring, p, smem, gs, st, tag, q);
    }
    return m5_gemv_small<false, 1, XF_BF16T, true, 2, 2>(a, ring, p, smem, gs, st, tag, q);            // 0x22, T <= 8 only
  } else if (p.xf == XF_GATHER) {
    if (nt1) return m5_gemv_small<false, 1, XF_GATHER, false, 1, 2>(a, ring, p, smem, gs, st, tag, q);
    return m5_gemv_small<false, 2, XF_GATHER, false, 1, 2>(a, ring, p, smem, gs, st, tag, q);
  } else if (sm == 0x12) {
    if (nt1) return m5_gemv_small<false, 1, XF_BF16T, false, 1, 2>(a, ring, p, smem, gs, st, tag, q);
    return m5_gemv_small<false, 2, XF_BF16T, false, 1, 2>(a, ring, p, smem, gs, st, tag, q);
  }
  if (nt1) return m5_gemv_small<false, 1, XF_BF16T, false, 1, 3>(a, ring, p, smem, gs, st, tag, q);     // 0x13
  return m5_gemv_small<false, 2, XF_BF16T, false, 1, 3>(a, ring, p, smem, gs, st, tag, q);
}

// the host marks a phase PF_RING when this returns true (q3tts.cu m2_build_program)
__host__ __device__ inline bool m5_ring_phase(int kind, int small, int K) { return kind == M2_GEMV && small != 0 && (K & 1023) == 0; }

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_frames_mega5_kernel(const M2Args args) {
  extern __shared__ __align__(128) unsigned char m5_dyn[];
  __shared__ M2Args sa;
  __shared__ M5Ring ring;
  __shared__ uint32_t s_codes[16];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&args);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sa);
    for (int i = threadIdx.x; i < (int)(sizeof(M2Args) / 4); i += MEGA_THREADS) dst[i] = src[i];
    if (threadIdx.x == 0) {
      s_prof2_idx = g_prof2_idx;
      m4_mbar_init(m4_smem(&ring.full), 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      ring.q_prod = 0u; ring.issued_q = -1; ring.pend_q = -1; ring.ready_q = -1;
      ring.q_total = (unsigned)args.n_frames * (unsigned)args.n_ph;
      ring.n_ph = args.n_ph; ring.do_sample = args.do_sample;
      ring.buf_s = m4_smem(m5_dyn);
      ring.err = args.err;
    }
    // the first descriptors of the program -> the descriptor ring
    const int first = min(M5_DESC_AHEAD, args.n_frames * args.n_ph);
    for (int i = threadIdx.x; i < first * (int)(sizeof(M2Phase) / 16); i += MEGA_THREADS) {
      const int ph = i / (int)(sizeof(M2Phase) / 16), w = i % (int)(sizeof(M2Phase) / 16);
      reinterpret_cast<uint4*>(&ring.desc[ph % M5_DESCS])[w] = __ldg(reinterpret_cast<const uint4*>(args.prog + (ph % args.n_ph)) + w);
    }
    if (threadIdx.x == 0) ring.staged = (unsigned)first;
  }
  __syncthreads();
  const M2Args& a = sa;
  unsigned char* work = m5_dyn + (size_t)a.m4_slots;        // m4_slots: bytes of the prefetch buffer in this generation
  M2Sync gs{a.bar, a.err, 0u, gridDim.x, false, nullptr, nullptr};
  M4State st{0u, 0u, 0u, false};
  const int B = a.B;
  const uint32_t tag0 = __ldcg(a.tag_ctr);
  uint32_t q = 0;
  bool stop = false;
  m2_arrive(gs, PF_ARRIVE_REL);       // every phase waits for its predecessor's arrive; this is the first phase's
  for (int frame = 0; frame < a.n_frames && !stop; ++frame) {
    for (int i = 0; i < a.n_ph; ++i, ++q) {
      // descriptor of phase q + 8 -> registers of warp 1 now, shared memory at the end of this phase
      const unsigned qs = q + (unsigned)M5_DESC_AHEAD;
      const bool stager = (threadIdx.x >> 5) == 1 && (threadIdx.x & 31) < (int)(sizeof(M2Phase) / 16) && qs < ring.q_total;
      uint4 dnext = make_uint4(0, 0, 0, 0);
      if (stager) dnext = __ldg(reinterpret_cast<const uint4*>(a.prog + (qs % (unsigned)a.n_ph)) + (threadIdx.x & 31));
      const M2Phase& p = ring.desc[q % M5_DESCS];
      const uint32_t tag = tag0 + q + 1u;
      if (p.kind == M2_GEMV && (p.flags & PF_RING)) {
        M2Sync g2 = gs;
        g2.epoch = st.epoch; g2.dead = st.dead;
        st = m4_unpack(m5_gemv_dispatch(a, ring, p, work, g2, m4_pack(st), tag, q));
      } else {
        M2Sync g2 = gs;
        g2.epoch = st.epoch; g2.dead = st.dead;
        switch (p.kind) {
          case M2_GEMV: {
            const int nxi = p.next_gemv;
            const M2Phase* nx = nullptr;       // (next-phase L2 prefetch: only towards phases the ring does not serve)
            if (nxi >= 0) {
              // the descriptor of the next skinny-GEMM phase is at most 2 phases ahead: already staged
              const unsigned qn = q + (unsigned)((nxi > i) ? (nxi - i) : (a.n_ph - i + nxi));
              if (qn < ring.q_total && qn < ring.staged && !(ring.desc[qn % M5_DESCS].flags & PF_RING)) nx = &ring.desc[qn % M5_DESCS];
            }
            m2_unpack(g2, m2_gemv_dispatch(a, p, nx, work, g2, tag));
          } break;
          case M2_ATTN: {
            const int nxi = p.next_gemv;
            if (nxi >= 0) {
              const unsigned qn = q + (unsigned)((nxi > i) ? (nxi - i) : (a.n_ph - i + nxi));
              if (qn < ring.q_total && qn < ring.staged && !(ring.desc[qn % M5_DESCS].flags & PF_RING)) m2_prefetch(a, &ring.desc[qn % M5_DESCS]);
            }
            m2_unpack(g2, m2_attn(a, p, work, g2, tag));
          } break;
          case M2_PROLOGUE: {
            m2_wait(g2, p.flags);
            if (frame > 0 && a.do_sample) {
              int active = 0;
              for (int b = 0; b < B; ++b) active += __ldcg(a.fs.done + b) ? 0 : 1;
              if (active == 0) stop = true;
            }
            if (!stop && blockIdx.x == 0)
              for (int k = threadIdx.x; k < a.n_ac * B; k += MEGA_THREADS) a.fs.amax[k] = 0ull;
            __syncthreads();
            m2_arrive(g2, p.flags);
          } break;
          case M2_GATHER: {
            m2_wait(g2, p.flags);
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
            m2_arrive(g2, p.flags);
          } break;
          case M2_FINISH: {
            m2_wait(g2, p.flags);
            m2_finish(a, p, s_codes, tag);
            __syncthreads();
            m2_arrive(g2, p.flags);
          } break;
          case M2_COPYIN: {
            m2_wait(g2, p.flags);
            for (int k = blockIdx.x * MEGA_THREADS + threadIdx.x; k < B * (a.H >> 3); k += gridDim.x * MEGA_THREADS)
              m2_store_row8(reinterpret_cast<u64*>(p.Y) + (size_t)k * 4, ldcg16(reinterpret_cast<const uint4*>(p.X) + k), tag);
            __syncthreads();
            m2_arrive(g2, p.flags);
          } break;
          case M2_SAMPLE: {
            SampleSmem& sm = *reinterpret_cast<SampleSmem*>(work);
            m2_wait(g2, p.flags);
            for (int b = blockIdx.x; b < B; b += gridDim.x) m2_sample(a.smp, b, sm);
            __syncthreads();
            m2_arrive(g2, p.flags);
          } break;
          default: break;
        }
        st.epoch = g2.epoch; st.dead = g2.dead;
      }
      // stage the descriptor fetched at the top of the phase (its ring entry held phase q - 8, finished long ago)
      if (stager) reinterpret_cast<uint4*>(&ring.desc[qs % M5_DESCS])[threadIdx.x & 31] = dnext;
      if ((threadIdx.x >> 5) == 1) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0 && qs < ring.q_total) ring.staged = qs + 1u;
      }
      if (stop) break;
    }
  }
  // the tag counter of the session: read by every CTA at the start of the NEXT launch
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.tag_ctr = tag0 + q + (stop ? 1u : 0u);
}

// the host marks a phase PF_RING when its rows fit the prefetch buffer; buffer bytes needed by a phase on the busiest CTA
static size_t m5_region_bytes(int N, int K, bool dual, int grid) {
  const int units = N / 8, per_cta = (units + grid - 1) / grid;
  return (size_t)(dual ? 2 : 1) * per_cta * 8 * K * 2;
}
// room for the prefetch buffer beside mega2's work area (no program in shared memory); 0 when even 32 KB do not fit
static size_t mega5_buffer_cap(const q3_model_desc& d, int B, int max_seq, int grid, size_t* work_out) {
  const size_t work = mega2_smem_bytes(d, B, max_seq, grid, 0);
  if (work == 0) return 0;
  const size_t avail = 227 * 1024 - 5120;          // static shared memory: M2Args + M5Ring (descriptor ring 2.5 KB)
  const size_t w = (work + 127) & ~(size_t)127;
  *work_out = w;
  if (w + 32768 > avail) return 0;
  return (avail - w) & ~(size_t)127;
}
