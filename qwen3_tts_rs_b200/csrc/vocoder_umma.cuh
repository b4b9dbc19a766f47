// The vocoder's implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulator in tensor memory).
// Same operation, arguments and F32-class numerics as voc_conv_mma_kernel (vocoder_mma.cuh: every F32 operand split into two
// bf16 terms, hi*hi + hi*lo + lo*hi with F32 accumulation); ref: the convolutions of Decoder12Hz::decode
// (src/models/codec/decoder_12hz.rs:185-505), which the reference runs in F32 (src/lib.rs:344-345).
//
// CTA tile: 128 output channels x 128 output positions, accumulator = 128 lanes x 128 columns of tensor memory.  Per
// 32-channel chunk and tap the tensor core executes six 128x128x16 MMAs; one thread issues them.
//   weights      packed at model load into the exact shared-memory image the MMA reads -- for every (tap, chunk, 128-row
//                tile) 16 KB = [hi | lo][4 k-groups of 8 channels][128 rows][8 bf16], the canonical K-major no-swizzle layout
//                (core matrix = 8 rows x 16 bytes, contiguous) -- so one cp.async.bulk per step fills a stage of a 3-deep ring;
//   activations  staged by 256 threads: F32 loads (24 in flight per thread), SnakeBeta, hi/lo split, 16-byte shared stores
//                into [hi | lo][4 k-groups][window rows][8 bf16].  A tap is a ROW OFFSET into that window: with the
//                no-swizzle layout consecutive rows are 16 bytes apart, so the B descriptor of a tap is the window's
//                descriptor with its start address advanced by shift * 16 bytes -- no copy per tap.  The window is double
//                buffered: chunk c+1 is staged while the tensor core works on chunk c;
//   epilogue     tcgen05.ld (32 lanes x 32 columns per warp and load), bias / GELU / scale / residual / clamp, stores.
// Warp roles: warps 0-7 stage activations and run the epilogue, warp 8 allocates tensor memory and issues the MMAs,
// warp 9 feeds the weight ring.  Two CTAs per SM (96 KB shared memory, 128 of 512 tensor-memory columns each), so one CTA's
// staging overlaps the other's MMAs as well.
#pragma once
#include "vocoder_mma.cuh"

constexpr int UC_STAGES = 3;
constexpr int UC_THREADS = 320;
constexpr int UC_STAGERS = 256;
constexpr int UC_A_BYTES = 16384;        // one ring stage: hi (8 KB) + lo (8 KB)
constexpr int UC_A_ELEMS = UC_A_BYTES / 2;
constexpr int UC_NI = 3;                 // (window row, k-group) items per staging thread: 4 * (128 + halo) <= 3 * 256
constexpr int UC_MAX_BROWS = UC_NI * UC_STAGERS / 4;   // 192
constexpr unsigned UC_SPIN = 1u << 27;   // mbarrier polls before a wait gives up (traps instead of hanging the GPU)

__device__ __forceinline__ uint32_t uc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void uc_mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ bool uc_mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void uc_mbar_wait(uint32_t addr, uint32_t parity) {
  unsigned it = 0;
  while (!uc_mbar_try(addr, parity))
    if (++it > UC_SPIN) asm volatile("trap;");
}
__device__ __forceinline__ void uc_mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void uc_mbar_expect(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void uc_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// tcgen05.commit: the mbarrier receives one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void uc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void uc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void uc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle: core matrices of 8 rows x 16 bytes; lbo = bytes between the two
// k-groups of one MMA, sbo = bytes between 8-row groups; version 1 (Blackwell) in bits 46-47
__device__ __forceinline__ uint64_t uc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor of kind::f16: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 in bits 17-22, M >> 4 in bits 24-28
constexpr uint32_t UC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MC_BN >> 3) << 17) | ((uint32_t)(MC_BM >> 4) << 24);
__device__ __forceinline__ void uc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(UC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void uc_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// args: MmaConvArgs with w_um set (see voc_pack_umma_weights_kernel); requires Cin % 32 == 0 and 128 + halo <= 192
__global__ void __launch_bounds__(UC_THREADS, 2) voc_conv_umma_kernel(const MmaConvArgs a) {
  extern __shared__ __align__(128) unsigned char uc_dyn[];
  __shared__ __align__(8) unsigned long long bars[2 * UC_STAGES + 5];   // a_full[3] a_empty[3] b_full[2] b_free[2] acc_full
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int halo = a.max_shift - a.min_shift;
  const int brows = MC_BN + halo;
  const uint32_t plane_b = (uint32_t)brows * 16u;         // one k-group of the window: [rows][8 bf16]
  const uint32_t bbuf_b = 8u * plane_b;                   // hi (4 k-groups) + lo (4 k-groups)
  unsigned char* Bs = uc_dyn + UC_STAGES * UC_A_BYTES;
  const uint32_t As_s = uc_smem(uc_dyn), Bs_s = uc_smem(Bs);
  const uint32_t a_full = uc_smem(&bars[0]), a_empty = uc_smem(&bars[UC_STAGES]);
  const uint32_t b_full = uc_smem(&bars[2 * UC_STAGES]), b_free = uc_smem(&bars[2 * UC_STAGES + 2]);
  const uint32_t acc_full = uc_smem(&bars[2 * UC_STAGES + 4]);
  const int b = blockIdx.z / a.phases, phase = blockIdx.z - b * a.phases;
  const int co0 = blockIdx.y * MC_BM, q0 = blockIdx.x * MC_BN;
  const int chunks = a.Cin / MC_BK;
  const int total_steps = chunks * a.ntaps;

  if (tid == 0) {
    for (int i = 0; i < UC_STAGES; ++i) { uc_mbar_init(a_full + 8 * i, 1u); uc_mbar_init(a_empty + 8 * i, 1u); }
    for (int i = 0; i < 2; ++i) { uc_mbar_init(b_full + 8 * i, UC_STAGERS); uc_mbar_init(b_free + 8 * i, 1u); }
    uc_mbar_init(acc_full, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(uc_smem(&tmem_base_s)), "r"((uint32_t)MC_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  uc_fence_before();
  __syncthreads();
  uc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 9) {
    // ===== weight ring producer =====
    if (lane == 0) {
      const size_t tile_stride = UC_A_ELEMS;
      const int ntile = gridDim.y;
      for (int s = 0; s < total_steps; ++s) {
        const int stage = s % UC_STAGES, it = s / UC_STAGES;
        uc_mbar_wait(a_empty + 8 * stage, (uint32_t)((it & 1) ^ 1));
        const int chunk = s / a.ntaps, tap = s - chunk * a.ntaps;
        const int wt = a.tap_w[tap] + phase * a.phase_tap_step;
        const bf16* src = a.w_um + (((size_t)wt * chunks + chunk) * ntile + blockIdx.y) * tile_stride;
        uc_mbar_expect(a_full + 8 * stage, UC_A_BYTES);
        uc_bulk_g2s(As_s + stage * UC_A_BYTES, src, UC_A_BYTES, a_full + 8 * stage);
      }
    }
  } else if (warp == 8) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t accumulate = 0u;
      for (int chunk = 0; chunk < chunks; ++chunk) {
        const int buf = chunk & 1, u = chunk >> 1;
        uc_mbar_wait(b_full + 8 * buf, (uint32_t)(u & 1));
        uc_fence_after();
        for (int tap = 0; tap < a.ntaps; ++tap) {
          const int s = chunk * a.ntaps + tap;
          const int stage = s % UC_STAGES, it = s / UC_STAGES;
          uc_mbar_wait(a_full + 8 * stage, (uint32_t)(it & 1));
          uc_fence_after();
          const uint32_t roff = (uint32_t)(a.tap_shift[tap] - a.min_shift);
          const uint32_t a_hi = As_s + stage * UC_A_BYTES, a_lo = a_hi + UC_A_BYTES / 2;
          const uint32_t b_hi = Bs_s + buf * bbuf_b + roff * 16u, b_lo = b_hi + 4u * plane_b;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {                 // K = 16 per MMA: k-groups 2 ks, 2 ks + 1
            const uint64_t ad_hi = uc_desc(a_hi + ks * 2 * (MC_BM * 16), MC_BM * 16, 128);
            const uint64_t ad_lo = uc_desc(a_lo + ks * 2 * (MC_BM * 16), MC_BM * 16, 128);
            const uint64_t bd_hi = uc_desc(b_hi + ks * 2 * plane_b, plane_b, 128);
            const uint64_t bd_lo = uc_desc(b_lo + ks * 2 * plane_b, plane_b, 128);
            uc_mma(tmem, ad_hi, bd_hi, accumulate);
            accumulate = 1u;
            uc_mma(tmem, ad_hi, bd_lo, 1u);
            uc_mma(tmem, ad_lo, bd_hi, 1u);
          }
          uc_commit(a_empty + 8 * stage);                  // the stage may be refilled once these MMAs have read it
        }
        uc_commit(b_free + 8 * buf);                       // the window buffer may be restaged
      }
      uc_commit(acc_full);
    }
  } else {
    // ===== activation staging (warps 0-7) =====
    const float* xb = a.x + (size_t)b * a.Cin * a.Tin;
    const int n_items = 4 * brows;
    for (int chunk = 0; chunk < chunks; ++chunk) {
      const int buf = chunk & 1, u = chunk >> 1;
      const int c0 = chunk * MC_BK;
      float v[UC_NI][8];
#pragma unroll
      for (int j = 0; j < UC_NI; ++j) {
        const int i = tid + j * UC_STAGERS;
        const int kc = i / brows, r = i - kc * brows;
        const int t = q0 + a.min_shift + r;
        const bool ok = i < n_items && t >= 0 && t < a.Tin;
        const float* src = xb + (size_t)(c0 + 8 * kc) * a.Tin + t;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j][e] = ok ? __ldg(src + (size_t)e * a.Tin) : 0.f;
      }
      uc_mbar_wait(b_free + 8 * buf, (uint32_t)((u & 1) ^ 1));     // the MMAs of chunk - 2 have read this buffer
      unsigned char* bb = Bs + (size_t)buf * bbuf_b;
#pragma unroll
      for (int j = 0; j < UC_NI; ++j) {
        const int i = tid + j * UC_STAGERS;
        if (i < n_items) {
          const int kc = i / brows, r = i - kc * brows;
          const int t = q0 + a.min_shift + r;
          const bool ok = t >= 0 && t < a.Tin;               // out-of-range positions are zero padding, not snake(0)
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float v0 = v[j][e], v1 = v[j][e + 1];
            if (a.snake_a && ok) {
              const int ci = c0 + 8 * kc + e;
              v0 = snake_f(v0, a.snake_a[ci], a.snake_ib[ci]);
              v1 = snake_f(v1, a.snake_a[ci + 1], a.snake_ib[ci + 1]);
            }
            const bf16 h0 = f2bf(v0), h1 = f2bf(v1);
            const bf16 l0 = f2bf(v0 - bf2f(h0)), l1 = f2bf(v1 - bf2f(h1));
            hi[e >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo[e >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          unsigned char* dst = bb + (size_t)kc * plane_b + (size_t)r * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + 4 * plane_b) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      // generic-proxy stores -> visible to the tensor core's (async-proxy) reads, then one arrival per thread
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uc_mbar_arrive(b_full + 8 * buf);
    }
    // ===== epilogue: warp w reads lanes 32 (w % 4) .. +32 (the only ones it may), columns 64 (w / 4) .. +64 =====
    uc_mbar_wait(acc_full, 0u);
    uc_fence_after();
    const int lq = warp & 3, chalf = warp >> 2;
    const int mrow = co0 + lq * 32 + lane;                 // row of the GEMM: output channel, or (channel, phase) = (mrow / rdiv, mrow % rdiv)
    const int rdiv = a.rdiv > 1 ? a.rdiv : 1;
    const int co = mrow / rdiv;
    const bool co_ok = co < a.Cout;
    const float bv = (co_ok && a.bias) ? a.bias[co] : 0.f;
    const float sc = (co_ok && a.scale) ? a.scale[co] : 1.f;
    const int oo = a.out_off + phase + (mrow - co * rdiv);
    const bool vec = a.out_stride == 1 && (a.Tout & 3) == 0 && oo == 0;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int col = chalf * 64 + h * 32;
      uint32_t r[32];
      __syncwarp();
      uc_tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)col, r);
      const size_t row = ((size_t)b * a.Cout + (co_ok ? co : 0)) * a.Tout;
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 4) {
        if (!co_ok) break;
        const int q = q0 + col + j0;
        float o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float val = __uint_as_float(r[j0 + e]) + bv;
          if (a.epi == CEPI_GELU) val = 0.5f * val * (1.0f + erff(val * 0.70710678118654752440f));
          if (a.epi == CEPI_RELU) val = fmaxf(val, 0.f);
          if (a.scale) val = val * sc;
          o4[e] = val;
        }
        if (vec && q + 3 < a.Q) {
          const size_t o = row + q;
          if (a.res) {
            const float4 rv = *reinterpret_cast<const float4*>(a.res + o);
            o4[0] += rv.x; o4[1] += rv.y; o4[2] += rv.z; o4[3] += rv.w;
          }
          if (a.epi == CEPI_CLAMP) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o4[e] = fminf(fmaxf(o4[e], -1.0f), 1.0f);
          }
          *reinterpret_cast<float4*>(a.y + o) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (q + e >= a.Q) continue;
            const int t = (q + e) * a.out_stride + oo;
            if (t >= a.Tout) continue;
            const size_t o = row + t;
            float val = o4[e];
            if (a.res) val = a.res[o] + val;
            if (a.epi == CEPI_CLAMP) val = fminf(fmaxf(val, -1.0f), 1.0f);
            a.y[o] = val;
          }
        }
      }
    }
  }
  // every tcgen05.ld has completed (wait::ld) before the barrier; the allocating warp frees the columns
  uc_fence_before();
  __syncthreads();
  if (warp == 8) {
    uc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)MC_BN) : "memory");
  }
}

static size_t umma_conv_smem_bytes(int halo) { return (size_t)UC_STAGES * UC_A_BYTES + (size_t)2 * 8 * (MC_BN + halo) * 16; }

// weights [Cout][Cin][k] (conv) or [Cin][Cout][k] (transposed) -> for every (tap slot, chunk, 128-row tile) the 16 KB
// shared-memory image: [hi | lo][k-group 0..3][row 0..127][8 bf16].  rdiv = 1: row = output channel, tap slot = tap.
// rdiv = s > 1 (transposed conv of stride s, k = ntp * s): row = co * s + r, and tap slot j (input shift j - (ntp - 1)) holds
// weight tap r + (ntp - 1 - j) * s -- y[co][s q + r] = sum_ci x[ci][q] w[ci][co][r] + x[ci][q - 1] w[ci][co][r + s].
__global__ void voc_pack_umma_weights_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int k,
                                             int rows_pad, int chunks, int transposed, int rdiv, int ntp) {
  const int ntile = rows_pad / MC_BM;
  const size_t n = (size_t)ntp * chunks * ntile * (UC_A_ELEMS / 2);      // one thread per (hi, lo) element pair
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7);
    size_t r = i >> 3;
    const int row = (int)(r % MC_BM);
    r /= MC_BM;
    const int kc = (int)(r & 3);
    r >>= 2;
    const int tile = (int)(r % ntile);
    r /= ntile;
    const int chunk = (int)(r % chunks), j = (int)(r / chunks);
    const int mrow = tile * MC_BM + row, ci = chunk * MC_BK + kc * 8 + e;
    const int co = mrow / rdiv, ph = mrow - co * rdiv;
    const int wt = rdiv == 1 ? j : ph + (ntp - 1 - j) * rdiv;
    float v = 0.f;
    if (co < Cout && ci < Cin && wt < k) v = transposed ? w[((size_t)ci * Cout + co) * k + wt] : w[((size_t)co * Cin + ci) * k + wt];
    const bf16 h = f2bf(v);
    const size_t base = (((size_t)j * chunks + chunk) * ntile + tile) * UC_A_ELEMS + ((size_t)kc * MC_BM + row) * 8 + e;
    out[base] = h;
    out[base + UC_A_ELEMS / 2] = f2bf(v - bf2f(h));
  }
}
