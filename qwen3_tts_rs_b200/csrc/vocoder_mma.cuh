// Tensor-core implicit-GEMM convolution for the vocoder (causal conv1d with dilation, 1x1 projections and the
// phases of the causal transposed conv), F32-accurate.
//
//   Y[co][q*os + oo] = epi( bias[co] + sum_{j < ntaps} sum_{ci} W[j][co][ci] * act(X[ci][q + shift_j]) )
//
// The reference runs the vocoder in F32 (src/lib.rs:344-345) and the parity bar is 1e-3 RMS on PCM, so a plain
// TF32/bf16 GEMM is not accurate enough through ~40 chained layers with sin^2 activations.  Each F32 operand is
// split into two bf16 terms (hi = bf16(x), lo = bf16(x - hi)) and the product is evaluated as
// hi*hi + hi*lo + lo*hi with F32 accumulation on the tensor cores (mma.sync.m16n8k16.bf16): relative error
// ~2^-16 per product -- F32-class for this purpose -- at one third of the bf16 tensor rate, which is still an
// order of magnitude above the FP32 CUDA-core ceiling.  Weights are split and re-packed once at model load;
// activations are split while they are staged into shared memory (where the SnakeBeta prologue is also applied),
// so HBM traffic is the plain F32 activations, read once per 128-wide output-channel tile.
//
// Tiling: CTA = 128 (co) x 128 (q) outputs, 8 warps as 4 (co) x 2 (q), warp tile 32 x 64;  K loop over
// 32-channel chunks x taps.  The staged activation window [128 + halo positions][32 ch] is shared by all taps
// (a tap is just a row offset), stored position-major so that ldmatrix feeds the B fragments for any shift.
#pragma once
#include "common.cuh"
#include "vocoder_kernels.cuh"

constexpr int MC_BM = 128, MC_BN = 128, MC_BK = 32, MC_PITCH = 40;   // bf16 elements per smem row (32 + 8 pad)
constexpr int MC_MAX_TAPS = 8;

struct MmaConvArgs {
  const float* x;         // [B][Cin][Tin]
  const bf16* w_hi;       // packed [ntaps_total][chunks][Cout_pad][32]  (Cout_pad = multiple of 128)
  const bf16* w_lo;
  const bf16* w_um;       // the same weights as shared-memory images of the tcgen05 kernel (vocoder_umma.cuh), or null
  const float* bias;      // [Cout] or null
  const float* snake_a;   // [Cin] or null
  const float* snake_ib;
  const float* res;       // [B][Cout][Tout] or null
  const float* scale;     // [Cout] or null
  float* y;               // [B][Cout][Tout]
  int B, Cin, Cout, Cout_pad, Tin, Tout, Q;   // Q = number of q positions (outputs per phase)
  int ntaps;
  int tap_w[MC_MAX_TAPS];      // weight tap index
  int tap_shift[MC_MAX_TAPS];  // input position = q + shift
  int min_shift, max_shift;
  int out_stride, out_off;     // output position = q*out_stride + out_off
  int epi;
  int phases;                  // transposed conv: gridDim.z = B * phases, phase r adds r to tap_w / out_off
  int phase_tap_step;          // weight tap index += r * phase_tap_step (=1 for tconv)
  int rdiv;                    // tcgen05 kernel only: GEMM row = co * rdiv + phase (transposed conv as one GEMM), 0 / 1 = plain
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* smem_ptr) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem_ptr, const void* gptr) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(addr), "l"(gptr));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// smem: Bs_hi/Bs_lo [MC_BN + halo][MC_PITCH] bf16 ; As_hi/As_lo [2 stages][MC_BM][MC_PITCH] bf16
__global__ void __launch_bounds__(256, 2) voc_conv_mma_kernel(const MmaConvArgs a) {
  extern __shared__ __align__(16) unsigned char mc_smem[];
  const int halo = a.max_shift - a.min_shift;
  const int brows = MC_BN + halo;
  bf16* Bs_hi = reinterpret_cast<bf16*>(mc_smem);
  bf16* Bs_lo = Bs_hi + (size_t)brows * MC_PITCH;
  bf16* As_hi = Bs_lo + (size_t)brows * MC_PITCH;          // [2][MC_BM][MC_PITCH]
  bf16* As_lo = As_hi + 2 * MC_BM * MC_PITCH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;                  // warp tile origin: rows wm*32, cols wn*64
  const int b = blockIdx.z / a.phases, phase = blockIdx.z - b * a.phases;
  const int co0 = blockIdx.y * MC_BM, q0 = blockIdx.x * MC_BN;
  const int chunks = (a.Cin + MC_BK - 1) / MC_BK;
  const float* xb = a.x + (size_t)b * a.Cin * a.Tin;

  float acc[2][8][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  auto load_A = [&](int stage, int chunk, int tap) {
    // one A tile = 128 rows x 64 bytes, contiguous in the packed weights; 512 16-byte pieces per plane
    const int wt = a.tap_w[tap] + phase * a.phase_tap_step;
    const size_t base = (((size_t)wt * chunks + chunk) * a.Cout_pad + co0) * MC_BK;
    for (int i = tid; i < MC_BM * 4; i += 256) {
      const int row = i >> 2, piece = i & 3;
      cp_async16(As_hi + ((size_t)stage * MC_BM + row) * MC_PITCH + piece * 8, a.w_hi + base + (size_t)row * MC_BK + piece * 8);
      cp_async16(As_lo + ((size_t)stage * MC_BM + row) * MC_PITCH + piece * 8, a.w_lo + base + (size_t)row * MC_BK + piece * 8);
    }
    cp_async_commit();
  };

  const int total_steps = chunks * a.ntaps;
  load_A(0, 0, 0);
  // A thread's (up to 12) loads of the activation window are issued back to back before any of them is used: ncu on
  // the late decoder blocks showed 32% (k=7 convs) to 64% (1x1 convs) of the warp stalls on these loads when each
  // one was consumed right after it was issued (a chain of ~11 exposed global-memory latencies per channel chunk).
  // Keeping them in flight ACROSS the tap loop instead spills (64 accumulators + fragments + 24 values > 128 regs).
  constexpr int MC_NP = 12;                                 // >= 16 channel pairs * (128 + 54 halo) rows / 256 threads
  const int n_pairs = (MC_BK / 2) * brows;
  float pv0[MC_NP], pv1[MC_NP];
  auto fetch_window = [&](int chunk) {
    const int c0 = chunk * MC_BK;
#pragma unroll
    for (int j = 0; j < MC_NP; ++j) {
      const int i = tid + j * 256;
      pv0[j] = 0.f; pv1[j] = 0.f;
      if (i < n_pairs) {
        const int cp = i / brows, r = i - cp * brows;       // channel pair, window row (consecutive threads: consecutive t)
        const int t = q0 + a.min_shift + r;
        const int ci = c0 + 2 * cp;
        if (t >= 0 && t < a.Tin) {
          if (ci < a.Cin) pv0[j] = __ldg(xb + (size_t)ci * a.Tin + t);
          if (ci + 1 < a.Cin) pv1[j] = __ldg(xb + (size_t)(ci + 1) * a.Tin + t);
        }
      }
    }
  };
  for (int chunk = 0; chunk < chunks; ++chunk) {
    // ---- stage the activation window of this channel chunk: positions q0+min_shift .. q0+127+max_shift ----
    fetch_window(chunk);                                     // all of this thread's loads in flight together
    __syncthreads();                                         // previous chunk's readers are done with Bs
    const int c0 = chunk * MC_BK;
#pragma unroll
    for (int j = 0; j < MC_NP; ++j) {
      const int i = tid + j * 256;
      if (i < n_pairs) {
        const int cp = i / brows, r = i - cp * brows;
        const int t = q0 + a.min_shift + r;
        const int ci = c0 + 2 * cp;
        float v0 = pv0[j], v1 = pv1[j];
        if (a.snake_a && t >= 0 && t < a.Tin) {             // out-of-range positions are zero padding, not snake(0)
          if (ci < a.Cin) v0 = snake_f(v0, a.snake_a[ci], a.snake_ib[ci]);
          if (ci + 1 < a.Cin) v1 = snake_f(v1, a.snake_a[ci + 1], a.snake_ib[ci + 1]);
        }
        const bf16 h0 = f2bf(v0), h1 = f2bf(v1);
        const bf16 l0 = f2bf(v0 - bf2f(h0)), l1 = f2bf(v1 - bf2f(h1));
        __nv_bfloat162 hh, ll;
        hh.x = h0; hh.y = h1; ll.x = l0; ll.y = l1;
        *reinterpret_cast<__nv_bfloat162*>(Bs_hi + (size_t)r * MC_PITCH + 2 * cp) = hh;
        *reinterpret_cast<__nv_bfloat162*>(Bs_lo + (size_t)r * MC_PITCH + 2 * cp) = ll;
      }
    }
    for (int tap = 0; tap < a.ntaps; ++tap) {
      const int step = chunk * a.ntaps + tap;
      const int stage = step & 1;
      // prefetch the next A tile (next tap, or tap 0 of the next chunk)
      if (step + 1 < total_steps) {
        const int nt = tap + 1 < a.ntaps ? tap + 1 : 0;
        const int nc = tap + 1 < a.ntaps ? chunk : chunk + 1;
        load_A(stage ^ 1, nc, nt);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();                                       // A(stage) landed; Bs of this chunk is complete
      const int roff = a.tap_shift[tap] - a.min_shift;       // window row of output column 0 for this tap
#pragma unroll
      for (int kk = 0; kk < MC_BK; kk += 16) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          const int row = wm * 32 + mi * 16 + (lane & 15);
          const int col = kk + (lane >> 4) * 8;
          ldmatrix_x4(ah[mi][0], ah[mi][1], ah[mi][2], ah[mi][3], As_hi + ((size_t)stage * MC_BM + row) * MC_PITCH + col);
          ldmatrix_x4(al[mi][0], al[mi][1], al[mi][2], al[mi][3], As_lo + ((size_t)stage * MC_BM + row) * MC_PITCH + col);
        }
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {                     // two n8 tiles per ldmatrix.x4
          const int n = wn * 64 + nj * 16 + (lane & 7) + ((lane >> 4) << 3);
          const int col = kk + ((lane >> 3) & 1) * 8;
          uint32_t bh[4], bl[4];
          ldmatrix_x4(bh[0], bh[1], bh[2], bh[3], Bs_hi + (size_t)(roff + n) * MC_PITCH + col);
          ldmatrix_x4(bl[0], bl[1], bl[2], bl[3], Bs_lo + (size_t)(roff + n) * MC_PITCH + col);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            mma16816(acc[mi][2 * nj], ah[mi], bh[0], bh[1]);
            mma16816(acc[mi][2 * nj], ah[mi], bl[0], bl[1]);
            mma16816(acc[mi][2 * nj], al[mi], bh[0], bh[1]);
            mma16816(acc[mi][2 * nj + 1], ah[mi], bh[2], bh[3]);
            mma16816(acc[mi][2 * nj + 1], ah[mi], bl[2], bl[3]);
            mma16816(acc[mi][2 * nj + 1], al[mi], bh[2], bh[3]);
          }
        }
      }
      __syncthreads();                                       // everyone is done with A(stage) before it is refilled
    }
  }
  // ---- epilogue ----
  const int g = lane >> 2, tg = lane & 3;
  const int oo = a.out_off + phase;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int co = co0 + wm * 32 + mi * 16 + g + half * 8;
      if (co >= a.Cout) continue;
      const float bv = a.bias ? a.bias[co] : 0.f;
      const float sc = a.scale ? a.scale[co] : 1.f;
#pragma unroll
      for (int nj = 0; nj < 8; ++nj)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int q = q0 + wn * 64 + nj * 8 + 2 * tg + e;
          if (q >= a.Q) continue;
          const int t = q * a.out_stride + oo;
          if (t >= a.Tout) continue;
          const size_t o = ((size_t)b * a.Cout + co) * a.Tout + t;
          float v = acc[mi][nj][half * 2 + e] + bv;
          if (a.epi == CEPI_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
          if (a.epi == CEPI_RELU) v = fmaxf(v, 0.f);
          if (a.scale) v = v * sc;
          if (a.res) v = a.res[o] + v;
          if (a.epi == CEPI_CLAMP) v = fminf(fmaxf(v, -1.0f), 1.0f);
          a.y[o] = v;
        }
    }
}

static size_t mma_conv_smem_bytes(int halo) {
  return ((size_t)2 * (MC_BN + halo) * MC_PITCH + (size_t)4 * MC_BM * MC_PITCH) * sizeof(bf16);
}

// weights [Cout][Cin][k] (conv) or [Cin][Cout][k] (transposed) -> hi/lo bf16 packed [k][chunks][Cout_pad][32]
__global__ void voc_pack_mma_weights_kernel(const float* __restrict__ w, bf16* __restrict__ hi, bf16* __restrict__ lo, int Cout,
                                            int Cin, int k, int Cout_pad, int chunks, int transposed) {
  const size_t n = (size_t)k * chunks * Cout_pad * MC_BK;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % MC_BK);
    size_t r = i / MC_BK;
    const int co = (int)(r % Cout_pad);
    r /= Cout_pad;
    const int chunk = (int)(r % chunks), j = (int)(r / chunks);
    const int ci = chunk * MC_BK + c;
    float v = 0.f;
    if (co < Cout && ci < Cin) v = transposed ? w[((size_t)ci * Cout + co) * k + j] : w[((size_t)co * Cin + ci) * k + j];
    const bf16 h = f2bf(v);
    hi[i] = h;
    lo[i] = f2bf(v - bf2f(h));
  }
}
