// tcgen05 + TMA GEMM for the multi-token projections of the path (prompt prefill, text projection):
//   Y[t][n] = epi( sum_k W[n][k] X[t][k] ),  W: [N][K] bf16 (checkpoint layout, K-major), X: [T][K] bf16, T > 16.
// ref: candle Linear::forward -> cuBLAS GEMM at transformer.rs:258-260, 371, 408-413 and talker.rs:294-321 when the
// sequence dimension is the prompt length (run_prefill_layers, talker.rs:823-841).
//
// Blackwell-native structure (one 128 x 128 output tile per CTA, swap-AB so the weight rows fill the 128-lane M
// dimension and the tokens are N):
//   warp 0 / one lane : TMA producer -- cp.async.bulk.tensor.2d of a [128 x 64] bf16 tile of W (and of the
//                       SwiGLU partner W2) and a [128 x 64] tile of X per k-block into a 3/4-stage shared-memory
//                       ring, 128B-swizzled, completion on an mbarrier (expect_tx);
//   warp 1 / one lane : MMA issuer -- tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> f32), M = 128, N = 128,
//                       K = 16, four per k-block, accumulators in TMEM (128 or 256 columns); tcgen05.commit frees
//                       the stage (empty barrier) and finally signals the epilogue;
//   all 4 warps       : epilogue -- tcgen05.ld 32x32b (each warp its 32 TMEM lanes = 32 weight rows), fused
//                       bias / SiLU / SwiGLU / residual with the reference's rounding points, coalesced bf16 stores.
// Out-of-range tokens are zero-filled by TMA; N must be a multiple of 128 and K of 64 (true for every Qwen3-TTS
// projection), otherwise the caller uses the CUDA-core GEMV.
#pragma once
#include <mutex>
#include <cuda.h>

#include "common.cuh"
#include "gemv.cuh"

constexpr int TC_BM_ = 128, TC_BN_ = 128, TC_BK_ = 64;
constexpr uint32_t TC_TILE_BYTES = TC_BM_ * TC_BK_ * 2;   // 16 KB

struct GemmTcArgs {
  int N, K, T;
  int epi;
  bf16* Y;
  int ldy;
  const bf16* bias;
  const bf16* R;
  int ldr;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a protocol bug traps instead of hanging the GPU
  for (uint32_t it = 0; it < (1u << 28); ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 128B-swizzled, K-major operand tile [rows][64 bf16]: 8-row atoms of 1024 B (stride byte offset), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address
  d |= (uint64_t)1 << 16;                               // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset
  d |= (uint64_t)1 << 46;                               // descriptor version
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <bool DUAL>
__global__ void __launch_bounds__(128, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                        const __grid_constant__ CUtensorMap tmap_w2,
                                                        const __grid_constant__ CUtensorMap tmap_x, const GemmTcArgs a) {
  constexpr int STAGES = DUAL ? 3 : 4;
  constexpr uint32_t STAGE_BYTES = (DUAL ? 3 : 2) * TC_TILE_BYTES;
  constexpr uint32_t TMEM_COLS = DUAL ? 256 : 128;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B the 128B swizzle needs
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[4], empty_bar[4], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TC_BM_, t0 = blockIdx.y * TC_BN_;
  const int kblocks = a.K / TC_BK_;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  if (warp == 0) {    // TMEM allocation by one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int kb = 0; kb < kblocks; ++kb) {
      const int s = kb % STAGES;
      if (kb >= STAGES) mbar_wait(&empty_bar[s], ((kb / STAGES) - 1) & 1);
      unsigned char* st = tiles + (size_t)s * STAGE_BYTES;
      mbar_expect_tx(&full_bar[s], STAGE_BYTES);
      tma_load_2d(st, &tmap_w, &full_bar[s], kb * TC_BK_, n0);
      tma_load_2d(st + TC_TILE_BYTES, &tmap_x, &full_bar[s], kb * TC_BK_, t0);
      if (DUAL) tma_load_2d(st + 2 * TC_TILE_BYTES, &tmap_w2, &full_bar[s], kb * TC_BK_, n0);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    // instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN_ >> 3) << 17) | ((uint32_t)(TC_BM_ >> 4) << 24);
    for (int kb = 0; kb < kblocks; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full_bar[s], (kb / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE_BYTES);
      const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + TC_TILE_BYTES);
      const uint64_t a2desc = umma_desc_sw128(sa + 2 * TC_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < TC_BK_ / 16; ++k) {
        const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
        umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);            // +32 B per K = 16
        if (DUAL) umma_bf16(tmem_base + 128, a2desc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
      }
      umma_commit(&empty_bar[s]);                       // stage reusable once these MMAs have read it
    }
    umma_commit(&done_bar);                             // accumulators complete
  }
  __syncwarp();
  // ===== epilogue: every warp drains its 32 TMEM lanes (= 32 weight rows) =====
  mbar_wait(&done_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int n = n0 + warp * 32 + lane;
  const float bv = (a.bias != nullptr) ? bf2f(a.bias[n]) : 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < TC_BN_; c0 += 32) {
    uint32_t r[32], r2[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    tmem_ld32(taddr, r);
    if (DUAL) tmem_ld32(taddr + 128, r2);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int t = t0 + c0 + j;
      if (t < a.T) {
        const float v = rbf(__uint_as_float(r[j]));
        float y;
        switch (a.epi) {
          case EPI_BIAS: y = v + bv; break;
          case EPI_BIAS_SILU: y = silu_f(rbf(v + bv)); break;
          case EPI_RESIDUAL: y = bf2f(a.R[(size_t)t * a.ldr + n]) + v; break;
          case EPI_SWIGLU: y = rbf(silu_f(v)) * rbf(__uint_as_float(r2[j])); break;
          default: y = v; break;
        }
        a.Y[(size_t)t * a.ldy + n] = f2bf(y);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    Q3_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    Q3_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, Q3_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// row-major [rows][K] bf16 tensor, box = [128 rows][64 k], 128B swizzle, zero fill out of range
static CUtensorMap make_tmap_2d(const bf16* ptr, int rows, int K, int ld) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK_, (cuuint32_t)TC_BM_};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  Q3_REQUIRE(r == CUDA_SUCCESS, Q3_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return m;
}

static bool gemm_tc_supported(int N, int K, int T, int ldx, int epi) {
  if (N % TC_BM_ != 0 || K % TC_BK_ != 0 || T < 1 || (ldx * 2) % 16 != 0) return false;
  return epi == EPI_STORE || epi == EPI_BIAS || epi == EPI_BIAS_SILU || epi == EPI_RESIDUAL || epi == EPI_SWIGLU;
}

// Y = epi(X W^T); X: [T][ldx] bf16 (already normalised when the op has an RMSNorm prologue)
static void gemm_tc_launch(const bf16* W, const bf16* W2, const bf16* X, int ldx, const GemmTcArgs& a, cudaStream_t st) {
  const bool dual = a.epi == EPI_SWIGLU;
  CUtensorMap tw = make_tmap_2d(W, a.N, a.K, a.K);
  CUtensorMap tw2 = dual ? make_tmap_2d(W2, a.N, a.K, a.K) : tw;
  CUtensorMap tx = make_tmap_2d(X, a.T, a.K, ldx);
  dim3 grid(a.N / TC_BM_, ceil_div(a.T, TC_BN_));
  const size_t smem_single = 4 * 2 * TC_TILE_BYTES + 1024, smem_dual = 3 * 3 * TC_TILE_BYTES + 1024;
  {
    // the shared-memory opt-in is a per-DEVICE function attribute: one process may hold models on several GPUs
    // (q3_model_desc.device) and sessions launch from several threads
    static std::mutex mu;
    static bool configured[64] = {};
    int dev = 0;
    Q3_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      Q3_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_single));
      Q3_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dual));
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  if (dual) gemm_tc_kernel<true><<<grid, 128, smem_dual, st>>>(tw, tw2, tx, a);
  else gemm_tc_kernel<false><<<grid, 128, smem_single, st>>>(tw, tw2, tx, a);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
}
