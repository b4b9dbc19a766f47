// Shared device/host helpers for libq3tts_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/q3tts.h"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// host-side error plumbing: C++ exceptions are used internally and converted to q3_status at the
// ABI boundary (abi.cu); none escapes.
struct Q3Error : public std::runtime_error {
  q3_status code;
  Q3Error(q3_status c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define Q3_CHECK_CUDA(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw Q3Error(Q3_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +    \
                                     __FILE__ + ":" + std::to_string(__LINE__) + ")");         \
  } while (0)

#define Q3_REQUIRE(cond, code, msg)                                                             \
  do {                                                                                          \
    if (!(cond)) throw Q3Error(code, std::string(msg));                                        \
  } while (0)

extern std::atomic<uint64_t> g_q3_launches;   // kernels launched by this library (abi.cu)
#define Q3_COUNT_LAUNCH() (g_q3_launches.fetch_add(1, std::memory_order_relaxed))
#define Q3_LAUNCH_CHECK() Q3_CHECK_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------------------------
// device helpers
__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }
// round an f32 value to the nearest bf16 and widen again: the "every candle op writes a bf16
// tensor" rounding point of the reference's CUDA path.
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// two packed bf16 (little endian: low half = element 0) -> two floats, exact.
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  // weights are read exactly once per pass: bypass L1 allocation
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum_xor(float v) {
  // xor butterfly, offsets 16,8,4,2,1: the tree of kernels/fused_residual_rmsnorm.cu:29-34
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
