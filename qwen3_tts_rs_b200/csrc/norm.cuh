// RMSNorm building blocks.
//
// The reference's one native kernel, kernels/fused_residual_rmsnorm.cu (and candle's rmsnorm it was
// adapted from), fixes the floating-point SUMMATION ORDER of sum(x^2): with block_size
// bs = ncols < 1024 ? 32 : 1024 (src/models/fused_ops.rs:161-166) "thread" tid accumulates columns
// tid, tid+bs, tid+2bs, ... with an FMA chain, each group of 32 consecutive tids is combined by an
// xor-butterfly (offsets 16,8,4,2,1), and for bs = 1024 the 32 group sums are combined by a second
// butterfly (kernels/fused_residual_rmsnorm.cu:57-82).  The shipped PTX was built with fast-math:
// mean = div.approx.ftz(sum, ncols); scale = rsqrt.approx.ftz(mean + eps); out = (scale*x)*w.
//
// The routines here reproduce that exact tree -- so results are bit-identical to the reference kernel
// -- while loading 8 consecutive bf16 (128 bit) per thread: thread j of a 128-thread group owns the
// partial sums of reference tids 8j..8j+7, the butterfly levels 16 and 8 become lane shuffles by 2 and
// 1, and levels 4,2,1 are in-register adds.
#pragma once
#include "common.cuh"

__device__ __forceinline__ float ref_mean_rsqrt(float sumsq, int ncols, float eps) {
  // The reference source says rsqrtf(tmp / ncols + eps); built with fast-math (as the shipped PTX was)
  // ptxas lowers div.approx to MUFU.RCP and contracts the multiply with the following add:
  //   t = fma(rcp(ncols), sumsq, eps);  scale = MUFU.RSQ(t).
  // Spelled out here so the result does not depend on this translation unit's contraction choices.
  float rn, t, sc;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"((float)ncols));
  asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(rn), "f"(sumsq), "f"(eps));
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(sc) : "f"(t));
  return sc;
}

// ---- ncols < 1024: one warp, lane == reference tid ------------------------------------------------
// get(col) returns the f32 value of column col.  All 32 lanes must call; result in every lane.
template <typename F>
__device__ __forceinline__ float sumsq_ref_small(int ncols, F get) {
  const int lane = threadIdx.x & 31;
  float tmp = 0.f;
  for (int c = lane; c < ncols; c += 32) {
    float v = get(c);
    tmp = fmaf(v, v, tmp);
  }
  return warp_sum_xor(tmp);
}

// ---- ncols >= 1024: 128 cooperating threads (4 full warps), g = index within the group -------------
// p[e] holds the FMA-chain partial of reference tid 8g+e.  s_part: 32 floats of shared memory private
// to this group.  bar_id: named barrier id (1..15) private to this group.  Result in every thread.
__device__ __forceinline__ float sumsq_ref_large_finish(float p[8], int g, float* s_part, int bar_id) {
#pragma unroll
  for (int e = 0; e < 8; ++e) p[e] += __shfl_xor_sync(0xffffffffu, p[e], 2);   // butterfly level 16
#pragma unroll
  for (int e = 0; e < 8; ++e) p[e] += __shfl_xor_sync(0xffffffffu, p[e], 1);   // level 8
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = p[e] + p[e ^ 4];                            // level 4
#pragma unroll
  for (int e = 0; e < 8; ++e) b[e] = a[e] + a[e ^ 2];                            // level 2
  float sw = b[0] + b[1];                                                       // level 1
  if ((g & 3) == 0) s_part[g >> 2] = sw;                                         // s_sum[warp_id]
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(128));
  float t = s_part[threadIdx.x & 31];
  t = warp_sum_xor(t);
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(128));   // s_part may be reused after this
  return t;
}

// bf16 row held as packed uint4 chunks: chunk index q covers columns 8q..8q+7.
__device__ __forceinline__ void unpack8(const uint4& u, float f[8]) {
  f[0] = bf_lo(u.x); f[1] = bf_hi(u.x); f[2] = bf_lo(u.y); f[3] = bf_hi(u.y);
  f[4] = bf_lo(u.z); f[5] = bf_hi(u.z); f[6] = bf_lo(u.w); f[7] = bf_hi(u.w);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float f[8]) {
  uint4 u;
  u.x = pack2(f[0], f[1]); u.y = pack2(f[2], f[3]); u.z = pack2(f[4], f[5]); u.w = pack2(f[6], f[7]);
  return u;
}
