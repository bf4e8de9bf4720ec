// Activation I/O helpers: a tensor is either plain fp32 or the split-bf16 pair (hi, lo) with hi + lo ~= fp32.
// All accessors move 4 consecutive channels (16 B fp32 / 2 x 8 B bf16).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace hfagp {

__device__ __forceinline__ void split2(float v, __nv_bfloat16& h, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(v);
  l = __float2bfloat16_rn(v - __bfloat162float(h));
}

__device__ __forceinline__ float4 bf16x4_sum(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t q) {
  const uint2 a = __ldg(reinterpret_cast<const uint2*>(hi) + q), b = __ldg(reinterpret_cast<const uint2*>(lo) + q);
  float4 r;
  r.x = __uint_as_float(a.x << 16) + __uint_as_float(b.x << 16);
  r.y = __uint_as_float(a.x & 0xffff0000u) + __uint_as_float(b.x & 0xffff0000u);
  r.z = __uint_as_float(a.y << 16) + __uint_as_float(b.y << 16);
  r.w = __uint_as_float(a.y & 0xffff0000u) + __uint_as_float(b.y & 0xffff0000u);
  return r;
}

// q = index in units of 4 elements
__device__ __forceinline__ float4 ld4_any(const float* x, const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t q) {
  return x ? __ldg(reinterpret_cast<const float4*>(x) + q) : bf16x4_sum(hi, lo, q);
}

// two fp32 -> packed bf16 pair (element 0 in the low half), round to nearest even: one cvt for both
__device__ __forceinline__ uint32_t pack2_bf16(float e0, float e1) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
  return r;
}
// same values as split2 on each element (hi = rn(v), lo = rn(v - hi)), 6 instructions per pair
__device__ __forceinline__ void st4_split(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t q, const float v[4]) {
  uint32_t hw[2], lw[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    hw[e] = pack2_bf16(v[2 * e], v[2 * e + 1]);
    lw[e] = pack2_bf16(v[2 * e] - __uint_as_float(hw[e] << 16), v[2 * e + 1] - __uint_as_float(hw[e] & 0xffff0000u));
  }
  reinterpret_cast<uint2*>(hi)[q] = make_uint2(hw[0], hw[1]);
  reinterpret_cast<uint2*>(lo)[q] = make_uint2(lw[0], lw[1]);
}

__device__ __forceinline__ void st4_any(float* y, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t q, const float v[4]) {
  if (y) reinterpret_cast<float4*>(y)[q] = make_float4(v[0], v[1], v[2], v[3]);
  else st4_split(hi, lo, q, v);
}

}  // namespace hfagp
