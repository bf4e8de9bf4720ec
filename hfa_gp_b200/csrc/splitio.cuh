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

__device__ __forceinline__ void st4_split(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t q, const float v[4]) {
  uint32_t hw[2], lw[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    __nv_bfloat16 h0, h1, l0, l1;
    split2(v[2 * e], h0, l0);
    split2(v[2 * e + 1], h1, l1);
    hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  reinterpret_cast<uint2*>(hi)[q] = make_uint2(hw[0], hw[1]);
  reinterpret_cast<uint2*>(lo)[q] = make_uint2(lw[0], lw[1]);
}

__device__ __forceinline__ void st4_any(float* y, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t q, const float v[4]) {
  if (y) reinterpret_cast<float4*>(y)[q] = make_float4(v[0], v[1], v[2], v[3]);
  else st4_split(hi, lo, q, v);
}

}  // namespace hfagp
