// Shared helpers for libhfagp_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include "../../include/hfagp.h"

namespace hfagp {

// thread-local last-error string (hfagp_last_error)
char* err_buf();
int fail(int code, const char* fmt, ...);

#define HFAGP_CHECK_ARG(cond, ...)                               \
  do {                                                           \
    if (!(cond)) return ::hfagp::fail(HFAGP_E_INVALID, __VA_ARGS__); \
  } while (0)

#define HFAGP_CHECK_LAUNCH(name)                                                         \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess)                                                              \
      return ::hfagp::fail(HFAGP_E_CUDA, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define HFAGP_CUDA(call)                                                                 \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ::hfagp::fail(HFAGP_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));   \
  } while (0)

// programmatic dependent launch of the kernels that support it (tc_common.cuh); opt-in with HFAGP_PDL=1
bool pdl_enabled();

// Per-DEVICE one-time setup (cudaFuncSetAttribute applies to the device that is current when it is called, so a
// process that drives several GPUs must repeat it on each): runs `f` the first time the calling thread's current
// device is seen through `mask`; returns f's error (cudaSuccess afterwards).  f is idempotent, so a benign race between
// two first callers only repeats it.
template <typename F>
inline cudaError_t per_device_once(std::atomic<uint64_t>& mask, F&& f) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = 1ull << (dev & 63);
  if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = f();
  if (e == cudaSuccess) mask.fetch_or(bit, std::memory_order_release);
  return e;
}
// SM count of the current device (cached per device ordinal)
int device_sm_count();

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }
// negative-side slope of an HFAGP_ACT_* activation (linear 1, leaky-ReLU 0.2, ReLU 0): act(v) = max(v, slope*v)
__host__ __device__ __forceinline__ float act_slope(int act) { return act == HFAGP_ACT_LRELU ? 0.2f : (act == HFAGP_ACT_RELU ? 0.f : 1.f); }

// softplus with torch semantics (beta=1, threshold=20)
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2: two fp32 operations per issue slot; a pair is a 64-bit register,
// element 0 in the low half)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hfagp
