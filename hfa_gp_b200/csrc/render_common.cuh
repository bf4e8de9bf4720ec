// Shared pieces of the tri-plane renderer kernels (render.cu: mma.sync forward + backward; render_tc.cu: tcgen05 forward).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace hfagp {

constexpr int RC = 32;        // channels per plane == decoder input width
constexpr int RH = 64;        // decoder hidden width
constexpr int RO = 33;        // 1 sigma + 32 colour features
constexpr int R_WARPS = 16;   // rays per strip / warps per CTA (8 when 16 rays' scratch does not fit in 227 KB)
constexpr int TILE = 16;      // samples per MMA tile and per gather batch
constexpr int MLP_FLOATS = RH * RC + RH + RO * RH + RO;  // 4257

// shared-memory weight image (per CTA): B fragments of both layers as split bf16 + fp32 biases
constexpr int W0F_U2 = 8 * 2 * 2 * 32;   // [ntile 8][kstep 2][hi|lo][lane] uint2
constexpr int W1F_U2 = 5 * 4 * 2 * 32;   // [ntile 5][kstep 4][hi|lo][lane] uint2
constexpr int WEIGHT_BYTES = (W0F_U2 + W1F_U2) * 8 + (RH + 40) * 4;
// backward adds the transposed operands: dh = dout . W1 (K = 48 padded outputs, N = 64) and df = dpre . W0 (K = 64, N = 32)
constexpr int W1T_U2 = 8 * 3 * 2 * 32;   // [ntile 8][kstep 3][hi|lo][lane] uint2
constexpr int W0T_U2 = 4 * 4 * 2 * 32;   // [ntile 4][kstep 4][hi|lo][lane] uint2
constexpr int WEIGHT_BYTES_BWD = ((WEIGHT_BYTES + 15) & ~15) + (W1T_U2 + W0T_U2) * 8;

struct RenderParams {
  HfagpRenderDesc d;
  const float* planes;
  const float* cam;
  const float* mlp;
  const float* lin;
  const float* jitter;
  const float* u_fine;
  const float* depth_range;
  float* feat;
  float* depth;
  float* wsum;
  int32_t* inds;
  int32_t* below;
  int32_t* above;
  int32_t* sort_idx;
  float* depths_sorted;
  const float* dfeat;     // backward only: gradient of feat [n][res][res][32]
  float* dplanes;         // backward only: gradient of planes (accumulated with red.global.add)
  // backward only, optional: per-sample operands of the decoder-MLP weight gradient, in storage order
  // (coarse samples first): dump_f[(ray*T + j)][32] mean tri-plane features, dump_do[(ray*T + j)][33] gradient of the
  // decoder's raw outputs (column 0 = sigma, 1 + c = colour c)
  float* dump_f;
  float* dump_do;
};

struct Tap {
  uint32_t off;   // byte offset of the texel's channel 0 inside this sample's frame (0 when outside)
  float w;        // bilinear weight (0 when outside: grid_sample padding_mode='zeros')
};

__host__ __device__ inline int round16(int v) { return (v + 15) & ~15; }

// Per-warp shared memory:
//   colq  [round16(S) + round16(SF)] rows x 32 u16   decoded colours as 16-bit fixed point (see quantise note)
//   ftile [16][32] fp32                               features of the tile being decoded (MMA A operand)
//   dep, sig, sdep, ssig [Tp] fp32 ; order [Tp] u8 ; cdf [S+2], zmid [S] fp32 ; taps [16*12]
// wts (march weights) aliases sdep during the coarse pass (sdep is first written by the sort) and dep during the
// final pass (unsorted depths are dead after the sort).
__host__ __device__ inline size_t render_warp_bytes(int S, int SF, bool bwd = false) {
  const size_t Tp = (size_t)(S + SF + 3) & ~(size_t)3;   // per-sample arrays padded so each stays 16 B aligned
  size_t b = (size_t)(round16(S) + round16(SF)) * RC * 2 + TILE * RC * 4 + 4 * Tp * 4 + Tp + (size_t)(2 * S + 2) * 4;
  if (bwd) b += 5 * Tp * 4 + 32 * 4;     // tarr (T_k), aarr (alpha_k), dsg (d sigma), pj (dfeat . colour), final weights + dfeat row
  b = (b + 15) & ~(size_t)15;
  return b + TILE * 12 * sizeof(Tap);
}

// exclusive product scan over n values held as v(k) for k = lane + 32q; returns weights into wts[k] = alpha*T
// and the sum of weights.  alpha(k) supplied through a lambda.
template <typename FA>
__device__ __forceinline__ float march_weights(int nint, int lane, float* wts, FA alpha_of, float* tarr = nullptr,
                                               float* aarr = nullptr) {
  float carry = 1.f, wsum = 0.f;
  for (int base = 0; base < nint; base += 32) {
    int k = base + lane;
    float a = k < nint ? alpha_of(k) : 0.f;
    float v = k < nint ? (1.f - a + 1e-10f) : 1.f;
    float incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= t;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    float w = a * (carry * excl);
    if (k < nint) wts[k] = w;
    if (tarr && k < nint) { tarr[k] = carry * excl; aarr[k] = a; }
    wsum += k < nint ? w : 0.f;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
  return warp_sum(wsum);
}

// ---- split-bf16 helpers (packed pairs: element 0 in the low half)
__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));   // first source -> upper half
  return r;
}
// (a, b) -> hi pair and residual lo pair
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16x2(a - ah, b - bh);
}
// split_pair on a packed pair
__device__ __forceinline__ void split_pair2(f32x2 ab, uint32_t& hi, uint32_t& lo) {
  float a, b;
  upk2(ab, a, b);
  hi = pack_bf16x2(a, b);
  float ra, rb;
  upk2(sub2(ab, pk2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u))), ra, rb);
  lo = pack_bf16x2(ra, rb);
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
// softplus(x) = max(x,0) + ln2 * log2(1 + 2^(-|x| log2e)); equals torch's (beta 1, threshold 20) to fp32 rounding
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float softplus_fast(float x) {
  return fmaf(lg2f(1.f + ex2f(-1.4426950408889634f * fabsf(x))), 0.6931471805599453f, fmaxf(x, 0.f));
}
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid01(float x) { return rcpf(1.f + ex2f(-1.4426950408889634f * x)); }

// physical column of logical channel c in row r of the feature tile (row stride 32 floats): XOR-ing bits 3..4
// with the row keeps the gather's float4 stores and the MMA fragment float2 loads conflict-free
__device__ __forceinline__ int colx(int r, int c) { return c ^ ((r & 3) << 3); }
// physical 32-bit word (2 channels) of channel pair cw = c/2 in row r of the colour buffer (row stride 16 words)
__device__ __forceinline__ int colqx(int r, int cw) { return cw ^ (((r >> 1) & 3) << 2); }
// Quantise note: colour = sigmoid*1.002 - 0.001 is stored between decode and composite as q = round(65535*sigmoid)
// (step 1.53e-5, |error| <= 7.7e-6 per colour, <= 1.6e-5 on the composited feature): half the bytes per row lets
// 16 instead of 8 rays live on an SM, which is what hides the gather latency.
constexpr float QSCALE = 65535.f;
constexpr float QSTEP = 1.002f / 65535.f;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- integer bookkeeping of the importance resampling, shared by every renderer kernel and by
// hfagp_render_bookkeeping (which runs exactly this code on caller-supplied floats)
// torch.searchsorted(cdf[0..n), u, right=True) = #{ i : cdf[i] <= u } for a non-decreasing cdf
__device__ __forceinline__ int searchsorted_right(const float* cdf, int n, float u) {
  int ind = 0;
  for (int len = n; len > 0;) {
    const int half = len >> 1;
    if (cdf[ind + half] <= u) { ind += half + 1; len -= half + 1; } else { len = half; }
  }
  return ind;
}
// Stable ranks of the T depths dep[0..T) (coarse samples first, as torch.cat + sort sees them): lane holds elements
// lane + 32 e.  Fast path counts strictly smaller depths against one broadcast read of each depth; if any two depths are
// equal the ranks no longer sum to T(T-1)/2 and the exact (value, index) order is recomputed.
// The first S depths are the coarse samples: strictly increasing whenever the jitter stays inside its stratum (checked,
// not assumed), so a coarse sample's rank among them is its index and a fine sample's is a lower bound — only the SF
// fine depths are counted by comparison (half the work of the all-pairs count).  S = 0 selects the all-pairs count.
template <int NE>
__device__ __forceinline__ void stable_ranks(const float* dep, int T, int lane, float (&de)[NE], int (&rk)[NE], int S = 0) {
#pragma unroll
  for (int e = 0; e < NE; ++e) { de[e] = lane + 32 * e < T ? dep[lane + 32 * e] : 0.f; rk[e] = 0; }
  bool inc = S > 1;
  if (inc) {
    bool ok = true;
    for (int j = lane; j + 1 < S; j += 32) ok = ok && dep[j] < dep[j + 1];
    inc = __all_sync(0xffffffffu, ok);
  }
  const int j0 = inc ? S : 0;
  if (inc) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int i = lane + 32 * e;
      if (i < S) {
        rk[e] = i;
      } else {                                       // #{coarse depths < de[e]}: lower bound in the sorted prefix
        int lo = 0;
        for (int len = S; len > 0;) {
          const int half = len >> 1;
          if (dep[lo + half] < de[e]) { lo += half + 1; len -= half + 1; } else { len = half; }
        }
        rk[e] = lo;
      }
    }
  }
  for (int j = j0; j < T; ++j) {
    const float dj = dep[j];
#pragma unroll
    for (int e = 0; e < NE; ++e) rk[e] += dj < de[e] ? 1 : 0;
  }
  int rsum = 0;
#pragma unroll
  for (int e = 0; e < NE; ++e) rsum += lane + 32 * e < T ? rk[e] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
  if (rsum != T * (T - 1) / 2) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int i = lane + 32 * e;
      int rank = 0;
      for (int j = 0; j < T; ++j) {
        const float dj = dep[j];
        rank += (dj < de[e] || (dj == de[e] && j < i)) ? 1 : 0;
      }
      rk[e] = rank;
    }
  }
}

// render_tc.cu: the tcgen05 forward renderer (used by hfagp_render_fwd whenever its shared-memory plan fits)
bool render_tc_supported(const HfagpRenderDesc& d);
int render_tc_launch(const RenderParams& p, int sms, cudaStream_t stream);

}  // namespace hfagp
