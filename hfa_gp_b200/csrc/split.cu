// Split-bf16 operand producers for the tensor-core path: x (fp32) -> hi = bf16(x), lo = bf16(x - hi).
#include <cuda_bf16.h>
#include "common.cuh"
#include "splitio.cuh"

namespace hfagp {

__global__ void split_kernel(size_t count, const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  split2(__ldg(x + i), hi[i], lo[i]);
}

// 4 values per thread (16-byte loads, 8-byte stores)
__global__ void split4_kernel(size_t count4, const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                              __nv_bfloat16* __restrict__ lo) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
  const float v[4] = {a.x, a.y, a.z, a.w};
  st4_split(hi, lo, i, v);
}

// The encoder's first layer at inference: ConvLayer(3, C, 1) = 1x1 convolution of the NCHW frame + FusedLeakyReLU
// (encoder3d.py:142-179 with kernel_size 1), written straight as the split-bf16 channels-last operand of the next
// tensor-core convolution: replaces nchw_to_nhwc + the SIMT implicit GEMM + the fp32 -> split pass (three round trips of the
// 67 MB activation at batch 4) by one kernel that reads the 3 MB frame and writes the operand once.
// thread = (pixel, 4 output channels); w [cout][cin] fp32 (equalised-lr scale folded in), cin <= 4.
__global__ void __launch_bounds__(256) stem_conv1x1_kernel(int hw, int cin, int cout4, const float* __restrict__ x,
                                                          const float* __restrict__ w, const float* __restrict__ bias, float slope,
                                                          float gain, __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= (uint32_t)hw * cout4) return;
  const int pix = i / cout4, cq = i - pix * cout4;
  const int n = blockIdx.y;
  const float* xp = x + (size_t)n * cin * hw + pix;
  float xin[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) xin[c] = c < cin ? __ldg(xp + (size_t)c * hw) : 0.f;
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = cq * 4 + k;
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < cin) a = fmaf(xin[c], __ldg(w + co * cin + c), a);
    a += bias ? __ldg(bias + co) : 0.f;
    o[k] = fmaxf(a, slope * a) * gain;
  }
  st4_split(y_hi, y_lo, ((size_t)n * hw + pix) * cout4 + cq, o);
}

// block per (o, n): wmod = w * s written as split bf16, dcoef from the fp32 products.  A thread owns 4 consecutive
// input channels (its style values stay in registers) and walks the taps: 16-byte loads, 8-byte stores.
__global__ void modulate_split_kernel(int ntaps, int cout, int cin, const float* __restrict__ w,
                                      const float* __restrict__ styles, __nv_bfloat16* __restrict__ whi,
                                      __nv_bfloat16* __restrict__ wlo, float* __restrict__ dcoef) {
  const int o = blockIdx.x, n = blockIdx.y;
  const int c4 = cin >> 2;
  float ss = 0.f;
  for (int q = threadIdx.x; q < c4; q += blockDim.x) {
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(styles + (size_t)n * cin) + q);
    for (int t = 0; t < ntaps; ++t) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + ((size_t)t * cout + o) * cin) + q);
      const float v[4] = {w4.x * s4.x, w4.y * s4.y, w4.z * s4.z, w4.w * s4.w};
      st4_split(whi, wlo, ((((size_t)n * ntaps + t) * cout + o) * cin >> 2) + q, v);
      ss = fmaf(v[0], v[0], ss); ss = fmaf(v[1], v[1], ss); ss = fmaf(v[2], v[2], ss); ss = fmaf(v[3], v[3], ss);
    }
  }
  if (dcoef) {
    __shared__ float red[32];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      v = warp_sum(v);
      if (threadIdx.x == 0) dcoef[(size_t)n * cout + o] = rsqrtf(v + 1e-8f);
    }
  }
}

// All layers of a network in ONE launch: a block per (layer, output channel, sample).  The weights of a frame are
// modulated by styles that are all known before the first convolution runs, so the 24 per-layer launches of the
// inference path collapse into one HBM-streaming pass over the 113 MB of generator weights.
constexpr int MAX_MOD_LAYERS = 40;
struct ModTable {
  const float* w[MAX_MOD_LAYERS];
  __nv_bfloat16* hi[MAX_MOD_LAYERS];
  __nv_bfloat16* lo[MAX_MOD_LAYERS];
  float* dcoef[MAX_MOD_LAYERS];          // null: no demodulation (ToRGB)
  long long soff[MAX_MOD_LAYERS];        // offset of the layer's styles [batch][cin] in the flat styles buffer
  int ntaps[MAX_MOD_LAYERS], cout[MAX_MOD_LAYERS], cin[MAX_MOD_LAYERS];
  int row_start[MAX_MOD_LAYERS + 1];     // prefix sum of cout
  int nlayers;
};

__global__ void __launch_bounds__(128) modulate_split_multi_kernel(const __grid_constant__ ModTable tb,
                                                                  const float* __restrict__ styles) {
  const int row = blockIdx.x, n = blockIdx.y;
  int l = 0;
  while (row >= tb.row_start[l + 1]) ++l;
  const int o = row - tb.row_start[l];
  const int ntaps = tb.ntaps[l], cout = tb.cout[l], cin = tb.cin[l], c4 = cin >> 2;
  const float* w = tb.w[l];
  __nv_bfloat16* whi = tb.hi[l];
  __nv_bfloat16* wlo = tb.lo[l];
  const float* sn = styles + tb.soff[l] + (size_t)n * cin;
  float ss = 0.f;
  for (int q = threadIdx.x; q < c4; q += blockDim.x) {
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(sn) + q);
    for (int t = 0; t < ntaps; ++t) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + ((size_t)t * cout + o) * cin) + q);
      const float v[4] = {w4.x * s4.x, w4.y * s4.y, w4.z * s4.z, w4.w * s4.w};
      st4_split(whi, wlo, ((((size_t)n * ntaps + t) * cout + o) * cin >> 2) + q, v);
      ss = fmaf(v[0], v[0], ss); ss = fmaf(v[1], v[1], ss); ss = fmaf(v[2], v[2], ss); ss = fmaf(v[3], v[3], ss);
    }
  }
  if (tb.dcoef[l]) {
    __shared__ float red[4];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) tb.dcoef[l][(size_t)n * cout + o] = rsqrtf(red[0] + red[1] + red[2] + red[3] + 1e-8f);
  }
}

}  // namespace hfagp

using namespace hfagp;

extern "C" int hfagp_modulate_split_multi_fwd(int nlayers, int batch, const float* const* w_host, const int32_t* ntaps_host,
                                              const int32_t* cout_host, const int32_t* cin_host,
                                              const int64_t* styles_off_host, const float* styles,
                                              uint16_t* const* hi_host, uint16_t* const* lo_host,
                                              float* const* dcoef_host, void* stream) {
  HFAGP_CHECK_ARG(nlayers > 0 && nlayers <= MAX_MOD_LAYERS && batch > 0 && batch <= 65535, "modulate_split_multi_fwd: bad dims");
  HFAGP_CHECK_ARG(w_host && ntaps_host && cout_host && cin_host && styles_off_host && styles && hi_host && lo_host && dcoef_host,
                  "modulate_split_multi_fwd: null pointer");
  ModTable tb;
  int rows = 0;
  for (int l = 0; l < nlayers; ++l) {
    HFAGP_CHECK_ARG(w_host[l] && hi_host[l] && lo_host[l] && ntaps_host[l] > 0 && cout_host[l] > 0 && cin_host[l] > 0 &&
                        (cin_host[l] & 3) == 0, "modulate_split_multi_fwd: layer %d: bad entry (cin must be a multiple of 4)", l);
    tb.w[l] = w_host[l];
    tb.hi[l] = reinterpret_cast<__nv_bfloat16*>(hi_host[l]);
    tb.lo[l] = reinterpret_cast<__nv_bfloat16*>(lo_host[l]);
    tb.dcoef[l] = dcoef_host[l];
    tb.soff[l] = styles_off_host[l];
    tb.ntaps[l] = ntaps_host[l]; tb.cout[l] = cout_host[l]; tb.cin[l] = cin_host[l];
    tb.row_start[l] = rows;
    rows += cout_host[l];
  }
  tb.row_start[nlayers] = rows;
  tb.nlayers = nlayers;
  modulate_split_multi_kernel<<<dim3(rows, batch), 128, 0, (cudaStream_t)stream>>>(tb, styles);
  HFAGP_CHECK_LAUNCH("modulate_split_multi_kernel");
  return HFAGP_OK;
}

// A convolution weight in torch layout w[O][I][T] (T = kh*kw taps) times the equalised-lr scale -> every operand form
// the encoder's forward and backward read, in ONE pass: pk[t][O][I] fp32, its split-bf16 pair, and the transposed
// split-bf16 pair pkT[t][I][O] of the data-gradient convolution (each may be NULL).  Replaces, per layer and training
// step, mul + permute-copy + split + transpose-copy + split (five passes over the weight, five launches).
// block = a 32 (o) x 32 (i) tile with all its taps staged in shared memory ([T][32][33]); TC > 0: tap count known at
// compile time (the index arithmetic of the contiguous runs is then shifts and constant divisions); a thread writes
// two neighbouring elements (4-byte bf16x2 / 8-byte fp32 stores).
// tap stride of the staged tile: 4 banks of skew per tap, so that the run-order walk (consecutive threads = consecutive
// taps of one (o, i)) does not put the 9 taps of an element on one bank
constexpr int PW_TS = 32 * 33 + 4;
template <int TC>
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(int O, int I, int T_rt, const float* __restrict__ w, float scale,
                                                              float* __restrict__ pk, __nv_bfloat16* __restrict__ hi,
                                                              __nv_bfloat16* __restrict__ lo, __nv_bfloat16* __restrict__ hiT,
                                                              __nv_bfloat16* __restrict__ loT) {
  extern __shared__ float pw_tile[];
  const int T = TC > 0 ? TC : T_rt;
  const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  const int ni = min(32, I - i0), no = min(32, O - o0);
  const int run = ni * T;                              // contiguous floats of one output channel inside the tile
  if (TC > 0 && ni == 32) {
#pragma unroll 6                                       // (a block is latency bound: keep several loads in flight per thread)
    for (int idx = threadIdx.x; idx < no * (32 * TC); idx += 256) {
      const int ol = idx / (32 * TC), rem = idx - ol * (32 * TC);
      const int il = rem / TC, t = rem - il * TC;
      pw_tile[t * PW_TS + ol * 33 + il] = __ldg(w + ((size_t)(o0 + ol) * I + i0) * TC + rem) * scale;
    }
  } else {
    for (int idx = threadIdx.x; idx < no * run; idx += 256) {
      const int ol = idx / run, rem = idx - ol * run;
      const int il = rem / T, t = rem - il * T;
      pw_tile[t * PW_TS + ol * 33 + il] = __ldg(w + ((size_t)(o0 + ol) * I + i0) * T + rem) * scale;
    }
  }
  __syncthreads();
  const bool pairs = ((I | O) & 1) == 0;               // even extents: 2 elements per thread, aligned vector stores
  if (pairs) {
#pragma unroll 2
    for (int idx = threadIdx.x; idx < T * 512; idx += 256) {
      const int t = idx >> 9, a = (idx >> 4) & 31, b = (idx & 15) * 2;
      if (a < no && b < ni) {                          // (o, i..i+1)
        const float v0 = pw_tile[t * PW_TS + a * 33 + b], v1 = pw_tile[t * PW_TS + a * 33 + b + 1];
        const size_t at = ((size_t)t * O + o0 + a) * I + i0 + b;
        if (pk) *reinterpret_cast<float2*>(pk + at) = make_float2(v0, v1);
        if (hi) {
          __nv_bfloat16 h0, l0, h1, l1;
          split2(v0, h0, l0);
          split2(v1, h1, l1);
          *reinterpret_cast<__nv_bfloat162*>(hi + at) = __nv_bfloat162(h0, h1);
          *reinterpret_cast<__nv_bfloat162*>(lo + at) = __nv_bfloat162(l0, l1);
        }
      }
      if (hiT && a < ni && b < no) {                   // (i, o..o+1)
        const float v0 = pw_tile[t * PW_TS + b * 33 + a], v1 = pw_tile[t * PW_TS + (b + 1) * 33 + a];
        const size_t at = ((size_t)t * I + i0 + a) * O + o0 + b;
        __nv_bfloat16 h0, l0, h1, l1;
        split2(v0, h0, l0);
        split2(v1, h1, l1);
        *reinterpret_cast<__nv_bfloat162*>(hiT + at) = __nv_bfloat162(h0, h1);
        *reinterpret_cast<__nv_bfloat162*>(loT + at) = __nv_bfloat162(l0, l1);
      }
    }
    return;
  }
  for (int idx = threadIdx.x; idx < T * 1024; idx += 256) {
    const int t = idx >> 10, a = (idx >> 5) & 31, b = idx & 31;
    if (a < no && b < ni) {                            // (o, i) = (a, b): consecutive threads along i
      const float v = pw_tile[t * PW_TS + a * 33 + b];
      const size_t at = ((size_t)t * O + o0 + a) * I + i0 + b;
      if (pk) pk[at] = v;
      if (hi) split2(v, hi[at], lo[at]);
    }
    if (hiT && a < ni && b < no) {                     // (i, o) = (a, b): consecutive threads along o
      const float v = pw_tile[t * PW_TS + b * 33 + a];
      const size_t at = ((size_t)t * I + i0 + a) * O + o0 + b;
      split2(v, hiT[at], loT[at]);
    }
  }
}

// The inverse walk for the weight gradient: dw[t][O][Ip] (the layout the wgrad kernels accumulate in; Ip >= I padded input
// channels) -> grad[O][I][T] += dw, torch's layout, in place.  Replaces a strided AccumulateGrad add per layer.
template <int TC>
__global__ void __launch_bounds__(256) unpack_conv_wgrad_kernel(int O, int I, int Ip, int T_rt, const float* __restrict__ dw,
                                                               float* __restrict__ grad) {
  extern __shared__ float pw_tile[];
  const int T = TC > 0 ? TC : T_rt;
  const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  const int ni = min(32, I - i0), no = min(32, O - o0);
  if ((Ip & 3) == 0 && ni == 32) {
#pragma unroll 3
    for (int idx = threadIdx.x; idx < T * 256; idx += 256) {         // 16-byte loads along i
      const int t = idx >> 8, a = (idx >> 3) & 31, b = (idx & 7) * 4;
      if (a < no) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dw + ((size_t)t * O + o0 + a) * Ip + i0 + b));
        float* d = pw_tile + t * PW_TS + a * 33 + b;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    }
  } else {
    for (int idx = threadIdx.x; idx < T * 1024; idx += 256) {
      const int t = idx >> 10, a = (idx >> 5) & 31, b = idx & 31;
      if (a < no && b < ni) pw_tile[t * PW_TS + a * 33 + b] = __ldg(dw + ((size_t)t * O + o0 + a) * Ip + i0 + b);
    }
  }
  __syncthreads();
  if (TC > 0 && ni == 32) {
    const int n = no * (32 * TC);
    for (int base = threadIdx.x; base < n; base += 256 * 6) {       // six independent read-modify-writes per thread and trip
      float g[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int idx = base + u * 256;
        if (idx < n) {
          const int ol = idx / (32 * TC), rem = idx - ol * (32 * TC);
          g[u] = grad[((size_t)(o0 + ol) * I + i0) * TC + rem];
        }
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int idx = base + u * 256;
        if (idx < n) {
          const int ol = idx / (32 * TC), rem = idx - ol * (32 * TC);
          const int il = rem / TC, t = rem - il * TC;
          grad[((size_t)(o0 + ol) * I + i0) * TC + rem] = g[u] + pw_tile[t * PW_TS + ol * 33 + il];
        }
      }
    }
    return;
  }
  const int run = ni * T;
  for (int idx = threadIdx.x; idx < no * run; idx += 256) {
    const int ol = idx / run, rem = idx - ol * run;
    const int il = rem / T, t = rem - il * T;
    grad[((size_t)(o0 + ol) * I + i0) * T + rem] += pw_tile[t * PW_TS + ol * 33 + il];
  }
}

template <int TC>
static int pack_launch(int cout, int cin, int taps, const float* w, float scale, float* pk, uint16_t* pk_hi, uint16_t* pk_lo,
                       uint16_t* pkt_hi, uint16_t* pkt_lo, cudaStream_t st) {
  const size_t smem = (size_t)taps * PW_TS * sizeof(float);
  static std::atomic<uint64_t> attr_done{0};
  HFAGP_CUDA(per_device_once(attr_done, [] { return cudaFuncSetAttribute(pack_conv_weight_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49 * PW_TS * 4); }));
  pack_conv_weight_kernel<TC><<<dim3(cdiv(cin, 32), cdiv(cout, 32)), 256, smem, st>>>(
      cout, cin, taps, w, scale, pk, reinterpret_cast<__nv_bfloat16*>(pk_hi), reinterpret_cast<__nv_bfloat16*>(pk_lo),
      reinterpret_cast<__nv_bfloat16*>(pkt_hi), reinterpret_cast<__nv_bfloat16*>(pkt_lo));
  HFAGP_CHECK_LAUNCH("pack_conv_weight_kernel");
  return HFAGP_OK;
}
template <int TC>
static int unpack_launch(int cout, int cin, int cin_padded, int taps, const float* dw, float* grad, cudaStream_t st) {
  const size_t smem = (size_t)taps * PW_TS * sizeof(float);
  static std::atomic<uint64_t> attr_done{0};
  HFAGP_CUDA(per_device_once(attr_done, [] { return cudaFuncSetAttribute(unpack_conv_wgrad_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49 * PW_TS * 4); }));
  unpack_conv_wgrad_kernel<TC><<<dim3(cdiv(cin, 32), cdiv(cout, 32)), 256, smem, st>>>(cout, cin, cin_padded, taps, dw, grad);
  HFAGP_CHECK_LAUNCH("unpack_conv_wgrad_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_unpack_conv_wgrad(int cout, int cin, int cin_padded, int taps, const float* dw, float* grad, void* stream) {
  HFAGP_CHECK_ARG(dw && grad && cout > 0 && cin > 0 && cin_padded >= cin && taps > 0 && taps <= 49, "unpack_conv_wgrad: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (taps == 9) return unpack_launch<9>(cout, cin, cin_padded, taps, dw, grad, st);
  if (taps == 1) return unpack_launch<1>(cout, cin, cin_padded, taps, dw, grad, st);
  if (taps == 16) return unpack_launch<16>(cout, cin, cin_padded, taps, dw, grad, st);
  return unpack_launch<0>(cout, cin, cin_padded, taps, dw, grad, st);
}

extern "C" int hfagp_pack_conv_weight(int cout, int cin, int taps, const float* w, float scale, float* pk, uint16_t* pk_hi,
                                      uint16_t* pk_lo, uint16_t* pkt_hi, uint16_t* pkt_lo, void* stream) {
  HFAGP_CHECK_ARG(w && cout > 0 && cin > 0 && taps > 0 && taps <= 49, "pack_conv_weight: bad args");
  HFAGP_CHECK_ARG((pk_hi != nullptr) == (pk_lo != nullptr) && (pkt_hi != nullptr) == (pkt_lo != nullptr),
                  "pack_conv_weight: give both halves of a split pair");
  cudaStream_t st = (cudaStream_t)stream;
  if (taps == 9) return pack_launch<9>(cout, cin, taps, w, scale, pk, pk_hi, pk_lo, pkt_hi, pkt_lo, st);
  if (taps == 1) return pack_launch<1>(cout, cin, taps, w, scale, pk, pk_hi, pk_lo, pkt_hi, pkt_lo, st);
  if (taps == 16) return pack_launch<16>(cout, cin, taps, w, scale, pk, pk_hi, pk_lo, pkt_hi, pkt_lo, st);
  return pack_launch<0>(cout, cin, taps, w, scale, pk, pk_hi, pk_lo, pkt_hi, pkt_lo, st);
}

extern "C" int hfagp_split_bf16(long long count, const float* x, uint16_t* hi, uint16_t* lo, void* stream) {
  HFAGP_CHECK_ARG(x && hi && lo && count > 0, "split_bf16: bad args");
  if ((count & 3) == 0 && (((uintptr_t)x | (uintptr_t)hi | (uintptr_t)lo) & 15) == 0) {
    split4_kernel<<<cdiv(count >> 2, 256), 256, 0, (cudaStream_t)stream>>>((size_t)(count >> 2), x,
                                                                          reinterpret_cast<__nv_bfloat16*>(hi),
                                                                          reinterpret_cast<__nv_bfloat16*>(lo));
  } else {
    split_kernel<<<cdiv(count, 256), 256, 0, (cudaStream_t)stream>>>((size_t)count, x, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                     reinterpret_cast<__nv_bfloat16*>(lo));
  }
  HFAGP_CHECK_LAUNCH("split_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_modulate_split_fwd(int batch, int ntaps, int cout, int cin, const float* w, const float* styles,
                                        uint16_t* wmod_hi, uint16_t* wmod_lo, float* dcoef, void* stream) {
  HFAGP_CHECK_ARG(w && styles && wmod_hi && wmod_lo, "modulate_split_fwd: null pointer");
  HFAGP_CHECK_ARG(batch > 0 && batch <= 65535 && ntaps > 0 && cout > 0 && cin > 0 && (cin & 3) == 0,
                  "modulate_split_fwd: bad dims (cin must be a multiple of 4)");
  int threads = cin >= 512 ? 128 : (cin >= 256 ? 64 : 32);
  modulate_split_kernel<<<dim3(cout, batch), threads, 0, (cudaStream_t)stream>>>(
      ntaps, cout, cin, w, styles, reinterpret_cast<__nv_bfloat16*>(wmod_hi), reinterpret_cast<__nv_bfloat16*>(wmod_lo),
      dcoef);
  HFAGP_CHECK_LAUNCH("modulate_split_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_stem_conv1x1_fwd(int batch, int h, int w_, int cin, int cout, const float* x_nchw, const float* w,
                                      const float* bias, int act, float act_gain, uint16_t* y_hi, uint16_t* y_lo, void* stream) {
  HFAGP_CHECK_ARG(x_nchw && w && y_hi && y_lo && batch > 0 && batch <= 65535 && h > 0 && w_ > 0, "stem_conv1x1_fwd: bad args");
  HFAGP_CHECK_ARG(cin >= 1 && cin <= 4 && cout >= 4 && (cout & 3) == 0, "stem_conv1x1_fwd: cin <= 4 and cout % 4 == 0 required");
  HFAGP_CHECK_ARG((long long)h * w_ * (cout >> 2) < (1ll << 31), "stem_conv1x1_fwd: frame too large");
  const int hw = h * w_;
  stem_conv1x1_kernel<<<dim3(cdiv((long long)hw * (cout >> 2), 256), batch), 256, 0, (cudaStream_t)stream>>>(
      hw, cin, cout >> 2, x_nchw, w, bias, act_slope(act), act_gain, reinterpret_cast<__nv_bfloat16*>(y_hi),
      reinterpret_cast<__nv_bfloat16*>(y_lo));
  HFAGP_CHECK_LAUNCH("stem_conv1x1_kernel");
  return HFAGP_OK;
}
