// LPIPS(net='alex') pieces that are not convolutions (trainer_rgb.py:62,86-87 -> pip `lpips`).
// The five AlexNet convolutions run on conv_tc_kernel (bias + ReLU epilogue); this file holds what sits between them:
//   stem      ScalingLayer (x - shift) / scale, zero-pad 2, space-to-depth by 4 and the split-bf16 conversion in one
//             pass: the 11x11 stride-4 convolution of 3 channels becomes a 3x3 stride-1 convolution of 48 channels
//             (kernel index 11 zero-padded), i.e. a K = 432 tensor-core GEMM instead of 121 three-channel taps
//   maxpool   MaxPool2d(3, stride 2) forward (split-bf16 out) and backward (first-maximum routing, as ATen)
//   head      per pixel: unit-normalise both feature vectors over channels, squared difference, 1x1 `lin` weights,
//             spatial mean -> one scalar per sample; backward to the generated image's features
// Channels-last activations; one warp per pixel for the head (coalesced 16-byte loads, shuffle reductions).
#include <cuda_bf16.h>
#include "common.cuh"
#include "splitio.cuh"

namespace hfagp {

// out[n][Y][X][(py*4+px)*3 + c] = (x[n][c][4Y+py-2][4X+px-2] - shift[c]) / scale[c]   (0 outside the image)
__global__ void lpips_stem_fwd_kernel(int batch, int h, int w_, const float* __restrict__ x, float3 shift, float3 inv_scale,
                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int oh = (h + 4) / 4, ow = (w_ + 4) / 4;
  const size_t total = (size_t)batch * oh * ow * 12;          // 12 groups of 4 channels
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int q = idx % 12;
  size_t r = idx / 12;
  const int X = r % ow;
  r /= ow;
  const int Y = r % oh, n = r / oh;
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch48 = q * 4 + k;
    const int c = ch48 % 3, pp = ch48 / 3, py = pp >> 2, px = pp & 3;
    const int iy = 4 * Y + py - 2, ix = 4 * X + px - 2;
    float t = 0.f;
    if (iy >= 0 && iy < h && ix >= 0 && ix < w_) {
      const float sh = c == 0 ? shift.x : (c == 1 ? shift.y : shift.z);
      const float is = c == 0 ? inv_scale.x : (c == 1 ? inv_scale.y : inv_scale.z);
      t = (__ldg(x + (((size_t)n * 3 + c) * h + iy) * w_ + ix) - sh) * is;
    }
    v[k] = t;
  }
  st4_split(hi, lo, idx, v);
}

// dimg[n][c][iy][ix] = dx48[n][(iy+2)/4][(ix+2)/4][((iy+2)%4*4 + (ix+2)%4)*3 + c] / scale[c]
__global__ void lpips_stem_bwd_kernel(int batch, int h, int w_, const float* __restrict__ dx48, float3 inv_scale,
                                      float* __restrict__ dimg) {
  const int oh = (h + 4) / 4, ow = (w_ + 4) / 4;
  const size_t total = (size_t)batch * 3 * h * w_;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ix = idx % w_;
  size_t r = idx / w_;
  const int iy = r % h;
  r /= h;
  const int c = r % 3, n = r / 3;
  const int Y = (iy + 2) >> 2, py = (iy + 2) & 3, X = (ix + 2) >> 2, px = (ix + 2) & 3;
  const float is = c == 0 ? inv_scale.x : (c == 1 ? inv_scale.y : inv_scale.z);
  dimg[idx] = __ldg(dx48 + (((size_t)n * oh + Y) * ow + X) * 48 + (py * 4 + px) * 3 + c) * is;
}

// MaxPool2d(kernel 3, stride 2, no padding), channels-last, 4 channels per thread
__global__ void maxpool3s2_fwd_kernel(int batch, int h, int w_, int c, const float* __restrict__ x,
                                      const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo,
                                      float* __restrict__ y, __nv_bfloat16* __restrict__ y_hi,
                                      __nv_bfloat16* __restrict__ y_lo) {
  const int oh = (h - 3) / 2 + 1, ow = (w_ - 3) / 2 + 1, c4 = c >> 2;
  const size_t total = (size_t)batch * oh * ow * c4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cq = idx % c4;
  size_t r = idx / c4;
  const int ox = r % ow;
  r /= ow;
  const int oy = r % oh, n = r / oh;
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float4 v = ld4_any(x, x_hi, x_lo, (((size_t)n * h + 2 * oy + ky) * w_ + 2 * ox + kx) * c4 + cq);
      m[0] = fmaxf(m[0], v.x); m[1] = fmaxf(m[1], v.y); m[2] = fmaxf(m[2], v.z); m[3] = fmaxf(m[3], v.w);
    }
  st4_any(y, y_hi, y_lo, idx, m);
}

// dx[n][iy][ix][c] = sum over the <= 4 windows containing (iy,ix) whose FIRST maximum (row-major scan, as ATen) is
// this element of dy[window].  Gather form: no atomics, dx is written.
__global__ void maxpool3s2_bwd_kernel(int batch, int h, int w_, int c, const float* __restrict__ x,
                                      const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo,
                                      const float* __restrict__ dy, float* __restrict__ dx) {
  const int oh = (h - 3) / 2 + 1, ow = (w_ - 3) / 2 + 1, c4 = c >> 2;
  const size_t total = (size_t)batch * h * w_ * c4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cq = idx % c4;
  size_t r = idx / c4;
  const int ix = r % w_;
  r /= w_;
  const int iy = r % h, n = r / h;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  const int oy_lo = max(0, (iy - 1) >> 1), oy_hi = min(oh - 1, iy >> 1);
  const int ox_lo = max(0, (ix - 1) >> 1), ox_hi = min(ow - 1, ix >> 1);
  for (int oy = oy_lo; oy <= oy_hi; ++oy)
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      // scan the window; remember for each of the 4 channels the linear position of its first maximum
      float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      int am[4] = {0, 0, 0, 0};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 v = ld4_any(x, x_hi, x_lo, (((size_t)n * h + 2 * oy + ky) * w_ + 2 * ox + kx) * c4 + cq);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (vv[k] > m[k]) { m[k] = vv[k]; am[k] = ky * 3 + kx; }
        }
      const int me = (iy - 2 * oy) * 3 + (ix - 2 * ox);
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(dy) + (((size_t)n * oh + oy) * ow + ox) * c4 + cq);
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (am[k] == me) g[k] += dv[k];
    }
  reinterpret_cast<float4*>(dx)[idx] = make_float4(g[0], g[1], g[2], g[3]);
}

// One warp per (sample b, pixel): f0 = feat[b][p], f1 = feat[B+b][p].
//   forward   out[b] += (1/hw) sum_c w_c (f0_c/n0 - f1_c/n1)^2,  n = sqrt(sum f^2) + 1e-10
//   backward  df1_j = gout[b]/hw * [ g_j/n1 - (sum_c g_c f1_c) f1_j / (r1 n1^2) ],  g_c = -2 w_c (f0_c/n0 - f1_c/n1)
constexpr int LPIPS_MAX_Q = 3;    // float4 groups per lane: channels <= 384

template <bool BWD>
__global__ void __launch_bounds__(256) lpips_head_kernel(int batch, int hw, int c, const float* __restrict__ f,
                                                        const __nv_bfloat16* __restrict__ f_hi,
                                                        const __nv_bfloat16* __restrict__ f_lo,
                                                        const float* __restrict__ lin, const float* __restrict__ gout,
                                                        float* __restrict__ out, float* __restrict__ df1) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (size_t)batch * hw) return;
  const int b = warp / hw, pix = warp - (size_t)b * hw;
  const int c4 = c >> 2;
  const size_t q0 = ((size_t)b * hw + pix) * c4, q1 = ((size_t)(batch + b) * hw + pix) * c4;
  float4 a[LPIPS_MAX_Q], v[LPIPS_MAX_Q], wl[LPIPS_MAX_Q];
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < LPIPS_MAX_Q; ++i) {
    const int q = lane + 32 * i;
    a[i] = v[i] = wl[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < c4) {
      a[i] = ld4_any(f, f_hi, f_lo, q0 + q);
      v[i] = ld4_any(f, f_hi, f_lo, q1 + q);
      wl[i] = __ldg(reinterpret_cast<const float4*>(lin) + q);
    }
    s0 += a[i].x * a[i].x + a[i].y * a[i].y + a[i].z * a[i].z + a[i].w * a[i].w;
    s1 += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  const float r1 = sqrtf(s1);
  const float i0 = 1.f / (sqrtf(s0) + 1e-10f), i1 = 1.f / (r1 + 1e-10f);
  float val = 0.f, gf = 0.f;        // gf = sum_c g_c f1_c
  float4 g[LPIPS_MAX_Q];
#pragma unroll
  for (int i = 0; i < LPIPS_MAX_Q; ++i) {
    const float d0 = a[i].x * i0 - v[i].x * i1, d1 = a[i].y * i0 - v[i].y * i1;
    const float d2 = a[i].z * i0 - v[i].z * i1, d3 = a[i].w * i0 - v[i].w * i1;
    val += wl[i].x * d0 * d0 + wl[i].y * d1 * d1 + wl[i].z * d2 * d2 + wl[i].w * d3 * d3;
    if (BWD) {
      g[i] = make_float4(-2.f * wl[i].x * d0, -2.f * wl[i].y * d1, -2.f * wl[i].z * d2, -2.f * wl[i].w * d3);
      gf += g[i].x * v[i].x + g[i].y * v[i].y + g[i].z * v[i].z + g[i].w * v[i].w;
    }
  }
  if (!BWD) {
    val = warp_sum(val);
    if (lane == 0) atomicAdd(out + b, val / (float)hw);
    return;
  }
  gf = warp_sum(gf);
  const float go = __ldg(gout + b) / (float)hw;
  const float k2 = r1 > 0.f ? gf * i1 * i1 / r1 : 0.f;
#pragma unroll
  for (int i = 0; i < LPIPS_MAX_Q; ++i) {
    const int q = lane + 32 * i;
    if (q < c4)
      reinterpret_cast<float4*>(df1)[((size_t)b * hw + pix) * c4 + q] =
          make_float4(go * (g[i].x * i1 - k2 * v[i].x), go * (g[i].y * i1 - k2 * v[i].y),
                      go * (g[i].z * i1 - k2 * v[i].z), go * (g[i].w * i1 - k2 * v[i].w));
  }
}

}  // namespace hfagp

using namespace hfagp;

extern "C" int hfagp_lpips_stem_fwd(int batch, int h, int w_, const float* x, const float* shift3_host,
                                    const float* scale3_host, uint16_t* y_hi, uint16_t* y_lo, void* stream) {
  HFAGP_CHECK_ARG(x && shift3_host && scale3_host && y_hi && y_lo && batch > 0, "lpips_stem_fwd: null pointer");
  HFAGP_CHECK_ARG(h >= 7 && w_ >= 7 && (h & 3) == 0 && (w_ & 3) == 0, "lpips_stem_fwd: image sides must be multiples of 4");
  const float3 sh = make_float3(shift3_host[0], shift3_host[1], shift3_host[2]);
  const float3 is = make_float3(1.f / scale3_host[0], 1.f / scale3_host[1], 1.f / scale3_host[2]);
  const size_t total = (size_t)batch * ((h + 4) / 4) * ((w_ + 4) / 4) * 12;
  lpips_stem_fwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      batch, h, w_, x, sh, is, reinterpret_cast<__nv_bfloat16*>(y_hi), reinterpret_cast<__nv_bfloat16*>(y_lo));
  HFAGP_CHECK_LAUNCH("lpips_stem_fwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_lpips_stem_bwd(int batch, int h, int w_, const float* dx48, const float* scale3_host, float* dimg,
                                    void* stream) {
  HFAGP_CHECK_ARG(dx48 && scale3_host && dimg && batch > 0 && (h & 3) == 0 && (w_ & 3) == 0, "lpips_stem_bwd: bad args");
  const float3 is = make_float3(1.f / scale3_host[0], 1.f / scale3_host[1], 1.f / scale3_host[2]);
  lpips_stem_bwd_kernel<<<cdiv((long long)batch * 3 * h * w_, 256), 256, 0, (cudaStream_t)stream>>>(batch, h, w_, dx48, is,
                                                                                                 dimg);
  HFAGP_CHECK_LAUNCH("lpips_stem_bwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_maxpool3s2_fwd(int batch, int h, int w_, int c, const float* x, const uint16_t* x_hi,
                                    const uint16_t* x_lo, float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream) {
  HFAGP_CHECK_ARG((x != nullptr) != (x_hi != nullptr && x_lo != nullptr), "maxpool3s2_fwd: give x or (x_hi, x_lo)");
  HFAGP_CHECK_ARG((y != nullptr) != (y_hi != nullptr && y_lo != nullptr), "maxpool3s2_fwd: give y or (y_hi, y_lo)");
  HFAGP_CHECK_ARG(batch > 0 && h >= 3 && w_ >= 3 && c > 0 && (c & 3) == 0, "maxpool3s2_fwd: bad dims");
  const size_t total = (size_t)batch * ((h - 3) / 2 + 1) * ((w_ - 3) / 2 + 1) * (c >> 2);
  maxpool3s2_fwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      batch, h, w_, c, x, reinterpret_cast<const __nv_bfloat16*>(x_hi), reinterpret_cast<const __nv_bfloat16*>(x_lo), y,
      reinterpret_cast<__nv_bfloat16*>(y_hi), reinterpret_cast<__nv_bfloat16*>(y_lo));
  HFAGP_CHECK_LAUNCH("maxpool3s2_fwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_maxpool3s2_bwd(int batch, int h, int w_, int c, const float* x, const uint16_t* x_hi,
                                    const uint16_t* x_lo, const float* dy, float* dx, void* stream) {
  HFAGP_CHECK_ARG((x != nullptr) != (x_hi != nullptr && x_lo != nullptr), "maxpool3s2_bwd: give x or (x_hi, x_lo)");
  HFAGP_CHECK_ARG(dy && dx && batch > 0 && h >= 3 && w_ >= 3 && c > 0 && (c & 3) == 0, "maxpool3s2_bwd: bad dims");
  const size_t total = (size_t)batch * h * w_ * (c >> 2);
  maxpool3s2_bwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      batch, h, w_, c, x, reinterpret_cast<const __nv_bfloat16*>(x_hi), reinterpret_cast<const __nv_bfloat16*>(x_lo), dy, dx);
  HFAGP_CHECK_LAUNCH("maxpool3s2_bwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_lpips_head_fwd(int batch, int hw, int c, const float* f, const uint16_t* f_hi, const uint16_t* f_lo,
                                    const float* lin, float* out, void* stream) {
  HFAGP_CHECK_ARG((f != nullptr) != (f_hi != nullptr && f_lo != nullptr), "lpips_head_fwd: give f or (f_hi, f_lo)");
  HFAGP_CHECK_ARG(lin && out && batch > 0 && hw > 0 && c > 0 && (c & 3) == 0 && c <= 128 * LPIPS_MAX_Q,
                  "lpips_head_fwd: c must be a multiple of 4, <= %d", 128 * LPIPS_MAX_Q);
  lpips_head_kernel<false><<<cdiv((long long)batch * hw * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      batch, hw, c, f, reinterpret_cast<const __nv_bfloat16*>(f_hi), reinterpret_cast<const __nv_bfloat16*>(f_lo), lin,
      nullptr, out, nullptr);
  HFAGP_CHECK_LAUNCH("lpips_head_kernel<fwd>");
  return HFAGP_OK;
}

extern "C" int hfagp_lpips_head_bwd(int batch, int hw, int c, const float* f, const uint16_t* f_hi, const uint16_t* f_lo,
                                    const float* lin, const float* gout, float* df1, void* stream) {
  HFAGP_CHECK_ARG((f != nullptr) != (f_hi != nullptr && f_lo != nullptr), "lpips_head_bwd: give f or (f_hi, f_lo)");
  HFAGP_CHECK_ARG(lin && gout && df1 && batch > 0 && hw > 0 && c > 0 && (c & 3) == 0 && c <= 128 * LPIPS_MAX_Q,
                  "lpips_head_bwd: c must be a multiple of 4, <= %d", 128 * LPIPS_MAX_Q);
  lpips_head_kernel<true><<<cdiv((long long)batch * hw * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      batch, hw, c, f, reinterpret_cast<const __nv_bfloat16*>(f_hi), reinterpret_cast<const __nv_bfloat16*>(f_lo), lin,
      gout, nullptr, df1);
  HFAGP_CHECK_LAUNCH("lpips_head_kernel<bwd>");
  return HFAGP_OK;
}
