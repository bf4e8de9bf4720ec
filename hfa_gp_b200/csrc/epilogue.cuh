// The fused convolution epilogue shared by the SIMT and the tcgen05 kernels (see hfagp_conv2d_fwd in
// include/hfagp.h):  demod -> noise -> bias -> leaky-ReLU -> gain -> clamp -> residual merge -> skip-image add.
#pragma once
#include "common.cuh"

namespace hfagp {

struct ConvParams {
  HfagpConvDesc d;
  const float* x;
  const float* w;
  const float* dcoef;
  const float* noise;
  const float* bias;
  const float* residual;
  const float* up_img;
  float* y;
};

// upsample2d(img)[oy][ox][co] for a channels-last low-res image [uh][uw][cstride]:
// zero-insert x2, pad [2,1,2,1], [1,3,3,1]^2/64 * 4  ==  separable {0.25, 0.75} polyphase.
__device__ __forceinline__ float upsample_tap(const float* __restrict__ img, int uh, int uw, int cstride, int oy,
                                              int ox, int co) {
  const int my = oy >> 1, mx = ox >> 1;
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  if (oy & 1) { y0 = my; y1 = my + 1; wy0 = 0.75f; wy1 = 0.25f; } else { y0 = my - 1; y1 = my; wy0 = 0.25f; wy1 = 0.75f; }
  if (ox & 1) { x0 = mx; x1 = mx + 1; wx0 = 0.75f; wx1 = 0.25f; } else { x0 = mx - 1; x1 = mx; wx0 = 0.25f; wx1 = 0.75f; }
  const bool vy0 = y0 >= 0 && y0 < uh, vy1 = y1 >= 0 && y1 < uh;
  const bool vx0 = x0 >= 0 && x0 < uw, vx1 = x1 >= 0 && x1 < uw;
  float v = 0.f;
  if (vy0 && vx0) v += wy0 * wx0 * __ldg(img + ((size_t)y0 * uw + x0) * cstride + co);
  if (vy0 && vx1) v += wy0 * wx1 * __ldg(img + ((size_t)y0 * uw + x1) * cstride + co);
  if (vy1 && vx0) v += wy1 * wx0 * __ldg(img + ((size_t)y1 * uw + x0) * cstride + co);
  if (vy1 && vx1) v += wy1 * wx1 * __ldg(img + ((size_t)y1 * uw + x1) * cstride + co);
  return v;
}

// The same interpolation prepared once per pixel: four tap offsets (in pixels of the low-res image) and weights
// (0 for taps outside the image), so the per-channel work is 4 vector loads + FMAs.
struct UpTaps {
  int off[4];
  float w[4];
};
__device__ __forceinline__ UpTaps upsample_taps(int uh, int uw, int oy, int ox) {
  const int my = oy >> 1, mx = ox >> 1;
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  if (oy & 1) { y0 = my; y1 = my + 1; wy0 = 0.75f; wy1 = 0.25f; } else { y0 = my - 1; y1 = my; wy0 = 0.25f; wy1 = 0.75f; }
  if (ox & 1) { x0 = mx; x1 = mx + 1; wx0 = 0.75f; wx1 = 0.25f; } else { x0 = mx - 1; x1 = mx; wx0 = 0.25f; wx1 = 0.75f; }
  const bool vy0 = y0 >= 0 && y0 < uh, vy1 = y1 >= 0 && y1 < uh;
  const bool vx0 = x0 >= 0 && x0 < uw, vx1 = x1 >= 0 && x1 < uw;
  UpTaps t;
  t.off[0] = (vy0 && vx0) ? y0 * uw + x0 : 0; t.w[0] = (vy0 && vx0) ? wy0 * wx0 : 0.f;
  t.off[1] = (vy0 && vx1) ? y0 * uw + x1 : 0; t.w[1] = (vy0 && vx1) ? wy0 * wx1 : 0.f;
  t.off[2] = (vy1 && vx0) ? y1 * uw + x0 : 0; t.w[2] = (vy1 && vx0) ? wy1 * wx0 : 0.f;
  t.off[3] = (vy1 && vx1) ? y1 * uw + x1 : 0; t.w[3] = (vy1 && vx1) ? wy1 * wx1 : 0.f;
  return t;
}

// per-output-pixel state
struct EpiCtx {
  int oy, ox;
  float nz;              // noise[oy][ox] * noise_gain
  size_t out_base;       // element offset of (n, oy, ox, 0) in the output tensor
  const float* dco;      // dcoef row of sample n (or null)
  const float* res;      // residual + out_base (or null)
  const float* up;       // low-res skip image of sample n (or null)
};

// (off_y, off_x): output offset of the sub-problem (an output parity class of a merged launch)
__device__ __forceinline__ void epi_setup_at(EpiCtx& e, const ConvParams& p, int n, int my, int mx, int off_y, int off_x) {
  const HfagpConvDesc& d = p.d;
  e.oy = my * d.out_stride + off_y;
  e.ox = mx * d.out_stride + off_x;
  e.nz = p.noise ? __ldg(p.noise + (size_t)e.oy * d.out_w + e.ox) * d.noise_gain : 0.f;
  e.out_base = (((size_t)n * d.out_h + e.oy) * d.out_w + e.ox) * d.cout;
  e.dco = p.dcoef ? p.dcoef + (size_t)n * d.cout : nullptr;
  e.res = p.residual ? p.residual + e.out_base : nullptr;
  e.up = p.up_img ? p.up_img + (size_t)n * d.up_h * d.up_w * d.cout : nullptr;
}

__device__ __forceinline__ void epi_setup(EpiCtx& e, const ConvParams& p, int n, int my, int mx) {
  epi_setup_at(e, p, n, my, mx, p.d.out_off_y, p.d.out_off_x);
}

__device__ __forceinline__ float epi_apply(const EpiCtx& e, const ConvParams& p, float v, int co) {
  const HfagpConvDesc& d = p.d;
  if (e.dco) v *= __ldg(e.dco + co);
  v += e.nz;
  if (p.bias) v += __ldg(p.bias + co);
  if (d.act != HFAGP_ACT_LINEAR) v = fmaxf(v, act_slope(d.act) * v);
  v *= d.act_gain;
  if (d.clamp > 0.f) v = fminf(fmaxf(v, -d.clamp), d.clamp);
  if (e.res) v = (v + __ldg(e.res + co)) * d.residual_scale;
  if (e.up) v += upsample_tap(e.up, d.up_h, d.up_w, d.cout, e.oy, e.ox, co);
  return v;
}

}  // namespace hfagp
