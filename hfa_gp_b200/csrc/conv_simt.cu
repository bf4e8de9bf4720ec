// fp32 SIMT implicit-GEMM convolution (channels-last) with the fused StyleGAN2 / encoder epilogue,
// plus the small bandwidth-bound companions (FIR after the transposed conv, small-N ToRGB, blur).
// This is the exact-fp32 path; see conv_tc.cu for the tcgen05 tensor-core path.
#include <cuda_bf16.h>
#include <cstdlib>
#include "common.cuh"
#include "epilogue.cuh"
#include "splitio.cuh"

namespace hfagp {

constexpr int BK = 16;
constexpr int LDK = BK + 4;  // row stride (floats): 8 consecutive rows hit 8 distinct 16B bank groups

template <int TM, int TN>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvParams p) {
  constexpr int BM = 16 * TM, BN = 16 * TN;
  constexpr int A_LD = BM / 64;  // float4 loads per thread per k-chunk
  constexpr int B_LD = BN / 64;
  __shared__ __align__(16) float As[2][BM * LDK];
  __shared__ __align__(16) float Bs[2][BN * LDK];

  const HfagpConvDesc& d = p.d;
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int M = d.oh * d.ow;
  const int cin = d.cin, cout = d.cout;
  const bool vec_ok = (cin & 3) == 0;

  const float* xn = p.x + (size_t)n * d.in_h * d.in_w * cin;
  const float* wn = p.w + (size_t)n * d.w_batch_stride;

  // fixed per-thread load coordinates
  int a_iy0[A_LD], a_ix0[A_LD];
  bool a_valid[A_LD];
#pragma unroll
  for (int i = 0; i < A_LD; ++i) {
    int idx = tid + i * 256;
    int row = idx >> 2;
    int m = m0 + row;
    a_valid[i] = m < M;
    int my = a_valid[i] ? m / d.ow : 0;
    int mx = a_valid[i] ? m - my * d.ow : 0;
    a_iy0[i] = my * d.in_stride;
    a_ix0[i] = mx * d.in_stride;
  }
  const int kq = tid & 3;
  const int chunks = (cin + BK - 1) / BK;
  const int iters = d.ntaps * chunks;

  float4 a_reg[A_LD], b_reg[B_LD];

  auto load_tiles = [&](int it) {
    int t = it / chunks;
    int c = (it - t * chunks) * BK + kq * 4;
    int dy = d.dy[t], dx = d.dx[t];
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int iy = a_iy0[i] + dy, ix = a_ix0[i] + dx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_valid[i] && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w && c < cin) {
        const float* ptr = xn + ((size_t)iy * d.in_w + ix) * cin + c;
        if (vec_ok) {
          v = __ldg(reinterpret_cast<const float4*>(ptr));
        } else {
          v.x = __ldg(ptr);
          if (c + 1 < cin) v.y = __ldg(ptr + 1);
          if (c + 2 < cin) v.z = __ldg(ptr + 2);
          if (c + 3 < cin) v.w = __ldg(ptr + 3);
        }
      }
      a_reg[i] = v;
    }
    const float* wt = wn + (size_t)d.wtap[t] * cout * cin;
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int row = (tid + i * 256) >> 2;
      int co = n0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co < cout && c < cin) {
        const float* ptr = wt + (size_t)co * cin + c;
        if (vec_ok) {
          v = __ldg(reinterpret_cast<const float4*>(ptr));
        } else {
          v.x = __ldg(ptr);
          if (c + 1 < cin) v.y = __ldg(ptr + 1);
          if (c + 2 < cin) v.z = __ldg(ptr + 2);
          if (c + 3 < cin) v.w = __ldg(ptr + 3);
        }
      }
      b_reg[i] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int row = (tid + i * 256) >> 2;
      *reinterpret_cast<float4*>(&As[buf][row * LDK + kq * 4]) = a_reg[i];
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int row = (tid + i * 256) >> 2;
      *reinterpret_cast<float4*>(&Bs[buf][row * LDK + kq * 4]) = b_reg[i];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int tx = tid & 15, ty = tid >> 4;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();

  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) load_tiles(it + 1);
    const float* as = As[buf];
    const float* bs = Bs[buf];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 a4[TM], b4[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a4[i] = *reinterpret_cast<const float4*>(&as[(ty + 16 * i) * LDK + q * 4]);
#pragma unroll
      for (int j = 0; j < TN; ++j) b4[j] = *reinterpret_cast<const float4*>(&bs[(tx + 16 * j) * LDK + q * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] = fmaf(a4[i].x, b4[j].x, acc[i][j]);
          acc[i][j] = fmaf(a4[i].y, b4[j].y, acc[i][j]);
          acc[i][j] = fmaf(a4[i].z, b4[j].z, acc[i][j]);
          acc[i][j] = fmaf(a4[i].w, b4[j].w, acc[i][j]);
        }
    }
    if (it + 1 < iters) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty + 16 * i;
    if (m >= M) continue;
    int my = m / d.ow, mx = m - my * d.ow;
    EpiCtx ec;
    epi_setup(ec, p, n, my, mx);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int co = n0 + tx + 16 * j;
      if (co >= cout) continue;
      p.y[ec.out_base + co] = epi_apply(ec, p, acc[i][j], co);
    }
  }
}

// ---------------------------------------------------------------- epilogue of a split-K convolution
// acc[n][oy][ox][c] holds the raw convolution sums (hfagp_conv2d_tc_acc_fwd); apply the fused epilogue elementwise.
__global__ void conv_epilogue_kernel(const ConvParams p, const float* __restrict__ acc, __nv_bfloat16* __restrict__ y_hi,
                                     __nv_bfloat16* __restrict__ y_lo) {
  const HfagpConvDesc& d = p.d;
  const int c4 = d.cout >> 2;
  const size_t total = (size_t)d.batch * d.out_h * d.out_w * c4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cq = idx % c4;
  size_t pix = idx / c4;
  const int ox = pix % d.out_w;
  pix /= d.out_w;
  const int oy = pix % d.out_h, n = pix / d.out_h;
  EpiCtx ec;
  epi_setup_at(ec, p, n, oy, ox, 0, 0);
  const float4 a = __ldg(reinterpret_cast<const float4*>(acc) + idx);
  float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = epi_apply(ec, p, v[k], cq * 4 + k);
  st4_any(p.y, y_hi, y_lo, idx, v);
}

// ---------------------------------------------------------------- FIR after transposed conv
// y[oy][ox] = sum_{ky,kx} g[ky] g[kx] t[oy+ky-1][ox+kx-1]  (g = [1,3,3,1]/4: the 4x4 filter with its gain of 4),
// evaluated separably in registers: a thread owns 4 channels of TWO adjacent output columns and UPFIR_R (4) output rows;
// it walks the R+3 input rows once (5 float4 loads per row), forms the two horizontal sums and scatters them into
// the vertical accumulators.  3.4 loads per output instead of 16; the epilogue is branch-free.
template <int UPFIR_R>
__global__ void __launch_bounds__(256) upfir_act_kernel(int batch, int h2, int w2, int c, const float* __restrict__ t,
                                                       const float* __restrict__ dcoef,
                                                       const float* __restrict__ noise, float noise_gain,
                                                       const float* __restrict__ bias, int act, float act_gain,
                                                       float clamp, float* __restrict__ y,
                                                       __nv_bfloat16* __restrict__ y_hi,
                                                       __nv_bfloat16* __restrict__ y_lo) {
  const int c4 = c >> 2;
  const int wp = (w2 + 1) >> 1;                       // column pairs
  const int hb = (h2 + UPFIR_R - 1) / UPFIR_R;        // row bands
  const size_t total = (size_t)batch * hb * wp * c4;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cq = idx % c4;
  size_t r = idx / c4;
  const int px = r % wp;
  r /= wp;
  const int band = r % hb, n = r / hb;
  const int ox0 = px * 2, oy0 = band * UPFIR_R;
  const int th = h2 + 1, tw = w2 + 1;
  const float4* tn = reinterpret_cast<const float4*>(t) + (size_t)n * th * tw * c4 + cq;
  const float g[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  float acc[UPFIR_R][2][4];
#pragma unroll
  for (int i = 0; i < UPFIR_R; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[i][0][k] = acc[i][1][k] = 0.f;

  // the five input columns ox0-1 .. ox0+3 are all inside the row except for the first and the last column pair:
  // interior threads walk a pointer, no per-load index arithmetic or bounds tests
  const bool interior_x = ox0 >= 1 && ox0 + 3 < tw;
  const long long rstride = (long long)tw * c4;
  const float4* rowp = tn + ((long long)(oy0 - 1) * tw + (ox0 - 1)) * c4;     // (row oy0-1, column ox0-1); may lie outside
#pragma unroll
  for (int rr = 0; rr < UPFIR_R + 3; ++rr) {
    const int iy = oy0 + rr - 1;
    if (iy < 0 || iy >= th) continue;
    const float4* rp = rowp + rr * rstride;
    float4 v[5];
    if (interior_x) {
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) v[kx] = __ldg(rp + kx * c4);
    } else {
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int ix = ox0 + kx - 1;
        v[kx] = (ix >= 0 && ix < tw) ? __ldg(rp + kx * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float ha[4], hb2[4];
    ha[0] = g[0] * v[0].x + g[1] * v[1].x + g[2] * v[2].x + g[3] * v[3].x;
    ha[1] = g[0] * v[0].y + g[1] * v[1].y + g[2] * v[2].y + g[3] * v[3].y;
    ha[2] = g[0] * v[0].z + g[1] * v[1].z + g[2] * v[2].z + g[3] * v[3].z;
    ha[3] = g[0] * v[0].w + g[1] * v[1].w + g[2] * v[2].w + g[3] * v[3].w;
    hb2[0] = g[0] * v[1].x + g[1] * v[2].x + g[2] * v[3].x + g[3] * v[4].x;
    hb2[1] = g[0] * v[1].y + g[1] * v[2].y + g[2] * v[3].y + g[3] * v[4].y;
    hb2[2] = g[0] * v[1].z + g[1] * v[2].z + g[2] * v[3].z + g[3] * v[4].z;
    hb2[3] = g[0] * v[1].w + g[1] * v[2].w + g[2] * v[3].w + g[3] * v[4].w;
    // input row rr (iy = oy0 + rr - 1) feeds output rows i = rr - ky, ky = 0..3
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int i = rr - ky;
      if (i < 0 || i >= UPFIR_R) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[i][0][k] = fmaf(g[ky], ha[k], acc[i][0][k]);
        acc[i][1][k] = fmaf(g[ky], hb2[k], acc[i][1][k]);
      }
    }
  }
  float sc[4], sh[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sc[k] = dcoef ? __ldg(dcoef + (size_t)n * c + cq * 4 + k) : 1.f;
    sh[k] = bias ? __ldg(bias + cq * 4 + k) : 0.f;
  }
  const float slope = act_slope(act);
  const float cl = clamp > 0.f ? clamp : __int_as_float(0x7f800000);
#pragma unroll
  for (int i = 0; i < UPFIR_R; ++i) {
    const int oy = oy0 + i;
    if (oy >= h2) break;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ox = ox0 + j;
      if (ox >= w2) continue;
      const float nz = noise ? __ldg(noise + (size_t)oy * w2 + ox) * noise_gain : 0.f;
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a = fmaf(acc[i][j][k], sc[k], sh[k] + nz);
        a = fmaxf(a, slope * a) * act_gain;
        o[k] = fminf(fmaxf(a, -cl), cl);
      }
      st4_any(y, y_hi, y_lo, (((size_t)n * h2 + oy) * w2 + ox) * c4 + cq, o);
    }
  }
}

// ---------------------------------------------------------------- small-N ToRGB (cout <= 4)
// 8 lanes per pixel (4 pixels per warp pass, TORGB_PASSES passes per warp); the per-sample weights sit in shared
// memory (every pixel group reads the same addresses: broadcast); 16-byte activation loads, 3-step shuffle reduce.
constexpr int TORGB_PASSES = 8;

__global__ void __launch_bounds__(256) torgb_small_kernel(int batch, int h, int w_, int cin, int cout,
                                                         const float* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ x_hi,
                                                         const __nv_bfloat16* __restrict__ x_lo,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float clamp, const float* __restrict__ up_img,
                                                         float* __restrict__ y) {
  extern __shared__ __align__(16) float wsm[];          // [cout][cin] of sample n
  const int n = blockIdx.y;
  const int c4 = cin >> 2;
  for (int i = threadIdx.x; i < cout * c4; i += blockDim.x)
    reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(w + (size_t)n * cout * cin) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, l8 = lane & 7, grp = lane >> 3;
  const int warp_in_block = threadIdx.x >> 5;
  const int hw = h * w_;
  const int pix0 = (blockIdx.x * 8 + warp_in_block) * (4 * TORGB_PASSES);
#pragma unroll 2
  for (int ps = 0; ps < TORGB_PASSES; ++ps) {
    const int pix = pix0 + ps * 4 + grp;
    const bool valid = pix < hw;
    const size_t gp = (size_t)n * hw + (valid ? pix : 0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = l8; q < c4; q += 8) {
      const float4 xv = x ? __ldg(reinterpret_cast<const float4*>(x) + gp * c4 + q) : bf16x4_sum(x_hi, x_lo, gp * c4 + q);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        if (o < cout) {
          const float4 wv = reinterpret_cast<const float4*>(wsm)[o * c4 + q];
          acc[o] = fmaf(xv.x, wv.x, acc[o]);
          acc[o] = fmaf(xv.y, wv.y, acc[o]);
          acc[o] = fmaf(xv.z, wv.z, acc[o]);
          acc[o] = fmaf(xv.w, wv.w, acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 4);
      acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 2);
      acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 1);
    }
    if (valid && l8 < cout) {
      float v = l8 == 0 ? acc[0] : l8 == 1 ? acc[1] : l8 == 2 ? acc[2] : acc[3];
      if (bias) v += __ldg(bias + l8);
      if (clamp > 0.f) v = fminf(fmaxf(v, -clamp), clamp);
      if (up_img) {
        const int oy = pix / w_, ox = pix - oy * w_;
        v += upsample_tap(up_img + (size_t)n * (h / 2) * (w_ / 2) * cout, h / 2, w_ / 2, cout, oy, ox, l8);
      }
      y[gp * cout + l8] = v;
    }
  }
}

// img[n][oy][ox][o] = clamp(acc + bias[o]) + upsample2d(prev)[..][o]: the tail of a ToRGB whose channel sums were
// accumulated by the producing convolution's epilogue (hfagp_conv2d_tc_rgb_fwd)
__global__ void torgb_finalize_kernel(int batch, int h, int w_, int k, const float* __restrict__ acc,
                                      const float* __restrict__ bias, float clamp, const float* __restrict__ up_img,
                                      float* __restrict__ y) {
  const size_t total = (size_t)batch * h * w_ * k;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int o = idx % k;
  size_t pix = idx / k;
  const int ox = pix % w_;
  pix /= w_;
  const int oy = pix % h, n = pix / h;
  float v = __ldg(acc + idx);
  if (bias) v += __ldg(bias + o);
  if (clamp > 0.f) v = fminf(fmaxf(v, -clamp), clamp);
  if (up_img) v += upsample_tap(up_img + (size_t)n * (h / 2) * (w_ / 2) * k, h / 2, w_ / 2, k, oy, ox, o);
  y[idx] = v;
}

// ---------------------------------------------------------------- encoder blur
__global__ void blur_kernel(int batch, int h, int w_, int c, int pad0, int stride, int oh, int ow, float gain,
                            const float* __restrict__ x, const __nv_bfloat16* __restrict__ x_hi,
                            const __nv_bfloat16* __restrict__ x_lo, float* __restrict__ y,
                            __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo) {
  const int c4 = c >> 2;
  size_t total = (size_t)batch * oh * ow * c4;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int cq = idx % c4;
  size_t pix = idx / c4;
  int ox = pix % ow;
  size_t r = pix / ow;
  int oy = r % oh;
  int n = r / oh;
  const float g[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  const size_t nb = (size_t)n * h * w_ * c4;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    int iy = oy * stride + ky - pad0;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      int ix = ox * stride + kx - pad0;
      if (ix < 0 || ix >= w_) continue;
      float wgt = g[ky] * g[kx] * gain;
      float4 v = ld4_any(x, x_hi, x_lo, nb + ((size_t)iy * w_ + ix) * c4 + cq);
      s[0] = fmaf(wgt, v.x, s[0]);
      s[1] = fmaf(wgt, v.y, s[1]);
      s[2] = fmaf(wgt, v.z, s[2]);
      s[3] = fmaf(wgt, v.w, s[3]);
    }
  }
  st4_any(y, y_hi, y_lo, idx, s);
}


// any channel count (3-channel images of the super-resolution skip path), fp32 only
__global__ void blur_scalar_kernel(int batch, int h, int w_, int c, int pad0, int stride, int oh, int ow, float gain,
                                   const float* __restrict__ x, float* __restrict__ y) {
  size_t total = (size_t)batch * oh * ow * c;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int ch = idx % c;
  size_t pix = idx / c;
  int ox = pix % ow;
  size_t r = pix / ow;
  int oy = r % oh;
  int n = r / oh;
  const float g[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  float s = 0.f;
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    int iy = oy * stride + ky - pad0;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      int ix = ox * stride + kx - pad0;
      if (ix < 0 || ix >= w_) continue;
      s = fmaf(g[ky] * g[kx] * gain, __ldg(x + (((size_t)n * h + iy) * w_ + ix) * c + ch), s);
    }
  }
  y[idx] = s;
}

__global__ void nchw_to_nhwc_kernel(int batch, int c, int hw, const float* __restrict__ x, float* __restrict__ y) {
  size_t total = (size_t)batch * c * hw;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int ch = idx % c;
  size_t r = idx / c;
  int p = r % hw;
  int n = r / hw;
  y[idx] = __ldg(x + ((size_t)n * c + ch) * hw + p);
}

__global__ void nhwc_to_nchw_kernel(int batch, int c, int hw, const float* __restrict__ x, float* __restrict__ y) {
  size_t total = (size_t)batch * c * hw;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int p = idx % hw;
  size_t r = idx / hw;
  int ch = r % c;
  int n = r / c;
  y[idx] = __ldg(x + ((size_t)n * hw + p) * c + ch);
}

}  // namespace hfagp

using namespace hfagp;

extern "C" int hfagp_conv2d_fwd(const HfagpConvDesc* desc, const float* x, const float* w, const float* dcoef,
                                const float* noise, const float* bias, const float* residual, const float* up_img,
                                float* y, void* stream) {
  HFAGP_CHECK_ARG(desc && x && w && y, "conv2d_fwd: null pointer");
  const HfagpConvDesc& d = *desc;
  HFAGP_CHECK_ARG(d.batch > 0 && d.cin > 0 && d.cout > 0 && d.oh > 0 && d.ow > 0, "conv2d_fwd: bad dims");
  HFAGP_CHECK_ARG(d.ntaps > 0 && d.ntaps <= HFAGP_MAX_TAPS, "conv2d_fwd: ntaps %d out of range", d.ntaps);
  HFAGP_CHECK_ARG(d.in_stride == 1 || d.in_stride == 2, "conv2d_fwd: in_stride must be 1 or 2");
  HFAGP_CHECK_ARG(d.out_stride >= 1 && (d.oh - 1) * d.out_stride + d.out_off_y < d.out_h &&
                      (d.ow - 1) * d.out_stride + d.out_off_x < d.out_w,
                  "conv2d_fwd: output window exceeds out_h/out_w");
  HFAGP_CHECK_ARG(!up_img || (d.up_h * 2 == d.out_h && d.up_w * 2 == d.out_w), "conv2d_fwd: up_img must be out/2");
  HFAGP_CHECK_ARG(d.batch <= 65535, "conv2d_fwd: batch too large");
  ConvParams p{d, x, w, dcoef, noise, bias, residual, up_img, y};
  const long long M = (long long)d.oh * d.ow;
  cudaStream_t st = (cudaStream_t)stream;
  // large tiles when they still fill the machine, small tiles otherwise
  long long big_ctas = (long long)cdiv(M, 128) * cdiv(d.cout, 128) * d.batch;
  if (big_ctas >= 148 && d.cout >= 96) {
    dim3 grid(cdiv(M, 128), cdiv(d.cout, 128), d.batch);
    conv_igemm_kernel<8, 8><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid(cdiv(M, 64), cdiv(d.cout, 64), d.batch);
    conv_igemm_kernel<4, 4><<<grid, 256, 0, st>>>(p);
  }
  HFAGP_CHECK_LAUNCH("conv_igemm_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_conv_epilogue_fwd(const HfagpConvDesc* desc, const float* acc, const float* dcoef, const float* noise,
                                       const float* bias, const float* residual, const float* up_img, float* y,
                                       uint16_t* y_hi, uint16_t* y_lo, void* stream) {
  HFAGP_CHECK_ARG(desc && acc && ((y != nullptr) != (y_hi != nullptr && y_lo != nullptr)),
                  "conv_epilogue_fwd: give acc and either y or (y_hi, y_lo)");
  const HfagpConvDesc& d = *desc;
  HFAGP_CHECK_ARG(d.batch > 0 && d.cout > 0 && (d.cout & 3) == 0, "conv_epilogue_fwd: cout must be a multiple of 4");
  HFAGP_CHECK_ARG(d.out_stride == 1 && d.out_off_y == 0 && d.out_off_x == 0 && d.oh == d.out_h && d.ow == d.out_w,
                  "conv_epilogue_fwd: the output must be dense");
  HFAGP_CHECK_ARG(!up_img || (d.up_h * 2 == d.out_h && d.up_w * 2 == d.out_w), "conv_epilogue_fwd: up_img must be out/2");
  ConvParams p{d, nullptr, nullptr, dcoef, noise, bias, residual, up_img, y};
  const size_t total = (size_t)d.batch * d.out_h * d.out_w * (d.cout >> 2);
  conv_epilogue_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(p, acc, reinterpret_cast<__nv_bfloat16*>(y_hi),
                                                                          reinterpret_cast<__nv_bfloat16*>(y_lo));
  HFAGP_CHECK_LAUNCH("conv_epilogue_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_upfir_act_fwd(int batch, int h2, int w2, int c, const float* t, const float* dcoef,
                                   const float* noise, float noise_gain, const float* bias, int act, float act_gain,
                                   float clamp, float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream) {
  HFAGP_CHECK_ARG(t && ((y != nullptr) != (y_hi != nullptr && y_lo != nullptr)),
                  "upfir_act_fwd: give t and either y or (y_hi, y_lo)");
  HFAGP_CHECK_ARG(batch > 0 && h2 > 0 && w2 > 0 && c > 0 && (c & 3) == 0, "upfir_act_fwd: c must be a multiple of 4");
  static const int r_env = getenv("HFAGP_UPFIR_R") ? atoi(getenv("HFAGP_UPFIR_R")) : 0;
  const int R = r_env ? r_env : 4;     // measured on B200 (tools/prof_upfir.py): 2: 100 us, 4: 85 us, 8: 91 us, 16: 142 us (512^2 x 128)
  size_t total = (size_t)batch * cdiv(h2, R) * cdiv(w2, 2) * (c >> 2);
  auto* hi_ = reinterpret_cast<__nv_bfloat16*>(y_hi);
  auto* lo_ = reinterpret_cast<__nv_bfloat16*>(y_lo);
  if (R == 4)
    upfir_act_kernel<4><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, h2, w2, c, t, dcoef, noise, noise_gain, bias, act, act_gain, clamp, y, hi_, lo_);
  else if (R == 2)
    upfir_act_kernel<2><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, h2, w2, c, t, dcoef, noise, noise_gain, bias, act, act_gain, clamp, y, hi_, lo_);
  else if (R == 16)
    upfir_act_kernel<16><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, h2, w2, c, t, dcoef, noise, noise_gain, bias, act, act_gain, clamp, y, hi_, lo_);
  else
    upfir_act_kernel<8><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, h2, w2, c, t, dcoef, noise, noise_gain, bias, act, act_gain, clamp, y, hi_, lo_);
  HFAGP_CHECK_LAUNCH("upfir_act_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_torgb_small_fwd(int batch, int h, int w_, int cin, int cout, const float* x, const uint16_t* x_hi,
                                     const uint16_t* x_lo, const float* w, const float* bias, float clamp,
                                     const float* up_img, float* y, void* stream) {
  HFAGP_CHECK_ARG(w && y && ((x != nullptr) != (x_hi != nullptr && x_lo != nullptr)),
                  "torgb_small_fwd: give w, y and either x or (x_hi, x_lo)");
  HFAGP_CHECK_ARG(cout >= 1 && cout <= 4 && (cin & 3) == 0, "torgb_small_fwd: cout<=4 and cin%%4==0 required");
  HFAGP_CHECK_ARG(!up_img || ((h & 1) == 0 && (w_ & 1) == 0), "torgb_small_fwd: odd size with up_img");
  HFAGP_CHECK_ARG(batch > 0 && batch <= 65535 && (size_t)cout * cin * 4 <= 48 * 1024, "torgb_small_fwd: bad dims");
  const int pix_per_block = 8 * 4 * TORGB_PASSES;
  dim3 grid(cdiv((long long)h * w_, pix_per_block), batch);
  torgb_small_kernel<<<grid, 256, (size_t)cout * cin * 4, (cudaStream_t)stream>>>(
      batch, h, w_, cin, cout, x, reinterpret_cast<const __nv_bfloat16*>(x_hi),
      reinterpret_cast<const __nv_bfloat16*>(x_lo), w, bias, clamp, up_img, y);
  HFAGP_CHECK_LAUNCH("torgb_small_kernel");
  return HFAGP_OK;
}

// stride-1 blur, two adjacent output columns per thread: the 4x5 input window is loaded once (10 loads per output
// instead of 16) and both horizontal sums are formed per row
__global__ void blur2_kernel(int batch, int h, int w_, int c, int pad0, int oh, int ow, float gain,
                             const float* __restrict__ x, const __nv_bfloat16* __restrict__ x_hi,
                             const __nv_bfloat16* __restrict__ x_lo, float* __restrict__ y,
                             __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo) {
  const int c4 = c >> 2, wp = (ow + 1) >> 1;
  const size_t total = (size_t)batch * oh * wp * c4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cq = idx % c4;
  size_t r = idx / c4;
  const int px = r % wp;
  r /= wp;
  const int oy = r % oh, n = r / oh;
  const int ox0 = px * 2;
  const float g[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  const size_t nb = (size_t)n * h * w_ * c4 + cq;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const int iy = oy + ky - pad0;
    if (iy < 0 || iy >= h) continue;
    float4 v[5];
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
      const int ix = ox0 + kx - pad0;
      v[kx] = (ix >= 0 && ix < w_) ? ld4_any(x, x_hi, x_lo, nb + ((size_t)iy * w_ + ix) * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float wy = g[ky] * gain;
    a[0] = fmaf(wy, g[0] * v[0].x + g[1] * v[1].x + g[2] * v[2].x + g[3] * v[3].x, a[0]);
    a[1] = fmaf(wy, g[0] * v[0].y + g[1] * v[1].y + g[2] * v[2].y + g[3] * v[3].y, a[1]);
    a[2] = fmaf(wy, g[0] * v[0].z + g[1] * v[1].z + g[2] * v[2].z + g[3] * v[3].z, a[2]);
    a[3] = fmaf(wy, g[0] * v[0].w + g[1] * v[1].w + g[2] * v[2].w + g[3] * v[3].w, a[3]);
    b[0] = fmaf(wy, g[0] * v[1].x + g[1] * v[2].x + g[2] * v[3].x + g[3] * v[4].x, b[0]);
    b[1] = fmaf(wy, g[0] * v[1].y + g[1] * v[2].y + g[2] * v[3].y + g[3] * v[4].y, b[1]);
    b[2] = fmaf(wy, g[0] * v[1].z + g[1] * v[2].z + g[2] * v[3].z + g[3] * v[4].z, b[2]);
    b[3] = fmaf(wy, g[0] * v[1].w + g[1] * v[2].w + g[2] * v[3].w + g[3] * v[4].w, b[3]);
  }
  const size_t o = (((size_t)n * oh + oy) * ow + ox0) * c4 + cq;
  st4_any(y, y_hi, y_lo, o, a);
  if (ox0 + 1 < ow) st4_any(y, y_hi, y_lo, o + c4, b);
}

extern "C" int hfagp_torgb_finalize_fwd(int batch, int h, int w_, int k, const float* acc, const float* bias, float clamp,
                                        const float* up_img, float* y, void* stream) {
  HFAGP_CHECK_ARG(acc && y && batch > 0 && h > 0 && w_ > 0 && k >= 1 && k <= 4, "torgb_finalize_fwd: bad args");
  HFAGP_CHECK_ARG(!up_img || ((h & 1) == 0 && (w_ & 1) == 0), "torgb_finalize_fwd: odd size with up_img");
  torgb_finalize_kernel<<<cdiv((long long)batch * h * w_ * k, 256), 256, 0, (cudaStream_t)stream>>>(batch, h, w_, k, acc, bias,
                                                                                               clamp, up_img, y);
  HFAGP_CHECK_LAUNCH("torgb_finalize_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_blur_fwd(int batch, int h, int w_, int c, int pad0, int pad1, int stride, float gain,
                              const float* x, const uint16_t* x_hi, const uint16_t* x_lo, float* y, uint16_t* y_hi,
                              uint16_t* y_lo, void* stream) {
  HFAGP_CHECK_ARG((x != nullptr) != (x_hi != nullptr && x_lo != nullptr), "blur_fwd: give x or (x_hi, x_lo)");
  HFAGP_CHECK_ARG((y != nullptr) != (y_hi != nullptr && y_lo != nullptr), "blur_fwd: give y or (y_hi, y_lo)");
  HFAGP_CHECK_ARG(stride == 1 || stride == 2, "blur_fwd: stride 1|2 required");
  int oh = (h + pad0 + pad1 - 4) / stride + 1;
  int ow = (w_ + pad0 + pad1 - 4) / stride + 1;
  HFAGP_CHECK_ARG(oh > 0 && ow > 0, "blur_fwd: empty output");
  if ((c & 3) == 0 && stride == 1) {
    size_t total = (size_t)batch * oh * cdiv(ow, 2) * (c >> 2);
    blur2_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        batch, h, w_, c, pad0, oh, ow, gain, x, reinterpret_cast<const __nv_bfloat16*>(x_hi),
        reinterpret_cast<const __nv_bfloat16*>(x_lo), y, reinterpret_cast<__nv_bfloat16*>(y_hi),
        reinterpret_cast<__nv_bfloat16*>(y_lo));
  } else if ((c & 3) == 0) {
    size_t total = (size_t)batch * oh * ow * (c >> 2);
    blur_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        batch, h, w_, c, pad0, stride, oh, ow, gain, x, reinterpret_cast<const __nv_bfloat16*>(x_hi),
        reinterpret_cast<const __nv_bfloat16*>(x_lo), y, reinterpret_cast<__nv_bfloat16*>(y_hi),
        reinterpret_cast<__nv_bfloat16*>(y_lo));
  } else {
    HFAGP_CHECK_ARG(x && y, "blur_fwd: split-bf16 I/O needs c%%4 == 0");
    size_t total = (size_t)batch * oh * ow * c;
    blur_scalar_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, h, w_, c, pad0, stride, oh, ow, gain, x, y);
  }
  HFAGP_CHECK_LAUNCH("blur_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_nchw_to_nhwc(int batch, int c, int h, int w_, const float* x, float* y, void* stream) {
  HFAGP_CHECK_ARG(x && y && batch > 0 && c > 0, "nchw_to_nhwc: bad args");
  size_t total = (size_t)batch * c * h * w_;
  nchw_to_nhwc_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, c, h * w_, x, y);
  HFAGP_CHECK_LAUNCH("nchw_to_nhwc_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_nhwc_to_nchw(int batch, int c, int h, int w_, const float* x, float* y, void* stream) {
  HFAGP_CHECK_ARG(x && y && batch > 0 && c > 0, "nhwc_to_nchw: bad args");
  size_t total = (size_t)batch * c * h * w_;
  nhwc_to_nchw_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, c, h * w_, x, y);
  HFAGP_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return HFAGP_OK;
}
