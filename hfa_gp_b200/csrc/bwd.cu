// Backward companions of the convolution path (training step, trainer_rgb.py:73-98):
//   act_bwd      fused "everything between two convolutions": sums the incoming gradients (each with its
//                per-(sample, channel) style scale), applies the activation / gain / clamp / residual-merge
//                derivative and the demodulation coefficient, writes the gradient of the convolution output
//                (fp32 or split-bf16 for the tensor-core dgrad), and reduces d(styles), d(bias), d(dcoef).
//   blur_up      transpose of the strided encoder blur; the stride-1 blurs and the FIR transposes reuse blur_fwd.
//   styles_bwd / demod_bwd   gradients of all style affines back to ws; demodulation coefficient backward.
//   linear_bwd   EqualLinear backward (dx, dW, db).
//   wgrad        fp32 SIMT weight gradient of the encoder convolutions (split-K, atomics).
// The data-gradient convolutions themselves are hfagp_conv2d_tc_fwd / hfagp_conv2d_fwd calls on transposed
// weights with mirrored tap lists (see hfa_gp_b200/autograd.py).
#include <cuda_bf16.h>
#include "common.cuh"
#include "splitio.cuh"

namespace hfagp {

struct ActBwdParams {
  HfagpActBwdDesc d;
  const float* y; const __nv_bfloat16* y_hi; const __nv_bfloat16* y_lo;
  const float* g0; const float* s0;
  const float* g1; const float* s1;
  const float* dimg; const float* wrgb; const float* srgb;
  const float* dcoef; const float* noise; const float* bias; const float* residual;
  float* dz; __nv_bfloat16* dz_hi; __nv_bfloat16* dz_lo;
  float* ds0; float* ds1; float* dsrgb; float* dbias; float* ddcoef;
  int pix_per_block;
};

__device__ __forceinline__ void atomic_add4(float* p, const float v[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) atomicAdd(p + k, v[k]);
}

// Specialised on which optional operands exist (second gradient stream, fused small-ToRGB gradient, residual merge): the
// common call — one gradient stream into a modulated layer — then fits 3 CTAs per SM instead of 1-2 (the generic form
// needs 128+ registers for operands it never touches, and this is a latency-bound stream kernel).
template <bool HAS_G1, bool HAS_DIMG, bool HAS_RES>
__global__ void __launch_bounds__(256, (HAS_G1 || HAS_DIMG) ? 2 : 3) act_bwd_kernel(const ActBwdParams p) {
  const HfagpActBwdDesc& d = p.d;
  const int c4 = d.c >> 2;
  const int planes = 256 / c4;                 // pixels processed side by side
  const int cq = threadIdx.x % c4, plane = threadIdx.x / c4;
  const bool idle = plane >= planes;             // (256 is not always a multiple of c/4; idle lanes only meet the barrier)
  const int n = blockIdx.y;
  const int hw = d.h * d.w;
  const int p_begin = blockIdx.x * p.pix_per_block;
  const int p_end = idle ? 0 : min(hw, p_begin + p.pix_per_block);
  const int c0 = cq * 4;
  const size_t nc = (size_t)n * d.c + c0;

  float sc0[4] = {1.f, 1.f, 1.f, 1.f}, sc1[4] = {1.f, 1.f, 1.f, 1.f}, scr[4] = {0.f, 0.f, 0.f, 0.f};
  float dco[4] = {1.f, 1.f, 1.f, 1.f}, bs[4] = {0.f, 0.f, 0.f, 0.f};
  float wr[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (p.s0) sc0[k] = __ldg(p.s0 + nc + k);
    if (HAS_G1 && p.s1) sc1[k] = __ldg(p.s1 + nc + k);
    if (HAS_DIMG) scr[k] = p.srgb ? __ldg(p.srgb + nc + k) : 1.f;
    if (p.dcoef) dco[k] = __ldg(p.dcoef + nc + k);
    if (p.bias) bs[k] = __ldg(p.bias + c0 + k);
#pragma unroll
    for (int o = 0; o < 4; ++o) wr[o][k] = (HAS_DIMG && o < d.rgb_k) ? __ldg(p.wrgb + (size_t)o * d.c + c0 + k) : 0.f;
  }
  const float slope_pos = d.act_gain, slope_neg = act_slope(d.act) * d.act_gain;
  // reciprocals once per thread instead of two fp32 divisions per element (ReLU's zero slope: that side has dpre = 0)
  const float inv_pos = 1.f / slope_pos, inv_neg = slope_neg != 0.f ? 1.f / slope_neg : 0.f;
  float inv_dco[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) inv_dco[k] = 1.f / dco[k];
  const float inv_rs = d.residual_scale != 0.f ? 1.f / d.residual_scale : 1.f;

  float r0[4] = {0.f, 0.f, 0.f, 0.f}, r1[4] = {0.f, 0.f, 0.f, 0.f}, rr[4] = {0.f, 0.f, 0.f, 0.f};
  float rb[4] = {0.f, 0.f, 0.f, 0.f}, rd[4] = {0.f, 0.f, 0.f, 0.f};

  // U pixels per iteration with every load issued before the first dependent use (the stores of one pixel may alias the
  // loads of the next as far as the compiler knows, so the batching is explicit): the kernel is a pure stream, and one
  // pixel at a time left a single 16 B load chain in flight per thread
  constexpr int U = 2;
  for (int pix0 = p_begin + plane; pix0 < p_end; pix0 += planes * U) {
    float4 y4[U], a0[U], a1[U], r4[U];
    float nzv[U];
    float di[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * planes;
      const bool ok = pix < p_end;
      const size_t q = ((size_t)n * hw + (ok ? pix : p_begin)) * c4 + cq;
      y4[u] = ld4_any(p.y, p.y_hi, p.y_lo, q);
      a0[u] = p.g0 ? __ldg(reinterpret_cast<const float4*>(p.g0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      a1[u] = HAS_G1 ? __ldg(reinterpret_cast<const float4*>(p.g1) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      r4[u] = HAS_RES ? __ldg(reinterpret_cast<const float4*>(p.residual) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      nzv[u] = p.noise ? __ldg(p.noise + (ok ? pix : p_begin)) * d.noise_gain : 0.f;
      if (HAS_DIMG) {
        const float* dp = p.dimg + ((size_t)n * hw + (ok ? pix : p_begin)) * d.rgb_k;
#pragma unroll
        for (int o = 0; o < 4; ++o) di[u][o] = o < d.rgb_k ? __ldg(dp + o) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * planes;
      if (pix >= p_end) break;
      const size_t q = ((size_t)n * hw + pix) * c4 + cq;
      const float yv[4] = {y4[u].x, y4[u].y, y4[u].z, y4[u].w};
      float gsum[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.g0) {
        const float av[4] = {a0[u].x, a0[u].y, a0[u].z, a0[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { gsum[k] = av[k] * sc0[k]; r0[k] = fmaf(av[k], yv[k], r0[k]); }
      }
      if (HAS_G1) {
        const float av[4] = {a1[u].x, a1[u].y, a1[u].z, a1[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { gsum[k] = fmaf(av[k], sc1[k], gsum[k]); r1[k] = fmaf(av[k], yv[k], r1[k]); }
      }
      if (HAS_DIMG) {
        float gr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int k = 0; k < 4; ++k) gr[k] = fmaf(di[u][o], wr[o][k], gr[k]);      // wr[o] = 0 beyond rgb_k
#pragma unroll
        for (int k = 0; k < 4; ++k) { gsum[k] = fmaf(gr[k], scr[k], gsum[k]); rr[k] = fmaf(gr[k], yv[k], rr[k]); }
      }
      // derivative of  y = merge(clamp(act(pre) * gain))  w.r.t. pre
      float av[4];
      if (HAS_RES) {
        av[0] = yv[0] * inv_rs - r4[u].x; av[1] = yv[1] * inv_rs - r4[u].y; av[2] = yv[2] * inv_rs - r4[u].z; av[3] = yv[3] * inv_rs - r4[u].w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) av[k] = yv[k];
      }
      const float nz = nzv[u];
      float dzv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool pos = av[k] > 0.f;
        const float slope = pos ? slope_pos : slope_neg;
        const bool pass = !(d.clamp > 0.f) || fabsf(av[k]) < d.clamp;
        const float dpre = pass ? gsum[k] * d.post_scale * slope : 0.f;
        rb[k] += dpre;
        if (p.ddcoef) {
          const float z = (av[k] * (pos ? inv_pos : inv_neg) - nz - bs[k]) * inv_dco[k];   // conv output before demodulation
          rd[k] = fmaf(dpre, z, rd[k]);
        }
        dzv[k] = dpre * dco[k];
      }
      if (p.dz || p.dz_hi) st4_any(p.dz, p.dz_hi, p.dz_lo, q, dzv);
    }
  }
  // per-channel reductions: first across the block's side-by-side pixel lanes in shared memory, then ONE atomic per
  // channel and block (every lane adding on its own put ~5 000 same-address atomics per channel in a row: the L2 unit
  // serialises them, which — not the streaming — was what the kernel's time went into)
  __shared__ float red[5][256 * 4];
  const float* vals[5] = {r0, r1, rr, rb, rd};
  const bool use[5] = {p.ds0 && p.g0, p.ds1 && HAS_G1, p.dsrgb && HAS_DIMG, p.dbias != nullptr, p.ddcoef != nullptr};
#pragma unroll
  for (int j = 0; j < 5; ++j)
    if (use[j]) {
#pragma unroll
      for (int k = 0; k < 4; ++k) red[j][(plane * c4 + cq) * 4 + k] = vals[j][k];
    }
  __syncthreads();
  if (plane == 0) {
    float* dst[5] = {p.ds0 ? p.ds0 + nc : nullptr, p.ds1 ? p.ds1 + nc : nullptr, p.dsrgb ? p.dsrgb + nc : nullptr,
                     p.dbias ? p.dbias + c0 : nullptr, p.ddcoef ? p.ddcoef + nc : nullptr};
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (use[j]) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int pl = 0; pl < planes; ++pl)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[k] += red[j][(pl * c4 + cq) * 4 + k];
        atomic_add4(dst[j], acc);
      }
  }
}

// dx[n][iy][ix][c] = gain * sum_{ky,kx} g[ky] g[kx] dy[n][(iy + pad0 - ky)/s][(ix + pad0 - kx)/s][c]   (exact divisions only)
template <int V>
__global__ void blur_up_kernel(int batch, int h, int w_, int c, int pad0, int stride, int oh, int ow, float gain,
                               const float* __restrict__ dy, float* __restrict__ dx) {
  const int cv = c / V;
  size_t total = (size_t)batch * h * w_ * cv;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int cq = idx % cv;
  size_t pix = idx / cv;
  int ix = pix % w_;
  size_t r = pix / w_;
  int iy = r % h;
  int n = r / h;
  const float g[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  float s[V];
#pragma unroll
  for (int k = 0; k < V; ++k) s[k] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    int ty = iy + pad0 - ky;
    if (ty < 0 || ty % stride) continue;
    int oy = ty / stride;
    if (oy >= oh) continue;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      int tx = ix + pad0 - kx;
      if (tx < 0 || tx % stride) continue;
      int ox = tx / stride;
      if (ox >= ow) continue;
      const float wgt = g[ky] * g[kx] * gain;
      const size_t o = (((size_t)n * oh + oy) * ow + ox) * cv + cq;
      if (V == 4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(dy) + o);
        s[0] = fmaf(wgt, v.x, s[0]); s[1 % V] = fmaf(wgt, v.y, s[1 % V]);
        s[2 % V] = fmaf(wgt, v.z, s[2 % V]); s[3 % V] = fmaf(wgt, v.w, s[3 % V]);
      } else {
        s[0] = fmaf(wgt, __ldg(dy + o), s[0]);
      }
    }
  }
  if (V == 4) reinterpret_cast<float4*>(dx)[idx] = make_float4(s[0], s[1 % V], s[2 % V], s[3 % V]);
  else dx[idx] = s[0];
}

// ---------------------------------------------------------------- styles / demodulation backward
constexpr int MAX_STYLE_LAYERS_B = 48;
struct StyleBwdTable {
  const float* aw[MAX_STYLE_LAYERS_B];
  long long off[MAX_STYLE_LAYERS_B];
  int cin[MAX_STYLE_LAYERS_B];
  int widx[MAX_STYLE_LAYERS_B];
  float gain[MAX_STYLE_LAYERS_B];
};

// block (layer, n): dws[n][widx][k] += gain/sqrt(wdim) * sum_i dstyles[l][n][i] * A_l[i][k]
__global__ void styles_bwd_kernel(const StyleBwdTable tb, int num_ws, int w_dim, float inv_sqrt,
                                  const float* __restrict__ dstyles, float* __restrict__ dws) {
  const int l = blockIdx.x, n = blockIdx.y;
  const int cin = tb.cin[l];
  const float* ds = dstyles + tb.off[l] + (size_t)n * cin;
  const float* a = tb.aw[l];
  // gridDim.z slices of the input channels: at batch 1 the 26 layers alone would leave most SMs idle behind 512-long
  // dependent chains
  const int per = (cin + gridDim.z - 1) / gridDim.z;
  const int i0 = blockIdx.z * per, i1 = min(cin, i0 + per);
  if (i0 >= i1) return;
  for (int k = threadIdx.x; k < w_dim; k += blockDim.x) {
    float acc = 0.f;
#pragma unroll 4
    for (int i = i0; i < i1; ++i) acc = fmaf(__ldg(ds + i), __ldg(a + (size_t)i * w_dim + k), acc);
    atomicAdd(dws + ((size_t)n * num_ws + tb.widx[l]) * w_dim + k, acc * tb.gain[l] * inv_sqrt);
  }
}

// ds[n][i] -= s[n][i] * sum_o ddcoef[n][o] * dcoef[n][o]^3 * w2[o][i]      (w2 = sum_taps w^2)
// block = 32 input channels x 8 slices of the output channels (coalesced w2 rows), shared-memory reduce over slices
__global__ void __launch_bounds__(256) demod_bwd_kernel(int cout, int cin, const float* __restrict__ w2,
                                                       const float* __restrict__ styles,
                                                       const float* __restrict__ dcoef, const float* __restrict__ ddcoef,
                                                       float* __restrict__ dstyles) {
  const int n = blockIdx.y;
  const int ix = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ix;
  float acc = 0.f;
  if (i < cin)
    for (int o = slice; o < cout; o += 8) {
      const float dc = __ldg(dcoef + (size_t)n * cout + o);
      acc = fmaf(__ldg(ddcoef + (size_t)n * cout + o) * dc * dc * dc, __ldg(w2 + (size_t)o * cin + i), acc);
    }
  __shared__ float red[8][33];
  red[slice][ix] = acc;
  __syncthreads();
  if (slice == 0 && i < cin) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][ix];
    dstyles[(size_t)n * cin + i] -= __ldg(styles + (size_t)n * cin + i) * t;
  }
}

// ---------------------------------------------------------------- EqualLinear backward
// gridDim.y slices of the output channels (1 slice: dx is written; more: atomically added into a zeroed dx)
__global__ void linear_bwd_dx_kernel(int batch, int cin, int cout, const float* __restrict__ dy,
                                     const float* __restrict__ w, float w_gain, float* __restrict__ dx) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * cin) return;
  int n = idx / cin, i = idx - n * cin;
  const int per = (cout + gridDim.y - 1) / gridDim.y;
  const int o0 = blockIdx.y * per, o1 = min(cout, o0 + per);
  float acc = 0.f;
  for (int o = o0; o < o1; ++o) acc = fmaf(__ldg(dy + (size_t)n * cout + o), __ldg(w + (size_t)o * cin + i), acc);
  if (gridDim.y == 1) dx[idx] = acc * w_gain;
  else atomicAdd(dx + idx, acc * w_gain);
}
__global__ void linear_bwd_dw_kernel(int batch, int cin, int cout, const float* __restrict__ dy,
                                     const float* __restrict__ x, float w_gain, float b_gain, float* __restrict__ dw,
                                     float* __restrict__ db) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cout * cin) return;
  int o = idx / cin, i = idx - o * cin;
  float acc = 0.f, bacc = 0.f;
  for (int n = 0; n < batch; ++n) {
    const float g = __ldg(dy + (size_t)n * cout + o);
    acc = fmaf(g, __ldg(x + (size_t)n * cin + i), acc);
    bacc += g;
  }
  dw[idx] += acc * w_gain;
  if (db && i == 0) db[o] += bacc * b_gain;
}

// ---------------------------------------------------------------- weight gradient (fp32 SIMT, split-K)
// dw[t][o][i] += scale * sum_{n,my,mx} dz[n][my][mx][o] * x[n][my*stride + dy_t][mx*stride + dx_t][i]
struct WgradParams {
  int batch, in_h, in_w, cin, cout, oh, ow, in_stride, ntaps;
  int dy[HFAGP_MAX_TAPS], dx[HFAGP_MAX_TAPS], wtap[HFAGP_MAX_TAPS];
  const float* x; const __nv_bfloat16* x_hi; const __nv_bfloat16* x_lo;
  const float* dz; const __nv_bfloat16* dz_hi; const __nv_bfloat16* dz_lo;
  float* dw;
  float scale;
  int k_per_split;
  const float* xscale;    // optional [batch][cin]: x is multiplied per (sample, channel) — the style of a modulated conv
  const float* dzscale;   // optional [batch][cout]: the same for dz (transposed-role use: the dense operand is x)
};

constexpr int WG_T = 64, WG_K = 16, WG_LD = WG_T + 4;

__global__ void __launch_bounds__(256) wgrad_kernel(const WgradParams p) {
  __shared__ __align__(16) float As[2][WG_K * WG_LD];   // dz tile: [pixel][cout]
  __shared__ __align__(16) float Bs[2][WG_K * WG_LD];   // x tile:  [pixel][cin]
  const int tid = threadIdx.x;
  const int tiles_i = (p.cin + WG_T - 1) / WG_T;
  const int o0 = (blockIdx.x / tiles_i) * WG_T, i0 = (blockIdx.x % tiles_i) * WG_T;
  const int t = blockIdx.y;
  const long long K = (long long)p.batch * p.oh * p.ow;
  const long long k_begin = (long long)blockIdx.z * p.k_per_split;
  const long long k_end = k_begin + p.k_per_split < K ? k_begin + p.k_per_split : K;
  const int row = tid >> 4, quad = tid & 15;          // load coordinates: pixel row of the chunk, channel quad
  const int ty = tid >> 4, tx = tid & 15;             // compute coordinates: 4 couts x 4 cins
  const int tdy = p.dy[t], tdx = p.dx[t];
  const int ohw = p.oh * p.ow;

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  float4 ra, rb;
  auto load = [&](long long kb) {
    const long long k = kb + row;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    if (k < k_end) {
      const int n = (int)(k / ohw);
      const int rem = (int)(k - (long long)n * ohw);
      const int my = rem / p.ow, mx = rem - my * p.ow;
      const int co = o0 + quad * 4;
      if (co < p.cout) {
        ra = ld4_any(p.dz, p.dz_hi, p.dz_lo, ((size_t)k * p.cout + co) >> 2);
        if (p.dzscale) {
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.dzscale + (size_t)n * p.cout + co));
          ra.x *= s4.x; ra.y *= s4.y; ra.z *= s4.z; ra.w *= s4.w;
        }
      }
      const int iy = my * p.in_stride + tdy, ix = mx * p.in_stride + tdx, ci = i0 + quad * 4;
      if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w && ci < p.cin) {
        rb = ld4_any(p.x, p.x_hi, p.x_lo, ((((size_t)n * p.in_h + iy) * p.in_w + ix) * p.cin + ci) >> 2);
        if (p.xscale) {
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.xscale + (size_t)n * p.cin + ci));
          rb.x *= s4.x; rb.y *= s4.y; rb.z *= s4.z; rb.w *= s4.w;
        }
      }
    }
  };
  auto store = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][row * WG_LD + quad * 4]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf][row * WG_LD + quad * 4]) = rb;
  };
  load(k_begin);
  store(0);
  __syncthreads();
  int buf = 0;
  for (long long kb = k_begin; kb < k_end; kb += WG_K) {
    if (kb + WG_K < k_end) load(kb + WG_K);
#pragma unroll
    for (int kk = 0; kk < WG_K; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk * WG_LD + ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk * WG_LD + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) acc[aa][bb] = fmaf(av[aa], bv[bb], acc[aa][bb]);
    }
    if (kb + WG_K < k_end) store(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  float* out = p.dw + (size_t)p.wtap[t] * p.cout * p.cin;
#pragma unroll
  for (int aa = 0; aa < 4; ++aa) {
    const int o = o0 + ty * 4 + aa;
    if (o >= p.cout) continue;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int i = i0 + tx * 4 + bb;
      if (i < p.cin) atomicAdd(out + (size_t)o * p.cin + i, acc[aa][bb] * p.scale);
    }
  }
}

// Weight gradient of a 1x1 convolution with FOUR (padded) output channels — the 3-channel ToRGB of the super-resolution
// blocks: dw[o][i] += scale * xscale[n][i] * sum_p dz[n][p][o] * x[n][p][i].  A pure stream over the layer input (134 MB at
// 512^2 x 128 x batch 2): a warp covers 32 channel quads of one pixel (coalesced), a block's warps take different pixels,
// partial sums stay in registers over the block's pixel range, then one shared-memory reduction and one atomic per element
// and block.  The generic tiled SIMT wgrad spent 0.4 ms on this shape (its 64 x 64 output tile is 1/16 full).
__global__ void __launch_bounds__(256) wgrad_cout4_kernel(int hw, int cin, int pix_per_block, const float* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo,
                                                         const float* __restrict__ dz, const float* __restrict__ xscale, float scale,
                                                         float* __restrict__ dw) {
  __shared__ float red[8][16][33];
  const int c4 = cin >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(hw, p0 + pix_per_block);
  for (int cq0 = 0; cq0 < c4; cq0 += 32) {           // 32 channel quads per sweep (one sweep for cin <= 128)
    const int cq = cq0 + lane;
    float acc[4][4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[o][k] = 0.f;
    if (cq < c4) {
      for (int pix = p0 + warp; pix < p1; pix += 8) {
        const size_t q = ((size_t)n * hw + pix) * c4 + cq;
        const float4 xv = ld4_any(x, x_hi, x_lo, q);
        const float4 g = __ldg(reinterpret_cast<const float4*>(dz) + (size_t)n * hw + pix);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[o][k] = fmaf(gs[o], xs[k], acc[o][k]);
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int k = 0; k < 4; ++k) red[warp][o * 4 + k][lane] = acc[o][k];
    __syncthreads();
    // 16 values x 32 lanes = 512 sums over the 8 warps: two per thread
    for (int e = threadIdx.x; e < 512; e += 256) {
      const int v = e >> 5, l = e & 31;
      float t = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) t += red[w8][v][l];
      const int o = v >> 2, k = v & 3, ci = (cq0 + l) * 4 + k;
      if (cq0 + l < c4) {
        const float sc = scale * (xscale ? __ldg(xscale + (size_t)n * cin + ci) : 1.f);
        atomicAdd(dw + (size_t)o * cin + ci, t * sc);
      }
    }
    __syncthreads();
  }
}

// Last step of a modulated convolution's weight gradient (generator unfrozen): the style-scaled wgrad dw (packed
// [tap][o][i], or [tap][i][o] when the layer is an up-sampling one and the wgrad ran with the roles swapped) plus the
// demodulation term  -W[t][o][i] * sum_n ddcoef[n][o] dcoef[n][o]^3 styles[n][i]^2,  unpacked into the parameter layout and
// ADDED to grad[o][i][tap].  A 32 x 32 (o, i) tile per block goes through shared memory so that both the reads (along the
// fast index of dw) and the writes (taps * 32 consecutive floats per o) are coalesced.  One launch replaces ~10 elementwise
// passes over the 9.4 MB weight tensor.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256) modconv_wgrad_finish_kernel(int taps, int cout, int cin, int batch, const float* __restrict__ dw,
                                                                  const float* __restrict__ w, const float* __restrict__ ddcoef,
                                                                  const float* __restrict__ dcoef, const float* __restrict__ styles,
                                                                  float* __restrict__ grad) {
  __shared__ float tile[32][33];
  __shared__ float a_s[8][32], s2_s[8][32];                   // per sample: ddcoef * dcoef^3 for the tile's o, styles^2 for its i
  const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 8 rows of 32
  const bool demod = ddcoef != nullptr;
  float coef[4] = {0.f, 0.f, 0.f, 0.f};                       // this thread's (o = o0 + ty + 8 r, i = i0 + tx)
  if (demod) {
    for (int n0 = 0; n0 < batch; n0 += 8) {
      __syncthreads();
      const int nn = n0 + ty;
      if (nn < batch) {
        const int o = o0 + tx, i = i0 + tx;
        float a = 0.f, s2 = 0.f;
        if (o < cout) { const float dc = __ldg(dcoef + (size_t)nn * cout + o); a = __ldg(ddcoef + (size_t)nn * cout + o) * dc * dc * dc; }
        if (i < cin) { const float sv = __ldg(styles + (size_t)nn * cin + i); s2 = sv * sv; }
        a_s[ty][tx] = a;
        s2_s[ty][tx] = s2;
      }
      __syncthreads();
      const int cnt = min(8, batch - n0);
      for (int n = 0; n < cnt; ++n)
#pragma unroll
        for (int r = 0; r < 4; ++r) coef[r] = fmaf(a_s[n][ty + 8 * r], s2_s[n][tx], coef[r]);
    }
  }
  for (int t = 0; t < taps; ++t) {
    if (TRANSPOSED) {
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 4; ++r) {                           // dw[t][i][o]: coalesced along o, transposed through the tile
        const int i = i0 + ty + 8 * r, o = o0 + tx;
        tile[ty + 8 * r][tx] = (i < cin && o < cout) ? __ldg(dw + ((size_t)t * cin + i) * cout + o) : 0.f;
      }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int o = o0 + ty + 8 * r, i = i0 + tx;
      if (o < cout && i < cin) {
        const size_t q = ((size_t)t * cout + o) * cin + i;
        float v = TRANSPOSED ? tile[tx][ty + 8 * r] : __ldg(dw + q);
        if (demod) v = fmaf(-__ldg(w + q), coef[r], v);
        grad[((size_t)o * cin + i) * taps + t] += v;          // (gathering the taps in registers first measured slower)
      }
    }
  }
}

}  // namespace hfagp

using namespace hfagp;

extern "C" int hfagp_act_bwd(const HfagpActBwdDesc* desc, const float* y, const uint16_t* y_hi, const uint16_t* y_lo,
                             const float* g0, const float* s0, const float* g1, const float* s1, const float* dimg,
                             const float* wrgb, const float* srgb, const float* dcoef, const float* noise,
                             const float* bias, const float* residual, float* dz, uint16_t* dz_hi, uint16_t* dz_lo,
                             float* ds0, float* ds1, float* dsrgb, float* dbias, float* ddcoef, void* stream) {
  HFAGP_CHECK_ARG(desc && (g0 || dimg), "act_bwd: null pointer (need g0 and/or dimg)");
  const HfagpActBwdDesc& d = *desc;
  HFAGP_CHECK_ARG((y != nullptr) != (y_hi != nullptr && y_lo != nullptr), "act_bwd: give y or (y_hi, y_lo)");
  HFAGP_CHECK_ARG(!(dz && dz_hi), "act_bwd: give dz or (dz_hi, dz_lo), not both");
  HFAGP_CHECK_ARG(d.batch > 0 && d.batch <= 65535 && d.h > 0 && d.w > 0 && d.c >= 4 && (d.c & 3) == 0 && d.c <= 1024,
                  "act_bwd: c must be a multiple of 4 in [4, 1024]");
  HFAGP_CHECK_ARG(!dimg || (wrgb && d.rgb_k >= 1 && d.rgb_k <= 4), "act_bwd: small-ToRGB fusion needs wrgb and rgb_k <= 4");
  HFAGP_CHECK_ARG(!ddcoef || dcoef, "act_bwd: ddcoef needs dcoef");
  ActBwdParams p{d, y, reinterpret_cast<const __nv_bfloat16*>(y_hi), reinterpret_cast<const __nv_bfloat16*>(y_lo),
                 g0, s0, g1, s1, dimg, wrgb, srgb, dcoef, noise, bias, residual, dz,
                 reinterpret_cast<__nv_bfloat16*>(dz_hi), reinterpret_cast<__nv_bfloat16*>(dz_lo),
                 ds0, ds1, dsrgb, dbias, ddcoef, 0};
  const int hw = d.h * d.w;
  const int planes = 256 / (d.c >> 2);
  // ~4 waves of CTAs, at least 8 pixels per side-by-side lane so the atomics stay a small share
  int ppb = cdiv((long long)hw * d.batch, device_sm_count() * 4);
  if (ppb < planes * 8) ppb = planes * 8;
  p.pix_per_block = ppb;
  dim3 grid(cdiv(hw, ppb), d.batch);
#define HFAGP_ACT_BWD(G1, DI, RS) act_bwd_kernel<G1, DI, RS><<<grid, 256, 0, (cudaStream_t)stream>>>(p)
  const int variant = (g1 ? 4 : 0) | (dimg ? 2 : 0) | (residual ? 1 : 0);
  switch (variant) {
    case 0: HFAGP_ACT_BWD(false, false, false); break;
    case 1: HFAGP_ACT_BWD(false, false, true); break;
    case 2: HFAGP_ACT_BWD(false, true, false); break;
    case 3: HFAGP_ACT_BWD(false, true, true); break;
    case 4: HFAGP_ACT_BWD(true, false, false); break;
    case 5: HFAGP_ACT_BWD(true, false, true); break;
    case 6: HFAGP_ACT_BWD(true, true, false); break;
    default: HFAGP_ACT_BWD(true, true, true); break;
  }
#undef HFAGP_ACT_BWD
  HFAGP_CHECK_LAUNCH("act_bwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_blur_up(int batch, int h, int w_, int c, int pad0, int pad1, int stride, float gain,
                             const float* dy, float* dx, void* stream) {
  HFAGP_CHECK_ARG(dy && dx && batch > 0 && c > 0 && (stride == 1 || stride == 2), "blur_up: bad args");
  const int oh = (h + pad0 + pad1 - 4) / stride + 1, ow = (w_ + pad0 + pad1 - 4) / stride + 1;
  HFAGP_CHECK_ARG(oh > 0 && ow > 0, "blur_up: empty gradient");
  if ((c & 3) == 0) {
    size_t total = (size_t)batch * h * w_ * (c >> 2);
    blur_up_kernel<4><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, h, w_, c, pad0, stride, oh, ow, gain, dy, dx);
  } else {
    size_t total = (size_t)batch * h * w_ * c;
    blur_up_kernel<1><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, h, w_, c, pad0, stride, oh, ow, gain, dy, dx);
  }
  HFAGP_CHECK_LAUNCH("blur_up_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_styles_bwd(int nlayers, int batch, int num_ws, int w_dim, const float* const* aff_w_host,
                                const int32_t* cin_host, const int32_t* widx_host, const float* post_gain_host,
                                const int64_t* off_host, const float* dstyles, float* dws, void* stream) {
  HFAGP_CHECK_ARG(nlayers > 0 && nlayers <= MAX_STYLE_LAYERS_B && batch > 0 && batch <= 65535, "styles_bwd: bad dims");
  HFAGP_CHECK_ARG(aff_w_host && cin_host && widx_host && post_gain_host && off_host && dstyles && dws, "styles_bwd: null pointer");
  StyleBwdTable tb;
  for (int l = 0; l < nlayers; ++l) {
    HFAGP_CHECK_ARG(widx_host[l] >= 0 && widx_host[l] < num_ws, "styles_bwd: ws index out of range");
    tb.aw[l] = aff_w_host[l];
    tb.off[l] = off_host[l];
    tb.cin[l] = cin_host[l];
    tb.widx[l] = widx_host[l];
    tb.gain[l] = post_gain_host[l];
  }
  const int slices = batch >= 8 ? 2 : 8;
  styles_bwd_kernel<<<dim3(nlayers, batch, slices), 256, 0, (cudaStream_t)stream>>>(tb, num_ws, w_dim, 1.0f / sqrtf((float)w_dim),
                                                                          dstyles, dws);
  HFAGP_CHECK_LAUNCH("styles_bwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_demod_bwd(int batch, int cout, int cin, const float* w2, const float* styles, const float* dcoef,
                               const float* ddcoef, float* dstyles, void* stream) {
  HFAGP_CHECK_ARG(w2 && styles && dcoef && ddcoef && dstyles && batch > 0 && batch <= 65535, "demod_bwd: bad args");
  demod_bwd_kernel<<<dim3(cdiv(cin, 32), batch), 256, 0, (cudaStream_t)stream>>>(cout, cin, w2, styles, dcoef, ddcoef, dstyles);
  HFAGP_CHECK_LAUNCH("demod_bwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_linear_bwd(int batch, int cin, int cout, const float* dy, const float* x, const float* w,
                                float w_gain, float b_gain, float* dx, float* dw, float* db, void* stream) {
  HFAGP_CHECK_ARG(dy && w && batch > 0 && cin > 0 && cout > 0, "linear_bwd: bad args");
  HFAGP_CHECK_ARG(!dw || x, "linear_bwd: dw needs x");
  if (dx) {
    // few long rows (the encoder's 8192-wide final map): slice the output channels over more CTAs
    const int slices = ((long long)batch * cin < (long long)device_sm_count() * 256 * 2 && cout >= 64) ? 8 : 1;
    if (slices > 1) HFAGP_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)batch * cin, (cudaStream_t)stream));
    linear_bwd_dx_kernel<<<dim3(cdiv((long long)batch * cin, 256), slices), 256, 0, (cudaStream_t)stream>>>(batch, cin, cout, dy, w, w_gain, dx);
    HFAGP_CHECK_LAUNCH("linear_bwd_dx_kernel");
  }
  if (dw) {
    linear_bwd_dw_kernel<<<cdiv((long long)cout * cin, 256), 256, 0, (cudaStream_t)stream>>>(batch, cin, cout, dy, x, w_gain,
                                                                                          b_gain, dw, db);
    HFAGP_CHECK_LAUNCH("linear_bwd_dw_kernel");
  }
  return HFAGP_OK;
}

namespace hfagp {
// wgrad_tc.cu: the tcgen05 weight gradient (MN-major operands straight from the channels-last activations)
bool wgrad_tc_supported(const HfagpConvDesc& d);
int wgrad_tc_launch(const HfagpConvDesc& d, const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* dz_hi,
                    const uint16_t* dz_lo, const float* xscale, const float* dzscale, float scale, float* dw, int num_sms,
                    cudaStream_t stream);
}  // namespace hfagp

static int wgrad_impl(const HfagpConvDesc* desc, const float* x, const uint16_t* x_hi, const uint16_t* x_lo,
                      const float* dz, const uint16_t* dz_hi, const uint16_t* dz_lo, const float* xscale,
                      const float* dzscale, float scale, float* dw, void* stream) {
  HFAGP_CHECK_ARG(desc && dw, "conv2d_wgrad: null pointer");
  HFAGP_CHECK_ARG((x != nullptr) != (x_hi != nullptr && x_lo != nullptr), "conv2d_wgrad: give x or (x_hi, x_lo)");
  HFAGP_CHECK_ARG((dz != nullptr) != (dz_hi != nullptr && dz_lo != nullptr), "conv2d_wgrad: give dz or (dz_hi, dz_lo)");
  const HfagpConvDesc& d = *desc;
  HFAGP_CHECK_ARG(d.batch > 0 && d.cin > 0 && d.cout > 0 && (d.cin & 3) == 0 && (d.cout & 3) == 0,
                  "conv2d_wgrad: cin and cout must be multiples of 4");
  HFAGP_CHECK_ARG(d.ntaps > 0 && d.ntaps <= HFAGP_MAX_TAPS && (d.in_stride == 1 || d.in_stride == 2), "conv2d_wgrad: bad taps/stride");
  HFAGP_CHECK_ARG(d.out_stride == 1 && d.out_h == d.oh && d.out_w == d.ow, "conv2d_wgrad: dz must be dense [n][oh][ow][cout]");
  if (d.ntaps == 1 && d.cout == 4 && d.in_stride == 1 && d.dy[0] == 0 && d.dx[0] == 0 && d.oh == d.in_h && d.ow == d.in_w &&
      dz && !dzscale && d.batch <= 65535) {
    // narrow 1x1 output (the 3-channel ToRGB, padded to 4): a streaming kernel instead of a 1/16-full tile
    const int hw = d.oh * d.ow;
    int ppb = cdiv((long long)hw * d.batch, (long long)device_sm_count() * 8);
    if (ppb < 64) ppb = 64;
    wgrad_cout4_kernel<<<dim3(cdiv(hw, ppb), d.batch), 256, 0, (cudaStream_t)stream>>>(
        hw, d.cin, ppb, x, reinterpret_cast<const __nv_bfloat16*>(x_hi), reinterpret_cast<const __nv_bfloat16*>(x_lo), dz, xscale,
        scale, dw + (size_t)d.wtap[0] * d.cout * d.cin);
    HFAGP_CHECK_LAUNCH("wgrad_cout4_kernel");
    return HFAGP_OK;
  }
  if (x_hi && dz_hi && wgrad_tc_supported(d)) {
    const int sms = device_sm_count();
    return wgrad_tc_launch(d, x_hi, x_lo, dz_hi, dz_lo, xscale, dzscale, scale, dw, sms, (cudaStream_t)stream);
  }
  WgradParams p;
  p.batch = d.batch; p.in_h = d.in_h; p.in_w = d.in_w; p.cin = d.cin; p.cout = d.cout; p.oh = d.oh; p.ow = d.ow;
  p.in_stride = d.in_stride; p.ntaps = d.ntaps;
  for (int t = 0; t < d.ntaps; ++t) { p.dy[t] = d.dy[t]; p.dx[t] = d.dx[t]; p.wtap[t] = d.wtap[t]; }
  p.x = x; p.x_hi = reinterpret_cast<const __nv_bfloat16*>(x_hi); p.x_lo = reinterpret_cast<const __nv_bfloat16*>(x_lo);
  p.dz = dz; p.dz_hi = reinterpret_cast<const __nv_bfloat16*>(dz_hi); p.dz_lo = reinterpret_cast<const __nv_bfloat16*>(dz_lo);
  p.dw = dw; p.scale = scale;
  p.xscale = xscale; p.dzscale = dzscale;
  const long long K = (long long)d.batch * d.oh * d.ow;
  const int tiles = cdiv(d.cout, WG_T) * cdiv(d.cin, WG_T);
  int splits = cdiv(device_sm_count() * 6, (long long)tiles * d.ntaps);
  const int max_splits = cdiv(K, 4 * WG_K);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.k_per_split = cdiv(cdiv(K, splits), WG_K) * WG_K;
  splits = cdiv(K, p.k_per_split);
  dim3 grid(tiles, d.ntaps, splits);
  wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  HFAGP_CHECK_LAUNCH("wgrad_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_conv2d_wgrad(const HfagpConvDesc* desc, const float* x, const uint16_t* x_hi, const uint16_t* x_lo,
                                  const float* dz, const uint16_t* dz_hi, const uint16_t* dz_lo, float scale, float* dw,
                                  void* stream) {
  return wgrad_impl(desc, x, x_hi, x_lo, dz, dz_hi, dz_lo, nullptr, nullptr, scale, dw, stream);
}

extern "C" int hfagp_conv2d_wgrad_mod(const HfagpConvDesc* desc, const float* x, const uint16_t* x_hi,
                                      const uint16_t* x_lo, const float* dz, const uint16_t* dz_hi, const uint16_t* dz_lo,
                                      const float* xscale, const float* dzscale, float scale, float* dw, void* stream) {
  return wgrad_impl(desc, x, x_hi, x_lo, dz, dz_hi, dz_lo, xscale, dzscale, scale, dw, stream);
}

extern "C" int hfagp_modconv_wgrad_finish(int taps, int cout, int cin, int batch, const float* dw, int dw_transposed, const float* w,
                                          const float* ddcoef, const float* dcoef, const float* styles, float* grad, void* stream) {
  HFAGP_CHECK_ARG(dw && grad && taps > 0 && cout > 0 && cin > 0, "modconv_wgrad_finish: bad args");
  HFAGP_CHECK_ARG(!ddcoef || (w && dcoef && styles && batch > 0), "modconv_wgrad_finish: the demodulation term needs w, dcoef, styles");
  HFAGP_CHECK_ARG(cdiv(cout, 32) <= 65535, "modconv_wgrad_finish: cout too large");
  const dim3 grid(cdiv(cin, 32), cdiv(cout, 32));
  if (dw_transposed)
    modconv_wgrad_finish_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(taps, cout, cin, batch, dw, w, ddcoef, dcoef, styles, grad);
  else
    modconv_wgrad_finish_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(taps, cout, cin, batch, dw, w, ddcoef, dcoef, styles, grad);
  HFAGP_CHECK_LAUNCH("modconv_wgrad_finish_kernel");
  return HFAGP_OK;
}
