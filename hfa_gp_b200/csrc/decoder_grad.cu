// Weight gradient of the OSG decoder MLP (32 -> 64 softplus -> 1 + 32) — the post-tune_iter regime, code/train_rgb.py:132-134.
//
// Inputs are the per-sample operands the renderer's backward wrote (hfagp_render_bwd_dec): f [S][32] mean tri-plane
// features, dout [S][33] gradient of the decoder's raw outputs.  Per 64-sample tile a CTA recomputes the hidden layer in
// fp32 and accumulates the four reductions over samples in REGISTERS across all of its tiles:
//
//   pre = f W0^T + b0 ; h = softplus(pre) ; dh = dout W1 ; dpre = dh * sigmoid(pre)
//   dW1[o][j] += dout[s][o] h[s][j]      db1[o] += dout[s][o]
//   dW0[j][c] += dpre[s][j] f[s][c]      db0[j] += dpre[s][j]
//
// One atomicAdd per gradient element and CTA at the end.  Plain fp32 FMAs (exact-fp32 class, no operand splitting): the
// whole job is 16.6 kFLOP per sample and the operands are read once from HBM (260 B per sample).
// Replaces: autograd of OSGDecoder (eg3d triplane.py) w.r.t. its parameters; round 1/2 used library GEMMs here.
#include "common.cuh"
#include "render_common.cuh"

namespace hfagp {

constexpr int DG_TS = 64;            // samples per tile
constexpr int DG_LD = DG_TS + 4;     // row stride of the [feature][sample] arrays (16 B aligned rows)
constexpr int DG_THREADS = 256;

struct DgSmem {
  float fT[RC][DG_LD];       // features, [channel][sample]
  float fS[DG_TS][RC];       // features, [sample][channel]
  float doT[RO][DG_LD];      // d(raw output), [output][sample]
  float hS[DG_TS][RH];       // hidden activations, [sample][unit]
  float dpT[RH][DG_LD];      // d(pre-activation), [unit][sample]
  float w0T[RC][RH];         // W0^T: [channel][unit]
  float w1[RO][RH];          // W1:   [output][unit]
  float b0[RH];
  float red[RH];             // db0 staging
};

__global__ void __launch_bounds__(DG_THREADS, 2) decoder_wgrad_kernel(long long samples, const float* __restrict__ f,
                                                                     const float* __restrict__ dout,
                                                                     const float* __restrict__ mlp, float* __restrict__ dmlp) {
  extern __shared__ __align__(16) uint8_t dg_raw[];
  DgSmem& sm = *reinterpret_cast<DgSmem*>(dg_raw);
  const int tid = threadIdx.x;
  const float* W0 = mlp;
  const float* B0 = W0 + RH * RC;
  const float* W1 = B0 + RH;
  for (int i = tid; i < RH * RC; i += DG_THREADS) sm.w0T[i & 31][i >> 5] = __ldg(W0 + i);       // W0[j][c]
  for (int i = tid; i < RO * RH; i += DG_THREADS) sm.w1[i >> 6][i & 63] = __ldg(W1 + i);        // W1[o][j]
  if (tid < RH) { sm.b0[tid] = __ldg(B0 + tid); sm.red[tid] = 0.f; }

  // phase A tile: samples 4b..4b+3, hidden units 4a..4a+3
  const int a = tid & 15, b = tid >> 4;
  // dW1 tile: outputs {b, b+16, (b == 0) 32}, units 4a..4a+3 ; dW0 tile: units {jg, jg+32}, channels 4cq..4cq+3
  const int cq = tid & 7, jg = tid >> 3;
  float acc1[3][4], acc0[2][4], accb0[4], accb1 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { acc1[0][k] = acc1[1][k] = acc1[2][k] = 0.f; acc0[0][k] = acc0[1][k] = 0.f; accb0[k] = 0.f; }

  const long long tiles = (samples + DG_TS - 1) / DG_TS;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long s0 = tile * DG_TS;
    const int ns = (int)(samples - s0 < DG_TS ? samples - s0 : DG_TS);
    __syncthreads();                                   // previous tile fully consumed (also orders the weight staging)
    // ---- stage the operands (zero tail: dout = 0 contributes nothing)
    for (int i = tid; i < DG_TS * (RC / 4); i += DG_THREADS) {
      const int s = i >> 3, c4 = i & 7;
      const float4 v = s < ns ? __ldg(reinterpret_cast<const float4*>(f + (s0 + s) * RC) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&sm.fS[s][c4 * 4]) = v;
      sm.fT[c4 * 4 + 0][s] = v.x; sm.fT[c4 * 4 + 1][s] = v.y; sm.fT[c4 * 4 + 2][s] = v.z; sm.fT[c4 * 4 + 3][s] = v.w;
    }
    for (int i = tid; i < DG_TS * RO; i += DG_THREADS) {
      const int s = i / RO, o = i - s * RO;
      sm.doT[o][s] = s < ns ? __ldg(dout + s0 * RO + i) : 0.f;
    }
    __syncthreads();
    if (tid < RO) {                                    // db1
      float t = 0.f;
#pragma unroll 8
      for (int s = 0; s < DG_TS; ++s) t += sm.doT[tid][s];
      accb1 += t;
    }
    // ---- phase A: 4 samples x 4 units per thread
    float pre[4][4], dh[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) { pre[i][k] = sm.b0[4 * a + k]; dh[i][k] = 0.f; }
#pragma unroll 8
    for (int c = 0; c < RC; ++c) {
      const float4 fv = *reinterpret_cast<const float4*>(&sm.fT[c][4 * b]);
      const float4 wv = *reinterpret_cast<const float4*>(&sm.w0T[c][4 * a]);
      const float fa[4] = {fv.x, fv.y, fv.z, fv.w}, wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) pre[i][k] = fmaf(fa[i], wa[k], pre[i][k]);
    }
#pragma unroll 3
    for (int o = 0; o < RO; ++o) {
      const float4 dv = *reinterpret_cast<const float4*>(&sm.doT[o][4 * b]);
      const float4 wv = *reinterpret_cast<const float4*>(&sm.w1[o][4 * a]);
      const float da[4] = {dv.x, dv.y, dv.z, dv.w}, wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) dh[i][k] = fmaf(da[i], wa[k], dh[i][k]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float hv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x = pre[i][k];
        const float e = __expf(-fabsf(x));                           // softplus = max(x,0) + log1p(e), sigmoid from the same e
        hv[k] = fmaxf(x, 0.f) + __logf(1.f + e);
        const float r = 1.f / (1.f + e);
        const float sg = x >= 0.f ? r : e * r;
        const float dp = dh[i][k] * sg;
        sm.dpT[4 * a + k][4 * b + i] = dp;
        accb0[k] += dp;
      }
      *reinterpret_cast<float4*>(&sm.hS[4 * b + i][4 * a]) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
    __syncthreads();
    // ---- dW1[o][j] += sum_s dout[s][o] h[s][j]
#pragma unroll 2
    for (int s = 0; s < DG_TS; s += 4) {
      float4 hv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) hv[i] = *reinterpret_cast<const float4*>(&sm.hS[s + i][4 * a]);
      const float4 d0 = *reinterpret_cast<const float4*>(&sm.doT[b][s]);
      const float4 d1 = *reinterpret_cast<const float4*>(&sm.doT[b + 16][s]);
      const float4 d2 = b == 0 ? *reinterpret_cast<const float4*>(&sm.doT[32][s]) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float dd[3][4] = {{d0.x, d0.y, d0.z, d0.w}, {d1.x, d1.y, d1.z, d1.w}, {d2.x, d2.y, d2.z, d2.w}};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float hh[4] = {hv[i].x, hv[i].y, hv[i].z, hv[i].w};
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc1[q][k] = fmaf(dd[q][i], hh[k], acc1[q][k]);
      }
    }
    // ---- dW0[j][c] += sum_s dpre[s][j] f[s][c]
#pragma unroll 2
    for (int s = 0; s < DG_TS; s += 4) {
      float4 fv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) fv[i] = *reinterpret_cast<const float4*>(&sm.fS[s + i][4 * cq]);
      const float4 p0 = *reinterpret_cast<const float4*>(&sm.dpT[jg][s]);
      const float4 p1 = *reinterpret_cast<const float4*>(&sm.dpT[jg + 32][s]);
      const float pp[2][4] = {{p0.x, p0.y, p0.z, p0.w}, {p1.x, p1.y, p1.z, p1.w}};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float ff[4] = {fv[i].x, fv[i].y, fv[i].z, fv[i].w};
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc0[q][k] = fmaf(pp[q][i], ff[k], acc0[q][k]);
      }
    }
  }

  // ---- one atomic per element and CTA (packing of the mlp buffer: W0 [64][32], b0 [64], W1 [33][64], b1 [33])
  float* dW0 = dmlp;
  float* dB0 = dW0 + RH * RC;
  float* dW1 = dB0 + RH;
  float* dB1 = dW1 + RO * RH;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(dW0 + (jg + 32 * q) * RC + 4 * cq + k, acc0[q][k]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    atomicAdd(dW1 + b * RH + 4 * a + k, acc1[0][k]);
    atomicAdd(dW1 + (b + 16) * RH + 4 * a + k, acc1[1][k]);
    if (b == 0) atomicAdd(dW1 + 32 * RH + 4 * a + k, acc1[2][k]);
    atomicAdd(&sm.red[4 * a + k], accb0[k]);           // 16 sample groups per unit
  }
  if (tid < RO) atomicAdd(dB1 + tid, accb1);
  __syncthreads();
  if (tid < RH) atomicAdd(dB0 + tid, sm.red[tid]);
}

}  // namespace hfagp

using namespace hfagp;

extern "C" int hfagp_decoder_wgrad(long long samples, const float* f, const float* dout, const float* mlp, float* dmlp,
                                   void* stream) {
  HFAGP_CHECK_ARG(f && dout && mlp && dmlp && samples > 0, "decoder_wgrad: null pointer or no samples");
  HFAGP_CHECK_ARG((reinterpret_cast<uintptr_t>(f) & 15) == 0, "decoder_wgrad: dump_f must be 16-byte aligned");
  const int sms = device_sm_count();
  static std::atomic<uint64_t> attr_done{0};
  HFAGP_CUDA(per_device_once(attr_done, [] { return cudaFuncSetAttribute(decoder_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DgSmem)); }));
  const long long tiles = (samples + DG_TS - 1) / DG_TS;
  const int blocks = (int)(tiles < 2ll * sms ? tiles : 2ll * sms);
  decoder_wgrad_kernel<<<blocks, DG_THREADS, sizeof(DgSmem), (cudaStream_t)stream>>>(samples, f, dout, mlp, dmlp);
  HFAGP_CHECK_LAUNCH("decoder_wgrad_kernel");
  return HFAGP_OK;
}
