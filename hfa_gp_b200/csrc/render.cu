// Tri-plane volume renderer (forward).  One warp owns one ray end to end; a CTA is a vertical strip of 16 rays
// (one persistent CTA per SM looping over strips).
//
//   ray generation -> stratified depths -> [ tri-plane bilinear gather -> OSG decoder MLP ] (coarse)
//   -> mid-point march -> smoothed pdf / cdf / searchsorted -> fine depths -> [ gather -> MLP ] (fine)
//   -> stable rank-sort merge -> final mid-point march -> composite (lane = channel)
//
// Gather: planes are channels-last, so one bilinear tap of one plane is one 128 B line.  Four samples are
// gathered per pass with lane = (sample, channel quad): each lane issues 12 independent 16 B loads (8 lanes cover
// a line) and no cross-lane reduction is needed.  The rays of a strip share their x pixel coordinate, so at equal
// sample index they hit (nearly) the same texels of planes 1 and 2 (both are functions of world x and z only);
// along a ray, plane 0 moves < 1 texel per sample.  The planes themselves (25 MB) stay L2-resident.
//
// Decoder MLP (32 -> 64 softplus -> 1+32) runs on the tensor pipe per 16-sample tile with warp-level
// mma.sync.m16n8k16 (bf16 operands, fp32 accumulate): rays have 48 samples = 3 x 16 rows, so the 16-row MMA
// keeps every warp autonomous (no CTA-wide barrier in the ray loop; tcgen05's 128-row tile would couple 2.7 rays
// and its share of the kernel is ~5 % of issue slots either way).  fp32-class accuracy comes from the same
// split-bf16 scheme as the convolutions (hi*hi + lo*hi + hi*lo); the layer-1 accumulator fragments of two
// n-tiles are, after softplus, exactly the layer-2 A fragment of one k-step, so the hidden layer never leaves
// registers.
#include <cuda_bf16.h>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include "common.cuh"
#include "render_common.cuh"

namespace hfagp {

// BWD = false: the forward renderer.  BWD = true: recompute the forward per ray (same code), then back-propagate
// d(feat) through the composite / march / decoder MLP / bilinear gather into d(planes) (8 rays per CTA).
template <bool BWD>
__global__ void __launch_bounds__(BWD ? 384 : R_WARPS * 32, 1) render_kernel(const RenderParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const HfagpRenderDesc& d = p.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int S = d.s_coarse, SF = d.s_fine, T = S + SF;

  // ---- CTA-shared decoder weights as MMA B fragments (split bf16)
  uint2* w0f = reinterpret_cast<uint2*>(smem_raw);
  uint2* w1f = w0f + W0F_U2;
  float* b0s = reinterpret_cast<float*>(w1f + W1F_U2);
  float* b1s = b0s + RH;     // permuted: [0..31] colour biases, [32] sigma bias, rest 0
  {
    const float* W0 = p.mlp;
    const float* B0 = W0 + RH * RC;
    const float* W1 = B0 + RH;
    const float* B1 = W1 + RO * RH;
    for (int i = threadIdx.x; i < 8 * 2 * 32; i += blockDim.x) {       // (ntile j, kstep s, lane)
      const int l = i & 31, s = (i >> 5) & 1, j = i >> 6;
      const int n = 8 * j + (l >> 2), k0 = 16 * s + 2 * (l & 3);
      const float* r = W0 + n * RC + k0;
      uint2 hi, lo;
      split_pair(__ldg(r), __ldg(r + 1), hi.x, lo.x);
      split_pair(__ldg(r + 8), __ldg(r + 9), hi.y, lo.y);
      w0f[((j * 2 + s) * 2 + 0) * 32 + l] = hi;
      w0f[((j * 2 + s) * 2 + 1) * 32 + l] = lo;
    }
    for (int i = threadIdx.x; i < 5 * 4 * 32; i += blockDim.x) {
      const int l = i & 31, s = (i >> 5) & 3, j = i >> 7;
      const int np = 8 * j + (l >> 2), k0 = 16 * s + 2 * (l & 3);
      const int o = np < 32 ? np + 1 : (np == 32 ? 0 : -1);              // output row permutation: colours first
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (o >= 0) {
        const float* r = W1 + o * RH + k0;
        v[0] = __ldg(r); v[1] = __ldg(r + 1); v[2] = __ldg(r + 8); v[3] = __ldg(r + 9);
      }
      uint2 hi, lo;
      split_pair(v[0], v[1], hi.x, lo.x);
      split_pair(v[2], v[3], hi.y, lo.y);
      w1f[((j * 4 + s) * 2 + 0) * 32 + l] = hi;
      w1f[((j * 4 + s) * 2 + 1) * 32 + l] = lo;
    }
    for (int i = threadIdx.x; i < RH; i += blockDim.x) b0s[i] = __ldg(B0 + i);
    for (int i = threadIdx.x; i < 40; i += blockDim.x) b1s[i] = i < 32 ? __ldg(B1 + i + 1) : (i == 32 ? __ldg(B1) : 0.f);
    if constexpr (BWD) {
      uint2* w1t = reinterpret_cast<uint2*>(smem_raw + ((WEIGHT_BYTES + 15) & ~15));
      uint2* w0t = w1t + W1T_U2;
      // dh[s][j] = sum_op dout[s][op] W1[perm(op)][j]:  B(k = op, n = j)
      for (int i = threadIdx.x; i < 8 * 3 * 32; i += blockDim.x) {
        const int l = i & 31, ks = (i >> 5) % 3, jn = i / 96;
        const int nn = 8 * jn + (l >> 2), k0 = 16 * ks + 2 * (l & 3);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int op = k0 + (e & 1) + (e >> 1) * 8;
          const int o = op < 32 ? op + 1 : (op == 32 ? 0 : -1);
          v[e] = o >= 0 ? __ldg(W1 + o * RH + nn) : 0.f;
        }
        uint2 hi, lo;
        split_pair(v[0], v[1], hi.x, lo.x);
        split_pair(v[2], v[3], hi.y, lo.y);
        w1t[((jn * 3 + ks) * 2 + 0) * 32 + l] = hi;
        w1t[((jn * 3 + ks) * 2 + 1) * 32 + l] = lo;
      }
      // df[s][c] = sum_j dpre[s][j] W0[j][c]:  B(k = j, n = c)
      for (int i = threadIdx.x; i < 4 * 4 * 32; i += blockDim.x) {
        const int l = i & 31, ks = (i >> 5) & 3, cn = i >> 7;
        const int nn = 8 * cn + (l >> 2), k0 = 16 * ks + 2 * (l & 3);
        uint2 hi, lo;
        split_pair(__ldg(W0 + k0 * RC + nn), __ldg(W0 + (k0 + 1) * RC + nn), hi.x, lo.x);
        split_pair(__ldg(W0 + (k0 + 8) * RC + nn), __ldg(W0 + (k0 + 9) * RC + nn), hi.y, lo.y);
        w0t[((cn * 4 + ks) * 2 + 0) * 32 + l] = hi;
        w0t[((cn * 4 + ks) * 2 + 1) * 32 + l] = lo;
      }
    }
  }
  __syncthreads();

  // ---- per-warp scratch
  const size_t wbytes = render_warp_bytes(S, SF, BWD);
  uint8_t* wbase = smem_raw + (BWD ? WEIGHT_BYTES_BWD : ((WEIGHT_BYTES + 15) & ~15)) + (size_t)warp * wbytes;
  const int S16 = round16(S);
  const int crows = S16 + round16(SF);
  const int Tp = (T + 3) & ~3;
  uint32_t* colq = reinterpret_cast<uint32_t*>(wbase);                  // [crows][16] words of 2 x u16
  float* ftile = reinterpret_cast<float*>(colq + (size_t)crows * 16);   // [16][32]
  float* dep = ftile + TILE * RC;
  float* sig = dep + Tp;
  float* sdep = sig + Tp;
  float* ssig = sdep + Tp;
  uint8_t* order = reinterpret_cast<uint8_t*>(ssig + Tp);
  float* cdf = reinterpret_cast<float*>(order + Tp);
  float* zmid = cdf + (S + 2);
  float* tarr = zmid + S;          // backward only (5 * Tp + 32 floats)
  float* aarr = tarr + Tp;
  float* dsg = aarr + Tp;
  float* pj = dsg + Tp;
  float* wts_b = pj + Tp;
  float* dfs = wts_b + Tp;
  Tap* taps = reinterpret_cast<Tap*>(wbase + wbytes - TILE * 12 * sizeof(Tap));

  const int PW = d.plane_w, PH = d.plane_h;
  const int texel_stride = 3 * RC;
  const int res = d.res;
  const int nwarps = blockDim.x >> 5;            // rays per strip
  const int ytiles = (res + nwarps - 1) / nwarps;
  const long long strips = (long long)d.batch * ytiles * res;

  for (long long strip = blockIdx.x; strip < strips; strip += gridDim.x) {
    const int n = (int)(strip / ((long long)ytiles * res));
    const int rem = (int)(strip - (long long)n * ytiles * res);
    const int yt = rem / res, px = rem - yt * res;
    const int py = yt * nwarps + warp;
    if (py >= res) continue;                       // warp-uniform; no CTA barrier inside the loop
    const long long ray = ((long long)n * res + py) * res + px;
    const float* cam = p.cam + (size_t)n * 25;
    const float* pl = p.planes + (size_t)n * PH * PW * texel_stride;

    // ---- ray generation (uniform across the warp)
    float ox_, oy_, oz_, dx_, dy_, dz_;
    {
      const float inv = 1.0f / res, half = 0.5f / res;
      const float xc = px * inv + half, yc = py * inv + half;
      const float fx = __ldg(cam + 16), sk = __ldg(cam + 17), cx = __ldg(cam + 18);
      const float fy = __ldg(cam + 20), cy = __ldg(cam + 21);
      const float xl = (xc - cx + cy * sk / fy - sk * yc / fy) / fx;
      const float yl = (yc - cy) / fy;
      float m[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) m[i] = __ldg(cam + i);
      ox_ = m[3]; oy_ = m[7]; oz_ = m[11];
      float wx = m[0] * xl + m[1] * yl + m[2] + m[3];
      float wy = m[4] * xl + m[5] * yl + m[6] + m[7];
      float wz = m[8] * xl + m[9] * yl + m[10] + m[11];
      float vx = wx - ox_, vy = wy - oy_, vz = wz - oz_;
      float nrm = fmaxf(sqrtf(vx * vx + vy * vy + vz * vz), 1e-12f);
      dx_ = vx / nrm; dy_ = vy / nrm; dz_ = vz / nrm;
    }

    // ---- coarse depths
    for (int s = lane; s < S; s += 32)
      dep[s] = __ldg(p.lin + s) + __ldg(p.jitter + (size_t)ray * S + s) * d.delta;
    __syncwarp();

    // ---- gather the features of samples [base, base+cnt) (depths in dep[]) into ftile rows 0..cnt-1
    auto gather_tile = [&](int base, int cnt) {
        // phase A: lanes 0-15 build the taps of planes 0 and 1, lanes 16-31 those of plane 2
        {
          const int sl = lane & 15;
          if (sl < cnt) {
            const float t = dep[base + sl];
            const float qx = (ox_ + t * dx_) * d.box_scale;
            const float qy = (oy_ + t * dy_) * d.box_scale;
            const float qz = (oz_ + t * dz_) * d.box_scale;
            const int p_begin = lane < 16 ? 0 : 2, p_end = lane < 16 ? 2 : 3;
            for (int pidx = p_begin; pidx < p_end; ++pidx) {
              const float gx = pidx == 2 ? qz : qx;
              const float gy = pidx == 0 ? qy : (pidx == 1 ? qz : qx);
              const float ix = ((gx + 1.f) * PW - 1.f) * 0.5f;
              const float iy = ((gy + 1.f) * PH - 1.f) * 0.5f;
              const float fx0 = floorf(ix), fy0 = floorf(iy);
              const float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
              const float wl = fx1 - ix, wr = ix - fx0, wt = fy1 - iy, wb = iy - fy0;
              // clamp before the int conversion so far-away samples cannot overflow
              const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)PW + 1.f);
              const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)PH + 1.f);
              const int x1 = x0 + 1, y1 = y0 + 1;
              const bool vx0 = x0 >= 0 && x0 < PW, vx1 = x1 >= 0 && x1 < PW;
              const bool vy0 = y0 >= 0 && y0 < PH, vy1 = y1 >= 0 && y1 < PH;
              Tap* tp = taps + (sl * 12 + pidx * 4);
              const int cbase = pidx * RC;
              const int b00 = ((y0 * PW + x0) * texel_stride + cbase) * 4, dxb = texel_stride * 4, dyb = PW * texel_stride * 4;
              tp[0] = (vx0 && vy0) ? Tap{(uint32_t)b00, wl * wt} : Tap{0u, 0.f};
              tp[1] = (vx1 && vy0) ? Tap{(uint32_t)(b00 + dxb), wr * wt} : Tap{0u, 0.f};
              tp[2] = (vx0 && vy1) ? Tap{(uint32_t)(b00 + dyb), wl * wb} : Tap{0u, 0.f};
              tp[3] = (vx1 && vy1) ? Tap{(uint32_t)(b00 + dyb + dxb), wr * wb} : Tap{0u, 0.f};
            }
          }
        }
        __syncwarp();
        // phase B: four samples per pass, lane = (sample sq, channel quad cg): 12 unconditional 16 B loads per lane
        // (8 lanes cover one 128 B texel line; outside taps carry weight 0), no cross-lane reduction needed
        {
          const int sq = lane >> 3, cg = lane & 7;
          const char* lb = reinterpret_cast<const char*>(pl) + cg * 16;
#pragma unroll 1
          for (int q = 0; q < TILE; q += 4) {
            if (q >= cnt) break;
            const int sl = min(q + sq, cnt - 1);       // lanes past the end redo the last sample (not stored)
            const Tap* tp = taps + sl * 12;
            Tap rec[12];
            float4 v[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
              rec[k] = tp[k];
              v[k] = __ldg(reinterpret_cast<const float4*>(lb + rec[k].off));
            }
            float4 a[3];
#pragma unroll
            for (int pi = 0; pi < 3; ++pi) {
              a[pi] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int k = 4 * pi; k < 4 * pi + 4; ++k) {
                a[pi].x = fmaf(rec[k].w, v[k].x, a[pi].x);
                a[pi].y = fmaf(rec[k].w, v[k].y, a[pi].y);
                a[pi].z = fmaf(rec[k].w, v[k].z, a[pi].z);
                a[pi].w = fmaf(rec[k].w, v[k].w, a[pi].w);
              }
            }
            const float third = 1.f / 3.f;
            if (q + sq < cnt) {
              const int r = q + sq;
              *reinterpret_cast<float4*>(ftile + r * RC + colx(r, 4 * cg)) =
                  make_float4((a[0].x + a[1].x + a[2].x) * third, (a[0].y + a[1].y + a[2].y) * third,
                              (a[0].z + a[1].z + a[2].z) * third, (a[0].w + a[1].w + a[2].w) * third);
            }
          }
        }
        __syncwarp();
    };
    // feature tile -> split-bf16 MMA A fragments of layer 1 (rows g, g+8; k-steps = channels 0-15, 16-31)
    auto feature_frags = [&](uint32_t (&ah)[2][4], uint32_t (&al)[2][4]) {
      const int r0 = g, r1 = g + 8;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int c0 = 16 * s + 2 * t4;
        const float2 f00 = *reinterpret_cast<const float2*>(ftile + r0 * RC + colx(r0, c0));
        const float2 f10 = *reinterpret_cast<const float2*>(ftile + r1 * RC + colx(r1, c0));
        const float2 f01 = *reinterpret_cast<const float2*>(ftile + r0 * RC + colx(r0, c0 + 8));
        const float2 f11 = *reinterpret_cast<const float2*>(ftile + r1 * RC + colx(r1, c0 + 8));
        split_pair(f00.x, f00.y, ah[s][0], al[s][0]);
        split_pair(f10.x, f10.y, ah[s][1], al[s][1]);
        split_pair(f01.x, f01.y, ah[s][2], al[s][2]);
        split_pair(f11.x, f11.y, ah[s][3], al[s][3]);
      }
    };

    // ---- gather + decode the samples [s_begin, s_end) whose depths are in dep[], 16 at a time
    auto shade = [&](int s_begin, int s_end, int crow_begin) {
      for (int base = s_begin; base < s_end; base += TILE) {
        const int crow0 = crow_begin + (base - s_begin);          // colour-buffer row of the tile's first sample
        const int cnt = min(TILE, s_end - base);
        gather_tile(base, cnt);
        // decoder MLP on the 16-row tile [base, base+16) with mma.sync (rows beyond cnt are don't-care)
        uint32_t ah[2][4], al[2][4];
        feature_frags(ah, al);
        // layer 1 two n-tiles at a time: after softplus their accumulator fragments are exactly the A fragment of
        // layer-2 k-step s, so the hidden layer never leaves registers (and only 8 of its 32 values are live)
        float o2[5][4];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const float2 bb = *reinterpret_cast<const float2*>(b1s + 8 * j + 2 * t4);
          o2[j][0] = bb.x; o2[j][1] = bb.y; o2[j][2] = bb.x; o2[j][3] = bb.y;
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          float h[2][4];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * s + jj;
            const float2 bb = *reinterpret_cast<const float2*>(b0s + 8 * j + 2 * t4);
            h[jj][0] = bb.x; h[jj][1] = bb.y; h[jj][2] = bb.x; h[jj][3] = bb.y;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint2 bh = w0f[((j * 2 + ks) * 2 + 0) * 32 + lane];
              const uint2 bl = w0f[((j * 2 + ks) * 2 + 1) * 32 + lane];
              mma_bf16(h[jj], ah[ks], bh);
              mma_bf16(h[jj], al[ks], bh);
              mma_bf16(h[jj], ah[ks], bl);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) h[jj][e] = softplus_fast(h[jj][e]);
          }
          uint32_t a2h[4], a2l[4];
          split_pair(h[0][0], h[0][1], a2h[0], a2l[0]);
          split_pair(h[0][2], h[0][3], a2h[1], a2l[1]);
          split_pair(h[1][0], h[1][1], a2h[2], a2l[2]);
          split_pair(h[1][2], h[1][3], a2h[3], a2l[3]);
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const uint2 bh = w1f[((j * 4 + s) * 2 + 0) * 32 + lane];
            const uint2 bl = w1f[((j * 4 + s) * 2 + 1) * 32 + lane];
            mma_bf16(o2[j], a2h, bh);
            mma_bf16(o2[j], a2l, bh);
            mma_bf16(o2[j], a2h, bl);
          }
        }
        // colours (n-tiles 0..3 = channels 8j+2t, +1) into the 16-bit colour rows; sigma = column 32 (n-tile 4, t == 0)
        {
          const int q0 = crow0 + g, q1 = crow0 + g + 8;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cw = 4 * j + t4;
            const uint32_t a = __float2uint_rn(sigmoid01(o2[j][0]) * QSCALE), b = __float2uint_rn(sigmoid01(o2[j][1]) * QSCALE);
            const uint32_t c = __float2uint_rn(sigmoid01(o2[j][2]) * QSCALE), e = __float2uint_rn(sigmoid01(o2[j][3]) * QSCALE);
            colq[q0 * 16 + colqx(q0, cw)] = a | (b << 16);
            colq[q1 * 16 + colqx(q1, cw)] = c | (e << 16);
          }
        }
        if (t4 == 0) {
          if (g < cnt) sig[base + g] = o2[4][0];
          if (g + 8 < cnt) sig[base + g + 8] = o2[4][2];
        }
        __syncwarp();   // the feature tile and tap records are reused by the next tile
      }
    };

    shade(0, S, 0);

    if (SF > 0) {
      // ---- coarse march (weights only) -> smoothed pdf -> inverse-CDF fine depths
      float* wts = sdep;                           // sdep is not live until the sort
      march_weights(S - 1, lane, wts, [&](int k) {
        float sm = softplus_t(0.5f * (sig[k] + sig[k + 1]) - 1.f);
        return 1.f - expf(-(sm * (dep[k + 1] - dep[k])));
      });
      __syncwarp();
      const int NB = S - 3;  // number of pdf bins actually used (upstream: weights[:, 1:-1])
      // a[j] = 0.5*(max(w[j-1],w[j]) + max(w[j],w[j+1])) + 0.01 for j in [0, S-2]; bins use j = 1..S-3
      float psum = 0.f;
      for (int j = lane; j < NB; j += 32) {
        const int jj = j + 1;
        float m0 = fmaxf(wts[jj - 1], wts[jj]);
        float m1 = jj + 1 <= S - 2 ? fmaxf(wts[jj], wts[jj + 1]) : wts[jj];
        float a = 0.5f * (m0 + m1) + 0.01f + 1e-5f;
        cdf[1 + j] = a;
        psum += a;
      }
      for (int j = lane; j < S - 1; j += 32) zmid[j] = 0.5f * (dep[j] + dep[j + 1]);
      psum = warp_sum(psum);
      __syncwarp();
      // inclusive cumsum of pdf = a / psum into cdf[1..NB], cdf[0] = 0
      {
        float carry = 0.f;
        for (int base = 0; base < NB; base += 32) {
          int j = base + lane;
          float v = j < NB ? cdf[1 + j] / psum : 0.f;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            float t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
          }
          if (j < NB) cdf[1 + j] = carry + v;
          carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) cdf[0] = 0.f;
      }
      __syncwarp();
      for (int k = lane; k < SF; k += 32) {
        const float u = __ldg(p.u_fine + (size_t)ray * SF + k);
        // searchsorted(cdf[0..NB], u, right=True) = #{cdf[i] <= u}; cdf is non-decreasing (running sum of positives)
        const int ind = searchsorted_right(cdf, NB + 1, u);
        const int lo = max(ind - 1, 0), hi = min(ind, NB);
        const float c0 = cdf[lo], c1 = cdf[hi];
        float den = c1 - c0;
        if (den < 1e-5f) den = 1.f;
        const float z0 = zmid[lo], z1 = zmid[hi];
        dep[S + k] = z0 + (u - c0) / den * (z1 - z0);
        if (p.inds) {
          p.inds[(size_t)ray * SF + k] = ind;
          p.below[(size_t)ray * SF + k] = lo;
          p.above[(size_t)ray * SF + k] = hi;
        }
      }
      __syncwarp();
      shade(S, T, S16);
      // ---- stable rank sort of the T depths (coarse first, as torch.cat + sort sees them).  Fast path counts
      // strictly-smaller depths for the lane's (up to 4) elements against one broadcast read of each depth;
      // if any two depths are equal the ranks no longer sum to T(T-1)/2 and the exact tie-aware path redoes it.
      auto rank_sort = [&](auto ne_tag) {
        constexpr int NE = decltype(ne_tag)::value;       // elements per lane = ceil(T / 32)
        float de[NE];
        int rk[NE];
        stable_ranks<NE>(dep, T, lane, de, rk, S);
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int i = lane + 32 * e;
          if (i < T) {
            order[rk[e]] = (uint8_t)i;
            sdep[rk[e]] = de[e];
            ssig[rk[e]] = sig[i];
          }
        }
      };
      if (T <= 32) rank_sort(std::integral_constant<int, 1>{});
      else if (T <= 64) rank_sort(std::integral_constant<int, 2>{});
      else if (T <= 96) rank_sort(std::integral_constant<int, 3>{});
      else rank_sort(std::integral_constant<int, 4>{});
    } else {
      for (int i = lane; i < T; i += 32) {
        order[i] = (uint8_t)i;
        sdep[i] = dep[i];
        ssig[i] = sig[i];
      }
    }
    __syncwarp();

    // ---- final march
    float* wts = BWD ? wts_b : dep;                // forward: unsorted depths are dead after the sort; backward re-gathers
    const float wtot = march_weights(T - 1, lane, wts, [&](int k) {
      float sm = softplus_t(0.5f * (ssig[k] + ssig[k + 1]) - 1.f);
      return 1.f - expf(-(sm * (sdep[k + 1] - sdep[k])));
    }, BWD ? tarr : nullptr, BWD ? aarr : nullptr);
    __syncwarp();
    if constexpr (BWD) {
      // ================= backward of this ray =================
      // feat = 2 rgb - 1, rgb = sum_k w_k (c_k + c_{k+1})/2, w_k = alpha_k T_k, T_k = prod_{m<k} (1 - alpha_m + 1e-10),
      // alpha_k = 1 - exp(-softplus(sigma_mid_k - 1) delta_k).  Sample positions carry no gradient (upstream
      // computes the importance samples under no_grad); depth and weight-sum outputs are not differentiated.
      const int shift = S16 - S;
      auto crow_of = [&](int j) { return j < S ? j : j + shift; };
      dfs[lane] = 2.f * __ldg(p.dfeat + (size_t)ray * RC + lane);
      // omega_j (colour weights, storage order) -> sig[]
      for (int k = lane; k < T; k += 32)
        sig[order[k]] = 0.5f * ((k > 0 ? wts[k - 1] : 0.f) + (k < T - 1 ? wts[k] : 0.f));
      __syncwarp();
      // P_j = dfeat2 . colour_j  (lane = sample)
      for (int j = lane; j < T; j += 32) {
        const int cr = crow_of(j);
        float acc = 0.f;
#pragma unroll 4
        for (int cw = 0; cw < 16; ++cw) {
          const uint32_t wq = colq[cr * 16 + colqx(cr, cw)];
          acc = fmaf(dfs[2 * cw], fmaf((float)(wq & 0xffffu), QSTEP, -0.001f), acc);
          acc = fmaf(dfs[2 * cw + 1], fmaf((float)(wq >> 16), QSTEP, -0.001f), acc);
        }
        pj[j] = acc;
      }
      __syncwarp();
      // d sigma_mid per interval (lane = interval), suffix sums R_k = sum_{m>k} G_m w_m by a forward scan
      float total = 0.f;
      for (int k = lane; k < T - 1; k += 32) total = fmaf(0.5f * (pj[order[k]] + pj[order[k + 1]]), wts[k], total);
      total = warp_sum(total);
      float carry = 0.f;
      for (int base = 0; base < T - 1; base += 32) {
        const int k = base + lane;
        const bool ok = k < T - 1;
        const float G = ok ? 0.5f * (pj[order[k]] + pj[order[k + 1]]) : 0.f;
        float v = ok ? G * wts[k] : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float t = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v += t;
        }
        const float R = total - (carry + v);
        carry += __shfl_sync(0xffffffffu, v, 31);
        if (ok) {
          const float a = aarr[k];
          const float dalpha = G * tarr[k] - R / (1.f - a + 1e-10f);
          const float dsp = dalpha * (sdep[k + 1] - sdep[k]) * (1.f - a);
          const float xm = 0.5f * (ssig[k] + ssig[k + 1]) - 1.f;
          tarr[k] = dsp / (1.f + expf(-xm));              // d sigma_mid_k (softplus' = sigmoid); overwrites T_k
        }
      }
      __syncwarp();
      for (int k = lane; k < T; k += 32)
        dsg[order[k]] = 0.5f * ((k > 0 ? tarr[k - 1] : 0.f) + (k < T - 1 ? tarr[k] : 0.f));
      __syncwarp();

      const uint2* w1t = reinterpret_cast<const uint2*>(smem_raw + ((WEIGHT_BYTES + 15) & ~15));
      const uint2* w0t = w1t + W1T_U2;
      float dfv[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        dfv[ks][0] = dfs[16 * ks + 2 * t4]; dfv[ks][1] = dfs[16 * ks + 2 * t4 + 1];
        dfv[ks][2] = dfs[16 * ks + 2 * t4 + 8]; dfv[ks][3] = dfs[16 * ks + 2 * t4 + 9];
      }
      float* dpl = p.dplanes + (size_t)n * PH * PW * texel_stride;

      auto bwd_tiles = [&](int s_begin, int s_end, int crow_begin) {
        for (int base = s_begin; base < s_end; base += TILE) {
          const int crow0 = crow_begin + (base - s_begin);
          const int cnt = min(TILE, s_end - base);
          gather_tile(base, cnt);
          if (p.dump_f) {
            // lane = channel: row r of the tile is sample base + r of this ray
            float* fd = p.dump_f + ((size_t)ray * T + base) * RC;
            for (int r = 0; r < cnt; ++r) fd[(size_t)r * RC + lane] = ftile[r * RC + colx(r, lane)];
          }
          uint32_t ah[2][4], al[2][4];
          feature_frags(ah, al);
          // layer 1 forward again: h = softplus(pre), all 8 n-tiles stay in registers
          float h[8][4];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 bb = *reinterpret_cast<const float2*>(b0s + 8 * j + 2 * t4);
            h[j][0] = bb.x; h[j][1] = bb.y; h[j][2] = bb.x; h[j][3] = bb.y;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint2 bh = w0f[((j * 2 + ks) * 2 + 0) * 32 + lane];
              const uint2 bl = w0f[((j * 2 + ks) * 2 + 1) * 32 + lane];
              mma_bf16(h[j], ah[ks], bh);
              mma_bf16(h[j], al[ks], bh);
              mma_bf16(h[j], ah[ks], bl);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) h[j][e] = softplus_fast(h[j][e]);
          }
          // d(out) rows of this thread: samples base+g and base+g+8 (zero beyond cnt)
          const bool v0 = g < cnt, v1 = g + 8 < cnt;
          const float om0 = v0 ? sig[base + g] : 0.f, om1 = v1 ? sig[base + g + 8] : 0.f;
          const float ds0 = v0 ? dsg[base + g] : 0.f, ds1 = v1 ? dsg[base + g + 8] : 0.f;
          const int q0 = crow0 + g, q1 = crow0 + g + 8;
          float dh[8][4];
#pragma unroll
          for (int j = 0; j < 8; ++j) { dh[j][0] = dh[j][1] = dh[j][2] = dh[j][3] = 0.f; }
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            float av[8];
            if (ks < 2) {
              // colour columns: d(out) = omega * dfeat2 * 1.002 * s (1 - s), s = q / 65535
              const uint32_t w00 = colq[q0 * 16 + colqx(q0, 8 * ks + t4)], w10 = colq[q1 * 16 + colqx(q1, 8 * ks + t4)];
              const uint32_t w01 = colq[q0 * 16 + colqx(q0, 8 * ks + t4 + 4)], w11 = colq[q1 * 16 + colqx(q1, 8 * ks + t4 + 4)];
              auto dsg_ = [](uint32_t q16) { const float sg = (float)q16 * (1.f / 65535.f); return 1.002f * sg * (1.f - sg); };
              av[0] = om0 * dfv[ks][0] * dsg_(w00 & 0xffffu); av[1] = om0 * dfv[ks][1] * dsg_(w00 >> 16);
              av[2] = om1 * dfv[ks][0] * dsg_(w10 & 0xffffu); av[3] = om1 * dfv[ks][1] * dsg_(w10 >> 16);
              av[4] = om0 * dfv[ks][2] * dsg_(w01 & 0xffffu); av[5] = om0 * dfv[ks][3] * dsg_(w01 >> 16);
              av[6] = om1 * dfv[ks][2] * dsg_(w11 & 0xffffu); av[7] = om1 * dfv[ks][3] * dsg_(w11 >> 16);
            } else {
              // permuted output 32 is sigma; outputs 33..47 are padding
#pragma unroll
              for (int e = 0; e < 8; ++e) av[e] = 0.f;
              if (t4 == 0) { av[0] = ds0; av[2] = ds1; }
            }
            if (p.dump_do) {
              float* d0 = p.dump_do + ((size_t)ray * T + base + g) * 33;
              float* d1 = d0 + 8 * 33;
              if (ks < 2) {
                const int c0 = 1 + 16 * ks + 2 * t4;
                if (v0) { d0[c0] = av[0]; d0[c0 + 1] = av[1]; d0[c0 + 8] = av[4]; d0[c0 + 9] = av[5]; }
                if (v1) { d1[c0] = av[2]; d1[c0 + 1] = av[3]; d1[c0 + 8] = av[6]; d1[c0 + 9] = av[7]; }
              } else if (t4 == 0) {
                if (v0) d0[0] = av[0];
                if (v1) d1[0] = av[2];
              }
            }
            uint32_t a2h[4], a2l[4];
            split_pair(av[0], av[1], a2h[0], a2l[0]);
            split_pair(av[2], av[3], a2h[1], a2l[1]);
            split_pair(av[4], av[5], a2h[2], a2l[2]);
            split_pair(av[6], av[7], a2h[3], a2l[3]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint2 bh = w1t[((j * 3 + ks) * 2 + 0) * 32 + lane];
              const uint2 bl = w1t[((j * 3 + ks) * 2 + 1) * 32 + lane];
              mma_bf16(dh[j], a2h, bh);
              mma_bf16(dh[j], a2l, bh);
              mma_bf16(dh[j], a2h, bl);
            }
          }
          // d(pre) = dh * softplus'(pre) = dh * (1 - exp(-h)); then df = d(pre) . W0
          float df[4][4];
#pragma unroll
          for (int cn = 0; cn < 4; ++cn) { df[cn][0] = df[cn][1] = df[cn][2] = df[cn][3] = 0.f; }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            float dp[2][4];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                dp[jj][e] = dh[2 * ks + jj][e] * (1.f - ex2f(-1.4426950408889634f * h[2 * ks + jj][e]));
            uint32_t a1h[4], a1l[4];
            split_pair(dp[0][0], dp[0][1], a1h[0], a1l[0]);
            split_pair(dp[0][2], dp[0][3], a1h[1], a1l[1]);
            split_pair(dp[1][0], dp[1][1], a1h[2], a1l[2]);
            split_pair(dp[1][2], dp[1][3], a1h[3], a1l[3]);
#pragma unroll
            for (int cn = 0; cn < 4; ++cn) {
              const uint2 bh = w0t[((cn * 4 + ks) * 2 + 0) * 32 + lane];
              const uint2 bl = w0t[((cn * 4 + ks) * 2 + 1) * 32 + lane];
              mma_bf16(df[cn], a1h, bh);
              mma_bf16(df[cn], a1l, bh);
              mma_bf16(df[cn], a1h, bl);
            }
          }
          __syncwarp();      // every lane holds its feature fragments; the tile can now carry d(feature) / 3
          {
            const float third = 1.f / 3.f;
#pragma unroll
            for (int cn = 0; cn < 4; ++cn) {
              const int c0 = 8 * cn + 2 * t4;
              *reinterpret_cast<float2*>(ftile + g * RC + colx(g, c0)) = make_float2(df[cn][0] * third, df[cn][1] * third);
              *reinterpret_cast<float2*>(ftile + (g + 8) * RC + colx(g + 8, c0)) = make_float2(df[cn][2] * third, df[cn][3] * third);
            }
          }
          __syncwarp();
          // scatter: lane = (sample, channel quad); one 16 B reduction per tap
          {
            const int sq = lane >> 3, cg = lane & 7;
            char* lbw = reinterpret_cast<char*>(dpl) + cg * 16;
            for (int q = 0; q < cnt; q += 4) {
              const int sl = q + sq;
              if (sl < cnt) {
                const float4 d4 = *reinterpret_cast<const float4*>(ftile + sl * RC + colx(sl, 4 * cg));
                const Tap* tp = taps + sl * 12;
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                  const Tap rec = tp[k];
#ifdef HFAGP_DBG_SCATTER_MASK
                  if (!((HFAGP_DBG_SCATTER_MASK >> (k >> 2)) & 1)) continue;
#endif
                  if (rec.w != 0.f)
                    red_add_v4(reinterpret_cast<float*>(lbw + rec.off), rec.w * d4.x, rec.w * d4.y, rec.w * d4.z, rec.w * d4.w);
                }
              }
            }
          }
          __syncwarp();
        }
      };
      bwd_tiles(0, S, 0);
      if (SF > 0) bwd_tiles(S, T, S16);
      continue;
    }
    float dacc = 0.f;
    for (int k = lane; k < T - 1; k += 32) dacc = fmaf(wts[k], 0.5f * (sdep[k] + sdep[k + 1]), dacc);
    dacc = warp_sum(dacc);
    // composite: sum_k w_k (c_k + c_{k+1})/2 over the sorted samples == sum_j omega_j c_j over the samples in
    // storage order with omega(order[k]) = (w_{k-1} + w_k)/2; omega overwrites sig[] (no longer needed).
    // c_j = q_j * QSTEP - 0.001 and sum_j omega_j = sum_k w_k, so the dequantisation is applied once at the end.
    float osum = 0.f;
    for (int k = lane; k < T; k += 32) {
      const float om = 0.5f * ((k > 0 ? wts[k - 1] : 0.f) + (k < T - 1 ? wts[k] : 0.f));
      sig[order[k]] = om;
      osum += om;
    }
    osum = warp_sum(osum);
    __syncwarp();
    float acc = 0.f;
    {
      const int cw = lane >> 1, sh = (lane & 1) * 16;
      auto qcol = [&](int crow) { return (float)((colq[crow * 16 + colqx(crow, cw)] >> sh) & 0xffffu); };
      int j = 0;
      for (; j + 4 <= S; j += 4) {
        const float4 om = *reinterpret_cast<const float4*>(sig + j);
        acc = fmaf(om.x, qcol(j), acc);
        acc = fmaf(om.y, qcol(j + 1), acc);
        acc = fmaf(om.z, qcol(j + 2), acc);
        acc = fmaf(om.w, qcol(j + 3), acc);
      }
      for (; j < S; ++j) acc = fmaf(sig[j], qcol(j), acc);
      const int shift = S16 - S;                   // fine sample j lives in colour row j + shift
      for (; j < T && (j & 3); ++j) acc = fmaf(sig[j], qcol(j + shift), acc);
      for (; j + 4 <= T; j += 4) {
        const float4 om = *reinterpret_cast<const float4*>(sig + j);
        acc = fmaf(om.x, qcol(j + shift), acc);
        acc = fmaf(om.y, qcol(j + shift + 1), acc);
        acc = fmaf(om.z, qcol(j + shift + 2), acc);
        acc = fmaf(om.w, qcol(j + shift + 3), acc);
      }
      for (; j < T; ++j) acc = fmaf(sig[j], qcol(j + shift), acc);
    }
    acc = fmaf(acc, QSTEP, -0.001f * osum);
    p.feat[(size_t)ray * RC + lane] = acc * 2.f - 1.f;
    if (lane == 0) {
      float dv = dacc / wtot;
      if (isnan(dv)) dv = INFINITY;
      dv = fminf(fmaxf(dv, __ldg(p.depth_range)), __ldg(p.depth_range + 1));
      p.depth[ray] = dv;
      p.wsum[ray] = wtot;
    }
    if (p.sort_idx)
      for (int i = lane; i < T; i += 32) p.sort_idx[(size_t)ray * T + i] = order[i];
    if (p.depths_sorted)
      for (int i = lane; i < T; i += 32) p.depths_sorted[(size_t)ray * T + i] = sdep[i];
    __syncwarp();
  }
}

}  // namespace hfagp

using namespace hfagp;

namespace hfagp {
// The integer stage of the renderer in isolation, one warp per ray, on caller-supplied floats (the very device
// functions the render kernels call): searchsorted(right=True) / below / above of u against cdf, and the stable sort
// permutation of the merged depth list.
__global__ void __launch_bounds__(128) render_bookkeeping_kernel(int rays, int ncdf, int s_fine, int T, const float* __restrict__ cdf,
                                                                 const float* __restrict__ u, const float* __restrict__ depths,
                                                                 int32_t* __restrict__ inds, int32_t* __restrict__ below,
                                                                 int32_t* __restrict__ above, int32_t* __restrict__ sort_idx) {
  extern __shared__ float bk_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ray = blockIdx.x * 4 + warp;
  if (ray >= rays) return;
  float* c = bk_smem + warp * (ncdf + T);
  float* dep = c + ncdf;
  for (int i = lane; i < ncdf; i += 32) c[i] = cdf[(size_t)ray * ncdf + i];
  for (int i = lane; i < T; i += 32) dep[i] = depths[(size_t)ray * T + i];
  __syncwarp();
  if (inds)
    for (int k = lane; k < s_fine; k += 32) {
      const int ind = searchsorted_right(c, ncdf, u[(size_t)ray * s_fine + k]);
      inds[(size_t)ray * s_fine + k] = ind;
      below[(size_t)ray * s_fine + k] = max(ind - 1, 0);
      above[(size_t)ray * s_fine + k] = min(ind, ncdf - 1);
    }
  if (sort_idx) {
    float de[4];
    int rk[4];
    stable_ranks<4>(dep, T, lane, de, rk, T - s_fine);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (lane + 32 * e < T) sort_idx[(size_t)ray * T + rk[e]] = lane + 32 * e;
  }
}
}  // namespace hfagp

extern "C" int hfagp_render_bookkeeping(int rays, int ncdf, int s_fine, int t_total, const float* cdf, const float* u,
                                        const float* depths, int32_t* inds, int32_t* below, int32_t* above,
                                        int32_t* sort_idx, void* stream) {
  HFAGP_CHECK_ARG(rays > 0 && ncdf >= 1 && ncdf <= 128 && t_total >= 1 && t_total <= 128 && s_fine >= 0, "render_bookkeeping: bad dims");
  HFAGP_CHECK_ARG(cdf && depths && (!inds || (u && below && above)), "render_bookkeeping: null pointer");
  render_bookkeeping_kernel<<<(rays + 3) / 4, 128, 4 * (ncdf + t_total) * sizeof(float), (cudaStream_t)stream>>>(
      rays, ncdf, s_fine, t_total, cdf, u, depths, inds, below, above, sort_idx);
  HFAGP_CHECK_LAUNCH("render_bookkeeping_kernel");
  return HFAGP_OK;
}

static int render_fwd_impl(bool allow_tc, const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                           const float* lin, const float* jitter, const float* u_fine, const float* depth_range,
                           float* feat, float* depth, float* wsum, int32_t* inds, int32_t* below, int32_t* above, int32_t* sort_idx,
                           float* depths_sorted, void* stream) {
  HFAGP_CHECK_ARG(desc && planes && c && mlp && lin && jitter && depth_range && feat && depth && wsum, "render_fwd: null pointer");
  const HfagpRenderDesc& d = *desc;
  HFAGP_CHECK_ARG(d.batch > 0 && d.res > 0 && d.plane_h > 0 && d.plane_w > 0, "render_fwd: bad dims");
  HFAGP_CHECK_ARG(d.s_coarse >= 4 && d.s_coarse <= 64 && d.s_fine >= 0 && d.s_fine <= 64,
                  "render_fwd: samples per ray must be 4..64 coarse, 0..64 fine");
  HFAGP_CHECK_ARG(d.s_fine == 0 || u_fine, "render_fwd: u_fine required when s_fine > 0");
  HFAGP_CHECK_ARG(!inds || (below && above), "render_fwd: inds/below/above go together");
  HFAGP_CHECK_ARG((long long)d.plane_h * d.plane_w * 96 < (1ll << 31), "render_fwd: plane too large for 32-bit tap offsets");
  RenderParams p{d, planes, c, mlp, lin, jitter, u_fine, depth_range, feat, depth, wsum, inds, below, above, sort_idx, depths_sorted,
                 nullptr, nullptr};
  const int sms = device_sm_count();
  if (allow_tc && render_tc_supported(d)) return render_tc_launch(p, sms, (cudaStream_t)stream);
  int nwarps = R_WARPS;
  size_t smem = ((WEIGHT_BYTES + 15) & ~15) + nwarps * render_warp_bytes(d.s_coarse, d.s_fine);
  if (smem > 227 * 1024) {
    nwarps = R_WARPS / 2;
    smem = ((WEIGHT_BYTES + 15) & ~15) + nwarps * render_warp_bytes(d.s_coarse, d.s_fine);
  }
  static std::atomic<uint64_t> attr_done{0};   // opt in to the full 227 KB once per device; not repeated on the (graph-captured) hot path
  HFAGP_CUDA(per_device_once(attr_done, [] { return cudaFuncSetAttribute(render_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  HFAGP_CHECK_ARG(smem <= 227 * 1024, "render_fwd: shared memory need exceeds 227 KB");
  const long long strips = (long long)d.batch * ((d.res + nwarps - 1) / nwarps) * d.res;
  const int blocks = (int)(strips < sms ? strips : sms);   // persistent: one CTA per SM
  render_kernel<false><<<blocks, nwarps * 32, smem, (cudaStream_t)stream>>>(p);
  HFAGP_CHECK_LAUNCH("render_kernel<fwd>");
  return HFAGP_OK;
}

extern "C" int hfagp_render_fwd(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                                const float* lin, const float* jitter, const float* u_fine, const float* depth_range,
                                float* feat, float* depth, float* wsum, int32_t* inds, int32_t* below, int32_t* above, int32_t* sort_idx,
                                float* depths_sorted, void* stream) {
  return render_fwd_impl(true, desc, planes, c, mlp, lin, jitter, u_fine, depth_range, feat, depth, wsum, inds, below, above,
                         sort_idx, depths_sorted, stream);
}

extern "C" int hfagp_render_fwd_simt(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                                     const float* lin, const float* jitter, const float* u_fine, const float* depth_range,
                                     float* feat, float* depth, float* wsum, int32_t* inds, int32_t* below, int32_t* above,
                                     int32_t* sort_idx, float* depths_sorted, void* stream) {
  return render_fwd_impl(false, desc, planes, c, mlp, lin, jitter, u_fine, depth_range, feat, depth, wsum, inds, below, above,
                         sort_idx, depths_sorted, stream);
}

static int render_bwd_impl(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                           const float* lin, const float* jitter, const float* u_fine, const float* dfeat,
                           float* dplanes, float* dump_f, float* dump_do, void* stream) {
  HFAGP_CHECK_ARG(desc && planes && c && mlp && lin && jitter && dfeat && dplanes, "render_bwd: null pointer");
  const HfagpRenderDesc& d = *desc;
  HFAGP_CHECK_ARG(d.batch > 0 && d.res > 0 && d.plane_h > 0 && d.plane_w > 0, "render_bwd: bad dims");
  HFAGP_CHECK_ARG(d.s_coarse >= 4 && d.s_coarse <= 64 && d.s_fine >= 0 && d.s_fine <= 64,
                  "render_bwd: samples per ray must be 4..64 coarse, 0..64 fine");
  HFAGP_CHECK_ARG(d.s_fine == 0 || u_fine, "render_bwd: u_fine required when s_fine > 0");
  HFAGP_CHECK_ARG((long long)d.plane_h * d.plane_w * 96 < (1ll << 31), "render_bwd: plane too large for 32-bit tap offsets");
  RenderParams p{d, planes, c, mlp, lin, jitter, u_fine, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                 nullptr, dfeat, dplanes, dump_f, dump_do};
  // rays (warps) per CTA: as many as the shared-memory plan allows, at most 12 (167 registers per thread at 384 threads)
  static const int nw_env = [] { const char* e = getenv("HFAGP_RBWD_WARPS"); return e ? atoi(e) : 0; }();      // profiling override
  int nwarps = nw_env >= 1 && nw_env <= 12 ? nw_env : 12;
  while (nwarps > 1 && WEIGHT_BYTES_BWD + nwarps * render_warp_bytes(d.s_coarse, d.s_fine, true) > 227 * 1024) --nwarps;
  const size_t smem = WEIGHT_BYTES_BWD + nwarps * render_warp_bytes(d.s_coarse, d.s_fine, true);
  static std::atomic<uint64_t> attr_done{0};
  HFAGP_CUDA(per_device_once(attr_done, [] { return cudaFuncSetAttribute(render_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  HFAGP_CHECK_ARG(smem <= 227 * 1024, "render_bwd: shared memory need exceeds 227 KB");
  const int sms = device_sm_count();
  const long long strips = (long long)d.batch * ((d.res + nwarps - 1) / nwarps) * d.res;
  const int blocks = (int)(strips < sms ? strips : sms);
  render_kernel<true><<<blocks, nwarps * 32, smem, (cudaStream_t)stream>>>(p);
  HFAGP_CHECK_LAUNCH("render_kernel<bwd>");
  return HFAGP_OK;
}

extern "C" int hfagp_render_bwd(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                                const float* lin, const float* jitter, const float* u_fine, const float* dfeat,
                                float* dplanes, void* stream) {
  return render_bwd_impl(desc, planes, c, mlp, lin, jitter, u_fine, dfeat, dplanes, nullptr, nullptr, stream);
}

extern "C" int hfagp_render_bwd_dec(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                                    const float* lin, const float* jitter, const float* u_fine, const float* dfeat,
                                    float* dplanes, float* dump_f, float* dump_do, void* stream) {
  HFAGP_CHECK_ARG(dump_f && dump_do, "render_bwd_dec: null dump buffers");
  return render_bwd_impl(desc, planes, c, mlp, lin, jitter, u_fine, dfeat, dplanes, dump_f, dump_do, stream);
}
