// Tri-plane volume renderer, one warp per ray: ray generation -> stratified depths -> tri-plane bilinear
// gather (lane = channel, one 128 B texel line per tap) -> OSG decoder MLP (lane = sample, weights broadcast
// from shared memory) -> mid-point march -> importance resampling -> second gather/MLP -> stable rank-sort
// merge -> final compositing (lane = channel).  All per-ray state lives in shared memory / registers; HBM
// traffic is the plane reads (L2-resident) and one 128 B feature row + 2 scalars per ray.
#include <mutex>
#include "common.cuh"

namespace hfagp {

constexpr int RC = 32;        // channels per plane == decoder input width
constexpr int RH = 64;        // decoder hidden width
constexpr int RO = 33;        // 1 sigma + 32 colour features
constexpr int COL_LD = 33;    // padded row stride of the per-sample feature/colour rows
constexpr int R_WARPS = 4;
constexpr int MLP_FLOATS = RH * RC + RH + RO * RH + RO;  // 4257
constexpr int MLP_PAD = 4260;

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
};

struct Tap {
  int off;   // float offset of the texel's channel 0 inside this sample's frame, or -1 (outside -> zero)
  float w;
};

__host__ __device__ inline size_t render_warp_floats(int T, int s_coarse) {
  // taps[32*12*2] + col[T][33] + dep,sig,sdep,ssig,wts [T] + order[T] + cdf[s_coarse+2] + zmid[s_coarse]
  size_t f = (size_t)T * COL_LD + 6 * (size_t)T + (s_coarse + 2) + s_coarse + 32 * 12 * 2;
  return (f + 3) & ~(size_t)3;
}

// exclusive product scan over n values held as v(k) for k = lane + 32q; returns weights into wts[k] = alpha*T
// and the sum of weights.  alpha(k) supplied through a lambda.
template <typename FA>
__device__ __forceinline__ float march_weights(int nint, int lane, float* wts, FA alpha_of) {
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
    wsum += k < nint ? w : 0.f;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
  return warp_sum(wsum);
}

__global__ void __launch_bounds__(R_WARPS * 32) render_fwd_kernel(const RenderParams p) {
  extern __shared__ __align__(16) float smem[];
  const HfagpRenderDesc& d = p.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = d.s_coarse, SF = d.s_fine, T = S + SF;
  const int rays_per_frame = d.res * d.res;
  const long long total_rays = (long long)d.batch * rays_per_frame;

  // block-shared decoder weights
  float* mlp = smem;
  for (int i = threadIdx.x; i < MLP_FLOATS; i += blockDim.x) mlp[i] = __ldg(p.mlp + i);
  const float* w0 = mlp;
  const float* b0 = mlp + RH * RC;
  const float* w1 = b0 + RH;
  const float* b1 = w1 + RO * RH;
  __syncthreads();

  float* ws = smem + MLP_PAD + (size_t)warp * render_warp_floats(T, S);
  Tap* taps = reinterpret_cast<Tap*>(ws);  // first: keeps the 8-byte records aligned for any T
  float* col = ws + 32 * 12 * 2;
  float* dep = col + (size_t)T * COL_LD;
  float* sig = dep + T;
  float* sdep = sig + T;
  float* ssig = sdep + T;
  float* wts = ssig + T;
  int* order = reinterpret_cast<int*>(wts + T);
  float* cdf = reinterpret_cast<float*>(order + T);
  float* zmid = cdf + (S + 2);

  const int PW = d.plane_w, PH = d.plane_h;
  const int texel_stride = 3 * RC;

  for (long long ray = (long long)blockIdx.x * R_WARPS + warp; ray < total_rays; ray += (long long)gridDim.x * R_WARPS) {
    const int n = (int)(ray / rays_per_frame);
    const int r = (int)(ray - (long long)n * rays_per_frame);
    const float* cam = p.cam + (size_t)n * 25;
    const float* pl = p.planes + (size_t)n * PH * PW * texel_stride;

    // ---- ray generation (uniform across the warp)
    float ox_, oy_, oz_, dx_, dy_, dz_;
    {
      const int py = r / d.res, px = r - py * d.res;
      const float inv = 1.0f / d.res, half = 0.5f / d.res;
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

    // ---- gather + decode a run of samples [s_begin, s_end) whose depths are in dep[]
    auto shade = [&](int s_begin, int s_end) {
      for (int base = s_begin; base < s_end; base += 32) {
        const int cnt = min(32, s_end - base);
        // phase A: lane = sample, 12 (offset, weight) taps
        if (lane < cnt) {
          const float t = dep[base + lane];
          const float qx = (ox_ + t * dx_) * d.box_scale;
          const float qy = (oy_ + t * dy_) * d.box_scale;
          const float qz = (oz_ + t * dz_) * d.box_scale;
#pragma unroll
          for (int pidx = 0; pidx < 3; ++pidx) {
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
            Tap* tp = taps + (lane * 12 + pidx * 4);
            tp[0].off = (vx0 && vy0) ? (y0 * PW + x0) * texel_stride + pidx * RC : -1; tp[0].w = wl * wt;
            tp[1].off = (vx1 && vy0) ? (y0 * PW + x1) * texel_stride + pidx * RC : -1; tp[1].w = wr * wt;
            tp[2].off = (vx0 && vy1) ? (y1 * PW + x0) * texel_stride + pidx * RC : -1; tp[2].w = wl * wb;
            tp[3].off = (vx1 && vy1) ? (y1 * PW + x1) * texel_stride + pidx * RC : -1; tp[3].w = wr * wb;
          }
        }
        __syncwarp();
        // phase B: lane = channel
        for (int s = 0; s < cnt; ++s) {
          float acc[3] = {0.f, 0.f, 0.f};
          const Tap* tp = taps + s * 12;
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            const Tap tk = tp[k];
            if (tk.off >= 0) acc[k >> 2] = fmaf(tk.w, __ldg(pl + tk.off + lane), acc[k >> 2]);
          }
          col[(size_t)(base + s) * COL_LD + lane] = (acc[0] + acc[1] + acc[2]) / 3.f;
        }
        __syncwarp();
        // phase C: lane = sample, decoder MLP 32 -> 64 (softplus) -> 33
        if (lane < cnt) {
          float* row = col + (size_t)(base + lane) * COL_LD;
          float f[RC];
#pragma unroll
          for (int c = 0; c < RC; ++c) f[c] = row[c];
          float out[RO];
#pragma unroll
          for (int o = 0; o < RO; ++o) out[o] = b1[o];
#pragma unroll 1
          for (int jc = 0; jc < RH; jc += 8) {
            float h[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) h[jj] = b0[jc + jj];
#pragma unroll
            for (int c = 0; c < RC; c += 4) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const float4 w4 = *reinterpret_cast<const float4*>(w0 + (jc + jj) * RC + c);
                h[jj] = fmaf(w4.x, f[c], h[jj]);
                h[jj] = fmaf(w4.y, f[c + 1], h[jj]);
                h[jj] = fmaf(w4.z, f[c + 2], h[jj]);
                h[jj] = fmaf(w4.w, f[c + 3], h[jj]);
              }
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) h[jj] = softplus_t(h[jj]);
#pragma unroll
            for (int o = 0; o < RO; ++o) {
              const float4 wa = *reinterpret_cast<const float4*>(w1 + o * RH + jc);
              const float4 wb = *reinterpret_cast<const float4*>(w1 + o * RH + jc + 4);
              float v = out[o];
              v = fmaf(wa.x, h[0], v); v = fmaf(wa.y, h[1], v); v = fmaf(wa.z, h[2], v); v = fmaf(wa.w, h[3], v);
              v = fmaf(wb.x, h[4], v); v = fmaf(wb.y, h[5], v); v = fmaf(wb.z, h[6], v); v = fmaf(wb.w, h[7], v);
              out[o] = v;
            }
          }
          sig[base + lane] = out[0];
#pragma unroll
          for (int c = 0; c < RC; ++c) row[c] = (1.f / (1.f + expf(-out[1 + c]))) * 1.002f - 0.001f;
        }
        __syncwarp();
      }
    };

    shade(0, S);

    if (SF > 0) {
      // ---- coarse march (weights only) -> smoothed pdf -> inverse-CDF fine depths
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
        int ind = 0;  // searchsorted(cdf[0..NB], u, right=True) = #{cdf[i] <= u}
        for (int i = 0; i <= NB; ++i) ind += cdf[i] <= u ? 1 : 0;
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
      shade(S, T);
      // ---- stable rank sort of the T depths (coarse first, as torch.cat + sort sees them)
      for (int i = lane; i < T; i += 32) {
        const float di = dep[i];
        int rank = 0;
        for (int j = 0; j < T; ++j) {
          const float dj = dep[j];
          rank += (dj < di || (dj == di && j < i)) ? 1 : 0;
        }
        order[rank] = i;
        sdep[rank] = di;
        ssig[rank] = sig[i];
      }
    } else {
      for (int i = lane; i < T; i += 32) {
        order[i] = i;
        sdep[i] = dep[i];
        ssig[i] = sig[i];
      }
    }
    __syncwarp();

    // ---- final march
    const float wtot = march_weights(T - 1, lane, wts, [&](int k) {
      float sm = softplus_t(0.5f * (ssig[k] + ssig[k + 1]) - 1.f);
      return 1.f - expf(-(sm * (sdep[k + 1] - sdep[k])));
    });
    __syncwarp();
    float dacc = 0.f;
    for (int k = lane; k < T - 1; k += 32) dacc = fmaf(wts[k], 0.5f * (sdep[k] + sdep[k + 1]), dacc);
    dacc = warp_sum(dacc);
    float acc = 0.f;
    float prev = col[(size_t)order[0] * COL_LD + lane];
    for (int k = 0; k < T - 1; ++k) {
      const float cur = col[(size_t)order[k + 1] * COL_LD + lane];
      acc = fmaf(wts[k], 0.5f * (prev + cur), acc);
      prev = cur;
    }
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

extern "C" int hfagp_render_fwd(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
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
  RenderParams p{d, planes, c, mlp, lin, jitter, u_fine, depth_range, feat, depth, wsum, inds, below, above, sort_idx, depths_sorted};
  const int T = d.s_coarse + d.s_fine;
  size_t smem = (MLP_PAD + R_WARPS * render_warp_floats(T, d.s_coarse)) * sizeof(float);
  static std::once_flag attr_once;   // opt in to the full 227 KB once; not repeated on the (graph-captured) hot path
  std::call_once(attr_once, [] { cudaFuncSetAttribute(render_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  HFAGP_CHECK_ARG(smem <= 227 * 1024, "render_fwd: shared memory need exceeds 227 KB");
  long long total_rays = (long long)d.batch * d.res * d.res;
  int blocks = (int)((total_rays + R_WARPS - 1) / R_WARPS);
  render_fwd_kernel<<<blocks, R_WARPS * 32, smem, (cudaStream_t)stream>>>(p);
  HFAGP_CHECK_LAUNCH("render_fwd_kernel");
  return HFAGP_OK;
}
