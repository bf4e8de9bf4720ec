// What surrounds the render path inside one training step (Trainer.gen_update, trainer_rgb.py:73-98):
//   latent_bwd   backward of the latent-subspace map  ws = weights . Q^T + delta  (headnerf.py:96-100)
//   facepool     AdaptiveAvgPool2d(size) of the 512^2 image (trainer_rgb.py:63,84) fused with the NHWC -> NCHW
//                layout change, and its transpose
//   mse          MSELoss(reduction='mean') forward (one atomic per CTA) and backward (trainer_rgb.py:15,85)
//   adam         torch.optim.Adam step over one flat fp32 parameter buffer (trainer_rgb.py:57,95), with the
//                1/world_size of the gradient all-reduce folded in
// All of these are HBM-bound streaming kernels: 16-byte accesses, grid sized from the element count.
#include "common.cuh"

namespace hfagp {

// dq[j][kk] = sum_n dws[n][j] * w[n][kk] ; ddelta[j] = sum_n dws[n][j]          one thread per (j, kk)
__global__ void latent_bwd_q_kernel(int batch, int k, int dim, const float* __restrict__ dws,
                                    const float* __restrict__ w, float* __restrict__ dq, float* __restrict__ ddelta) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)dim * k) return;
  const int j = (int)(idx / k), kk = (int)(idx - (long long)j * k);
  float acc = 0.f, dsum = 0.f;
  for (int n = 0; n < batch; ++n) {
    const float g = __ldg(dws + (size_t)n * dim + j);
    acc = fmaf(g, __ldg(w + (size_t)n * k + kk), acc);
    dsum += g;
  }
  if (dq) dq[idx] = acc;
  if (ddelta && kk == 0) ddelta[j] = dsum;
}

// dweights[n][kk] = sum_j dws[n][j] * q[j][kk]        block (kk, n), tree reduction over j
__global__ void __launch_bounds__(256) latent_bwd_w_kernel(int k, int dim, const float* __restrict__ dws,
                                                          const float* __restrict__ q, float* __restrict__ dweights) {
  const int kk = blockIdx.x, n = blockIdx.y;
  float acc = 0.f;
  for (int j = threadIdx.x; j < dim; j += blockDim.x)
    acc = fmaf(__ldg(dws + (size_t)n * dim + j), __ldg(q + (size_t)j * k + kk), acc);
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) dweights[(size_t)n * k + kk] = v;
  }
}

// y[n][c][oy][ox] = mean_{f x f} x[n][oy*f+..][ox*f+..][c]        one thread per (n, oy, ox), all c (c <= 4)
__global__ void facepool_fwd_kernel(int batch, int h, int w_, int c, int f, const float* __restrict__ x,
                                    float* __restrict__ y) {
  const int oh = h / f, ow = w_ / f;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)batch * oh * ow) return;
  const int ox = (int)(idx % ow);
  const long long r = idx / ow;
  const int oy = (int)(r % oh), n = (int)(r / oh);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int dy = 0; dy < f; ++dy)
    for (int dx = 0; dx < f; ++dx) {
      const float* px = x + (((size_t)n * h + oy * f + dy) * w_ + ox * f + dx) * c;
      for (int k = 0; k < c; ++k) acc[k] += __ldg(px + k);
    }
  const float inv = 1.f / (float)(f * f);
  for (int k = 0; k < c; ++k) y[(((size_t)n * c + k) * oh + oy) * ow + ox] = acc[k] * inv;
}

// dx[n][iy][ix][c] = dy[n][c][iy/f][ix/f] / f^2
__global__ void facepool_bwd_kernel(int batch, int h, int w_, int c, int f, const float* __restrict__ dy,
                                    float* __restrict__ dx) {
  const int oh = h / f, ow = w_ / f;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)batch * h * w_) return;
  const int ix = (int)(idx % w_);
  const long long r = idx / w_;
  const int iy = (int)(r % h), n = (int)(r / h);
  const float inv = 1.f / (float)(f * f);
  for (int k = 0; k < c; ++k)
    dx[(size_t)idx * c + k] = __ldg(dy + (((size_t)n * c + k) * oh + iy / f) * ow + ix / f) * inv;
}

// loss[0] += scale * sum (a - b)^2
__global__ void __launch_bounds__(256) mse_fwd_kernel(long long count, const float* __restrict__ a,
                                                     const float* __restrict__ b, float scale, float* __restrict__ loss) {
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long c4 = count >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c4; i += stride) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(a) + i), v = __ldg(reinterpret_cast<const float4*>(b) + i);
    const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
    acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  if (blockIdx.x == 0 && threadIdx.x < (count & 3)) {
    const float d = a[(c4 << 2) + threadIdx.x] - b[(c4 << 2) + threadIdx.x];
    acc += d * d;
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss, v * scale);
  }
}

// da[i] (+)= 2 * scale * gout[0] * (a[i] - b[i])
__global__ void mse_bwd_kernel(long long count, const float* __restrict__ a, const float* __restrict__ b, float scale,
                               const float* __restrict__ gout, int accumulate, float* __restrict__ da) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float g = 2.f * scale * __ldg(gout) * (a[i] - b[i]);
  da[i] = accumulate ? da[i] + g : g;
}

// torch.optim.Adam (no amsgrad, L2 weight decay folded into the gradient), single tensor:
//   g = grad * grad_scale + wd * p ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g g
//   p -= step_size * m / (sqrt(v) / bc2_sqrt + eps)          step_size = lr / (1 - b1^t), bc2_sqrt = sqrt(1 - b2^t)
__global__ void __launch_bounds__(256) adam_kernel(long long count, float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ m, float* __restrict__ v, float grad_scale,
                                                  float omb1, float beta2, float omb2, float eps, float wd,
                                                  float step_size, float bc2_sqrt, const float* __restrict__ sched) {
  if (sched) {                                        // step-dependent scalars from device memory (CUDA-graph replays)
    step_size = __ldg(sched);
    bc2_sqrt = __ldg(sched + 1);
  }
  const long long c4 = count >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto upd = [&](float& pv, float gv, float& mv, float& vv) {
    gv = gv * grad_scale;
    if (wd != 0.f) gv = fmaf(wd, pv, gv);
    mv = mv + omb1 * (gv - mv);                       // torch: exp_avg.lerp_(grad, 1 - beta1)
    vv = fmaf(omb2 * gv, gv, beta2 * vv);             // torch: exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv = pv - step_size * (mv / denom);
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    upd(pv.x, gv.x, mv.x, vv.x); upd(pv.y, gv.y, mv.y, vv.y); upd(pv.z, gv.z, mv.z, vv.z); upd(pv.w, gv.w, mv.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (count & 3)) {
    const long long i = (c4 << 2) + threadIdx.x;
    upd(p[i], g[i], m[i], v[i]);
  }
}

}  // namespace hfagp

using namespace hfagp;

extern "C" int hfagp_latent_bwd(int batch, int k, int dim, const float* dws, const float* weights, const float* q,
                                float* dweights, float* dq, float* ddelta, void* stream) {
  HFAGP_CHECK_ARG(dws && batch > 0 && batch <= 65535 && k > 0 && dim > 0, "latent_bwd: bad args");
  HFAGP_CHECK_ARG(!dq || weights, "latent_bwd: dq needs weights");
  HFAGP_CHECK_ARG(!dweights || q, "latent_bwd: dweights needs q");
  if (dq || ddelta) {
    latent_bwd_q_kernel<<<cdiv((long long)dim * k, 256), 256, 0, (cudaStream_t)stream>>>(batch, k, dim, dws, weights, dq, ddelta);
    HFAGP_CHECK_LAUNCH("latent_bwd_q_kernel");
  }
  if (dweights) {
    latent_bwd_w_kernel<<<dim3(k, batch), 256, 0, (cudaStream_t)stream>>>(k, dim, dws, q, dweights);
    HFAGP_CHECK_LAUNCH("latent_bwd_w_kernel");
  }
  return HFAGP_OK;
}

extern "C" int hfagp_facepool_fwd(int batch, int h, int w_, int c, int f, const float* x, float* y, void* stream) {
  HFAGP_CHECK_ARG(x && y && batch > 0 && c >= 1 && c <= 4 && f >= 1 && h % f == 0 && w_ % f == 0,
                  "facepool_fwd: need c <= 4 and an integer pooling factor");
  facepool_fwd_kernel<<<cdiv((long long)batch * (h / f) * (w_ / f), 256), 256, 0, (cudaStream_t)stream>>>(batch, h, w_, c, f, x, y);
  HFAGP_CHECK_LAUNCH("facepool_fwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_facepool_bwd(int batch, int h, int w_, int c, int f, const float* dy, float* dx, void* stream) {
  HFAGP_CHECK_ARG(dy && dx && batch > 0 && c >= 1 && c <= 4 && f >= 1 && h % f == 0 && w_ % f == 0,
                  "facepool_bwd: need c <= 4 and an integer pooling factor");
  facepool_bwd_kernel<<<cdiv((long long)batch * h * w_, 256), 256, 0, (cudaStream_t)stream>>>(batch, h, w_, c, f, dy, dx);
  HFAGP_CHECK_LAUNCH("facepool_bwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_mse_fwd(long long count, const float* a, const float* b, float scale, float* loss, void* stream) {
  HFAGP_CHECK_ARG(a && b && loss && count > 0, "mse_fwd: bad args");
  int blocks = cdiv(count >> 2, 256 * 4);
  if (blocks > device_sm_count() * 8) blocks = device_sm_count() * 8;
  if (blocks < 1) blocks = 1;
  mse_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(count, a, b, scale, loss);
  HFAGP_CHECK_LAUNCH("mse_fwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_mse_bwd(long long count, const float* a, const float* b, float scale, const float* gout,
                             int accumulate, float* da, void* stream) {
  HFAGP_CHECK_ARG(a && b && gout && da && count > 0, "mse_bwd: bad args");
  mse_bwd_kernel<<<cdiv(count, 256), 256, 0, (cudaStream_t)stream>>>(count, a, b, scale, gout, accumulate, da);
  HFAGP_CHECK_LAUNCH("mse_bwd_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_adam_step(long long count, float* p, const float* g, float* m, float* v, float grad_scale,
                               double lr, double beta1, double beta2, double eps, double weight_decay,
                               long long step, void* stream) {
  HFAGP_CHECK_ARG(p && g && m && v && count > 0 && step >= 1, "adam_step: bad args");
  HFAGP_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
  // scalar arithmetic in double on the host, as torch does in Python, then rounded once to fp32
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  int blocks = cdiv(count >> 2, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(count, p, g, m, v, grad_scale, (float)(1.0 - beta1), (float)beta2,
                                                        (float)(1.0 - beta2), (float)eps, (float)weight_decay, step_size,
                                                        bc2_sqrt, nullptr);
  HFAGP_CHECK_LAUNCH("adam_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_adam_sched(double lr, double beta1, double beta2, long long step, float* sched_host) {
  HFAGP_CHECK_ARG(sched_host && step >= 1, "adam_sched: bad args");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  sched_host[0] = (float)(lr / bc1);
  sched_host[1] = (float)sqrt(bc2);
  return HFAGP_OK;
}

extern "C" int hfagp_adam_step_dev(long long count, float* p, const float* g, float* m, float* v, float grad_scale,
                                   double beta1, double beta2, double eps, double weight_decay, const float* sched,
                                   void* stream) {
  HFAGP_CHECK_ARG(p && g && m && v && sched && count > 0, "adam_step_dev: bad args");
  HFAGP_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam_step_dev: buffers must be 16-byte aligned");
  int blocks = cdiv(count >> 2, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(count, p, g, m, v, grad_scale, (float)(1.0 - beta1), (float)beta2,
                                                        (float)(1.0 - beta2), (float)eps, (float)weight_decay, 0.f, 1.f, sched);
  HFAGP_CHECK_LAUNCH("adam_kernel");
  return HFAGP_OK;
}
