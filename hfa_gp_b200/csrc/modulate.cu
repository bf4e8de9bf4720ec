// Small bandwidth/latency-bound ops: style affines for the whole network in one launch, per-sample
// weight modulation + demodulation coefficients, EqualLinear, latent-subspace map.
#include <cstdlib>
#include "common.cuh"

namespace hfagp {

constexpr int MAX_STYLE_LAYERS = 48;

struct StyleTable {
  const float* aw[MAX_STYLE_LAYERS];
  const float* ab[MAX_STYLE_LAYERS];
  long long off[MAX_STYLE_LAYERS];
  int cin[MAX_STYLE_LAYERS];
  int widx[MAX_STYLE_LAYERS];
  float gain[MAX_STYLE_LAYERS];
  int row_start[MAX_STYLE_LAYERS + 1];  // prefix sum of cin: one warp per (layer, i)
  int nlayers;
};

// one warp per output row (layer l, channel i); loops over the batch
__global__ void styles_kernel(const StyleTable tb, int batch, int num_ws, int w_dim, float inv_sqrt,
                              const float* __restrict__ ws, float* __restrict__ styles) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= tb.row_start[tb.nlayers]) return;
  int l = 0;
  while (row >= tb.row_start[l + 1]) ++l;
  const int i = row - tb.row_start[l];
  const float* a = tb.aw[l] + (size_t)i * w_dim;
  const float b = __ldg(tb.ab[l] + i);
  for (int n = 0; n < batch; ++n) {
    const float* wv = ws + ((size_t)n * num_ws + tb.widx[l]) * w_dim;
    float s = 0.f;
    for (int k = lane; k < w_dim; k += 32) s = fmaf(__ldg(wv + k), __ldg(a + k) * inv_sqrt, s);
    s = warp_sum(s);
    if (lane == 0) styles[tb.off[l] + (size_t)n * tb.cin[l] + i] = (s + b) * tb.gain[l];
  }
}

// block per (o, n): wmod[n][t][o][i] = w[t][o][i] * s[n][i]; dcoef[n][o] = rsqrt(sum wmod^2 + 1e-8)
__global__ void modulate_kernel(int ntaps, int cout, int cin, const float* __restrict__ w,
                                const float* __restrict__ styles, float* __restrict__ wmod,
                                float* __restrict__ dcoef) {
  const int o = blockIdx.x, n = blockIdx.y;
  const float* sn = styles + (size_t)n * cin;
  float ss = 0.f;
  for (int t = 0; t < ntaps; ++t) {
    const float* wr = w + ((size_t)t * cout + o) * cin;
    float* wo = wmod + (((size_t)n * ntaps + t) * cout + o) * cin;
    for (int i = threadIdx.x; i < cin; i += blockDim.x) {
      float v = __ldg(wr + i) * __ldg(sn + i);
      wo[i] = v;
      ss = fmaf(v, v, ss);
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

// block per (n, o) for long rows (the encoder's final 4x4 convolution as a linear map: cin = 8192): 16-byte loads,
// block reduction
__global__ void __launch_bounds__(128) linear_long_kernel(int batch, int cin, int cout, const float* __restrict__ x,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         float w_gain, float b_gain, float* __restrict__ y) {
  const int o = blockIdx.x, n = blockIdx.y;
  const float4* xv = reinterpret_cast<const float4*>(x + (size_t)n * cin);
  const float4* wv = reinterpret_cast<const float4*>(w + (size_t)o * cin);
  float s = 0.f;
  for (int k = threadIdx.x; k < (cin >> 2); k += blockDim.x) {
    const float4 a = __ldg(xv + k), c = __ldg(wv + k);
    s = fmaf(a.x, c.x * w_gain, s);
    s = fmaf(a.y, c.y * w_gain, s);
    s = fmaf(a.z, c.z * w_gain, s);
    s = fmaf(a.w, c.w * w_gain, s);
  }
  __shared__ float red[4];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    s = red[0] + red[1] + red[2] + red[3];
    y[(size_t)n * cout + o] = s + (b ? __ldg(b + o) * b_gain : 0.f);
  }
}

// warp per (n, o)
__global__ void linear_kernel(int batch, int cin, int cout, const float* __restrict__ x, const float* __restrict__ w,
                              const float* __restrict__ b, float w_gain, float b_gain, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= batch * cout) return;
  const int n = row / cout, o = row - n * cout;
  const float* xv = x + (size_t)n * cin;
  const float* wv = w + (size_t)o * cin;
  float s = 0.f;
  for (int k = lane; k < cin; k += 32) s = fmaf(__ldg(xv + k), __ldg(wv + k) * w_gain, s);
  s = warp_sum(s);
  if (lane == 0) y[row] = s + (b ? __ldg(b + o) * b_gain : 0.f);
}

__global__ void latent_kernel(int batch, int k, int dim, const float* __restrict__ weights,
                              const float* __restrict__ q, const float* __restrict__ delta, float* __restrict__ ws) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * dim) return;
  int n = idx / dim, j = idx - n * dim;
  const float* wv = weights + (size_t)n * k;
  const float* qv = q + (size_t)j * k;
  float s = 0.f;
  for (int i = 0; i < k; ++i) s = fmaf(__ldg(wv + i), __ldg(qv + i), s);
  ws[idx] = s + __ldg(delta + j);
}

}  // namespace hfagp

using namespace hfagp;

extern "C" int hfagp_styles_fwd(int nlayers, int batch, int num_ws, int w_dim, const float* ws,
                                const float* const* aff_w_host, const float* const* aff_b_host,
                                const int32_t* cin_host, const int32_t* widx_host, const float* post_gain_host,
                                const int64_t* out_off_host, float* styles, void* stream) {
  HFAGP_CHECK_ARG(nlayers > 0 && nlayers <= MAX_STYLE_LAYERS, "styles_fwd: nlayers %d out of range", nlayers);
  HFAGP_CHECK_ARG(ws && styles && aff_w_host && aff_b_host && cin_host && widx_host && post_gain_host && out_off_host,
                  "styles_fwd: null pointer");
  StyleTable tb;
  tb.nlayers = nlayers;
  tb.row_start[0] = 0;
  for (int l = 0; l < nlayers; ++l) {
    HFAGP_CHECK_ARG(widx_host[l] >= 0 && widx_host[l] < num_ws, "styles_fwd: ws index out of range");
    tb.aw[l] = aff_w_host[l];
    tb.ab[l] = aff_b_host[l];
    tb.off[l] = out_off_host[l];
    tb.cin[l] = cin_host[l];
    tb.widx[l] = widx_host[l];
    tb.gain[l] = post_gain_host[l];
    tb.row_start[l + 1] = tb.row_start[l] + cin_host[l];
  }
  int rows = tb.row_start[nlayers];
  styles_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      tb, batch, num_ws, w_dim, 1.0f / sqrtf((float)w_dim), ws, styles);
  HFAGP_CHECK_LAUNCH("styles_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_modulate_fwd(int batch, int ntaps, int cout, int cin, const float* w, const float* styles,
                                  float* wmod, float* dcoef, void* stream) {
  HFAGP_CHECK_ARG(w && styles && wmod, "modulate_fwd: null pointer");
  HFAGP_CHECK_ARG(batch > 0 && batch <= 65535 && ntaps > 0 && cout > 0 && cin > 0, "modulate_fwd: bad dims");
  int threads = cin >= 256 ? 256 : (cin >= 128 ? 128 : 64);
  modulate_kernel<<<dim3(cout, batch), threads, 0, (cudaStream_t)stream>>>(ntaps, cout, cin, w, styles, wmod, dcoef);
  HFAGP_CHECK_LAUNCH("modulate_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_linear_fwd(int batch, int cin, int cout, const float* x, const float* w, const float* b,
                                float w_gain, float b_gain, float* y, void* stream) {
  HFAGP_CHECK_ARG(x && w && y && batch > 0 && cin > 0 && cout > 0, "linear_fwd: bad args");
  if (cin >= 2048 && (cin & 3) == 0 && batch <= 65535 && (((uintptr_t)x | (uintptr_t)w) & 15) == 0) {
    linear_long_kernel<<<dim3(cout, batch), 128, 0, (cudaStream_t)stream>>>(batch, cin, cout, x, w, b, w_gain, b_gain, y);
    HFAGP_CHECK_LAUNCH("linear_long_kernel");
    return HFAGP_OK;
  }
  linear_kernel<<<cdiv((long long)batch * cout * 32, 256), 256, 0, (cudaStream_t)stream>>>(batch, cin, cout, x, w, b,
                                                                                          w_gain, b_gain, y);
  HFAGP_CHECK_LAUNCH("linear_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_latent_fwd(int batch, int k, int dim, const float* weights, const float* q, const float* delta,
                                float* ws, void* stream) {
  HFAGP_CHECK_ARG(weights && q && delta && ws && batch > 0 && k > 0 && dim > 0, "latent_fwd: bad args");
  latent_kernel<<<cdiv((long long)batch * dim, 256), 256, 0, (cudaStream_t)stream>>>(batch, k, dim, weights, q, delta,
                                                                                    ws);
  HFAGP_CHECK_LAUNCH("latent_kernel");
  return HFAGP_OK;
}

// ---- error plumbing (one definition for the library)
namespace hfagp {
char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
bool pdl_enabled() {
  // measured on B200 inside the frame graph: 2.108 ms/frame with, 2.118 without (noise): the graph's kernel-to-kernel
  // edges already hide what PDL would; kept opt-in (HFAGP_PDL=1)
  static const bool on = [] { const char* e = getenv("HFAGP_PDL"); return e && e[0] == '1'; }();
  return on;
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace hfagp

extern "C" int hfagp_abi_version(void) { return HFAGP_ABI_VERSION; }
extern "C" const char* hfagp_last_error(void) { return hfagp::err_buf(); }

namespace hfagp {
int device_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int n = cache[dev & 63].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cache[dev & 63].store(n, std::memory_order_relaxed);
  return n;
}
}  // namespace hfagp

extern "C" int hfagp_device_sm_count(void) { return hfagp::device_sm_count(); }
extern "C" size_t hfagp_conv2d_tc_acc_workspace_bytes(const HfagpConvDesc* d) {
  if (!d || d->batch <= 0 || d->out_h <= 0 || d->out_w <= 0 || d->cout <= 0) return 0;
  return sizeof(float) * (size_t)d->batch * d->out_h * d->out_w * d->cout;
}
extern "C" size_t hfagp_render_bwd_dec_workspace_bytes(const HfagpRenderDesc* d, size_t* dump_f_bytes, size_t* dump_do_bytes) {
  size_t samples = 0;
  if (d && d->batch > 0 && d->res > 0 && d->s_coarse + d->s_fine > 0)
    samples = (size_t)d->batch * d->res * d->res * (size_t)(d->s_coarse + d->s_fine);
  const size_t fb = samples * 32 * sizeof(float), db = samples * 33 * sizeof(float);
  if (dump_f_bytes) *dump_f_bytes = fb;
  if (dump_do_bytes) *dump_do_bytes = db;
  return fb + db;
}
extern "C" int hfagp_set_device(int device) {
  HFAGP_CUDA(cudaSetDevice(device));
  return HFAGP_OK;
}
