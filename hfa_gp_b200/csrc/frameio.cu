// Frame egress / ingress: the pixel-format conversions either side of the render path.
//   to_uint8     generated image [-1,1] fp32 channels-last -> uint8 [n][h][w][c] (what a PNG / video encoder takes),
//                0.79 MB per 512^2 frame crosses PCIe instead of 3.1 MB.  Two rounding conventions of the reference:
//                  mode 0  torchvision.utils.save_image(img, normalize=True, range=(-1,1))   run_recon_video_rgb.py:233-234
//                          v = clamp(x,-1,1); v = (v + 1) / 2; u8 = trunc(clamp(v*255 + 0.5, 0, 255))
//                  mode 1  layout_grid(float_to_uint8=True)                                   run_recon_video_rgb.py:34
//                          u8 = trunc(clamp(x*127.5 + 128, 0, 255))
//   from_uint8   decoded RGB frame uint8 [n][h][w][3] -> fp32 NCHW in [-1,1]: ToTensor() then Normalize(0.5, 0.5)
//                (train_rgb.py:78-81):  v = u8 / 255 ; v = (v - 0.5) / 0.5
//   resize       transforms.Resize on the decoded PIL frame (run_recon_video_3dmm.py:258-261, train_rgb.py:78-81) = Pillow's
//                two-pass fixed-point bilinear resampler: horizontal pass to a uint8 intermediate, vertical pass, each
//                u8 = clip8(((1 << 21) + sum_k pixel[xmin + k] * coeff[k]) >> 22)  with the 22-bit coefficient tables of
//                precompute_coeffs / normalize_coeffs_8bpc computed on the host (hfa_gp_b200/frameio.py); the vertical
//                pass writes uint8 and / or the ToTensor + Normalize result, so a 512^2 decoded frame becomes the
//                encoder's 256^2 fp32 input in two launches.  Integer arithmetic: bit-exact against PIL.
// Integer results are BIT-EXACT against torch: every float step is a separately rounded fp32 operation (no FMA
// contraction), in torch's order.  HBM-bound streaming kernels.
#include "common.cuh"

namespace hfagp {

__device__ __forceinline__ unsigned char quant_u8(float x, int mode) {
  float v;
  if (mode == 0) {
    v = fminf(fmaxf(x, -1.f), 1.f);
    v = __fadd_rn(v, 1.f);               // sub_(low) with low = -1
    v = __fdiv_rn(v, 2.f);               // div_(max(high - low, 1e-5))
    v = __fmul_rn(v, 255.f);
    v = __fadd_rn(v, 0.5f);
  } else {
    v = __fmul_rn(x, 127.5f);
    v = __fadd_rn(v, 128.f);
  }
  v = fminf(fmaxf(v, 0.f), 255.f);
  return (unsigned char)v;               // truncation, as Tensor.to(torch.uint8)
}

// one thread per 4 consecutive values (works on the flat channels-last array)
__global__ void to_uint8_kernel(size_t count4, size_t count, const float* __restrict__ x, int mode,
                                unsigned char* __restrict__ y) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
    uchar4 o;
    o.x = quant_u8(a.x, mode); o.y = quant_u8(a.y, mode); o.z = quant_u8(a.z, mode); o.w = quant_u8(a.w, mode);
    reinterpret_cast<uchar4*>(y)[i] = o;
  } else if (i == count4) {
    for (size_t k = count4 * 4; k < count; ++k) y[k] = quant_u8(x[k], mode);
  }
}

// y[n][c][h][w] = ((u8[n][h][w][c] / 255) - 0.5) / 0.5
__global__ void from_uint8_kernel(int batch, int h, int w_, int c, const unsigned char* __restrict__ x,
                                  float* __restrict__ y) {
  const size_t total = (size_t)batch * c * h * w_;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ox = i % w_;
  size_t r = i / w_;
  const int oy = r % h;
  r /= h;
  const int ch = r % c, n = r / c;
  float v = (float)x[(((size_t)n * h + oy) * w_ + ox) * c + ch];
  v = __fdiv_rn(v, 255.f);
  v = __fsub_rn(v, 0.5f);
  y[i] = __fdiv_rn(v, 0.5f);
}

constexpr int RESIZE_PRECISION_BITS = 32 - 8 - 2;       // Pillow's PRECISION_BITS

__device__ __forceinline__ unsigned char clip8(int acc) {
  const int v = acc >> RESIZE_PRECISION_BITS;
  return (unsigned char)min(max(v, 0), 255);
}

// y[n][row][ox][ch] from x[n][row][w][ch]: one thread per output byte (consecutive threads = consecutive channels / pixels)
__global__ void resize_h_u8_kernel(int rows, int w_, int c, int ow, int ksize, const int* __restrict__ bounds,
                                   const int* __restrict__ kk, const unsigned char* __restrict__ x, unsigned char* __restrict__ y) {
  const size_t total = (size_t)rows * ow * c;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = i % c;
  size_t r = i / c;
  const int ox = r % ow;
  const size_t row = r / ow;
  const int xmin = __ldg(bounds + 2 * ox), cnt = __ldg(bounds + 2 * ox + 1);
  const unsigned char* src = x + (row * w_ + xmin) * c + ch;
  const int* k = kk + (size_t)ox * ksize;
  int acc = 1 << (RESIZE_PRECISION_BITS - 1);
  for (int t = 0; t < cnt; ++t) acc += (int)src[(size_t)t * c] * __ldg(k + t);
  y[i] = clip8(acc);
}

// vertical pass over x[n][h][w][ch] -> oh rows; writes uint8 [n][oh][w][ch] and / or fp32 NCHW ((u8 / 255) - 0.5) / 0.5
__global__ void resize_v_u8_kernel(int batch, int h, int w_, int c, int oh, int ksize, const int* __restrict__ bounds,
                                   const int* __restrict__ kk, const unsigned char* __restrict__ x,
                                   unsigned char* __restrict__ y_u8, float* __restrict__ y_f32) {
  const size_t total = (size_t)batch * oh * w_ * c;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = i % c;
  size_t r = i / c;
  const int ox = r % w_;
  r /= w_;
  const int oy = r % oh, n = r / oh;
  const int ymin = __ldg(bounds + 2 * oy), cnt = __ldg(bounds + 2 * oy + 1);
  const unsigned char* src = x + (((size_t)n * h + ymin) * w_ + ox) * c + ch;
  const int* k = kk + (size_t)oy * ksize;
  int acc = 1 << (RESIZE_PRECISION_BITS - 1);
  for (int t = 0; t < cnt; ++t) acc += (int)src[(size_t)t * w_ * c] * __ldg(k + t);
  const unsigned char u = clip8(acc);
  if (y_u8) y_u8[i] = u;
  if (y_f32) {
    float v = __fdiv_rn((float)u, 255.f);
    v = __fsub_rn(v, 0.5f);
    y_f32[(((size_t)n * c + ch) * oh + oy) * w_ + ox] = __fdiv_rn(v, 0.5f);
  }
}

}  // namespace hfagp

using namespace hfagp;

extern "C" int hfagp_frame_resize_u8(int batch, int h, int w_, int c, int out_h, int out_w, int ksize_h, const int* bounds_h,
                                     const int* coeffs_h, int ksize_v, const int* bounds_v, const int* coeffs_v,
                                     const unsigned char* x, unsigned char* tmp, unsigned char* y_u8, float* y_f32, void* stream) {
  HFAGP_CHECK_ARG(x && (y_u8 || y_f32) && batch > 0 && h > 0 && w_ > 0 && c > 0 && out_h > 0 && out_w > 0, "frame_resize_u8: bad args");
  HFAGP_CHECK_ARG(bounds_v && coeffs_v && ksize_v > 0, "frame_resize_u8: the vertical pass needs its coefficient table");
  HFAGP_CHECK_ARG(out_w == w_ || (tmp && bounds_h && coeffs_h && ksize_h > 0), "frame_resize_u8: a width change needs tmp and the horizontal table");
  const unsigned char* mid = x;
  if (out_w != w_) {
    const size_t total = (size_t)batch * h * out_w * c;
    resize_h_u8_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch * h, w_, c, out_w, ksize_h, bounds_h, coeffs_h, x, tmp);
    HFAGP_CHECK_LAUNCH("resize_h_u8_kernel");
    mid = tmp;
  }
  const size_t total = (size_t)batch * out_h * out_w * c;
  resize_v_u8_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(batch, h, out_w, c, out_h, ksize_v, bounds_v, coeffs_v, mid,
                                                                       y_u8, y_f32);
  HFAGP_CHECK_LAUNCH("resize_v_u8_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_frame_to_uint8(long long count, const float* x, int mode, unsigned char* y, void* stream) {
  HFAGP_CHECK_ARG(x && y && count > 0 && (mode == 0 || mode == 1), "frame_to_uint8: bad args");
  HFAGP_CHECK_ARG((((uintptr_t)x) & 15) == 0 && (((uintptr_t)y) & 3) == 0, "frame_to_uint8: misaligned buffers");
  const size_t c4 = (size_t)count >> 2;
  to_uint8_kernel<<<cdiv(c4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(c4, (size_t)count, x, mode, y);
  HFAGP_CHECK_LAUNCH("to_uint8_kernel");
  return HFAGP_OK;
}

extern "C" int hfagp_frame_from_uint8(int batch, int h, int w_, int c, const unsigned char* x, float* y, void* stream) {
  HFAGP_CHECK_ARG(x && y && batch > 0 && h > 0 && w_ > 0 && c > 0, "frame_from_uint8: bad args");
  from_uint8_kernel<<<cdiv((long long)batch * c * h * w_, 256), 256, 0, (cudaStream_t)stream>>>(batch, h, w_, c, x, y);
  HFAGP_CHECK_LAUNCH("from_uint8_kernel");
  return HFAGP_OK;
}
