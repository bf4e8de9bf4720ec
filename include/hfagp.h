/*
 * hfagp.h — C ABI of libhfagp_sm100.so: the B200 (sm_100a) hot path of HFA-GP's per-frame render.
 *
 * The reference (bbaaii/HFA-GP) has no FFI of its own: its seam is the Python object protocol
 *   HeadNeRF_*.get_weights / get_latent / get_image                (code/networks/headnerf.py:76-134)
 *   generator.synthesis(ws, c=label, noise_mode='const')['image']  (code/networks/headnerf.py:112)
 * and every arithmetic op below that seam is a torch / NVlabs-eg3d call.  Each entry point here
 * replaces one of those calls; the comment on it cites the call it replaces.  The Python host
 * side (hfa_gp_b200/_cabi.py, ctypes) passes raw device pointers, sizes and a cudaStream_t —
 * no torch types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success, a negative HFAGP_E_* otherwise; never throws;
 *     hfagp_last_error() returns a thread-local description of the last failure.
 *   - all pointers are DEVICE pointers owned by the caller unless the name ends in _host.
 *   - never allocates device memory, never synchronises the stream; work is enqueued on `stream`.
 *   - activations are fp32, channels-last:  x[n][y][x][c]  ("NHWC").
 *   - conv weights are fp32 packed  w[tap][cout][cin]  (tap = ky*kw + kx), see hfagp_pack notes.
 *   - re-entrant: forward and backward may be driven from different host threads.
 *   - device: work goes to the device that is CURRENT on the calling thread (the one `stream` belongs to); the
 *     caller selects it with cudaSetDevice / hfagp_set_device before the call.  One-time kernel attributes (dynamic
 *     shared-memory opt-in) and the SM count used for grid sizing are kept per device ordinal, so one process may
 *     drive several GPUs.
 */
#ifndef HFAGP_H_
#define HFAGP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFAGP_ABI_VERSION 1

enum {
  HFAGP_OK = 0,
  HFAGP_E_INVALID = -1,   /* bad argument / unsupported shape */
  HFAGP_E_CUDA = -2,      /* a CUDA runtime call or launch failed */
  HFAGP_E_WORKSPACE = -3  /* caller workspace too small */
};

enum { HFAGP_ACT_LINEAR = 0, HFAGP_ACT_LRELU = 1 /* slope 0.2 */, HFAGP_ACT_RELU = 2 /* LPIPS AlexNet */ };

#define HFAGP_MAX_TAPS 32          /* 5x5 = 25 taps (LPIPS AlexNet conv2) is the largest tap list on the path */

int hfagp_abi_version(void);
const char* hfagp_last_error(void);
/* cudaSetDevice(device) for the calling thread (the `device` argument of SURVEY 8b, hoisted out of every signature). */
int hfagp_set_device(int device);
/* SM count of the current device: the persistent grids (one CTA per SM) and the split-K heuristic are sized by it. */
int hfagp_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution, channels-last.
 *
 *   acc[n][my][mx][co] = sum_t sum_ci x[n][my*in_stride + dy[t]][mx*in_stride + dx[t]][ci]
 *                                      * w[n*w_batch_stride + (wtap[t]*cout + co)*cin + ci]
 *   (input coordinates outside [0,H)x[0,W) read as zero), for my<oh, mx<ow, then the epilogue
 *
 *   v = acc * dcoef[n][co]           (if dcoef)        demodulation      eg3d modulated_conv2d
 *   v += noise[oy][ox] * noise_gain  (if noise)        per-pixel noise   SynthesisLayer 'const'
 *   v += bias[co]                    (if bias)
 *   v = lrelu_0.2(v) | relu(v)       (if act==LRELU | RELU)
 *   v *= act_gain ; clamp to +-clamp (if clamp > 0)                      bias_act
 *   v = (v + residual[n][oy][ox][co]) * residual_scale   (if residual)   ResBlock (out+skip)/sqrt2
 *   v += upsample2d(up_img)[n][oy][ox][co]               (if up_img)     'skip' image path
 *   y[n][oy][ox][co] = v     with  oy = my*out_stride + out_off_y, ox likewise, in a tensor of
 *                            out_h x out_w pixels.
 *
 * Replaces: F.conv2d in EqualConv2d.forward (code/networks/encoder3d.py:101-103) fused with
 * FusedLeakyReLU (:7-20) and the ResBlock merge (:191-198); and, inside generator.synthesis
 * (code/networks/headnerf.py:112 -> NVlabs/eg3d networks_stylegan2.py modulated_conv2d /
 * conv2d_resample / bias_act / ToRGBLayer / upfirdn2d.upsample2d).  The stride-2 transposed
 * convolution of the up-sampling layers is expressed as four (dy,dx,wtap) tap lists, one per
 * output parity class, with out_stride = 2.
 * ------------------------------------------------------------------------------------------- */
typedef struct HfagpConvDesc {
  int32_t batch, in_h, in_w, cin, cout;
  int32_t oh, ow;                 /* extent of the (my,mx) loop */
  int32_t in_stride;              /* 1 or 2 */
  int32_t out_h, out_w;           /* extent of the output tensor */
  int32_t out_stride, out_off_y, out_off_x;
  int32_t ntaps;
  int32_t dy[HFAGP_MAX_TAPS], dx[HFAGP_MAX_TAPS], wtap[HFAGP_MAX_TAPS];
  int64_t w_batch_stride;         /* 0 = weights shared by the batch */
  int32_t act;                    /* HFAGP_ACT_* */
  float act_gain, clamp;          /* clamp <= 0 : none */
  float noise_gain;               /* multiplies noise[] (noise_strength) */
  float residual_scale;
  int32_t up_h, up_w;             /* extent of up_img (= out/2) when given */
} HfagpConvDesc;

int hfagp_conv2d_fwd(const HfagpConvDesc* desc, const float* x, const float* w, const float* dcoef,
                     const float* noise, const float* bias, const float* residual, const float* up_img,
                     float* y, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core path (tcgen05 + TMEM, TMA-fed) of the same convolution, fp32-class accuracy.
 *
 * Operands are "split bf16": every fp32 value v is carried as two bf16 tensors of the same shape,
 * hi = bf16(v), lo = bf16(v - hi); the kernel accumulates hi*hi + lo*hi + hi*lo in fp32 (TMEM).
 *   x_hi/x_lo  [n][in_h][in_w][cin]      cin % 8 == 0 (K is walked in 64-channel chunks; TMA zero-fills a partial one)
 *   w_hi/w_lo  [n or 1][w_taps_total][cout][cin]   (per-sample when desc->w_batch_stride != 0)
 * Output: either y (fp32) or the split pair (y_hi, y_lo) for a following tensor-core layer.
 * Semantics, tap lists and the epilogue are exactly those of hfagp_conv2d_fwd.
 * ------------------------------------------------------------------------------------------- */
int hfagp_conv2d_tc_fwd(const HfagpConvDesc* desc, const uint16_t* x_hi, const uint16_t* x_lo,
                        const uint16_t* w_hi, const uint16_t* w_lo, int w_taps_total, const float* dcoef,
                        const float* noise, const float* bias, const float* residual, const float* up_img,
                        float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream);

/* hfagp_conv2d_tc_fwd that ALSO feeds the block's small ToRGB (cout <= 4, the super-resolution blocks): the finished
 * activations of a pixel are multiplied with the per-sample modulated ToRGB weights rgb_w[n][rgb_k][cout] while they
 * are still in registers and ADDED to rgb_acc[n][out_h][out_w][rgb_k] (caller zeroes it), so the layer output is not
 * read a second time; hfagp_torgb_finalize_fwd then applies bias, clamp and "+ upsample2d(previous image)". */
int hfagp_conv2d_tc_rgb_fwd(const HfagpConvDesc* desc, const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi,
                            const uint16_t* w_lo, int w_taps_total, const float* dcoef, const float* noise,
                            const float* bias, float* y, uint16_t* y_hi, uint16_t* y_lo, const float* rgb_w, int rgb_k,
                            float* rgb_acc, void* stream);
int hfagp_torgb_finalize_fwd(int batch, int h, int w_, int k, const float* acc, const float* bias, float clamp,
                             const float* up_img, float* y, void* stream);

/* Several sub-problems in ONE launch: descs[0..ndesc) (ndesc <= 4) share operands, channel counts, strides, the
 * output tensor and the epilogue, and differ only in their output window (oh, ow, out_off_*) and tap list — the
 * four output-parity classes of the stride-2 transposed convolution (up-sampling layers; data gradient of the
 * encoder's stride-2 convolutions) run as one persistent grid instead of four launches. */
int hfagp_conv2d_tc_multi_fwd(const HfagpConvDesc* descs, int ndesc, const uint16_t* x_hi, const uint16_t* x_lo,
                              const uint16_t* w_hi, const uint16_t* w_lo, int w_taps_total, const float* dcoef,
                              const float* noise, const float* bias, const float* residual, const float* up_img,
                              float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream);

/* The encoder's first layer at inference, ConvLayer(3, C, 1) (code/networks/encoder3d.py:142-179, kernel_size 1): 1x1
 * convolution of the NCHW frame x[n][cin][h][w] (cin <= 4) with w[cout][cin] (equalised-lr scale folded in) + bias +
 * activation * act_gain, written directly as the split-bf16 channels-last operand y[n][h][w][cout] of the following
 * tensor-core convolution (one pass instead of layout change + SIMT convolution + split). */
int hfagp_stem_conv1x1_fwd(int batch, int h, int w_, int cin, int cout, const float* x_nchw, const float* w,
                           const float* bias, int act, float act_gain, uint16_t* y_hi, uint16_t* y_lo, void* stream);

/* Caller-provided scratch, as pure functions of the shapes ("workspace is caller-provided"): the library never
 * allocates.  hfagp_conv2d_tc_acc_workspace_bytes = the zeroed fp32 accumulator acc[n][out_h][out_w][cout] of the split-K
 * form below (descs[0] carries the full output geometry); 0 for a null / degenerate descriptor.  (The renderer's
 * counterpart, hfagp_render_bwd_dec_workspace_bytes, is declared with its descriptor further down.) */
size_t hfagp_conv2d_tc_acc_workspace_bytes(const HfagpConvDesc* desc);

/* Split-K form for layers with few output tiles (the 4^2..32^2 blocks: a handful of 128-pixel tiles would leave
 * most SMs idle while each CTA streams megabytes of weights): every tile's K range (K chunks x tap groups) is dealt
 * out to up to `ksplit` CTAs, which ADD their raw fp32 partial sums into acc[n][out_h][out_w][cout] with 16-byte
 * atomics (caller zeroes acc).  No epilogue is applied: follow with hfagp_conv_epilogue_fwd (or, for the
 * up-sampling layers, hfagp_upfir_act_fwd, which reads the raw sums anyway). */
int hfagp_conv2d_tc_acc_fwd(const HfagpConvDesc* descs, int ndesc, const uint16_t* x_hi, const uint16_t* x_lo,
                            const uint16_t* w_hi, const uint16_t* w_lo, int w_taps_total, int ksplit, float* acc,
                            void* stream);

/* The epilogue of hfagp_conv2d_fwd (demod, noise, bias, leaky-ReLU, gain, clamp, residual merge, skip-image add)
 * applied elementwise to raw sums acc[n][out_h][out_w][cout] of a dense output (out_stride 1); writes y (fp32, may
 * alias acc) or the split pair. */
int hfagp_conv_epilogue_fwd(const HfagpConvDesc* desc, const float* acc, const float* dcoef, const float* noise,
                            const float* bias, const float* residual, const float* up_img, float* y, uint16_t* y_hi,
                            uint16_t* y_lo, void* stream);

/* fp32 -> split bf16 (elementwise). */
int hfagp_split_bf16(long long count, const float* x, uint16_t* hi, uint16_t* lo, void* stream);

/* hfagp_modulate_fwd writing the modulated weights as split bf16 (dcoef from the fp32 products). */
int hfagp_modulate_split_fwd(int batch, int ntaps, int cout, int cin, const float* w, const float* styles,
                             uint16_t* wmod_hi, uint16_t* wmod_lo, float* dcoef, void* stream);

/* hfagp_modulate_split_fwd for every layer of a network in ONE launch (host arrays of length nlayers: packed weight
 * pointer, taps, cout, cin, offset of the layer's styles inside the flat styles buffer of hfagp_styles_fwd, output
 * pointers; dcoef_host[l] may be NULL = no demodulation).  All styles of a frame exist before its first convolution,
 * so the inference path modulates the whole generator in one pass. */
int hfagp_modulate_split_multi_fwd(int nlayers, int batch, const float* const* w_host, const int32_t* ntaps_host,
                                   const int32_t* cout_host, const int32_t* cin_host, const int64_t* styles_off_host,
                                   const float* styles, uint16_t* const* hi_host, uint16_t* const* lo_host,
                                   float* const* dcoef_host, void* stream);

/* 4x4 [1,3,3,1]x[1,3,3,1]/64 FIR (gain 4, pad 1) over the (2H+1)x(2W+1) output of the stride-2
 * transposed convolution, fused with demodulation, noise, bias, leaky-ReLU, gain and clamp.
 * t[n][2H+1][2W+1][c] -> y[n][2H][2W][c] (fp32), or the split-bf16 pair (y_hi, y_lo) when y is NULL.
 * Replaces: upfirdn2d(..., padding=[1,1,1,1], gain=4) + fma(dcoef, noise) + bias_act in
 * eg3d conv2d_resample / SynthesisLayer.forward (reached via code/networks/headnerf.py:112). */
int hfagp_upfir_act_fwd(int batch, int h2, int w2, int c, const float* t, const float* dcoef,
                        const float* noise, float noise_gain, const float* bias, int act, float act_gain,
                        float clamp, float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream);

/* 1x1 modulated conv to <=4 output channels (the super-resolution ToRGB layers) fused with bias,
 * clamp and "+ upsample2d(previous image)".  w is the per-sample modulated weight [n][cout][cin];
 * the input is x (fp32) or, when x is NULL, the split-bf16 pair (x_hi, x_lo).
 * Replaces: eg3d ToRGBLayer.forward + SynthesisBlock skip add (via headnerf.py:112). */
int hfagp_torgb_small_fwd(int batch, int h, int w_, int cin, int cout, const float* x, const uint16_t* x_hi,
                          const uint16_t* x_lo, const float* w, const float* bias, float clamp,
                          const float* up_img, float* y, void* stream);
/* Training form of hfagp_torgb_small_fwd: also writes mask[n][h][w][cout] (one byte each) = |acc + bias| < clamp, the
 * derivative of the clamp the backward pass multiplies d(image) with (instead of a second, unclamped evaluation). */
int hfagp_torgb_small_mask_fwd(int batch, int h, int w_, int cin, int cout, const float* x, const uint16_t* x_hi,
                               const uint16_t* x_lo, const float* w, const float* bias, float clamp, const float* up_img,
                               float* y, unsigned char* mask, void* stream);

/* Per-layer styles for a whole network in one launch:
 *   styles[l][n][i] = (ws[n][widx[l]][:] . A_l[i][:] * inv_sqrt_wdim + b_l[i]) * post_gain[l]
 * Layer table (host arrays of length nlayers): affine weight / bias device pointers, cin, ws index,
 * post gain (1 for conv layers, 1/sqrt(cin) for ToRGB), and the offset of layer l inside `styles`.
 * Replaces: FullyConnectedLayer affine(w) of every SynthesisLayer/ToRGBLayer (eg3d). */
int hfagp_styles_fwd(int nlayers, int batch, int num_ws, int w_dim, const float* ws,
                     const float* const* aff_w_host, const float* const* aff_b_host,
                     const int32_t* cin_host, const int32_t* widx_host, const float* post_gain_host,
                     const int64_t* out_off_host, float* styles, void* stream);

/* Style modulation of one layer's weights, per sample:
 *   wmod[n][t][o][i] = w[t][o][i] * styles[n][i]
 *   dcoef[n][o]      = rsqrt(sum_{t,i} wmod^2 + 1e-8)        (only if dcoef != NULL)
 * Replaces: the `w = weight * styles; dcoefs = ...rsqrt()` head of eg3d modulated_conv2d. */
int hfagp_modulate_fwd(int batch, int ntaps, int cout, int cin, const float* w, const float* styles,
                       float* wmod, float* dcoef, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tri-plane volume renderer: ray generation, stratified + importance sampling, tri-plane bilinear
 * gather, OSG decoder MLP (32 -> 64 softplus -> 1+32), mid-point ray marching; one call per batch.
 *
 * planes  [n][ph][pw][3*32]  channels-last, plane p in channels [32p, 32p+32)
 * c       [n][25]            cam2world 4x4 row-major + intrinsics 3x3 (after the GL flip)
 * mlp     packed decoder weights with runtime gains folded in:
 *           w0[64][32], b0[64], w1[33][64], b1[33]   (contiguous, 2048+64+2112+33 floats)
 * lin     [s_coarse]   torch.linspace(ray_start, ray_end, s_coarse)
 * jitter  [n][rays][s_coarse]   upstream's torch.rand_like(depths_coarse)
 * u_fine  [n*rays][s_fine]      upstream's torch.rand(N_rays, N_importance)
 * feat    [n][res][res][32]  composited feature image, channels-last (rgb*2-1 applied)
 * depth   [n][rays], wsum [n][rays]
 * depth_range [2]  global min / max of all sample depths (upstream clamps the composite depth to it)
 * Optional integer bookkeeping (may be NULL), exactly upstream's tensors as int32:
 *   inds/below/above [n*rays][s_fine] (searchsorted right=True, clamps), sort_idx [n][rays][s_c+s_f],
 *   depths_sorted [n][rays][s_c+s_f] (fp32).
 * Replaces: RaySampler + ImportanceRenderer.forward + OSGDecoder + MipRayMarcher2 of eg3d
 * (reached via code/networks/headnerf.py:112).
 * ------------------------------------------------------------------------------------------- */
typedef struct HfagpRenderDesc {
  int32_t batch, res, plane_h, plane_w;
  int32_t s_coarse, s_fine;       /* <= 64 each; s_fine may be 0 */
  float delta;                    /* (ray_end - ray_start) / (s_coarse - 1) */
  float box_scale;                /* 2 / box_warp */
} HfagpRenderDesc;

int hfagp_render_fwd(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                     const float* lin, const float* jitter, const float* u_fine, const float* depth_range,
                     float* feat, float* depth, float* wsum, int32_t* inds, int32_t* below, int32_t* above, int32_t* sort_idx,
                     float* depths_sorted, void* stream);

/* Same contract as hfagp_render_fwd, but always on the legacy warp-per-ray kernel whose decoder MLP runs on mma.sync
 * (csrc/render.cu).  hfagp_render_fwd itself dispatches to the tcgen05 renderer (csrc/render_tc.cu: 128-sample tiles,
 * decoder layers as tcgen05.mma with TMEM accumulators) whenever its shared-memory plan fits; this entry point keeps the
 * second implementation reachable for cross-checks (tests/test_gpu_parity.py::test_render_tc_equals_simt). */
int hfagp_render_fwd_simt(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                          const float* lin, const float* jitter, const float* u_fine, const float* depth_range,
                          float* feat, float* depth, float* wsum, int32_t* inds, int32_t* below, int32_t* above,
                          int32_t* sort_idx, float* depths_sorted, void* stream);

/* The renderer's integer bookkeeping in isolation (north_star: "bit-exact for ray-index bookkeeping"): the same device
 * functions the render kernels call, run on caller-supplied floats, one warp per ray:
 *   inds = searchsorted(cdf[ray][0..ncdf), u[ray][k], right=True), below = max(inds-1, 0), above = min(inds, ncdf-1)
 *   sort_idx[ray] = stable argsort of depths[ray][0..t_total)                      (ncdf, t_total <= 128)
 * Given the reference's own cdf / depth floats the outputs must equal torch.searchsorted / clamp / torch.sort exactly
 * (tests/test_gpu_parity.py::test_bookkeeping_on_oracle_floats).  inds (with below/above) or sort_idx may be null. */
int hfagp_render_bookkeeping(int rays, int ncdf, int s_fine, int t_total, const float* cdf, const float* u,
                             const float* depths, int32_t* inds, int32_t* below, int32_t* above, int32_t* sort_idx,
                             void* stream);

/* Backward of hfagp_render_fwd w.r.t. the planes (decoder frozen, trainer_rgb.py:59-60): per ray the forward is
 * recomputed from the same inputs (jitter / u_fine make it deterministic), then d(feat) is pushed through the
 * composite, the mid-point march (d sigma), the decoder MLP (tensor-pipe, split bf16) and the bilinear gather;
 * dplanes[n][ph][pw][96] is ACCUMULATED with 16-byte red.global.add (caller zeroes it).  Sample positions carry
 * no gradient (upstream's importance sampling runs under no_grad); depth / wsum are not differentiated.
 * Replaces: autograd of ImportanceRenderer.forward + OSGDecoder + MipRayMarcher2 + grid_sample (eg3d). */
int hfagp_render_bwd(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                     const float* lin, const float* jitter, const float* u_fine, const float* dfeat,
                     float* dplanes, void* stream);

/* hfagp_render_bwd that additionally writes, per sample in storage order (coarse samples first, index
 * (n*rays + ray)*(s_coarse+s_fine) + j), the two operands of the decoder-MLP weight gradient: dump_f[...][32] the mean
 * tri-plane features fed to the decoder and dump_do[...][33] the gradient of its raw outputs (column 0 = sigma,
 * 1+c = colour c).  Used when tune_generator() has unfrozen the decoder (code/train_rgb.py:132-134). */
int hfagp_render_bwd_dec(const HfagpRenderDesc* desc, const float* planes, const float* c, const float* mlp,
                         const float* lin, const float* jitter, const float* u_fine, const float* dfeat,
                         float* dplanes, float* dump_f, float* dump_do, void* stream);
/* Bytes of the two per-sample operand dumps above (dump_f [samples][32], dump_do [samples][33]; samples = batch * res^2 *
 * (s_coarse + s_fine)): a pure function of the shapes, returns their sum. */
size_t hfagp_render_bwd_dec_workspace_bytes(const HfagpRenderDesc* desc, size_t* dump_f_bytes, size_t* dump_do_bytes);

/* Weight gradient of the OSG decoder MLP from the per-sample operands hfagp_render_bwd_dec wrote: dump_f [samples][32],
 * dump_do [samples][33].  The hidden layer is recomputed in fp32 per 64-sample tile, the four reductions over samples
 * (dW0, db0, dW1, db1) are accumulated in registers and ADDED to dmlp, which has the packing of `mlp`
 * (W0 [64][32], b0 [64], W1 [33][64], b1 [33] = 4257 floats; caller zeroes it).  Gradients are w.r.t. the effective
 * (gain-multiplied) weights in `mlp`.  Replaces: autograd of OSGDecoder's parameters (eg3d triplane.py), reached when
 * tune_generator() has unfrozen the generator (code/train_rgb.py:132-134). */
int hfagp_decoder_wgrad(long long samples, const float* dump_f, const float* dump_do, const float* mlp, float* dmlp,
                        void* stream);

/* [1,3,3,1]^2/64 FIR, zero-pad (pad0,pad1), optional output stride and gain:
 *   y[n][oy][ox][c] = gain * sum_{ky,kx} g[ky] g[kx] x[n][oy*stride + ky - pad0][ox*stride + kx - pad0][c]
 * with oh = (h + pad0 + pad1 - 4) / stride + 1.  Input is x (fp32) or the split-bf16 pair (x_hi, x_lo);
 * output is y (fp32) or the split pair (y_hi, y_lo) for a tensor-core consumer (split I/O needs c % 4 == 0).
 * Forward use: Blur.forward / upfirdn2d_native (code/networks/encoder3d.py:23-41,59-75), stride 2 = only the
 * positions a following stride-2 1x1 conv reads.  Backward uses (the filter is symmetric): transpose of a
 * stride-1 blur (pads 3-pad0, ...), transpose of the FIR after the up-sampling transposed conv (pads 2,2,
 * gain 4) and transpose of upsample2d (pads 1,1, stride 2, gain 4). */
int hfagp_blur_fwd(int batch, int h, int w_, int c, int pad0, int pad1, int stride, float gain, const float* x,
                   const uint16_t* x_hi, const uint16_t* x_lo, float* y, uint16_t* y_hi, uint16_t* y_lo,
                   void* stream);

/* Transpose of hfagp_blur_fwd (needed for stride 2): dy[n][oh][ow][c] -> dx[n][h][w][c],
 *   dx[iy][ix] = gain * sum_{ky,kx} g[ky] g[kx] dy[(iy + pad0 - ky)/stride][(ix + pad0 - kx)/stride]  (exact divisions)
 * Replaces: autograd of Blur + stride-2 conv sampling in the encoder skip path (encoder3d.py:165-171). */
int hfagp_blur_up(int batch, int h, int w_, int c, int pad0, int pad1, int stride, float gain, const float* dy,
                  float* dx, void* stream);

/* y[n][o] = (x[n][:] . w[o][:]) * w_gain + b[o] * b_gain   — EqualLinear with activation=None
 * (code/networks/encoder3d.py:128-136) and Weights_3DMM (code/networks/headnerf.py:152-158). */
int hfagp_linear_fwd(int batch, int cin, int cout, const float* x, const float* w, const float* b,
                     float w_gain, float b_gain, float* y, void* stream);

/* ws[n][j] = sum_k weights[n][k] * q[j][k] + delta[j]   (q = thin-QR basis [7168][k], row-major)
 * Replaces: diag_embed/matmul/sum + delta in get_latent (code/networks/headnerf.py:96-100). */
int hfagp_latent_fwd(int batch, int k, int dim, const float* weights, const float* q, const float* delta,
                     float* ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward path (training step: Trainer.gen_update, code/trainer_rgb.py:73-98 -> loss.backward()).
 * Data gradients of the convolutions are hfagp_conv2d[_tc]_fwd calls on transposed weights with mirrored tap
 * lists; the entries below are what sits between them.  Reductions ACCUMULATE into their outputs (atomics):
 * the caller zeroes them once per step.
 * ------------------------------------------------------------------------------------------- */
typedef struct HfagpActBwdDesc {
  int32_t batch, h, w, c;         /* geometry of the layer output y[n][h][w][c]; c % 4 == 0 */
  int32_t act;                    /* forward activation, HFAGP_ACT_* */
  float act_gain, clamp;          /* forward gain / clamp (clamp <= 0: none) */
  float noise_gain;               /* forward noise gain (only used to rebuild the pre-demod output for ddcoef) */
  float residual_scale;           /* forward y = (act_out + residual) * residual_scale when residual is given */
  float post_scale;               /* extra factor on the summed incoming gradient (residual_scale for a merge) */
  int32_t rgb_k;                  /* channels of dimg for the fused small-ToRGB gradient (<= 4) */
} HfagpActBwdDesc;

/* Everything between two data-gradient convolutions, for one layer output y = epilogue(conv(x)):
 *   g      = post_scale * ( g0*s0[n][c] + g1*s1[n][c] + (sum_o dimg[..][o] * wrgb[o][c]) * srgb[n][c] )
 *   dpre   = g * d(epilogue)/d(pre)            (leaky-ReLU slope, gain, clamp mask, residual merge)
 *   dz     = dpre * dcoef[n][c]                -> gradient of the raw convolution output (fp32 or split-bf16)
 *   ds0[n][c]   += sum_pix g0 * y              d(styles) of the consumer whose unscaled data gradient is g0
 *   ds1[n][c]   += sum_pix g1 * y              (second consumer: the block's ToRGB)
 *   dsrgb[n][c] += sum_pix (dimg . wrgb) * y   (small ToRGB of the super-resolution blocks, fused)
 *   dbias[c]    += sum_{n,pix} dpre
 *   ddcoef[n][c]+= sum_pix dpre * z            z = conv output before demodulation, rebuilt from y
 * g1/s1, dimg/wrgb/srgb, dcoef, noise, bias, residual and every reduction output may be NULL.
 * Replaces: autograd of bias_act / modulated_conv2d's x*styles and *dcoefs / ToRGBLayer (eg3d, via
 * headnerf.py:112) and of FusedLeakyReLU + ResBlock merge (encoder3d.py:7-20,191-198). */
int hfagp_act_bwd(const HfagpActBwdDesc* desc, const float* y, const uint16_t* y_hi, const uint16_t* y_lo,
                  const float* g0, const float* s0, const float* g1, const float* s1, const float* dimg,
                  const float* wrgb, const float* srgb, const float* dcoef, const float* noise,
                  const float* bias, const float* residual, float* dz, uint16_t* dz_hi, uint16_t* dz_lo,
                  float* ds0, float* ds1, float* dsrgb, float* dbias, float* ddcoef, void* stream);

/* dws[n][widx[l]][k] += post_gain[l]/sqrt(w_dim) * sum_i dstyles[l][n][i] * A_l[i][k]  for every layer l
 * (layer table as in hfagp_styles_fwd; dstyles uses the same flat layout as styles). */
int hfagp_styles_bwd(int nlayers, int batch, int num_ws, int w_dim, const float* const* aff_w_host,
                     const int32_t* cin_host, const int32_t* widx_host, const float* post_gain_host,
                     const int64_t* off_host, const float* dstyles, float* dws, void* stream);

/* Demodulation backward: dstyles[n][i] -= styles[n][i] * sum_o ddcoef[n][o] * dcoef[n][o]^3 * w2[o][i],
 * w2[o][i] = sum_taps w[t][o][i]^2. */
int hfagp_demod_bwd(int batch, int cout, int cin, const float* w2, const float* styles, const float* dcoef,
                    const float* ddcoef, float* dstyles, void* stream);

/* EqualLinear backward (forward: hfagp_linear_fwd): dx = (dy . w) * w_gain (written), dw += dy^T x * w_gain,
 * db += sum_n dy * b_gain.  dx / dw / db may be NULL. */
int hfagp_linear_bwd(int batch, int cin, int cout, const float* dy, const float* x, const float* w,
                     float w_gain, float b_gain, float* dx, float* dw, float* db, void* stream);

/* Weight gradient of hfagp_conv2d_fwd (fp32 SIMT, split-K with atomics), same desc / tap lists:
 *   dw[wtap[t]][co][ci] += scale * sum_{n,my,mx} dz[n][my][mx][co] * x[n][my*in_stride+dy[t]][mx*in_stride+dx[t]][ci]
 * x and dz may each be fp32 or a split-bf16 pair; dz is dense [n][oh][ow][cout].
 * Replaces: autograd of F.conv2d w.r.t. weight in EqualConv2d (encoder3d.py:101-103). */
int hfagp_conv2d_wgrad(const HfagpConvDesc* desc, const float* x, const uint16_t* x_hi, const uint16_t* x_lo,
                       const float* dz, const uint16_t* dz_hi, const uint16_t* dz_lo, float scale, float* dw,
                       void* stream);

/* hfagp_conv2d_wgrad for a style-MODULATED convolution (the generator's layers once tune_generator() has unfrozen
 * them, code/train_rgb.py:132-134): the per-sample style multiplies one operand inside the sum over samples,
 *   dw[t][co][ci] += scale * sum_{n,pix} (dz[n][pix][co] * dzscale[n][co]) * (x[n][pix+t][ci] * xscale[n][ci])
 * xscale [batch][cin] / dzscale [batch][cout] may each be NULL.  (The up-sampling layers use it with the roles of x
 * and dz swapped: the dense operand is the layer input, the strided one the gradient of the transposed conv.) */
int hfagp_conv2d_wgrad_mod(const HfagpConvDesc* desc, const float* x, const uint16_t* x_hi, const uint16_t* x_lo,
                           const float* dz, const uint16_t* dz_hi, const uint16_t* dz_lo, const float* xscale,
                           const float* dzscale, float scale, float* dw, void* stream);

/* Last step of a modulated convolution's weight gradient (generator unfrozen, code/train_rgb.py:132-134): the style-scaled
 * wgrad dw (packed [tap][cout][cin]; [tap][cin][cout] when dw_transposed — up-sampling layers, whose wgrad runs with the
 * roles of the operands swapped) plus, when ddcoef is given, the demodulation term
 *   - w[tap][o][i] * sum_n ddcoef[n][o] dcoef[n][o]^3 styles[n][i]^2        (w: the packed, unmodulated weight)
 * is unpacked into the parameter layout and ADDED to grad[cout][cin][taps] (= the parameter's .grad, [O][I][k][k]).
 * Replaces: autograd of modulated_conv2d w.r.t. its weight (eg3d networks_stylegan2.py). */
int hfagp_modconv_wgrad_finish(int taps, int cout, int cin, int batch, const float* dw, int dw_transposed, const float* w,
                               const float* ddcoef, const float* dcoef, const float* styles, float* grad, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Around the render path inside one training step (Trainer.gen_update, code/trainer_rgb.py:73-98).
 * ------------------------------------------------------------------------------------------- */

/* Backward of hfagp_latent_fwd: dweights[n][k] = sum_j dws[n][j] q[j][k] ; dq[j][k] = sum_n dws[n][j] weights[n][k] ;
 * ddelta[j] = sum_n dws[n][j].  Outputs are WRITTEN; each may be NULL.  d(bases) continues through
 * hfagp_basis_qr_bwd (headnerf.py:92-100). */
int hfagp_latent_bwd(int batch, int k, int dim, const float* dws, const float* weights, const float* q,
                     float* dweights, float* dq, float* ddelta, void* stream);

/* Orthonormal basis of the latent subspace: Q, _ = torch.qr(bases.T + eps) of get_latent (code/networks/headnerf.py:92,
 * :187, :247; eps = 1e-8 there) for bases[k][m] row-major, 1 <= k <= 64, m >= k (m = 14*512).  q[m][k] row-major is
 * LAPACK's reduced Q factor (same column signs: CholeskyQR2 + Householder sign reconstruction, Gram matrices in fp64
 * — see csrc/qr.cu); rinv[k][k] = R^-1 (upper triangular) is kept for the backward.  The factorisation needs
 * cond(bases) < ~1e3: a Cholesky pivot below that floor sets a device flag readable with hfagp_basis_qr_info (which synchronises the stream; validation use).
 * `workspace`: hfagp_basis_qr_workspace_bytes(k, m) bytes of device memory, 256-byte aligned; contents are scratch.
 * hfagp_basis_qr_bwd: the autograd backward of that factorisation for a gradient arriving at Q only (R is unused
 * downstream): gbases[k][m] (WRITTEN) = ((gQ + Q Y) R^-T)^T with Y = X + X^T - diag(X), X = triu(-Q^T gQ).
 * `deterministic` != 0: the Gram matrices are accumulated by one CTA in row order instead of by fp64 atomics from many (the
 * atomic order can move the last bit of the fp32 result from run to run); ~1 ms instead of ~40 us. */
size_t hfagp_basis_qr_workspace_bytes(int k, int m);
int hfagp_basis_qr_fwd(int k, int m, const float* bases, float eps, float* q, float* rinv, void* workspace,
                       int deterministic, void* stream);
int hfagp_basis_qr_bwd(int k, int m, const float* gq, const float* q, const float* rinv, float* gbases, void* workspace,
                       int deterministic, void* stream);
int hfagp_basis_qr_info(const void* workspace, int k, int m, int* info_host, void* stream);

/* face_pool = AdaptiveAvgPool2d((size,size)) of the generated image (code/trainer_rgb.py:63,84) for an integer
 * factor f = h/size, fused with the layout change: x[n][h][w][c] channels-last -> y[n][c][h/f][w/f]; c <= 4.
 * hfagp_facepool_bwd is its transpose: dy[n][c][h/f][w/f] -> dx[n][h][w][c] (written). */
int hfagp_facepool_fwd(int batch, int h, int w_, int c, int f, const float* x, float* y, void* stream);
int hfagp_facepool_bwd(int batch, int h, int w_, int c, int f, const float* dy, float* dx, void* stream);

/* MSELoss(reduction='mean') (code/trainer_rgb.py:15,65-67,85): loss[0] += scale * sum_i (a[i]-b[i])^2 with
 * scale = 1/count (caller zeroes loss); backward da[i] (+)= 2 * scale * gout[0] * (a[i]-b[i]), gout a device scalar. */
int hfagp_mse_fwd(long long count, const float* a, const float* b, float scale, float* loss, void* stream);
int hfagp_mse_bwd(long long count, const float* a, const float* b, float scale, const float* gout, int accumulate,
                  float* da, void* stream);

/* One torch.optim.Adam step (code/trainer_rgb.py:57,95; amsgrad off) over a flat fp32 buffer of `count` elements:
 *   g' = g*grad_scale + weight_decay*p ; m = lerp(m, g', 1-beta1) ; v = beta2 v + (1-beta2) g'^2
 *   p -= lr/(1-beta1^step) * m / (sqrt(v)/sqrt(1-beta2^step) + eps)
 * grad_scale carries the 1/world_size of the data-parallel gradient mean.  Hyper-parameters are doubles: the scalar
 * terms are evaluated in double on the host (as torch evaluates them in Python) and rounded once.  Buffers must be
 * 16-byte aligned. */
int hfagp_adam_step(long long count, float* p, const float* g, float* m, float* v, float grad_scale, double lr,
                    double beta1, double beta2, double eps, double weight_decay, long long step, void* stream);

/* The same update with its two step-dependent scalars read from DEVICE memory, so that a training step captured in a
 * CUDA graph can be replayed: hfagp_adam_sched (host only, no GPU work) evaluates sched_host[0] = lr / (1 - beta1^step)
 * and sched_host[1] = sqrt(1 - beta2^step) in double exactly as hfagp_adam_step does; the caller copies the pair to
 * `sched` (device, 2 floats) on the launching stream before each replay. */
int hfagp_adam_sched(double lr, double beta1, double beta2, long long step, float* sched_host);
int hfagp_adam_step_dev(long long count, float* p, const float* g, float* m, float* v, float grad_scale, double beta1,
                        double beta2, double eps, double weight_decay, const float* sched, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LPIPS(net='alex') (code/trainer_rgb.py:10,62,86-87 -> pip package `lpips`): what sits between its five AlexNet
 * convolutions, which run on hfagp_conv2d_tc_fwd with the bias + HFAGP_ACT_RELU epilogue.
 * ------------------------------------------------------------------------------------------- */

/* ScalingLayer + zero-pad 2 + space-to-depth by 4 + split-bf16, x[n][3][h][w] (NCHW fp32, sides % 4 == 0) ->
 * y[n][(h+4)/4][(w+4)/4][48] with channel (py*4+px)*3+c = (x[c][4Y+py-2][4X+px-2] - shift[c]) / scale[c]: the 11x11
 * stride-4 first convolution becomes a 3x3 stride-1 convolution over 48 channels.  shift/scale: 3 host floats. */
int hfagp_lpips_stem_fwd(int batch, int h, int w_, const float* x, const float* shift3_host, const float* scale3_host,
                         uint16_t* y_hi, uint16_t* y_lo, void* stream);
/* its transpose: dx48[n][(h+4)/4][(w+4)/4][48] -> dimg[n][3][h][w] (written) */
int hfagp_lpips_stem_bwd(int batch, int h, int w_, const float* dx48, const float* scale3_host, float* dimg, void* stream);

/* nn.MaxPool2d(3, stride=2), channels-last (c % 4 == 0); input / output fp32 or split bf16.  Backward routes each
 * window's gradient to its FIRST maximum (row-major scan, as ATen) in gather form: dx is written, no atomics. */
int hfagp_maxpool3s2_fwd(int batch, int h, int w_, int c, const float* x, const uint16_t* x_hi, const uint16_t* x_lo,
                         float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream);
int hfagp_maxpool3s2_bwd(int batch, int h, int w_, int c, const float* x, const uint16_t* x_hi, const uint16_t* x_lo,
                         const float* dy, float* dx, void* stream);

/* One LPIPS layer: features f[2*batch][hw][c] (first half: image 0, second half: image 1), lin[c] ->
 *   out[b] += mean_pix sum_c lin[c] * (f0/(|f0|+1e-10) - f1/(|f1|+1e-10))^2        (caller zeroes out)
 * Backward: df1[batch][hw][c] = gout[b] * d out[b] / d f1 (written); image 0 carries no gradient. */
int hfagp_lpips_head_fwd(int batch, int hw, int c, const float* f, const uint16_t* f_hi, const uint16_t* f_lo,
                         const float* lin, float* out, void* stream);
int hfagp_lpips_head_bwd(int batch, int hw, int c, const float* f, const uint16_t* f_hi, const uint16_t* f_lo,
                         const float* lin, const float* gout, float* df1, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Frame egress / ingress: the pixel-format conversions either side of the render path.
 * ------------------------------------------------------------------------------------------- */

/* Generated image (fp32, any layout, `count` values, nominally in [-1,1]) -> uint8, BIT-EXACT with torch:
 *   mode 0: torchvision.utils.save_image(img, normalize=True, range=(-1,1))  (code/run_recon_video_rgb.py:233-234)
 *           v = clamp(x,-1,1); v = (v+1)/2; u8 = trunc(clamp(v*255 + 0.5, 0, 255))
 *   mode 1: layout_grid(float_to_uint8=True)                                  (code/run_recon_video_rgb.py:34)
 *           u8 = trunc(clamp(x*127.5 + 128, 0, 255))
 * Applied to the channels-last image it yields the [h][w][3] bytes an image / video encoder takes. */
int hfagp_frame_to_uint8(long long count, const float* x, int mode, unsigned char* y, void* stream);

/* Decoded frame uint8 [n][h][w][c] -> fp32 NCHW in [-1,1]: ToTensor() + Normalize(0.5, 0.5) of the reference's
 * loader (code/train_rgb.py:78-81; code/dataset.py:205-214), bit-exact: v = u8/255 ; y = (v - 0.5) / 0.5. */
int hfagp_frame_from_uint8(int batch, int h, int w_, int c, const unsigned char* x, float* y, void* stream);

/* transforms.Resize on the decoded frame, as Pillow does it (the reference's ingress: Resize(args.size) -> ToTensor ->
 * Normalize on a PIL image, run_recon_video_3dmm.py:258-261, run_recon_video_audio.py:258-261, train_rgb.py:78-81):
 * two-pass fixed-point bilinear resampling of uint8 x[n][h][w][c] — horizontal pass into tmp[n][h][out_w][c] (skipped
 * when out_w == w), vertical pass into y_u8[n][out_h][out_w][c] and / or y_f32[n][c][out_h][out_w] = ((u8/255)-0.5)/0.5.
 * Each pass: u8 = clip8(((1 << 21) + sum_{k < count} pixel[first + k] * coeffs[o][k]) >> 22), with bounds[o] = (first,
 * count) and the int32 coefficient tables of Pillow's precompute_coeffs / normalize_coeffs_8bpc (22-bit fixed point),
 * computed by the caller (hfa_gp_b200/frameio.py: pil_bilinear_table).  Bit-exact against PIL.Image.resize(BILINEAR). */
int hfagp_frame_resize_u8(int batch, int h, int w_, int c, int out_h, int out_w, int ksize_h, const int* bounds_h,
                          const int* coeffs_h, int ksize_v, const int* bounds_v, const int* coeffs_v, const unsigned char* x,
                          unsigned char* tmp, unsigned char* y_u8, float* y_f32, void* stream);

/* A convolution weight in torch layout w[cout][cin][taps] (taps = kh*kw) times `scale` (EqualConv2d's equalised-lr
 * factor, code/networks/encoder3d.py:86-103) -> the operand forms of the encoder's forward and backward convolutions in one
 * pass: pk[t][cout][cin] fp32, its split-bf16 pair (pk_hi, pk_lo), the transposed split-bf16 pair pkt[t][cin][cout] of the
 * data-gradient convolution.  Every output (pair) may be NULL. */
int hfagp_pack_conv_weight(int cout, int cin, int taps, const float* w, float scale, float* pk, uint16_t* pk_hi,
                           uint16_t* pk_lo, uint16_t* pkt_hi, uint16_t* pkt_lo, void* stream);

/* The way back for the weight gradient: dw[t][cout][cin_padded] (the layout hfagp_conv2d_wgrad accumulates in) is ADDED to
 * grad[cout][cin][taps], the torch layout of EqualConv2d.weight.grad (the equalised-lr scale is already in dw). */
int hfagp_unpack_conv_wgrad(int cout, int cin, int cin_padded, int taps, const float* dw, float* grad, void* stream);

/* Layout helpers (elementwise, bandwidth-bound): NCHW <-> NHWC for the frame entering the encoder
 * and the image leaving the super-resolution head. */
int hfagp_nchw_to_nhwc(int batch, int c, int h, int w_, const float* x, float* y, void* stream);
int hfagp_nhwc_to_nchw(int batch, int c, int h, int w_, const float* x, float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HFAGP_H_ */
