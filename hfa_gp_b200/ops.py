"""Tensor-level wrappers over the C ABI.  Activations are fp32 channels-last ``[N,H,W,C]``.

Every function enqueues on ``torch.cuda.current_stream()`` and returns torch-owned outputs.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import ACT_LINEAR, ACT_LRELU, ActBwdDesc, ConvDesc, RenderDesc, check, ptr, stream

_desc_cache = {}
_launches = [0]
# bumped by every optimiser step that writes parameters through the C ABI (raw device writes do not advance
# torch's per-tensor version counters); part of the key of every packed-weight cache
param_epoch = [0]


# ---- zero-initialised scratch of a captured training step.  A step asks for ~60 small zeroed tensors (split-K and atomic
# accumulators, per-channel gradient sums, loss scalars); as separate fills each is a ~3 us node on the step graph's
# critical path.  While an arena is active (Trainer._capture only — eager steps never see it) they are carved out of ONE
# buffer that the graph zeroes with a single memset at its top.  Tensors from it live until the next replay, exactly like
# every other tensor allocated inside a captured step.
_zero_arena = None
ZERO_ARENA_MAX_ITEM = 1 << 20          # larger requests keep their own fill (bandwidth, not latency)


class ZeroArena:
    def __init__(self, device):
        self.device, self.buf, self.off, self.need = torch.empty(0, device=device).device, None, 0, 0

    def materialise(self):
        """Before the capture: allocate what the measured eager steps asked for."""
        if self.buf is None and self.need:
            self.buf = torch.empty(self.need, device=self.device, dtype=torch.uint8)

    def begin(self):
        """Top of a step: measuring (no buffer yet: requests are only counted) or, inside the capture, the one memset."""
        global _zero_arena
        if self.buf is not None:
            self.buf.zero_()
        self.off = 0
        _zero_arena = self

    def end(self):
        global _zero_arena
        self.need = max(self.need, self.off)
        _zero_arena = None


def zeros(shape, device, dtype=torch.float32):
    """``torch.zeros`` for scratch consumed inside the current step (see ZeroArena)."""
    a = _zero_arena
    if a is not None and (device == a.device or torch.empty(0, device=device).device == a.device):
        shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        n = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        if 0 < n <= ZERO_ARENA_MAX_ITEM:
            off = a.off
            a.off = off + (n + 255) // 256 * 256
            if a.buf is not None and a.off <= a.buf.numel():
                return a.buf[off:off + n].view(dtype).view(shape)
    return torch.zeros(shape, device=device, dtype=dtype)


def launch_count() -> int:
    """Kernels launched through the C ABI so far (every entry point enqueues exactly one kernel)."""
    return _launches[0]


def _ok(rc, what):
    check(rc, what)
    _launches[0] += 1


def _conv_desc(key, build):
    d = _desc_cache.get(key)
    if d is None:
        d = build()
        if len(_desc_cache) >= 4096:      # keys are shapes and constant epilogue settings: a long-running process that
            _desc_cache.clear()           # keeps seeing new shapes must not grow without bound
        _desc_cache[key] = d
    return d


def conv2d(x: torch.Tensor, w: torch.Tensor, taps: Sequence[Tuple[int, int, int]], cout: int, *,
           oh: int, ow: int, in_stride: int = 1, out: Optional[torch.Tensor] = None,
           out_hw: Optional[Tuple[int, int]] = None, out_stride: int = 1, out_off=(0, 0),
           w_batch_stride: int = 0, dcoef=None, noise=None, noise_gain: float = 0.0, bias=None,
           act: int = ACT_LINEAR, act_gain: float = 1.0, clamp: float = 0.0, residual=None,
           residual_scale: float = 1.0, up_img=None) -> torch.Tensor:
    """Generic implicit-GEMM convolution, see ``hfagp_conv2d_fwd`` in include/hfagp.h.
    taps: (dy, dx, weight-tap index) triples."""
    n, h, wd, cin = x.shape
    out_h, out_w = out_hw if out_hw is not None else (oh, ow)
    if out is None:
        out = torch.empty((n, out_h, out_w, cout), device=x.device, dtype=torch.float32)
    taps = tuple(taps)
    # per-call scalars that change during training (noise_strength after tune_generator) are NOT part of the key: they
    # are written into the cached descriptor right before the call
    key = (n, h, wd, cin, cout, oh, ow, in_stride, out_h, out_w, out_stride, out_off, taps, w_batch_stride, act,
           act_gain, clamp, residual_scale, up_img is not None)

    def build():
        d = ConvDesc()
        d.batch, d.in_h, d.in_w, d.cin, d.cout = n, h, wd, cin, cout
        d.oh, d.ow, d.in_stride = oh, ow, in_stride
        d.out_h, d.out_w, d.out_stride = out_h, out_w, out_stride
        d.out_off_y, d.out_off_x = out_off
        d.ntaps = len(taps)
        for i, (dy, dx, wt) in enumerate(taps):
            d.dy[i], d.dx[i], d.wtap[i] = dy, dx, wt
        d.w_batch_stride = w_batch_stride
        d.act, d.act_gain, d.clamp = act, act_gain, clamp
        d.noise_gain, d.residual_scale = noise_gain, residual_scale
        d.up_h, d.up_w = (out_h // 2, out_w // 2) if up_img is not None else (0, 0)
        return d

    d = _conv_desc(key, build)
    d.noise_gain = noise_gain
    _ok(_cabi.lib().hfagp_conv2d_fwd(C.byref(d), ptr(x), ptr(w), ptr(dcoef), ptr(noise), ptr(bias), ptr(residual),
                                       ptr(up_img), ptr(out), stream()), 'hfagp_conv2d_fwd')
    return out


TAPS_3X3 = tuple((ky - 1, kx - 1, ky * 3 + kx) for ky in range(3) for kx in range(3))
TAPS_1X1 = ((0, 0, 0),)


def _parity_taps(a: int, b: int):
    """Tap list of output parity class (a, b) of the stride-2 transposed 3x3 convolution:
    out[2m+a] += x[m - ky//2] * w[ky] for ky with ky % 2 == a."""
    kys = (0, 2) if a == 0 else (1,)
    kxs = (0, 2) if b == 0 else (1,)
    return tuple((-(ky // 2), -(kx // 2), ky * 3 + kx) for ky in kys for kx in kxs)


def conv_transpose_s2(x: torch.Tensor, w: torch.Tensor, cout: int, w_batch_stride: int,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv_transpose2d(stride=2, padding=0) of a 3x3 kernel, channels-last: [N,H,W,Ci] -> [N,2H+1,2W+1,Co]."""
    n, h, wd, _ = x.shape
    if out is None:
        out = torch.empty((n, 2 * h + 1, 2 * wd + 1, cout), device=x.device, dtype=torch.float32)
    for a in (0, 1):
        for b in (0, 1):
            conv2d(x, w, _parity_taps(a, b), cout, oh=h + 1 - a, ow=wd + 1 - b, out=out,
                   out_hw=(2 * h + 1, 2 * wd + 1), out_stride=2, out_off=(a, b), w_batch_stride=w_batch_stride)
    return out


def upfir_act(t: torch.Tensor, *, dcoef=None, noise=None, noise_gain=0.0, bias=None, act=ACT_LRELU,
              act_gain=math.sqrt(2.0), clamp=0.0, split_out: bool = False):
    n, th, tw, c = t.shape
    h2, w2 = th - 1, tw - 1
    if split_out:
        out = Split(torch.empty((n, h2, w2, c), device=t.device, dtype=torch.bfloat16),
                    torch.empty((n, h2, w2, c), device=t.device, dtype=torch.bfloat16))
        y, yh, yl = None, ptr(out.hi), ptr(out.lo)
    else:
        out = torch.empty((n, h2, w2, c), device=t.device, dtype=torch.float32)
        y, yh, yl = ptr(out), None, None
    _ok(_cabi.lib().hfagp_upfir_act_fwd(n, h2, w2, c, ptr(t), ptr(dcoef), ptr(noise), noise_gain, ptr(bias), act,
                                        act_gain, clamp, y, yh, yl, stream()), 'hfagp_upfir_act_fwd')
    return out


def torgb_small(x, wmod, bias, clamp, up_img, cout, want_mask=False):
    """``want_mask``: also return the clamp's derivative mask (bool [n,h,w,cout]) for the backward pass."""
    n, h, wd, cin = x.shape
    out = torch.empty((n, h, wd, cout), device=x.device, dtype=torch.float32)
    if isinstance(x, Split):
        xs = (None, ptr(x.hi), ptr(x.lo))
    else:
        xs = (ptr(x), None, None)
    if want_mask:
        mask = torch.empty((n, h, wd, cout), device=x.device, dtype=torch.bool)
        _ok(_cabi.lib().hfagp_torgb_small_mask_fwd(n, h, wd, cin, cout, *xs, ptr(wmod), ptr(bias), clamp, ptr(up_img),
                                                   ptr(out), ptr(mask), stream()), 'hfagp_torgb_small_mask_fwd')
        return out, mask
    _ok(_cabi.lib().hfagp_torgb_small_fwd(n, h, wd, cin, cout, *xs, ptr(wmod), ptr(bias), clamp, ptr(up_img),
                                          ptr(out), stream()), 'hfagp_torgb_small_fwd')
    return out


def torgb_finalize(acc, bias, clamp, up_img):
    """acc [n][h][w][k] (channel sums from conv2d_tc(..., rgb_acc=)) -> clamp(acc + bias) + upsample2d(up_img)."""
    n, h, wd, k = acc.shape
    out = torch.empty_like(acc)
    _ok(_cabi.lib().hfagp_torgb_finalize_fwd(n, h, wd, k, ptr(acc), ptr(bias), clamp, ptr(up_img), ptr(out), stream()),
        'hfagp_torgb_finalize_fwd')
    return out


class StyleTable:
    """Host-side layer table for ``hfagp_styles_fwd`` (built once per packed generator)."""

    def __init__(self, layers):
        # layers: list of (affine_weight, affine_bias, cin, ws_index, post_gain)
        n = len(layers)
        self.n = n
        self.keep = [(l[0], l[1]) for l in layers]
        self.aw = (C.c_void_p * n)(*[l[0].data_ptr() for l in layers])
        self.ab = (C.c_void_p * n)(*[l[1].data_ptr() for l in layers])
        self.cin = (C.c_int32 * n)(*[l[2] for l in layers])
        self.widx = (C.c_int32 * n)(*[l[3] for l in layers])
        self.gain = (C.c_float * n)(*[l[4] for l in layers])
        self.cins = [l[2] for l in layers]

    def offsets(self, b):
        offs, total = [], 0
        for c in self.cins:
            offs.append(total)
            total += b * c
        return offs, total

    def run_flat(self, ws: torch.Tensor):
        """-> (styles_flat [sum_l B*cin_l], offsets): layer l occupies [off_l, off_l + B*cin_l) as [B][cin_l]."""
        b, num_ws, w_dim = ws.shape
        offs, total = self.offsets(b)
        off_arr = (C.c_int64 * self.n)(*offs)
        styles = torch.empty(total, device=ws.device, dtype=torch.float32)
        _ok(_cabi.lib().hfagp_styles_fwd(self.n, b, num_ws, w_dim, ptr(ws), self.aw, self.ab, self.cin, self.widx,
                                           self.gain, off_arr, ptr(styles), stream()), 'hfagp_styles_fwd')
        return styles, offs

    def views(self, flat, offs, b):
        return [flat[o:o + b * c].view(b, c) for o, c in zip(offs, self.cins)]

    def run(self, ws: torch.Tensor):
        flat, offs = self.run_flat(ws)
        return self.views(flat, offs, ws.shape[0])


def modulate(w: torch.Tensor, styles: torch.Tensor, demodulate: bool):
    """w [taps][O][I] (shared), styles [B][I] -> wmod [B][taps][O][I], dcoef [B][O] or None."""
    taps, cout, cin = w.shape
    b = styles.shape[0]
    wmod = torch.empty((b, taps, cout, cin), device=w.device, dtype=torch.float32)
    dcoef = torch.empty((b, cout), device=w.device, dtype=torch.float32) if demodulate else None
    _ok(_cabi.lib().hfagp_modulate_fwd(b, taps, cout, cin, ptr(w), ptr(styles), ptr(wmod), ptr(dcoef), stream()),
          'hfagp_modulate_fwd')
    return wmod, dcoef


def render(planes, c, mlp, lin, jitter, u_fine, depth_range, *, res, s_coarse, s_fine, delta, box_scale,
           bookkeeping: bool = False, simt: bool = False):
    """planes [N,PH,PW,96] channels-last -> feat [N,res,res,32], depth [N,res*res], wsum [N,res*res].
    ``simt=True`` forces the legacy mma.sync kernel (``hfagp_render_fwd_simt``; cross-checks only)."""
    n, ph, pw, _ = planes.shape
    dev = planes.device
    rays = res * res
    feat = torch.empty((n, res, res, 32), device=dev, dtype=torch.float32)
    depth = torch.empty((n, rays), device=dev, dtype=torch.float32)
    wsum = torch.empty((n, rays), device=dev, dtype=torch.float32)
    book = {}
    if bookkeeping:
        t = s_coarse + s_fine
        if s_fine > 0:
            for k in ('inds', 'below', 'above'):
                book[k] = torch.empty((n * rays, s_fine), device=dev, dtype=torch.int32)
        book['sort_idx'] = torch.empty((n, rays, t), device=dev, dtype=torch.int32)
        book['depths_sorted'] = torch.empty((n, rays, t), device=dev, dtype=torch.float32)
    d = RenderDesc(n, res, ph, pw, s_coarse, s_fine, delta, box_scale)
    name = 'hfagp_render_fwd_simt' if simt else 'hfagp_render_fwd'
    _ok(getattr(_cabi.lib(), name)(C.byref(d), ptr(planes), ptr(c), ptr(mlp), ptr(lin), ptr(jitter), ptr(u_fine),
                                   ptr(depth_range), ptr(feat), ptr(depth), ptr(wsum), ptr(book.get('inds')), ptr(book.get('below')),
                                   ptr(book.get('above')), ptr(book.get('sort_idx')),
                                   ptr(book.get('depths_sorted')), stream()), name)
    return feat, depth, wsum, book


def render_bookkeeping(cdf, u, depths):
    """cdf [R,NC], u [R,SF], depths [R,T] (fp32, device) -> inds, below, above [R,SF], sort_idx [R,T] (int32): the
    renderer's integer stage on the given floats (``hfagp_render_bookkeeping``)."""
    r, nc = cdf.shape
    sf, t = u.shape[1], depths.shape[1]
    dev = cdf.device
    inds, below, above = (torch.empty((r, sf), device=dev, dtype=torch.int32) for _ in range(3))
    sort_idx = torch.empty((r, t), device=dev, dtype=torch.int32)
    _ok(_cabi.lib().hfagp_render_bookkeeping(r, nc, sf, t, ptr(cdf), ptr(u), ptr(depths), ptr(inds), ptr(below), ptr(above),
                                               ptr(sort_idx), stream()), 'hfagp_render_bookkeeping')
    return inds, below, above, sort_idx


def decoder_wgrad(f, do, mlp):
    """Per-sample decoder operands (f [S,32], do [S,33], see render_bwd(decoder=True)) -> gradient of the packed
    effective decoder weights [4257] (W0 [64,32], b0 [64], W1 [33,64], b1 [33]); see ``hfagp_decoder_wgrad``."""
    dmlp = zeros(mlp.shape, mlp.device)
    _ok(_cabi.lib().hfagp_decoder_wgrad(f.shape[0], ptr(f), ptr(do), ptr(mlp), ptr(dmlp), stream()), 'hfagp_decoder_wgrad')
    return dmlp


def render_bwd(planes, c, mlp, lin, jitter, u_fine, dfeat, *, res, s_coarse, s_fine, delta, box_scale, decoder=False):
    """d(feat) [N,res,res,32] -> d(planes) [N,PH,PW,96], see ``hfagp_render_bwd``.  ``decoder=True`` also returns the
    per-sample (features [S,32], d(raw decoder output) [S,33]) pair for the decoder's weight gradient."""
    n, ph, pw, _ = planes.shape
    dplanes = torch.zeros_like(planes)
    d = RenderDesc(n, res, ph, pw, s_coarse, s_fine, delta, box_scale)
    if decoder:
        total = n * res * res * (s_coarse + s_fine)
        # every sample of every VALID ray is written by the kernel; rays only go missing when res is not a multiple of the
        # 8 rays of a backward strip (then those rows must read as zero): skip 0.8 GB of memset per step otherwise
        alloc = torch.empty if res % 8 == 0 else torch.zeros
        f = alloc((total, 32), device=planes.device, dtype=torch.float32)
        do = alloc((total, 33), device=planes.device, dtype=torch.float32)
        _ok(_cabi.lib().hfagp_render_bwd_dec(C.byref(d), ptr(planes), ptr(c), ptr(mlp), ptr(lin), ptr(jitter), ptr(u_fine),
                                             ptr(dfeat), ptr(dplanes), ptr(f), ptr(do), stream()), 'hfagp_render_bwd_dec')
        return dplanes, f, do
    _ok(_cabi.lib().hfagp_render_bwd(C.byref(d), ptr(planes), ptr(c), ptr(mlp), ptr(lin), ptr(jitter), ptr(u_fine),
                                       ptr(dfeat), ptr(dplanes), stream()), 'hfagp_render_bwd')
    return dplanes


def blur(x, pad0, pad1, stride=1, split_out: bool = False, gain: float = 1.0):
    """gain * [1,3,3,1]^2/64 FIR of a channels-last activation (fp32 tensor or Split) -> fp32 or Split."""
    n, h, wd, c = x.shape
    oh = (h + pad0 + pad1 - 4) // stride + 1
    ow = (wd + pad0 + pad1 - 4) // stride + 1
    xin = (None, ptr(x.hi), ptr(x.lo)) if isinstance(x, Split) else (ptr(x), None, None)
    if split_out:
        out = Split(torch.empty((n, oh, ow, c), device=x.device, dtype=torch.bfloat16),
                    torch.empty((n, oh, ow, c), device=x.device, dtype=torch.bfloat16))
        yout = (None, ptr(out.hi), ptr(out.lo))
    else:
        out = torch.empty((n, oh, ow, c), device=x.device, dtype=torch.float32)
        yout = (ptr(out), None, None)
    _ok(_cabi.lib().hfagp_blur_fwd(n, h, wd, c, pad0, pad1, stride, gain, *xin, *yout, stream()), 'hfagp_blur_fwd')
    return out


def blur_up(dy, h, wd, pad0, pad1, stride, gain: float = 1.0):
    """Transpose of blur(x[n,h,wd,c], pad0, pad1, stride): dy [n,oh,ow,c] -> dx [n,h,wd,c]."""
    n, oh, ow, c = dy.shape
    assert oh == (h + pad0 + pad1 - 4) // stride + 1 and ow == (wd + pad0 + pad1 - 4) // stride + 1
    dx = torch.empty((n, h, wd, c), device=dy.device, dtype=torch.float32)
    _ok(_cabi.lib().hfagp_blur_up(n, h, wd, c, pad0, pad1, stride, gain, ptr(dy), ptr(dx), stream()), 'hfagp_blur_up')
    return dx


def linear(x, w, b, w_gain, b_gain):
    n, cin = x.shape
    cout = w.shape[0]
    out = torch.empty((n, cout), device=x.device, dtype=torch.float32)
    _ok(_cabi.lib().hfagp_linear_fwd(n, cin, cout, ptr(x), ptr(w), ptr(b), w_gain, b_gain, ptr(out), stream()),
          'hfagp_linear_fwd')
    return out


def _qr_workspace(k, m, dev):
    return torch.empty((_cabi.lib().hfagp_basis_qr_workspace_bytes(k, m),), device=dev, dtype=torch.uint8)


def basis_qr(bases, eps=1e-8, check_info=False):
    """``torch.qr(bases.T + eps)`` of get_latent without LAPACK (``hfagp_basis_qr_fwd``): bases [K, M] ->
    (q [M, K] with LAPACK's column signs, rinv [K, K] = R^-1 for the backward).  ``check_info`` reads the device flag a
    non-positive Cholesky pivot raises (synchronises; off inside a training step)."""
    k, m = bases.shape
    b = bases.detach().float().contiguous()
    q = torch.empty((m, k), device=b.device, dtype=torch.float32)
    rinv = torch.empty((k, k), device=b.device, dtype=torch.float32)
    ws = _qr_workspace(k, m, b.device)
    _ok(_cabi.lib().hfagp_basis_qr_fwd(k, m, ptr(b), float(eps), ptr(q), ptr(rinv), ptr(ws), int(_deterministic), stream()),
        'hfagp_basis_qr_fwd')
    _launches[0] += 4                     # Gram, k x k, apply + Gram, k x k + signs, apply
    if check_info:
        info = C.c_int(0)
        check(_cabi.lib().hfagp_basis_qr_info(ptr(ws), k, m, C.byref(info), stream()), 'hfagp_basis_qr_info')
        if info.value:
            raise _cabi.HfagpError('basis_qr: the basis is rank deficient or too ill conditioned for CholeskyQR2')
    return q, rinv


def basis_qr_bwd(gq, q, rinv):
    """Gradient of ``basis_qr``'s q w.r.t. bases: gq [M, K] -> gbases [K, M] (``hfagp_basis_qr_bwd``)."""
    m, k = q.shape
    gq = gq.float().contiguous()
    gb = torch.empty((k, m), device=q.device, dtype=torch.float32)
    ws = _qr_workspace(k, m, q.device)
    _ok(_cabi.lib().hfagp_basis_qr_bwd(k, m, ptr(gq), ptr(q), ptr(rinv), ptr(gb), ptr(ws), int(_deterministic), stream()),
        'hfagp_basis_qr_bwd')
    _launches[0] += 2
    return gb


def latent(weights, q, delta, dim_total):
    n, k = weights.shape
    out = torch.empty((n, dim_total), device=weights.device, dtype=torch.float32)
    _ok(_cabi.lib().hfagp_latent_fwd(n, k, dim_total, ptr(weights), ptr(q), ptr(delta), ptr(out), stream()),
          'hfagp_latent_fwd')
    return out


def stem_conv1x1(x_nchw, w, bias, act=ACT_LRELU, act_gain=math.sqrt(2.0)) -> 'Split':
    """NCHW frame [N,cin<=4,H,W] -> split-bf16 channels-last [N,H,W,cout] through a 1x1 convolution + bias + activation
    (``hfagp_stem_conv1x1_fwd``); ``w`` [cout][cin] fp32."""
    n, cin, h, wd = x_nchw.shape
    cout = w.shape[0]
    hi = torch.empty((n, h, wd, cout), device=x_nchw.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    _ok(_cabi.lib().hfagp_stem_conv1x1_fwd(n, h, wd, cin, cout, ptr(x_nchw), ptr(w), ptr(bias), act, act_gain, ptr(hi), ptr(lo),
                                           stream()), 'hfagp_stem_conv1x1_fwd')
    return Split(hi, lo)


def nchw_to_nhwc(x):
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), device=x.device, dtype=torch.float32)
    _ok(_cabi.lib().hfagp_nchw_to_nhwc(n, c, h, w, ptr(x), ptr(out), stream()), 'hfagp_nchw_to_nhwc')
    return out


def nhwc_to_nchw(x):
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
    _ok(_cabi.lib().hfagp_nhwc_to_nchw(n, c, h, w, ptr(x), ptr(out), stream()), 'hfagp_nhwc_to_nchw')
    return out


# ------------------------------------------------------------------ tensor-core (split-bf16) path

class Split:
    """An fp32 tensor carried as two bf16 tensors (hi + lo) — the operand format of the tcgen05 path."""
    __slots__ = ('hi', 'lo')

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    @property
    def device(self):
        return self.hi.device

    def float(self):
        return self.hi.float() + self.lo.float()


def split(x: torch.Tensor) -> Split:
    hi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    lo = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _ok(_cabi.lib().hfagp_split_bf16(x.numel(), ptr(x), ptr(hi), ptr(lo), stream()), 'hfagp_split_bf16')
    return Split(hi, lo)


def pack_conv_weight(weight: torch.Tensor, scale: float, want_split=True, want_t=True):
    """torch conv weight [O,I,kh,kw] * scale -> (pk [kh*kw,O,I] fp32, Split(pk) | None, Split(pk transposed to
    [kh*kw,I,O]) | None) in one pass (``hfagp_pack_conv_weight``)."""
    o, i, kh, kw = weight.shape
    t = kh * kw
    w = weight.detach().float().contiguous()
    dev = w.device
    pk = torch.empty((t, o, i), device=dev, dtype=torch.float32)
    sp = Split(torch.empty((t, o, i), device=dev, dtype=torch.bfloat16), torch.empty((t, o, i), device=dev, dtype=torch.bfloat16)) if want_split else None
    spt = Split(torch.empty((t, i, o), device=dev, dtype=torch.bfloat16), torch.empty((t, i, o), device=dev, dtype=torch.bfloat16)) if want_t else None
    _ok(_cabi.lib().hfagp_pack_conv_weight(o, i, t, ptr(w), float(scale), ptr(pk), ptr(sp.hi) if sp else None,
                                           ptr(sp.lo) if sp else None, ptr(spt.hi) if spt else None,
                                           ptr(spt.lo) if spt else None, stream()), 'hfagp_pack_conv_weight')
    return pk, sp, spt


def unpack_conv_wgrad(dwp: torch.Tensor, grad: torch.Tensor):
    """grad [O,I,kh,kw] += dwp [kh*kw,O,I_padded] in place (``hfagp_unpack_conv_wgrad``)."""
    o, i, kh, kw = grad.shape
    t, o2, ip = dwp.shape
    assert t == kh * kw and o2 == o and ip >= i
    _ok(_cabi.lib().hfagp_unpack_conv_wgrad(o, i, ip, t, ptr(dwp), ptr(grad), stream()), 'hfagp_unpack_conv_wgrad')


def modulate_split(w: torch.Tensor, styles: torch.Tensor, demodulate: bool):
    """w [taps][O][I] fp32 (shared), styles [B][I] -> Split wmod [B][taps][O][I], dcoef [B][O] or None."""
    taps, cout, cin = w.shape
    b = styles.shape[0]
    hi = torch.empty((b, taps, cout, cin), device=w.device, dtype=torch.bfloat16)
    lo = torch.empty((b, taps, cout, cin), device=w.device, dtype=torch.bfloat16)
    dcoef = torch.empty((b, cout), device=w.device, dtype=torch.float32) if demodulate else None
    _ok(_cabi.lib().hfagp_modulate_split_fwd(b, taps, cout, cin, ptr(w), ptr(styles), ptr(hi), ptr(lo), ptr(dcoef),
                                             stream()), 'hfagp_modulate_split_fwd')
    return Split(hi, lo), dcoef


NUM_SMS = 148          # B200; the value used when no device can be asked (host-logic tests)
_sm_count = {}
_deterministic = os.environ.get('HFAGP_DETERMINISTIC', '0') not in ('', '0')


def set_deterministic(on: bool = True) -> bool:
    """Bit-reproducible frames (or HFAGP_DETERMINISTIC=1): the forward path stops using the two features whose fp32
    atomics make the summation order vary run to run — split-K for the layers with few output tiles (they run as plain
    tcgen05 convolutions on the SMs their tiles cover) and the ToRGB fused into the super-resolution conv1 epilogue (the
    separate ToRGB kernel reads the layer output instead).  Results stay within the parity tolerance of the default
    path, ~10 % slower per frame at batch 1.  Decide BEFORE a frame graph is captured.  Returns the previous setting.
    (Gradient kernels still accumulate with red.global.add; the reference's own CUDA backward is not deterministic
    either.)"""
    global _deterministic
    prev, _deterministic = _deterministic, bool(on)
    return prev


def deterministic() -> bool:
    return _deterministic


def num_sms() -> int:
    """SM count of the current device (``hfagp_device_sm_count``, cached per device ordinal)."""
    if not torch.cuda.is_available():
        return NUM_SMS
    dev = torch.cuda.current_device()
    if dev not in _sm_count:
        _sm_count[dev] = int(_cabi.lib().hfagp_device_sm_count())
    return _sm_count[dev]


def _ksplit(tiles: int, cin: int, taps, in_stride: int, sms: int = None) -> int:
    """Split-K factor for a tensor-core convolution with ``tiles`` output tiles: 1 unless the tiles would leave
    three quarters of the SMs idle.  Units of K = 64-channel chunks x tap groups (taps sharing dx at stride 1)."""
    if _deterministic:
        return 1
    sms = num_sms() if sms is None else sms
    groups = len({t[1] for t in taps}) if in_stride == 1 else len(taps)
    units = -(-cin // 64) * groups
    if tiles * 4 > sms or units < 4:
        return 1
    return max(1, min(units, sms // tiles))


def modulate_split_multi(entries, styles_flat, batch):
    """entries: list of (w [taps][O][I] fp32, styles offset, demodulate) -> list of (Split wmod [B][taps][O][I], dcoef
    [B][O] or None), all layers in ONE launch (``hfagp_modulate_split_multi_fwd``)."""
    n = len(entries)
    dev = styles_flat.device
    outs, his, los, dcs = [], [], [], []
    for w, off, demod in entries:
        taps, cout, cin = w.shape
        hi = torch.empty((batch, taps, cout, cin), device=dev, dtype=torch.bfloat16)
        lo = torch.empty((batch, taps, cout, cin), device=dev, dtype=torch.bfloat16)
        dc = torch.empty((batch, cout), device=dev, dtype=torch.float32) if demod else None
        outs.append((Split(hi, lo), dc))
        his.append(hi.data_ptr()); los.append(lo.data_ptr()); dcs.append(dc.data_ptr() if demod else None)
    vp = C.c_void_p
    _ok(_cabi.lib().hfagp_modulate_split_multi_fwd(
        n, batch, (vp * n)(*[e[0].data_ptr() for e in entries]), (C.c_int32 * n)(*[e[0].shape[0] for e in entries]),
        (C.c_int32 * n)(*[e[0].shape[1] for e in entries]), (C.c_int32 * n)(*[e[0].shape[2] for e in entries]),
        (C.c_int64 * n)(*[e[1] for e in entries]), ptr(styles_flat), (vp * n)(*his), (vp * n)(*los), (vp * n)(*dcs),
        stream()), 'hfagp_modulate_split_multi_fwd')
    return outs


def conv2d_tc(x: Split, w: Split, taps, cout: int, *, oh: int, ow: int, in_stride: int = 1, out=None,
              out_hw=None, out_stride: int = 1, out_off=(0, 0), w_batched: bool = False, split_out: bool = False,
              dcoef=None, noise=None, noise_gain: float = 0.0, bias=None, act: int = ACT_LINEAR,
              act_gain: float = 1.0, clamp: float = 0.0, residual=None, residual_scale: float = 1.0, up_img=None,
              rgb_w=None, rgb_acc=None):
    """tcgen05 implicit-GEMM convolution on split-bf16 operands; same semantics as conv2d().
    ``rgb_w [n][k][cout]`` / ``rgb_acc [n][oh][ow][k]`` (zeroed): also accumulate the block's small ToRGB from the
    finished activations (``hfagp_conv2d_tc_rgb_fwd``)."""
    n, h, wd, cin = x.shape
    out_h, out_w = out_hw if out_hw is not None else (oh, ow)
    w_taps_total = w.shape[-3]
    taps = tuple(taps)
    if out is None:
        if split_out:
            out = Split(torch.empty((n, out_h, out_w, cout), device=x.device, dtype=torch.bfloat16),
                        torch.empty((n, out_h, out_w, cout), device=x.device, dtype=torch.bfloat16))
        else:
            out = torch.empty((n, out_h, out_w, cout), device=x.device, dtype=torch.float32)
    wbs = w_taps_total * cout * cin if w_batched else 0
    key = ('tc', n, h, wd, cin, cout, oh, ow, in_stride, out_h, out_w, out_stride, out_off, taps, wbs, act, act_gain,
           clamp, residual_scale, up_img is not None)

    def build():
        d = ConvDesc()
        d.batch, d.in_h, d.in_w, d.cin, d.cout = n, h, wd, cin, cout
        d.oh, d.ow, d.in_stride = oh, ow, in_stride
        d.out_h, d.out_w, d.out_stride = out_h, out_w, out_stride
        d.out_off_y, d.out_off_x = out_off
        d.ntaps = len(taps)
        for i, (dy, dx, wt) in enumerate(taps):
            d.dy[i], d.dx[i], d.wtap[i] = dy, dx, wt
        d.w_batch_stride = wbs
        d.act, d.act_gain, d.clamp = act, act_gain, clamp
        d.noise_gain, d.residual_scale = noise_gain, residual_scale
        d.up_h, d.up_w = (out_h // 2, out_w // 2) if up_img is not None else (0, 0)
        return d

    d = _conv_desc(key, build)
    d.noise_gain = noise_gain
    is_split = isinstance(out, Split)
    if rgb_acc is not None:
        _ok(_cabi.lib().hfagp_conv2d_tc_rgb_fwd(C.byref(d), ptr(x.hi), ptr(x.lo), ptr(w.hi), ptr(w.lo), w_taps_total,
                                                ptr(dcoef), ptr(noise), ptr(bias), None if is_split else ptr(out),
                                                ptr(out.hi) if is_split else None, ptr(out.lo) if is_split else None,
                                                ptr(rgb_w), rgb_w.shape[-2], ptr(rgb_acc), stream()),
            'hfagp_conv2d_tc_rgb_fwd')
        return out
    ksplit = _ksplit(n * -(-oh * ow // 128) * -(-cout // 128), cin, taps, in_stride)
    if ksplit > 1 and out_stride == 1 and (oh, ow) == (out_h, out_w) and cout % 4 == 0:
        # few output tiles, long K (the 4^2..32^2 layers): split K over the idle SMs, then one elementwise epilogue
        acc = zeros((n, out_h, out_w, cout), x.device)
        _ok(_cabi.lib().hfagp_conv2d_tc_acc_fwd(C.byref(d), 1, ptr(x.hi), ptr(x.lo), ptr(w.hi), ptr(w.lo), w_taps_total,
                                                ksplit, ptr(acc), stream()), 'hfagp_conv2d_tc_acc_fwd')
        y = None if is_split else out
        _ok(_cabi.lib().hfagp_conv_epilogue_fwd(C.byref(d), ptr(acc), ptr(dcoef), ptr(noise), ptr(bias), ptr(residual),
                                                ptr(up_img), ptr(y), ptr(out.hi) if is_split else None,
                                                ptr(out.lo) if is_split else None, stream()), 'hfagp_conv_epilogue_fwd')
        return out
    _ok(_cabi.lib().hfagp_conv2d_tc_fwd(C.byref(d), ptr(x.hi), ptr(x.lo), ptr(w.hi), ptr(w.lo), w_taps_total,
                                        ptr(dcoef), ptr(noise), ptr(bias), ptr(residual), ptr(up_img),
                                        None if is_split else ptr(out), ptr(out.hi) if is_split else None,
                                        ptr(out.lo) if is_split else None, stream()), 'hfagp_conv2d_tc_fwd')
    return out


# ------------------------------------------------------------------ backward path

def _act_in(y):
    return (None, ptr(y.hi), ptr(y.lo)) if isinstance(y, Split) else (ptr(y), None, None)


def act_bwd(y, g0=None, s0=None, g1=None, s1=None, dimg=None, wrgb=None, srgb=None, dcoef=None, noise=None,
            noise_gain: float = 0.0, bias=None, residual=None, residual_scale: float = 1.0, act: int = ACT_LRELU,
            act_gain: float = math.sqrt(2.0), clamp: float = 0.0, post_scale: float = 1.0, out: str = 'f32',
            ds0=None, ds1=None, dsrgb=None, dbias=None, ddcoef=None):
    """See ``hfagp_act_bwd`` in include/hfagp.h.  ``out``: 'f32' | 'split' | 'none' selects the form of dz."""
    if g0 is None and g1 is not None:        # a single incoming gradient always travels in slot 0
        g0, s0, ds0, g1, s1, ds1 = g1, s1, ds1, None, None, None
    n, h, wd, c = y.shape
    d = ActBwdDesc(n, h, wd, c, act, act_gain, clamp, noise_gain, residual_scale if residual is not None else 0.0,
                   post_scale, dimg.shape[-1] if dimg is not None else 0)
    dz, dzp = None, (None, None, None)
    if out == 'split':
        dz = Split(torch.empty((n, h, wd, c), device=y.device, dtype=torch.bfloat16),
                   torch.empty((n, h, wd, c), device=y.device, dtype=torch.bfloat16))
        dzp = (None, ptr(dz.hi), ptr(dz.lo))
    elif out == 'f32':
        dz = torch.empty((n, h, wd, c), device=y.device, dtype=torch.float32)
        dzp = (ptr(dz), None, None)
    _ok(_cabi.lib().hfagp_act_bwd(C.byref(d), *_act_in(y), ptr(g0), ptr(s0), ptr(g1), ptr(s1), ptr(dimg), ptr(wrgb),
                                    ptr(srgb), ptr(dcoef), ptr(noise), ptr(bias), ptr(residual), *dzp, ptr(ds0),
                                    ptr(ds1), ptr(dsrgb), ptr(dbias), ptr(ddcoef), stream()), 'hfagp_act_bwd')
    return dz


def demod_bwd(w2, styles, dcoef, ddcoef, dstyles):
    """dstyles[n][i] -= styles[n][i] * sum_o ddcoef[n][o] dcoef[n][o]^3 w2[o][i]   (in place)."""
    cout, cin = w2.shape
    _ok(_cabi.lib().hfagp_demod_bwd(styles.shape[0], cout, cin, ptr(w2), ptr(styles), ptr(dcoef), ptr(ddcoef),
                                      ptr(dstyles), stream()), 'hfagp_demod_bwd')


def modconv_wgrad_finish(dw, w, grad, *, transposed=False, ddcoef=None, dcoef=None, styles=None):
    """grad[O][I][k][k] += unpack(dw) - w * sum_n ddcoef dcoef^3 styles^2 (see ``hfagp_modconv_wgrad_finish``).
    ``dw``: packed [taps][O][I] ([taps][I][O] when ``transposed``); ``w``: the packed unmodulated weight [taps][O][I]."""
    taps, cout, cin = w.shape          # (only the shape of ``w`` is used when there is no demodulation term)
    batch = styles.shape[0] if styles is not None else 0
    _ok(_cabi.lib().hfagp_modconv_wgrad_finish(taps, cout, cin, batch, ptr(dw), 1 if transposed else 0, ptr(w), ptr(ddcoef),
                                                 ptr(dcoef), ptr(styles), ptr(grad), stream()), 'hfagp_modconv_wgrad_finish')


def linear_bwd(dy, x, w, w_gain, b_gain, need_dx=True, dw=None, db=None):
    n, cout = dy.shape
    cin = w.shape[1]
    dx = torch.empty((n, cin), device=dy.device, dtype=torch.float32) if need_dx else None
    _ok(_cabi.lib().hfagp_linear_bwd(n, cin, cout, ptr(dy), ptr(x), ptr(w), w_gain, b_gain, ptr(dx), ptr(dw), ptr(db),
                                       stream()), 'hfagp_linear_bwd')
    return dx


def conv2d_wgrad(x, dz, taps, dw, *, oh, ow, in_stride=1, scale=1.0, xscale=None, dzscale=None):
    """dw [taps_total][cout][cin] += scale * sum dz (x) x, see ``hfagp_conv2d_wgrad`` (``xscale`` [n][cin] / ``dzscale``
    [n][cout]: per-sample style factors of a modulated convolution, ``hfagp_conv2d_wgrad_mod``)."""
    n, h, wd, cin = x.shape
    cout = dz.shape[-1]
    taps = tuple(taps)
    key = ('wg', n, h, wd, cin, cout, oh, ow, in_stride, taps)

    def build():
        d = ConvDesc()
        d.batch, d.in_h, d.in_w, d.cin, d.cout = n, h, wd, cin, cout
        d.oh, d.ow, d.in_stride = oh, ow, in_stride
        d.out_h, d.out_w, d.out_stride = oh, ow, 1
        d.ntaps = len(taps)
        for i, (dy_, dx_, wt) in enumerate(taps):
            d.dy[i], d.dx[i], d.wtap[i] = dy_, dx_, wt
        return d

    d = _conv_desc(key, build)
    if xscale is not None or dzscale is not None:
        _ok(_cabi.lib().hfagp_conv2d_wgrad_mod(C.byref(d), *_act_in(x), *_act_in(dz), ptr(xscale), ptr(dzscale), scale,
                                               ptr(dw), stream()), 'hfagp_conv2d_wgrad_mod')
        return dw
    _ok(_cabi.lib().hfagp_conv2d_wgrad(C.byref(d), *_act_in(x), *_act_in(dz), scale, ptr(dw), stream()),
          'hfagp_conv2d_wgrad')
    return dw


class StyleTableBwd:
    """dstyles (flat, layout of StyleTable.run) -> dws, all layers in one launch."""

    def __init__(self, table: 'StyleTable'):
        self.t = table

    def run(self, dstyles_flat, offs, batch, num_ws, w_dim):
        t = self.t
        dws = zeros((batch, num_ws, w_dim), dstyles_flat.device)
        off_arr = (C.c_int64 * t.n)(*offs)
        _ok(_cabi.lib().hfagp_styles_bwd(t.n, batch, num_ws, w_dim, t.aw, t.cin, t.widx, t.gain, off_arr,
                                           ptr(dstyles_flat), ptr(dws), stream()), 'hfagp_styles_bwd')
        return dws


def conv_transpose_s2_tc(x: Split, w: Split, cout: int, w_batched: bool = False) -> torch.Tensor:
    """conv_transpose_s2() on the tcgen05 path: the four output-parity classes as ONE persistent launch
    (``hfagp_conv2d_tc_multi_fwd``) scattering into [N,2H+1,2W+1,Co]."""
    n, h, wd, cin = x.shape
    out = torch.empty((n, 2 * h + 1, 2 * wd + 1, cout), device=x.device, dtype=torch.float32)
    w_taps_total = w.shape[-3]
    wbs = w_taps_total * cout * cin if w_batched else 0
    key = ('tcup', n, h, wd, cin, cout, wbs)

    def build():
        arr = (ConvDesc * 4)()
        i = 0
        for a in (0, 1):
            for b in (0, 1):
                d = arr[i]
                i += 1
                d.batch, d.in_h, d.in_w, d.cin, d.cout = n, h, wd, cin, cout
                d.oh, d.ow, d.in_stride = h + 1 - a, wd + 1 - b, 1
                d.out_h, d.out_w, d.out_stride = 2 * h + 1, 2 * wd + 1, 2
                d.out_off_y, d.out_off_x = a, b
                taps = _parity_taps(a, b)
                d.ntaps = len(taps)
                for k, (dy, dx, wt) in enumerate(taps):
                    d.dy[k], d.dx[k], d.wtap[k] = dy, dx, wt
                d.w_batch_stride = wbs
                d.act, d.act_gain, d.clamp, d.noise_gain, d.residual_scale = ACT_LINEAR, 1.0, 0.0, 0.0, 1.0
        return arr

    arr = _conv_desc(key, build)
    tiles = n * -(-cout // 128) * sum(-(-(h + 1 - a) * (wd + 1 - b) // 128) for a in (0, 1) for b in (0, 1))
    ksplit = _ksplit(tiles, cin, ((0, 0, 0),), 1)          # the smallest class has one tap group: units = K chunks
    if ksplit > 1:
        out.zero_()
        _ok(_cabi.lib().hfagp_conv2d_tc_acc_fwd(arr, 4, ptr(x.hi), ptr(x.lo), ptr(w.hi), ptr(w.lo), w_taps_total,
                                                ksplit, ptr(out), stream()), 'hfagp_conv2d_tc_acc_fwd')
        return out
    _ok(_cabi.lib().hfagp_conv2d_tc_multi_fwd(arr, 4, ptr(x.hi), ptr(x.lo), ptr(w.hi), ptr(w.lo), w_taps_total, None,
                                              None, None, None, None, ptr(out), None, None, stream()),
        'hfagp_conv2d_tc_multi_fwd')
    return out


# ------------------------------------------------------------------ training step: latent backward, loss, optimiser

def latent_bwd(dws, weights, q, need_dweights=True, need_dq=True, need_ddelta=True):
    """dws [B,dim] -> (dweights [B,K], dq [dim,K], ddelta [dim]), see ``hfagp_latent_bwd``."""
    n, dim = dws.shape
    k = q.shape[1]
    dev = dws.device
    dweights = torch.empty((n, k), device=dev, dtype=torch.float32) if need_dweights else None
    dq = torch.empty((dim, k), device=dev, dtype=torch.float32) if need_dq else None
    ddelta = torch.empty((dim,), device=dev, dtype=torch.float32) if need_ddelta else None
    _ok(_cabi.lib().hfagp_latent_bwd(n, k, dim, ptr(dws), ptr(weights), ptr(q), ptr(dweights), ptr(dq), ptr(ddelta),
                                       stream()), 'hfagp_latent_bwd')
    return dweights, dq, ddelta


def facepool(x_nhwc, size):
    """AdaptiveAvgPool2d((size,size)) for an integer factor, channels-last in -> NCHW out."""
    n, h, wd, c = x_nhwc.shape
    if h % size or wd % size or h // size != wd // size:
        raise _cabi.HfagpError(f'face_pool {h}x{wd} -> {size} is not an integer-factor average')
    f = h // size
    out = torch.empty((n, c, size, size), device=x_nhwc.device, dtype=torch.float32)
    _ok(_cabi.lib().hfagp_facepool_fwd(n, h, wd, c, f, ptr(x_nhwc), ptr(out), stream()), 'hfagp_facepool_fwd')
    return out


def facepool_bwd(dy_nchw, h, wd):
    n, c, oh, ow = dy_nchw.shape
    dx = torch.empty((n, h, wd, c), device=dy_nchw.device, dtype=torch.float32)
    _ok(_cabi.lib().hfagp_facepool_bwd(n, h, wd, c, h // oh, ptr(dy_nchw), ptr(dx), stream()), 'hfagp_facepool_bwd')
    return dx


def mse(a, b):
    loss = zeros((1,), a.device).view(())
    _ok(_cabi.lib().hfagp_mse_fwd(a.numel(), ptr(a), ptr(b), 1.0 / a.numel(), ptr(loss), stream()), 'hfagp_mse_fwd')
    return loss


def mse_bwd(a, b, gout, out=None):
    """d(mse)/da * gout (device scalar); accumulated into ``out`` when given."""
    acc = out is not None
    if out is None:
        out = torch.empty_like(a)
    _ok(_cabi.lib().hfagp_mse_bwd(a.numel(), ptr(a), ptr(b), 1.0 / a.numel(), ptr(gout), int(acc), ptr(out), stream()),
          'hfagp_mse_bwd')
    return out


def adam_step(p, g, m, v, *, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """In-place torch.optim.Adam update of the flat buffer ``p`` (see ``hfagp_adam_step``)."""
    _ok(_cabi.lib().hfagp_adam_step(p.numel(), ptr(p), ptr(g), ptr(m), ptr(v), grad_scale, lr, beta1, beta2, eps,
                                      weight_decay, int(step), stream()), 'hfagp_adam_step')


def adam_sched(lr, beta1, beta2, step, sched_host, offset=0):
    """Host only: sched_host[offset:offset+2] = (lr / (1 - beta1^step), sqrt(1 - beta2^step)), as ``hfagp_adam_step``
    evaluates them (``hfagp_adam_sched``)."""
    check(_cabi.lib().hfagp_adam_sched(lr, beta1, beta2, int(step), sched_host.data_ptr() + 4 * offset), 'hfagp_adam_sched')


def adam_step_dev(p, g, m, v, sched, *, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """``adam_step`` with the step-dependent scalars read from the device pair ``sched`` (``hfagp_adam_step_dev``)."""
    _ok(_cabi.lib().hfagp_adam_step_dev(p.numel(), ptr(p), ptr(g), ptr(m), ptr(v), grad_scale, beta1, beta2, eps,
                                          weight_decay, ptr(sched), stream()), 'hfagp_adam_step_dev')
