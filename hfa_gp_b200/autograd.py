"""Backward of the generator's convolution stacks (StyleGAN2 backbone and super-resolution head) for the
training step ``Trainer.gen_update`` (``/root/reference/code/trainer_rgb.py:73-98``): forward through
``generator.synthesis`` (``headnerf.py:112``) then ``loss.backward()``.

With the generator frozen (the reference's state until ``tune_iter``, ``trainer_rgb.py:59-60``) the gradient
that has to reach the encoder / ``bases`` / ``delta`` flows  image -> super-resolution -> feature image ->
renderer -> planes -> backbone -> styles -> ws.  Each network stage is one ``torch.autograd.Function`` whose
forward is the inference code (recording a tape of the activations it produced anyway) and whose backward walks
the tape with the sm_100a kernels:

  * ``hfagp_act_bwd``     everything between two convolutions (activation/clamp derivative, demodulation, the
                          per-sample style scales of the consumers, and the d(styles)/d(dcoef) reductions);
  * data-gradient convs   the forward kernels (tcgen05 split-bf16 or fp32 SIMT) on transposed, *unmodulated*
                          weights: modulation is a per-(sample, channel) scale, applied by ``hfagp_act_bwd`` of
                          the layer below, so one shared weight tensor serves the whole batch;
  * ``hfagp_blur_fwd``    transposes of the [1,3,3,1] FIRs (symmetric filter: same kernel, other pads);
  * ``hfagp_demod_bwd`` / ``hfagp_styles_bwd``   back to ``ws``.
"""
from __future__ import annotations

import math

import torch

from . import ops
from ._cabi import ACT_LINEAR, ACT_LRELU, HfagpError

SQRT2 = math.sqrt(2.0)

TAPS_3X3_T = tuple((-(ky - 1), -(kx - 1), ky * 3 + kx) for ky in range(3) for kx in range(3))   # dgrad of a same-conv
TAPS_UP_T = tuple((ky, kx, ky * 3 + kx) for ky in range(3) for kx in range(3))                  # dgrad of the stride-2 transposed conv


def _dgrad(gen, dz, pl, taps, *, oh, ow, in_stride=1):
    """Unscaled data gradient  dxu = conv(dz, W^T)  of one layer: [B,*,*,cout] -> fp32 [B,oh,ow,cin]."""
    bw = pl.bwd()
    if gen.precision == 'tc' and bw['wT_split'] is not None:
        dzs = dz if isinstance(dz, ops.Split) else ops.split(dz)
        return ops.conv2d_tc(dzs, bw['wT_split'], taps, pl.cin, oh=oh, ow=ow, in_stride=in_stride)
    dzf = dz.float() if isinstance(dz, ops.Split) else dz
    return ops.conv2d(dzf, bw['wT'], taps, pl.cin, oh=oh, ow=ow, in_stride=in_stride)


def _acc(param, g):
    """param.grad += g (shape of the parameter); the way loss.backward() would leave it."""
    g = g.reshape(param.shape).to(param.dtype)
    if param.grad is None:
        param.grad = g.clone()
    else:
        param.grad.add_(g)


def _unpack_conv(dw, k):
    """packed [k*k][O][I] -> parameter layout [O][I][k][k]"""
    t, o, i = dw.shape
    return dw.view(k, k, o, i).permute(2, 3, 0, 1)


def _demod_term(pl, styles, dcoef, ddc):
    """d(dcoef)/dW folded back: -W[t][o][i] * sum_n ddcoef[n][o] dcoef[n][o]^3 styles[n][i]^2   (tiny tensors)"""
    coef = ((ddc * dcoef * dcoef * dcoef)[:, :, None] * (styles * styles)[:, None, :]).sum(0)   # [O, I]; batch is tiny
    return -pl.w * coef[None]


def _layer_param_grads(m, rec, dz, ddc, db, h, w, up):
    """Gradients of one modulated SynthesisLayer's own parameters (the post-tune_iter regime, train_rgb.py:132-134):
    weight (style-scaled wgrad + demodulation term), bias, noise_strength."""
    pl, styles, dcoef = rec['pl'], rec['styles'], rec['dcoef']
    x = rec['x']
    if up:
        # transposed conv: dW[(ky,kx)][o][i] = sum dt[2m+k][o] * x[m][i] * s[i]  — wgrad with the roles swapped
        # (dense operand = layer input, strided operand = gradient), result [t][i][o]
        hin, win = x.shape[1], x.shape[2]
        dwp = ops.zeros((9, pl.cin, pl.cout), pl.w.device)
        ops.conv2d_wgrad(dz, x, TAPS_UP_T, dwp, oh=hin, ow=win, in_stride=2, dzscale=styles)
    else:
        dwp = ops.zeros(pl.w.shape, pl.w.device)
        ops.conv2d_wgrad(x, dz, ops.TAPS_3X3, dwp, oh=h, ow=w, xscale=styles)
    # + demodulation term, unpacked into [O][I][k][k] and added to .grad in one pass (hfagp_modconv_wgrad_finish)
    if m.weight.grad is None:
        m.weight.grad = torch.zeros_like(m.weight)
    if m.weight.grad.is_contiguous() and m.weight.grad.dtype == torch.float32:
        ops.modconv_wgrad_finish(dwp, pl.w, m.weight.grad, transposed=up, ddcoef=ddc.contiguous(), dcoef=dcoef.contiguous(),
                                 styles=styles.contiguous())
    else:
        full = dwp.transpose(1, 2) if up else dwp
        _acc(m.weight, _unpack_conv(full + _demod_term(pl, styles, dcoef, ddc), 3))
    _acc(m.bias, db)
    if rec.get('noise_buf') is not None:
        dzf = dz.float() if isinstance(dz, ops.Split) else dz
        if up:
            raise HfagpError('internal: noise gradient of an up layer must be taken from the post-FIR gradient')
        dpre_sum = (dzf / dcoef[:, None, None, :]).sum(-1)
        _acc(m.noise_strength, (dpre_sum * rec['noise_buf'][None]).sum())


def _torgb_param_grads(m, rt, dimg_t, h, w):
    pl = rt['pl']
    x = rt['x']
    k = pl.cout
    kp = (k + 3) // 4 * 4
    dz = dimg_t.contiguous()
    if kp != k:
        dz = torch.nn.functional.pad(dz, (0, kp - k))
    dwp = ops.zeros((1, kp, pl.cin), dz.device)
    ops.conv2d_wgrad(x, dz, ops.TAPS_1X1, dwp, oh=h, ow=w, xscale=rt['styles'])
    _acc(m.weight, dwp[0, :k])
    _acc(m.bias, dimg_t.sum((0, 1, 2)))


def blocks_backward(gen, recs, dimg, dviews, wgrads=False):
    """Walk the recorded blocks in reverse.  ``dimg``: gradient of the last block's output image (fp32 NHWC).
    ``dviews[i]``: [B,cin] accumulator of d(styles) of layer i.  Returns (gx, dimg_in): the unscaled data gradient
    into the first block's input with its style scale / d(styles) view (or None), and the gradient of the first
    block's input image (or None).  ``wgrads``: also accumulate the gradients of the blocks' own parameters into
    their ``.grad`` (generator unfrozen by tune_generator())."""
    gx = None
    tc = gen.precision == 'tc'
    for blk in reversed(recs):
        r0, r1, rt = blk['conv0'], blk['conv1'], blk['torgb']
        pl1, plt = r1['pl'], rt['pl']
        y1 = r1['y']
        b, h, w = y1.shape[0], y1.shape[1], y1.shape[2]
        dev = dimg.device
        # ---- ToRGB branch
        dimg_t = dimg * rt['mask'] if rt['mask'] is not None else dimg
        kw = {}
        if rt['small']:
            kw = dict(dimg=dimg_t.contiguous(), wrgb=plt.w[0], srgb=rt['styles'], dsrgb=dviews[plt.index])
        else:
            dxu_rgb = _dgrad(gen, dimg_t, plt, ops.TAPS_1X1, oh=h, ow=w)
            kw = dict(g1=dxu_rgb, s1=rt['styles'], ds1=dviews[plt.index])
        if wgrads:
            _torgb_param_grads(blk['mods'][2], rt, dimg_t, h, w)
        if gx is not None:
            kw.update(g0=gx[0], s0=gx[1], ds0=gx[2])
        # ---- conv1: activation / demodulation backward, then its data gradient
        ddc = ops.zeros((b, pl1.cout), dev)
        db1 = ops.zeros((pl1.cout,), dev) if wgrads else None
        use_tc1 = tc and pl1.bwd()['wT_split'] is not None
        dz1 = ops.act_bwd(y1, dcoef=r1['dcoef'], noise=r1['noise'], noise_gain=pl1.noise_gain, bias=pl1.bias,
                          act=ACT_LRELU, act_gain=SQRT2, clamp=pl1.clamp, out='split' if use_tc1 else 'f32',
                          ddcoef=ddc, dbias=db1, **kw)
        ops.demod_bwd(pl1.bwd()['w2'], r1['styles'], r1['dcoef'], ddc, dviews[pl1.index])
        if wgrads:
            _layer_param_grads(blk['mods'][1], r1, dz1, ddc, db1, h, w, up=False)
        dxu1 = _dgrad(gen, dz1, pl1, TAPS_3X3_T, oh=h, ow=w)
        # ---- skip image path: img_out = upsample2d(img_prev) + torgb
        dimg = ops.blur(dimg, 1, 1, stride=2, gain=4.0) if blk['has_img_prev'] else None
        if r0 is None:
            # first backbone block: conv1 reads the learned constant; only its d(styles) is needed
            ops.act_bwd(blk['x_in'], g0=dxu1, ds0=dviews[pl1.index], act=ACT_LINEAR, act_gain=1.0, out='none')
            if wgrads:      # the learned constant input: d const[c][y][x] = sum_n dxu[n][y][x][c] * styles[n][c]
                _acc(blk['const'], (dxu1 * r1['styles'][:, None, None, :]).sum(0).permute(2, 0, 1))
            gx = None
            continue
        # ---- conv0 (x2 up): act/demod backward at 2H, FIR transpose, then the stride-2 data gradient
        pl0 = r0['pl']
        ddc0 = ops.zeros((b, pl0.cout), dev)
        db0 = ops.zeros((pl0.cout,), dev) if wgrads else None
        dz0 = ops.act_bwd(r0['y'], g0=dxu1, s0=r1['styles'], ds0=dviews[pl1.index], dcoef=r0['dcoef'],
                          noise=r0['noise'], noise_gain=pl0.noise_gain, bias=pl0.bias, act=ACT_LRELU, act_gain=SQRT2,
                          clamp=pl0.clamp, out='f32', ddcoef=ddc0, dbias=db0)
        ops.demod_bwd(pl0.bwd()['w2'], r0['styles'], r0['dcoef'], ddc0, dviews[pl0.index])
        use_tc0 = tc and pl0.bwd()['wT_split'] is not None
        dt = ops.blur(dz0, 2, 2, stride=1, gain=4.0, split_out=use_tc0)          # [B,2H+1,2W+1,cout]
        hin, win = h // 2, w // 2
        dxu0 = _dgrad(gen, dt, pl0, TAPS_UP_T, oh=hin, ow=win, in_stride=2)
        if wgrads:
            m0 = blk['mods'][0]
            nb0, r0['noise_buf'] = r0.get('noise_buf'), None           # noise enters after the FIR: use dz0, not dt
            _layer_param_grads(m0, r0, dt, ddc0, db0, h, w, up=True)
            if nb0 is not None:
                _acc(m0.noise_strength, ((dz0 / r0['dcoef'][:, None, None, :]).sum(-1) * nb0[None]).sum())
        gx = (dxu0, r0['styles'], dviews[pl0.index])
    return gx, dimg


class StylesFn(torch.autograd.Function):
    """ws [B,num_ws,w_dim] -> flat styles of all 32 layers (one launch each way)."""

    @staticmethod
    def forward(ctx, ws, gen):
        pk = gen._ensure_packed()
        flat, offs = pk['styles'].run_flat(ws)
        ctx.gen, ctx.offs, ctx.shape, ctx.ws = gen, offs, ws.shape, ws.detach()
        return flat

    @staticmethod
    def backward(ctx, dflat):
        gen = ctx.gen
        pk = gen._ensure_packed()
        b, num_ws, w_dim = ctx.shape
        dflat = dflat.contiguous()
        dws = ops.StyleTableBwd(pk['styles']).run(dflat, ctx.offs, b, num_ws, w_dim)
        if any(p.requires_grad for p in gen.parameters()):
            # affine layers of the unfrozen generator: styles_l = (ws[widx] A^T / sqrt(D) + b) * gain_l
            ws = ctx.ws
            views = pk['styles'].views(dflat, ctx.offs, b)
            for (kind, m, widx), ds in zip(pk['order'], views):
                gain = 1.0 if kind == 'conv' else 1.0 / math.sqrt(m.cin)
                a = m.affine
                dw = torch.zeros_like(a.weight)
                db = torch.zeros_like(a.bias)
                ops.linear_bwd(ds.contiguous(), ws[:, widx].contiguous(), a.weight.detach().contiguous(),
                               gain / math.sqrt(w_dim), gain, need_dx=False, dw=dw, db=db)
                _acc(a.weight, dw)
                _acc(a.bias, db)
        return dws, None


def _dviews(pk, dflat, b):
    table = pk['styles']
    offs, _ = table.offsets(b)
    return table.views(dflat, offs, b)


class BackboneFn(torch.autograd.Function):
    """styles -> tri-planes [B,R,R,96] (channels-last)."""

    @staticmethod
    def forward(ctx, styles_flat, gen, noise_mode, batch, tap):
        pk = gen._ensure_packed()
        cfg = gen.cfg
        offs, _ = pk['styles'].offsets(batch)
        views = pk['styles'].views(styles_flat, offs, batch)
        nb = sum(3 if r > 4 else 2 for r in cfg.block_resolutions)
        styles = iter(views[:nb])
        recs, x, img = [], None, None
        gen._batch = batch
        for r in cfg.block_resolutions:
            rec = {}
            x, img = gen._run_block(getattr(gen.backbone.synthesis, f'b{r}'), x, img, styles, noise_mode, pk, tap,
                                    f'b{r}', rec=rec)
            recs.append(rec)
        ctx.gen, ctx.recs, ctx.batch, ctx.total = gen, recs, batch, styles_flat.numel()
        return img

    @staticmethod
    def backward(ctx, dplanes):
        gen = ctx.gen
        pk = gen._ensure_packed()
        dflat = ops.zeros((ctx.total,), dplanes.device)
        blocks_backward(gen, ctx.recs, dplanes.contiguous(), _dviews(pk, dflat, ctx.batch),
                        wgrads=any(p.requires_grad for p in gen.parameters()))
        ctx.recs = None
        return dflat, None, None, None, None


class SuperresFn(torch.autograd.Function):
    """(feature image [B,r,r,32], styles) -> image [B,4r,4r,3] (channels-last)."""

    @staticmethod
    def forward(ctx, feat, styles_flat, gen, batch, tap):
        pk = gen._ensure_packed()
        offs, _ = pk['styles'].offsets(batch)
        views = pk['styles'].views(styles_flat, offs, batch)
        styles = iter(views[-6:])
        gen._batch = batch
        rgb_lo = feat[..., :3].contiguous()
        recs, x, img = [], feat, rgb_lo
        for i, blk in enumerate((gen.superresolution.block0, gen.superresolution.block1)):
            rec = {}
            x, img = gen._run_block(blk, x, img, styles, 'none', pk, tap, f'sr{i}', rec=rec)
            recs.append(rec)
        ctx.gen, ctx.recs, ctx.batch, ctx.total, ctx.feat = gen, recs, batch, styles_flat.numel(), feat
        return img

    @staticmethod
    def backward(ctx, dimg):
        gen = ctx.gen
        pk = gen._ensure_packed()
        dflat = ops.zeros((ctx.total,), dimg.device)
        gx, drgb_lo = blocks_backward(gen, ctx.recs, dimg.contiguous(), _dviews(pk, dflat, ctx.batch),
                                      wgrads=any(p.requires_grad for p in gen.parameters()))
        # first SR layer reads the feature image itself: dfeat = dxu * styles, d(styles) += sum dxu * feat
        dfeat = ops.act_bwd(ctx.feat, g0=gx[0], s0=gx[1], ds0=gx[2], act=ACT_LINEAR, act_gain=1.0, out='f32')
        dfeat[..., :3] += drgb_lo
        ctx.recs = ctx.feat = None
        return dfeat, dflat, None, None, None


def _decoder_param_grads(gen, f, do, mlp):
    """Weight gradient of the OSG decoder (32 -> 64 softplus -> 1+32) from the per-sample operands the render
    backward kernel wrote: features f [S,32] and d(raw output) do [S,33].  ``hfagp_decoder_wgrad`` recomputes the hidden
    layer and does the four reductions over S = batch * rays * samples; the result is w.r.t. the effective
    (gain-multiplied) weights, so each slice is scaled by its gain on the way into .grad."""
    d0, d2 = gen.decoder.net[0], gen.decoder.net[2]
    g = ops.decoder_wgrad(f, do, mlp)
    n0, n1 = d0.weight.numel(), d2.weight.numel()
    h, o = d0.weight.shape[0], d2.weight.shape[0]
    _acc(d0.weight, g[:n0] * d0.weight_gain)
    _acc(d0.bias, g[n0:n0 + h] * d0.bias_gain)
    _acc(d2.weight, g[n0 + h:n0 + h + n1] * d2.weight_gain)
    _acc(d2.bias, g[n0 + h + n1:n0 + h + n1 + o] * d2.bias_gain)


class RenderFn(torch.autograd.Function):
    """tri-planes [B,R,R,96] -> (feature image [B,r,r,32], depth [B,r*r], weight sum [B,r*r]); gradient to the
    planes only (camera, depths and the frozen decoder carry none)."""

    @staticmethod
    def forward(ctx, planes, gen, c, jitter, u_fine, depth_range, kw, bookkeeping):
        pk = gen._ensure_packed()
        feat, depth, wsum, book = ops.render(planes, c, pk['mlp'], pk['lin'], jitter, u_fine, depth_range,
                                             bookkeeping=bookkeeping, **kw)
        ctx.gen, ctx.kw, ctx.book = gen, kw, book
        ctx.save_for_backward(planes, c, jitter, u_fine if u_fine is not None else torch.empty(0, device=planes.device))
        ctx.mark_non_differentiable(depth, wsum)
        return feat, depth, wsum

    @staticmethod
    def backward(ctx, dfeat, _ddepth, _dwsum):
        planes, c, jitter, u_fine = ctx.saved_tensors
        gen = ctx.gen
        pk = gen._ensure_packed()
        dec = any(p.requires_grad for p in gen.decoder.parameters())
        out = ops.render_bwd(planes, c, pk['mlp'], pk['lin'], jitter, u_fine if u_fine.numel() else None,
                             dfeat.contiguous(), decoder=dec, **ctx.kw)
        if not dec:
            return out, None, None, None, None, None, None, None
        dplanes, f, do = out
        _decoder_param_grads(gen, f, do, pk['mlp'])
        return dplanes, None, None, None, None, None, None, None


# ------------------------------------------------------------------ encoder / latent / loss (trainer_rgb.py:79-91)

def _grad_slot(p):
    """True when leaf ``p`` has a dense fp32 ``.grad`` a kernel can accumulate into in place."""
    g = getattr(p, 'grad', None)
    return (g is not None and p.is_leaf and g.dtype == torch.float32 and g.is_contiguous() and g.shape == p.shape
            and g.device == p.device)


class LinearFn(torch.autograd.Function):
    """EqualLinear with activation=None (encoder3d.py:128-136): y = x W^T * scale + b * lr_mul."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale, lr_mul):
        w = weight.detach().contiguous()
        ctx.save_for_backward(x, w)
        ctx.scale, ctx.lr_mul, ctx.has_bias = scale, lr_mul, bias is not None
        ctx.leaves = (weight, bias)
        return ops.linear(x, w, None if bias is None else bias.detach(), scale, lr_mul)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_db = need_dw and ctx.has_bias
        weight, bias = ctx.leaves
        ctx.leaves = None
        # the kernel ACCUMULATES: when the leaves already own gradient storage (FlatAdam binds every .grad to its flat
        # buffer, zeroed once per step) it adds straight into it — no zero-fill, no AccumulateGrad add per tensor
        if (need_dw and _grad_slot(weight) and (not need_db or _grad_slot(bias))):
            dx = ops.linear_bwd(dy, x, w, ctx.scale, ctx.lr_mul, need_dx=need_dx, dw=weight.grad, db=bias.grad if need_db else None)
            return dx, None, None, None, None
        dw = ops.zeros(w.shape, w.device) if need_dw else None
        db = ops.zeros((w.shape[0],), w.device) if need_db else None
        dx = ops.linear_bwd(dy, x, w, ctx.scale, ctx.lr_mul, need_dx=need_dx, dw=dw, db=db)
        return dx, dw, db, None, None


class EncoderAppFn(torch.autograd.Function):
    """frame [B,3,S,S] -> [B, w_dim] through EncoderApp (encoder3d.py:231-239); gradients reach the convolution
    weights and activation biases (``*params`` = ``net.parameters()``, in that order), not the frame."""

    @staticmethod
    def forward(ctx, x, net, *params):
        tape = {}
        out = net._forward_impl(x, tape)
        ctx.net, ctx.tape, ctx.params = net, tape, params
        return out

    @staticmethod
    def backward(ctx, dout):
        grads = ctx.net._backward_impl(ctx.tape, dout.contiguous())
        out = tuple(grads.get(p) if need else None for p, need in zip(ctx.params, ctx.needs_input_grad[2:]))
        ctx.tape = None
        return (None, None) + out


class BasisQRFn(torch.autograd.Function):
    """Q of ``torch.qr(bases.T + 1e-8)`` (headnerf.py:92) and its backward on the C ABI (csrc/qr.cu): no LAPACK call,
    ~0.1 ms instead of cuSOLVER's 0.8 ms latency chain in front of every step's generator forward."""

    @staticmethod
    def forward(ctx, bases):
        q, rinv = ops.basis_qr(bases)
        ctx.save_for_backward(q, rinv)
        return q

    @staticmethod
    def backward(ctx, gq):
        q, rinv = ctx.saved_tensors
        return ops.basis_qr_bwd(gq, q, rinv)


class LatentFn(torch.autograd.Function):
    """ws = weights . Q^T + delta (headnerf.py:96-100); Q comes from BasisQRFn, which carries d(Q) on to ``bases``."""

    @staticmethod
    def forward(ctx, weights, q, delta):
        w, qc = weights.detach().float().contiguous(), q.detach().contiguous()
        ctx.save_for_backward(w, qc)
        return ops.latent(w, qc, delta.detach().contiguous(), qc.shape[0])

    @staticmethod
    def backward(ctx, dws):
        w, q = ctx.saved_tensors
        nw, nq, nd = ctx.needs_input_grad
        return ops.latent_bwd(dws.contiguous(), w, q, nw, nq, nd)


class FacePoolFn(torch.autograd.Function):
    """face_pool (trainer_rgb.py:63,84): channels-last image -> AdaptiveAvgPool2d(size) in NCHW."""

    @staticmethod
    def forward(ctx, img_nhwc, size):
        ctx.hw = img_nhwc.shape[1:3]
        return ops.facepool(img_nhwc.contiguous(), size)

    @staticmethod
    def backward(ctx, dy):
        return ops.facepool_bwd(dy.contiguous(), *ctx.hw), None


class MseFn(torch.autograd.Function):
    """MSELoss(reduction='mean')(real, generated) (trainer_rgb.py:15,85); gradient to ``generated`` only."""

    @staticmethod
    def forward(ctx, real, generated):
        real, generated = real.detach().float().contiguous(), generated.contiguous()
        ctx.save_for_backward(real, generated)
        return ops.mse(generated, real)

    @staticmethod
    def backward(ctx, gout):
        real, generated = ctx.saved_tensors
        return None, ops.mse_bwd(generated, real, gout.contiguous())
