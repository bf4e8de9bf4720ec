"""LPIPS(net='alex') as the trainers use it (``/root/reference/code/trainer_rgb.py:10,62,86-87``):
``loss = LPIPS(net='alex').to(device).eval()(real, generated)`` -> ``[B,1,1,1]``, on the sm_100a library.

The reference imports the un-vendored pip package ``lpips``; this module keeps its ``state_dict`` names
(``net.sliceK.N.weight``, ``linK.model.1.weight``, ``scaling_layer.shift/scale``) so real weights load when available,
and initialises from a seed for synthetic runs (no weights exist offline).  The torch modules below only HOLD
parameters; the arithmetic is:

  hfagp_lpips_stem_fwd    ScalingLayer + pad + space-to-depth(4) + split-bf16: the 11x11/stride-4 conv becomes a 3x3 conv
                          over 48 channels
  hfagp_conv2d_tc_fwd     the five AlexNet convolutions on tcgen05 (bias + ReLU epilogue, split-bf16 features; the small
                          15^2 / 31^2 layers take the split-K path automatically)
  hfagp_maxpool3s2_fwd    the two MaxPool2d(3, 2)
  hfagp_lpips_head_fwd    unit-normalise over channels, squared difference, ``lin`` weights, spatial mean, summed over layers

and the mirror image backwards (``LpipsFn.backward``): head_bwd -> act_bwd (ReLU) -> transposed-weight convs ->
maxpool_bwd -> stem_bwd, giving d(loss)/d(generated image).  Both images go through the trunk as one batch of 2B.
The LPIPS network itself is frozen (``.eval()``, ``requires_grad_(False)``), as in the reference.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import _cabi, ops
from ._cabi import ACT_RELU, HfagpError, ptr, stream


class _ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer('shift', torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('scale', torch.tensor([.458, .448, .450])[None, :, None, None])


class _Alex(nn.Module):
    """Parameter holders with torchvision alexnet().features module indices, cut into the five LPIPS slices."""

    def __init__(self):
        super().__init__()
        self.slice1, self.slice2, self.slice3 = nn.Sequential(), nn.Sequential(), nn.Sequential()
        self.slice4, self.slice5 = nn.Sequential(), nn.Sequential()
        self.slice1.add_module('0', nn.Conv2d(3, 64, 11, 4, 2))
        self.slice2.add_module('3', nn.Conv2d(64, 192, 5, 1, 2))
        self.slice3.add_module('6', nn.Conv2d(192, 384, 3, 1, 1))
        self.slice4.add_module('8', nn.Conv2d(384, 256, 3, 1, 1))
        self.slice5.add_module('10', nn.Conv2d(256, 256, 3, 1, 1))

    def convs(self):
        return [self.slice1[0], self.slice2[0], self.slice3[0], self.slice4[0], self.slice5[0]]


class _NetLinLayer(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.model = nn.Sequential(nn.Dropout(), nn.Conv2d(cin, 1, 1, 1, 0, bias=False))


def _taps(k, pad):
    return tuple((ky - pad, kx - pad, ky * k + kx) for ky in range(k) for kx in range(k))


def _mirror(taps):
    return tuple((-dy, -dx, t) for dy, dx, t in taps)


def pack_stem_weight(w1: torch.Tensor) -> torch.Tensor:
    """AlexNet conv1 weight [64,3,11,11] (stride 4, pad 2) -> [9 taps][64][48]: the same convolution as a 3x3 stride-1
    VALID convolution over the space-to-depth(4) image of hfagp_lpips_stem_fwd, whose channel (py*4+px)*3+c holds
    pixel (4Y+py-2, 4X+px-2) of input channel c; kernel row/column 11 is zero padding."""
    o = w1.shape[0]
    w = torch.nn.functional.pad(w1.detach().float(), (0, 1, 0, 1)).reshape(o, 3, 3, 4, 3, 4)       # o, c, ty, py, tx, px
    return w.permute(2, 4, 0, 3, 5, 1).reshape(9, o, 48).contiguous()


TAPS_STEM = _taps(3, 0)          # valid 3x3 over the space-to-depth grid
TAPS_5X5 = _taps(5, 2)


class LPIPS(nn.Module):
    CHNS = (64, 192, 384, 256, 256)

    def __init__(self, net='alex', seed=0, verbose=False, weights=None, synthetic=None):
        """``weights``: path of a ``state_dict`` of the pip ``lpips`` package's ``LPIPS(net='alex')`` (its key names are kept
        here: ``net.slice*.N.{weight,bias}``, ``lin*.model.1.weight``); default: env ``HFAGP_LPIPS_WEIGHTS``.
        Without weights the constructor RAISES — a randomly initialised AlexNet is a different loss — unless
        ``synthetic=True`` / env ``HFAGP_SYNTHETIC_LPIPS=1`` asks for the seeded random init of the synthetic benchmarks
        (same policy as ``load_G_official`` for the generator pickle)."""
        super().__init__()
        if net != 'alex':
            raise ValueError("only LPIPS(net='alex') is on the HFA-GP path (trainer_rgb.py:62)")
        weights = weights or os.environ.get('HFAGP_LPIPS_WEIGHTS')
        if synthetic is None:
            synthetic = os.environ.get('HFAGP_SYNTHETIC_LPIPS') == '1'
        if not weights and not synthetic:
            raise FileNotFoundError(
                "LPIPS(net='alex') needs the pretrained weights of the pip `lpips` package: pass weights=<state_dict file> "
                'or set HFAGP_LPIPS_WEIGHTS; for synthetic runs pass synthetic=True or set HFAGP_SYNTHETIC_LPIPS=1 '
                '(seeded random init, NOT the perceptual loss the reference trains with)')
        with torch.random.fork_rng():
            torch.manual_seed(seed)
            self.scaling_layer = _ScalingLayer()
            self.net = _Alex()
            self.lins = nn.ModuleList(_NetLinLayer(c) for c in self.CHNS)
            for lin in self.lins:                      # the trained heads are non-negative
                lin.model[1].weight.data.abs_()
        for k, lin in enumerate(self.lins):
            setattr(self, f'lin{k}', lin)
        if weights:
            sd = torch.load(weights, map_location='cpu', weights_only=True)
            self.load_state_dict(sd.get('state_dict', sd) if isinstance(sd, dict) else sd)
        self.synthetic = not weights
        self.requires_grad_(False)
        self._pk, self._pk_key = None, None

    # ------------------------------------------------------------------ kernel-layout weights (cached)
    def _packed(self):
        # (LPIPS weights are never trained: optimiser steps — ops.param_epoch — do not invalidate them)
        key = (tuple(p._version for p in self.parameters()), str(next(self.parameters()).device))
        if self._pk is not None and self._pk_key == key:
            return self._pk
        convs = self.net.convs()
        pk = {'w': [], 'wT': [], 'b': [], 'lin': [], 'taps': [TAPS_STEM, TAPS_5X5, ops.TAPS_3X3, ops.TAPS_3X3, ops.TAPS_3X3]}
        w1 = pack_stem_weight(convs[0].weight)                                 # [64,3,11,11] -> [9][64][48]
        packed = [w1] + [c.weight.detach().float().permute(2, 3, 0, 1).reshape(-1, c.weight.shape[0], c.weight.shape[1]).contiguous()
                         for c in convs[1:]]
        for w, c in zip(packed, convs):
            pk['w'].append(ops.split(w[None].contiguous()))
            pk['wT'].append(ops.split(w.transpose(1, 2)[None].contiguous()))
            pk['b'].append(c.bias.detach().float().contiguous())
        pk['lin'] = [l.model[1].weight.detach().float().reshape(-1).contiguous() for l in self.lins]
        pk['shift'] = (_cabi.C.c_float * 3)(*self.scaling_layer.shift.reshape(-1).tolist())
        pk['scale'] = (_cabi.C.c_float * 3)(*self.scaling_layer.scale.reshape(-1).tolist())
        self._pk, self._pk_key = pk, key
        return pk

    def forward(self, in0, in1, retPerLayer=False, normalize=False):
        if retPerLayer:
            raise HfagpError('LPIPS(retPerLayer=True) is not used on the HFA-GP path')
        if not (in0.is_cuda and in1.is_cuda):
            raise HfagpError('LPIPS needs CUDA tensors (there is no CPU fallback)')
        if self.training:
            raise HfagpError('LPIPS runs in eval mode on the HFA-GP path (trainer_rgb.py:62 calls .eval()); '
                             'train-mode Dropout is not implemented')
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        return LpipsFn.apply(in0, in1, self)


def _stem(x, pk):
    n, _, h, w = x.shape
    hi = torch.empty((n, (h + 4) // 4, (w + 4) // 4, 48), device=x.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    ops._ok(_cabi.lib().hfagp_lpips_stem_fwd(n, h, w, ptr(x), pk['shift'], pk['scale'], ptr(hi), ptr(lo), stream()),
            'hfagp_lpips_stem_fwd')
    return ops.Split(hi, lo)


def _maxpool(x: ops.Split):
    n, h, w, c = x.shape
    oh, ow = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    hi = torch.empty((n, oh, ow, c), device=x.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    ops._ok(_cabi.lib().hfagp_maxpool3s2_fwd(n, h, w, c, None, ptr(x.hi), ptr(x.lo), None, ptr(hi), ptr(lo), stream()),
            'hfagp_maxpool3s2_fwd')
    return ops.Split(hi, lo)


def _maxpool_bwd(x: ops.Split, dy):
    n, h, w, c = x.shape
    dx = torch.empty((n, h, w, c), device=dy.device, dtype=torch.float32)
    ops._ok(_cabi.lib().hfagp_maxpool3s2_bwd(n, h, w, c, None, ptr(x.hi), ptr(x.lo), ptr(dy.contiguous()), ptr(dx), stream()),
            'hfagp_maxpool3s2_bwd')
    return dx


def _half(f: ops.Split, b):
    return ops.Split(f.hi[b:], f.lo[b:])


class LpipsFn(torch.autograd.Function):
    """(real [B,3,H,W], generated [B,3,H,W]) -> [B,1,1,1]; gradient to ``generated`` only."""

    @staticmethod
    def forward(ctx, real, gen, mod):
        pk = mod._packed()
        b = real.shape[0]
        if real.shape != gen.shape or real.shape[1] != 3 or real.shape[2] % 4 or real.shape[3] % 4:
            raise HfagpError(f'LPIPS expects two [B,3,H,W] images with sides divisible by 4, got {tuple(real.shape)} / {tuple(gen.shape)}')
        x = torch.cat([real.detach().float(), gen.detach().float()]).contiguous()
        s = _stem(x, pk)
        feats = []
        cur = s
        for k in range(5):
            if k in (1, 2):
                cur = _maxpool(cur)
            n, h, w, _ = cur.shape
            oh, ow = (h - 2, w - 2) if k == 0 else (h, w)
            cur = ops.conv2d_tc(cur, pk['w'][k], pk['taps'][k], LPIPS.CHNS[k], oh=oh, ow=ow, bias=pk['b'][k],
                                act=ACT_RELU, split_out=True)
            feats.append(cur)
        out = ops.zeros((b,), x.device)
        for k, f in enumerate(feats):
            _, h, w, c = f.shape
            ops._ok(_cabi.lib().hfagp_lpips_head_fwd(b, h * w, c, None, ptr(f.hi), ptr(f.lo), ptr(pk['lin'][k]), ptr(out),
                                                     stream()), 'hfagp_lpips_head_fwd')
        ctx.mod, ctx.feats, ctx.stem_shape, ctx.img_shape = mod, feats, s.shape, gen.shape
        return out.view(b, 1, 1, 1)

    @staticmethod
    def backward(ctx, gout):
        pk = ctx.mod._packed()
        feats = ctx.feats
        b, _, ih, iw = ctx.img_shape
        go = gout.reshape(b).float().contiguous()
        dev = go.device

        def head_bwd(k):
            f = feats[k]
            _, h, w, c = f.shape
            d = torch.empty((b, h, w, c), device=dev, dtype=torch.float32)
            ops._ok(_cabi.lib().hfagp_lpips_head_bwd(b, h * w, c, None, ptr(f.hi), ptr(f.lo), ptr(pk['lin'][k]), ptr(go),
                                                     ptr(d), stream()), 'hfagp_lpips_head_bwd')
            return d

        g = None                                   # gradient flowing down the trunk into feats[k]'s ReLU output
        for k in (4, 3, 2, 1, 0):
            fk = _half(feats[k], b)
            dz = ops.act_bwd(fk, g0=head_bwd(k), g1=g, act=ACT_RELU, act_gain=1.0, out='split')
            _, h, w, _ = fk.shape
            cin = pk['w'][k].shape[-1]
            if k == 0:
                _, sh, sw, _ = ctx.stem_shape
                dx48 = ops.conv2d_tc(dz, pk['wT'][0], _mirror(TAPS_STEM), 48, oh=sh, ow=sw)
                dimg = torch.empty((b, 3, ih, iw), device=dev, dtype=torch.float32)
                ops._ok(_cabi.lib().hfagp_lpips_stem_bwd(b, ih, iw, ptr(dx48), pk['scale'], ptr(dimg), stream()),
                        'hfagp_lpips_stem_bwd')
                ctx.feats = None
                return None, dimg, None
            g = ops.conv2d_tc(dz, pk['wT'][k], _mirror(pk['taps'][k]), cin, oh=h, ow=w)
            if k in (1, 2):                        # the convolution's input was a max-pool of feats[k-1]
                g = _maxpool_bwd(_half(feats[k - 1], b), g)
