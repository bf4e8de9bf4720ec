"""B200-native EG3D tri-plane generator behind the call surface HFA-GP uses.

HFA-GP never looks inside the generator: it un-pickles NVlabs/eg3d's ``TriPlaneGenerator``
(``/root/reference/code/networks/headnerf.py:31-38``) and calls
``generator.synthesis(latent, c=label, noise_mode='const')['image']`` (``headnerf.py:112``),
iterates ``generator.parameters()`` to freeze/unfreeze it (``trainer_rgb.py:59-60,70-71``) and
saves / loads its tensors under ``generator.*`` (``trainer_rgb.py:146``, ``run_recon_video_rgb.py:211``).
This module provides an ``nn.Module`` with exactly that protocol and the EG3D ``state_dict`` names
(SURVEY.md App. A.9); the arithmetic is the sm_100a library behind ``include/hfagp.h``.

Activations are fp32 channels-last on the device; parameters are kept in their canonical
PyTorch shapes and re-packed into kernel layouts lazily whenever they change.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
from torch import nn

from . import ops
from ._cabi import ACT_LINEAR, ACT_LRELU, HfagpError

SQRT2 = math.sqrt(2.0)


@dataclass
class GeneratorConfig:
    """What the EG3D pickle carries as constructor kwargs / rendering_kwargs (ffhqrebalanced512-128)."""
    w_dim: int = 512
    c_dim: int = 25
    plane_res: int = 256
    plane_channels: int = 32
    channel_base: int = 32768
    channel_max: int = 512
    nrr: int = 128                    # neural_rendering_resolution
    img_resolution: int = 512
    sr_channels: tuple = (256, 128)
    sr_clamp: float = 256.0
    decoder_hidden: int = 64
    depth_res: int = 48
    depth_res_importance: int = 48
    ray_start: float = 2.25
    ray_end: float = 3.3
    box_warp: float = 1.0

    @property
    def block_resolutions(self) -> List[int]:
        return [2 ** i for i in range(2, int(math.log2(self.plane_res)) + 1)]

    def channels(self, res: int) -> int:
        return min(self.channel_base // res, self.channel_max)

    @property
    def num_ws(self) -> int:
        return 2 * len(self.block_resolutions)


def _resample_filter():
    f = torch.tensor([1.0, 3.0, 3.0, 1.0])
    f = torch.outer(f, f)
    return f / f.sum()


# ------------------------------------------------------------------ parameter holders (state_dict layout)

class _FC(nn.Module):
    def __init__(self, cin, cout, bias_init=0.0, lr_mul=1.0):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(cout, cin) / lr_mul)
        self.bias = nn.Parameter(torch.full([cout], float(bias_init)))
        self.weight_gain = lr_mul / math.sqrt(cin)
        self.bias_gain = lr_mul


class _SynthesisLayer(nn.Module):
    def __init__(self, cin, cout, w_dim, res, up, use_noise, clamp):
        super().__init__()
        self.cin, self.cout, self.res, self.up, self.use_noise, self.clamp = cin, cout, res, up, use_noise, clamp
        self.affine = _FC(w_dim, cin, bias_init=1.0)
        self.weight = nn.Parameter(torch.randn(cout, cin, 3, 3))
        self.register_buffer('resample_filter', _resample_filter())
        self.register_buffer('noise_const', torch.randn(res, res))
        self.noise_strength = nn.Parameter(torch.zeros([]))
        self.bias = nn.Parameter(torch.zeros(cout))


class _ToRGB(nn.Module):
    def __init__(self, cin, cout, w_dim, clamp):
        super().__init__()
        self.cin, self.cout, self.clamp = cin, cout, clamp
        self.affine = _FC(w_dim, cin, bias_init=1.0)
        self.weight = nn.Parameter(torch.randn(cout, cin, 1, 1))
        self.bias = nn.Parameter(torch.zeros(cout))


class _Block(nn.Module):
    def __init__(self, cin, cout, w_dim, res, img_channels, clamp=None, use_noise=True):
        super().__init__()
        self.cin, self.cout, self.res = cin, cout, res
        self.register_buffer('resample_filter', _resample_filter())
        if cin == 0:
            self.const = nn.Parameter(torch.randn(cout, res, res))
        else:
            self.conv0 = _SynthesisLayer(cin, cout, w_dim, res, 2, use_noise, clamp)
        self.conv1 = _SynthesisLayer(cout, cout, w_dim, res, 1, use_noise, clamp)
        self.torgb = _ToRGB(cout, img_channels, w_dim, clamp)
        self.num_conv = 1 if cin == 0 else 2


class _Synthesis(nn.Module):
    def __init__(self, cfg: GeneratorConfig):
        super().__init__()
        for res in cfg.block_resolutions:
            cin = cfg.channels(res // 2) if res > 4 else 0
            setattr(self, f'b{res}', _Block(cin, cfg.channels(res), cfg.w_dim, res, 3 * cfg.plane_channels))


class _Mapping(nn.Module):
    """Present only so checkpoints round-trip; HFA-GP feeds ws from get_latent and never runs it."""

    def __init__(self, cfg):
        super().__init__()
        self.embed = _FC(cfg.c_dim, cfg.w_dim)
        self.fc0 = _FC(2 * cfg.w_dim, cfg.w_dim, lr_mul=0.01)
        self.fc1 = _FC(cfg.w_dim, cfg.w_dim, lr_mul=0.01)
        self.register_buffer('w_avg', torch.zeros(cfg.w_dim))


class _Backbone(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.synthesis = _Synthesis(cfg)
        self.mapping = _Mapping(cfg)


class _Superresolution(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        c0, c1 = cfg.sr_channels
        self.block0 = _Block(cfg.plane_channels, c0, cfg.w_dim, 2 * cfg.nrr, 3, clamp=cfg.sr_clamp)
        self.block1 = _Block(c0, c1, cfg.w_dim, 4 * cfg.nrr, 3, clamp=cfg.sr_clamp)


class _Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.net = nn.Sequential(_FC(cfg.plane_channels, cfg.decoder_hidden), nn.Softplus(),
                                 _FC(cfg.decoder_hidden, 1 + cfg.plane_channels))


# ------------------------------------------------------------------ packed (kernel-layout) view

class _PackedLayer:
    __slots__ = ('w', 'bias', 'noise', 'noise_gain', 'cin', 'cout', 'up', 'clamp', 'use_noise', 'index', '_bwd')

    def bwd(self):
        """Operands of the backward pass, built on first use: transposed weights wT[tap][cin][cout] (fp32 and
        split bf16) for the data-gradient convolution and w2[cout][cin] = sum_taps w^2 for the demod backward."""
        if self._bwd is None:
            wt = self.w.transpose(1, 2).contiguous()
            self._bwd = dict(wT=wt, wT_split=ops.split(wt) if self.cout % 8 == 0 else None,
                             w2=self.w.square().sum(0).contiguous())
        return self._bwd


def _pack_conv(weight: torch.Tensor) -> torch.Tensor:
    """[O,I,kh,kw] -> [kh*kw][O][I] (cin contiguous: the K-major GEMM operand)."""
    o, i, kh, kw = weight.shape
    return weight.detach().permute(2, 3, 0, 1).reshape(kh * kw, o, i).contiguous().float()


class TriPlaneGenerator(nn.Module):
    """Drop-in for the object ``legacy.load_network_pkl(f)['G_ema']`` returns (headnerf.py:34)."""

    def __init__(self, cfg: Optional[GeneratorConfig] = None):
        super().__init__()
        self.cfg = cfg or GeneratorConfig()
        if self.cfg.plane_channels != 32 or self.cfg.decoder_hidden != 64:
            raise HfagpError('the render kernel is specialised for 32 plane channels and a 64-wide decoder')
        self.backbone = _Backbone(self.cfg)
        self.superresolution = _Superresolution(self.cfg)
        self.decoder = _Decoder(self.cfg)
        self.neural_rendering_resolution = self.cfg.nrr
        self.z_dim = self.w_dim = self.cfg.w_dim
        self.c_dim = self.cfg.c_dim
        self.img_resolution = self.cfg.img_resolution
        self.img_channels = 3
        self.rendering_kwargs = dict(depth_resolution=self.cfg.depth_res,
                                     depth_resolution_importance=self.cfg.depth_res_importance,
                                     ray_start=self.cfg.ray_start, ray_end=self.cfg.ray_end,
                                     box_warp=self.cfg.box_warp, clamp_mode='softplus', white_back=False,
                                     superresolution_noise_mode='none', avg_camera_radius=2.7,
                                     avg_camera_pivot=[0, 0, 0.2])
        self._packed = None
        self._packed_key = None
        self.fixed_draws = None    # (jitter_coarse, u_fine) used instead of torch.rand when synthesis() gets none
        self.precision = 'tc'      # 'tc': tcgen05 split-bf16 convolutions (fp32-class); 'fp32': SIMT kernels only
        self._premod = None        # inference: per-layer (modulated weights, dcoef) of the current synthesis() call
        self.overlap_streams = True   # inference: the ToRGB / skip-image chain of the backbone on a second CUDA stream
        self._side = None
        self._keep = None

    @property
    def num_ws(self) -> int:
        return self.cfg.num_ws

    # ---------------------------------------------------------------- packing
    def _fingerprint(self):
        dev = None
        ver = 0
        live = False
        for p in self.parameters():
            ver += p._version
            dev = p.device
            live = live or p.requires_grad
        for b in self.buffers():
            ver += b._version
        # optimiser steps write parameters through the C ABI without touching torch's version counters (param_epoch); a
        # FROZEN generator is not among the parameters they update, so its packed weights survive the step (re-packing
        # 26 layers — and reading ~20 noise strengths back to the host — every step is what made training host-bound)
        return (str(dev), ver, ops.param_epoch[0] if live else -1)

    def _layers_in_order(self):
        cfg = self.cfg
        out = []       # (kind, module, ws index)
        w_idx = 0
        for res in cfg.block_resolutions:
            blk = getattr(self.backbone.synthesis, f'b{res}')
            i = w_idx
            if blk.cin != 0:
                out.append(('conv', blk.conv0, i)); i += 1
            out.append(('conv', blk.conv1, i)); i += 1
            out.append(('torgb', blk.torgb, i))
            w_idx += blk.num_conv
        last = cfg.num_ws - 1
        for blk in (self.superresolution.block0, self.superresolution.block1):
            out.append(('conv', blk.conv0, last))
            out.append(('conv', blk.conv1, last))
            out.append(('torgb', blk.torgb, last))
        return out

    def _ensure_packed(self):
        key = self._fingerprint()
        if self._packed is not None and self._packed_key == key:
            return self._packed
        layers = self._layers_in_order()
        table, packed = [], {}
        for kind, m, widx in layers:
            gain = 1.0 if kind == 'conv' else 1.0 / math.sqrt(m.cin)
            table.append((m.affine.weight.detach().contiguous(), m.affine.bias.detach().contiguous(), m.cin, widx, gain))
            pl = _PackedLayer()
            pl._bwd = None
            pl.index = len(table) - 1
            pl.w = _pack_conv(m.weight)
            pl.bias = m.bias.detach().contiguous().float()
            pl.cin, pl.cout = m.cin, m.cout
            pl.clamp = float(m.clamp) if m.clamp is not None else 0.0
            if kind == 'conv':
                pl.up = m.up
                pl.use_noise = m.use_noise
                pl.noise = m.noise_const.detach().contiguous().float()
                pl.noise_gain = float(m.noise_strength.detach())
            packed[id(m)] = pl
        d0, d2 = self.decoder.net[0], self.decoder.net[2]
        mlp = torch.cat([(d0.weight.detach() * d0.weight_gain).reshape(-1), d0.bias.detach() * d0.bias_gain,
                         (d2.weight.detach() * d2.weight_gain).reshape(-1), d2.bias.detach() * d2.bias_gain]).float().contiguous()
        dev = mlp.device
        lin = torch.linspace(self.cfg.ray_start, self.cfg.ray_end, self.cfg.depth_res).to(dev)
        const = self.backbone.synthesis.b4.const.detach().permute(1, 2, 0).contiguous().float()
        self._packed = dict(order=layers, styles=ops.StyleTable(table), layers=packed, mlp=mlp, lin=lin, const=const,
                            const_split=ops.split(const) if const.is_cuda else None)
        self._packed_key = key
        return self._packed

    # ---------------------------------------------------------------- layer drivers (channels-last)
    # Kernel selection per layer: the tcgen05 path (split-bf16 operands, fp32 accumulate in TMEM) whenever the
    # layer's cin is a multiple of 8 (TMA's 16-byte stride rule; K is walked in 64-channel chunks and a partial
    # chunk is zero-filled), else the exact-fp32 SIMT kernel.  ``self.precision = 'fp32'`` forces SIMT everywhere.
    def _use_tc(self, cin):
        return self.precision == 'tc' and cin % 8 == 0

    @staticmethod
    def _as_f32(x):
        return x.float() if isinstance(x, ops.Split) else x

    def _conv_layer(self, x, m, styles, noise_mode, pk, split_out, rec=None, rgb=None):
        pl: _PackedLayer = pk['layers'][id(m)]
        noise = None
        if pl.use_noise and noise_mode == 'const' and pl.noise_gain != 0.0:
            noise = pl.noise
        elif pl.use_noise and noise_mode == 'random':
            raise HfagpError("noise_mode='random' is not part of the HFA-GP path (headnerf.py:112 passes 'const')")
        epi = dict(dcoef=None, noise=noise, noise_gain=pl.noise_gain, bias=pl.bias, act=ACT_LRELU, act_gain=SQRT2,
                   clamp=pl.clamp)
        h, w = x.shape[1], x.shape[2]
        if self._use_tc(pl.cin):
            xs = x if isinstance(x, ops.Split) else ops.split(x)
            pre = self._premod.get(id(m)) if self._premod else None
            wmod, epi['dcoef'] = pre if pre is not None else ops.modulate_split(pl.w, styles, True)
            if pl.up == 1:
                extra = dict(rgb_w=rgb[0], rgb_acc=rgb[1]) if rgb is not None else {}
                y = ops.conv2d_tc(xs, wmod, ops.TAPS_3X3, pl.cout, oh=h, ow=w, w_batched=True,
                                  split_out=split_out, **epi, **extra)
            else:
                t = ops.conv_transpose_s2_tc(xs, wmod, pl.cout, w_batched=True)
                y = ops.upfir_act(t, split_out=split_out, **epi)
        else:
            xs = self._as_f32(x)
            wmod, epi['dcoef'] = ops.modulate(pl.w, styles, True)
            wbs = wmod.stride(0)
            if pl.up == 1:
                y = ops.conv2d(xs, wmod, ops.TAPS_3X3, pl.cout, oh=h, ow=w, w_batch_stride=wbs, **epi)
                y = ops.split(y) if split_out else y
            else:
                t = ops.conv_transpose_s2(xs, wmod, pl.cout, wbs)
                y = ops.upfir_act(t, split_out=split_out, **epi)
        if rec is not None:
            rec.update(pl=pl, x=xs, y=y, styles=styles, dcoef=epi['dcoef'], noise=noise,
                       noise_buf=pl.noise if (pl.use_noise and noise_mode == 'const') else None)
        return y

    def _torgb_layer(self, x, m, styles, img, pk, rec=None):
        pl: _PackedLayer = pk['layers'][id(m)]
        h, w = x.shape[1], x.shape[2]
        if rec is not None:
            rec.update(pl=pl, x=x, styles=styles, small=pl.cout <= 4, mask=None)
        if pl.cout <= 4:
            wmod, _ = ops.modulate(pl.w, styles, False)
            if rec is not None and pl.clamp > 0:
                # training: the clamp's derivative mask (taken before the skip-image add) comes out of the same pass
                out, rec['mask'] = ops.torgb_small(x, wmod, pl.bias, pl.clamp, img, pl.cout, want_mask=True)
                return out
            return ops.torgb_small(x, wmod, pl.bias, pl.clamp, img, pl.cout)
        if rec is not None and pl.clamp > 0:
            raise HfagpError('backward of a clamped wide ToRGB is not implemented (EG3D clamps only the 3-channel SR ToRGB)')
        if self._use_tc(pl.cin):
            xs = x if isinstance(x, ops.Split) else ops.split(x)
            pre = self._premod.get(id(m)) if self._premod else None
            wmod, _ = pre if pre is not None else ops.modulate_split(pl.w, styles, False)
            return ops.conv2d_tc(xs, wmod, ops.TAPS_1X1, pl.cout, oh=h, ow=w, w_batched=True, bias=pl.bias,
                                 clamp=pl.clamp, up_img=img)
        wmod, _ = ops.modulate(pl.w, styles, False)
        return ops.conv2d(self._as_f32(x), wmod, ops.TAPS_1X1, pl.cout, oh=h, ow=w, w_batch_stride=wmod.stride(0),
                          bias=pl.bias, clamp=pl.clamp, up_img=img)

    def _run_block(self, blk, x, img, styles_iter, noise_mode, pk, tap, name, rec=None):
        # activations between layers of a block stay in the operand format of their consumer
        tc_next = self._use_tc(blk.cout)
        r0 = r1 = rt = None
        if rec is not None:
            r0, r1, rt = ({} if blk.cin != 0 else None), {}, {}
            rec.update(conv0=r0, conv1=r1, torgb=rt, has_img_prev=img is not None, x_in=x,
                       mods=(getattr(blk, 'conv0', None), blk.conv1, blk.torgb), const=getattr(blk, 'const', None))
        if blk.cin == 0:
            c = pk['const_split'] if tc_next else pk['const']
            if tc_next:
                x = ops.Split(c.hi[None].expand(self._batch, -1, -1, -1).contiguous(),
                              c.lo[None].expand(self._batch, -1, -1, -1).contiguous())
            else:
                x = c[None].expand(self._batch, -1, -1, -1).contiguous()
            if rec is not None:
                rec['x_in'] = x
        else:
            x = self._conv_layer(x, blk.conv0, next(styles_iter), noise_mode, pk, split_out=tc_next, rec=r0)
            if tap is not None:
                tap[name + '.conv0'] = self._as_f32(x)
        s1, st = next(styles_iter), next(styles_iter)
        plt = pk['layers'][id(blk.torgb)]
        if plt.cout <= 4 and self._use_tc(blk.cout) and blk.cout % 32 == 0 and not ops.deterministic():
            # super-resolution blocks: the 3-channel ToRGB rides on conv1's epilogue (its activations are still in
            # registers there) instead of re-reading the whole layer output — at inference and in the training forward
            wrgb, _ = ops.modulate(plt.w, st, False)                       # [n][1][k][cout]
            acc = ops.zeros((x.shape[0], blk.res, blk.res, plt.cout), wrgb.device)
            x = self._conv_layer(x, blk.conv1, s1, noise_mode, pk, split_out=tc_next, rgb=(wrgb[:, 0], acc), rec=r1)
            if tap is not None:
                tap[name + '.conv1'] = self._as_f32(x)
            if rt is not None:
                # what the backward of the ToRGB branch reads: conv1's output, the styles, the clamp's derivative mask
                rt.update(pl=plt, x=x, styles=st, small=True,
                          mask=((acc + plt.bias).abs() < plt.clamp) if plt.clamp > 0 else None)
            img = ops.torgb_finalize(acc, plt.bias, plt.clamp, img)
            if tap is not None:
                tap[name + '.img'] = img
            return x, img
        x = self._conv_layer(x, blk.conv1, s1, noise_mode, pk, split_out=tc_next, rec=r1)
        if tap is not None:
            tap[name + '.conv1'] = self._as_f32(x)
        if self._keep is not None and rec is None:
            # inference: ToRGB_b and the skip-image chain only meet the main chain again at the planes, so they run
            # on the side stream next to block b+1's convolutions (small, latency-bound launches at batch 1)
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            self._keep.append(x)                      # x must outlive the side stream's read of it
            with torch.cuda.stream(self._side):
                img = self._torgb_layer(x, blk.torgb, st, img, pk)
        else:
            img = self._torgb_layer(x, blk.torgb, st, img, pk, rec=rt)
        if tap is not None:
            tap[name + '.img'] = img
        return x, img

    # ---------------------------------------------------------------- public protocol
    def synthesis(self, ws, c, neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                  use_cached_backbone=False, noise_mode='const', jitter_coarse=None, u_fine=None,
                  tap: Optional[Dict] = None, **synthesis_kwargs):
        """``ws [B,num_ws,w_dim]``, ``c [B,25]`` -> ``{'image','image_raw','image_depth'}`` (NCHW fp32).

        ``jitter_coarse [B,rays,S,1]`` / ``u_fine [B*rays,S_imp]`` are upstream's two random draws; when
        omitted they are drawn here with the same calls, shapes and order on the device.
        ``tap`` (tests only) collects channels-last intermediates.
        """
        cfg = self.cfg
        wgrads = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        train = torch.is_grad_enabled() and (ws.requires_grad or wgrads)
        if not ws.is_cuda:
            raise HfagpError('TriPlaneGenerator.synthesis needs CUDA tensors (there is no CPU fallback)')
        if ws.dim() != 3 or ws.shape[1] != cfg.num_ws or ws.shape[2] != cfg.w_dim:
            raise HfagpError(f'ws must be [B,{cfg.num_ws},{cfg.w_dim}], got {tuple(ws.shape)}')
        if c.dim() != 2 or c.shape[1] != cfg.c_dim or c.shape[0] != ws.shape[0]:
            raise HfagpError(f'c must be [B,{cfg.c_dim}], got {tuple(c.shape)}')
        res = neural_rendering_resolution or self.neural_rendering_resolution
        if train:
            return self._synthesis_train(ws, c, res, noise_mode, jitter_coarse, u_fine, tap)
        ws = ws.detach().float().contiguous()
        c = c.detach().float().contiguous()
        b = ws.shape[0]
        self._batch = b
        pk = self._ensure_packed()
        pe = getattr(self, 'profile_events', None)      # bench.py: per-stage CUDA events on the launching stream

        def mark():
            if pe is None:
                return None
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e

        t0 = mark()
        flat, offs = pk['styles'].run_flat(ws)
        styles = iter(pk['styles'].views(flat, offs, b))
        self._premod = self._modulate_all(pk, flat, offs, b)

        x = img = None
        self._keep = None                     # (a previous call may have been aborted by an exception)
        if self.overlap_streams and tap is None and pe is None:
            if self._side is None or self._side.device != ws.device:
                self._side = torch.cuda.Stream(device=ws.device)
            self._keep = []
        ri = None
        if self._keep is not None:
            # the renderer's random draws and depth range depend on nothing: first work of the side stream
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                ri = self._render_inputs(b, res, ws.device, jitter_coarse, u_fine, pk)
        for r in cfg.block_resolutions:
            blk = getattr(self.backbone.synthesis, f'b{r}')
            x, img = self._run_block(blk, x, img, styles, noise_mode, pk, tap, f'b{r}')
        if self._keep is not None:
            torch.cuda.current_stream().wait_stream(self._side)     # join: the renderer reads the planes
            self._keep = None
        planes = img                                             # [B,256,256,96] channels-last
        t1 = mark()
        if tap is not None:
            tap['planes'] = planes

        if ri is None:
            ri = self._render_inputs(b, res, ws.device, jitter_coarse, u_fine, pk)
        jitter, u_f, depth_range, rkw = ri
        s, sf = cfg.depth_res, cfg.depth_res_importance
        t2 = mark()
        feat, depth, wsum, book = ops.render(planes, c, pk['mlp'], pk['lin'], jitter, u_f, depth_range,
                                             bookkeeping=tap is not None, **rkw)
        t3 = mark()
        if tap is not None:
            tap.update(book)
            tap.update(feature_image=feat, weight_sum=wsum)

        rgb_lo = feat[..., :3].contiguous()                       # [B,res,res,3]
        x, img = feat, rgb_lo
        for i, blk in enumerate((self.superresolution.block0, self.superresolution.block1)):
            x, img = self._run_block(blk, x, img, styles, 'none', pk, tap, f'sr{i}')
        out = {'image': ops.nhwc_to_nchw(img),
               'image_raw': ops.nhwc_to_nchw(rgb_lo),
               'image_depth': depth.view(b, 1, res, res)}
        self._premod = None
        if pe is not None:
            t4 = mark()
            pe.extend([('backbone', t0, t1), ('render', t2, t3), ('superres', t3, t4), ('synthesis', t0, t4)])
        return out

    def _modulate_all(self, pk, flat, offs, batch):
        """Inference: every tensor-core layer's weights modulated (and demodulation coefficients computed) in one
        launch, before the first convolution — all styles of a frame are known up front."""
        if self.precision != 'tc':
            return None
        sel = []
        for (kind, m, _), off in zip(pk['order'], offs):
            pl = pk['layers'][id(m)]
            if self._use_tc(pl.cin) and not (kind == 'torgb' and pl.cout <= 4):
                sel.append((m, (pl.w, off, kind == 'conv')))
        if not sel:
            return None
        outs = ops.modulate_split_multi([e for _, e in sel], flat, batch)
        return {id(m): o for (m, _), o in zip(sel, outs)}

    def _render_inputs(self, b, res, device, jitter_coarse, u_fine, pk):
        cfg = self.cfg
        rays = res * res
        s, sf = cfg.depth_res, cfg.depth_res_importance
        if jitter_coarse is None and u_fine is None and self.fixed_draws is not None:
            jitter_coarse, u_fine = self.fixed_draws
        if jitter_coarse is None:
            jitter_coarse = torch.rand((b, rays, s, 1), device=device)
        if u_fine is None and sf > 0:
            u_fine = torch.rand((b * rays, sf), device=device)
        jitter = jitter_coarse.detach().reshape(b, rays, s).float().contiguous()
        delta = (cfg.ray_end - cfg.ray_start) / (s - 1)
        lin = pk['lin']
        depth_range = torch.stack([lin[0] + jitter[:, :, 0].min() * delta, lin[-1] + jitter[:, :, -1].max() * delta]).float().contiguous()
        kw = dict(res=res, s_coarse=s, s_fine=sf, delta=delta, box_scale=2.0 / cfg.box_warp)
        return jitter, (u_fine.detach().float().contiguous() if sf > 0 else None), depth_range, kw

    def _synthesis_train(self, ws, c, res, noise_mode, jitter_coarse, u_fine, tap):
        """synthesis() with autograd: same kernels, each stage an autograd.Function (hfa_gp_b200/autograd.py)
        that keeps the activations it produced and walks them backwards.  Gradient reaches ``ws`` only."""
        from . import autograd as ag
        b = ws.shape[0]
        self._batch = b
        self._premod = self._keep = None      # inference-only state (a previous call may have been aborted)
        pk = self._ensure_packed()
        c = c.detach().float().contiguous()
        if not ws.requires_grad:
            # only generator parameters are being trained: the stage Functions carry the graph through ws
            ws = ws.detach().requires_grad_(True)
        # the renderer's random draws and depth range depend on nothing: a side stream runs their ~10 small kernels next
        # to the backbone (as the inference path does)
        cur = torch.cuda.current_stream(ws.device)
        if self._side is None or self._side.device != ws.device:
            self._side = torch.cuda.Stream(device=ws.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            jitter, u, depth_range, kw = self._render_inputs(b, res, ws.device, jitter_coarse, u_fine, pk)
        styles_flat = ag.StylesFn.apply(ws.float().contiguous(), self)
        # every tensor-core layer's weights modulated in ONE launch (as at inference) instead of one launch per layer
        offs, _ = pk['styles'].offsets(b)
        self._premod = self._modulate_all(pk, styles_flat.detach(), offs, b)
        planes = ag.BackboneFn.apply(styles_flat, self, noise_mode, b, tap)
        if tap is not None:
            tap['planes'] = planes
        cur.wait_stream(self._side)
        for t_ in (jitter, u, depth_range):
            if t_ is not None:
                t_.record_stream(cur)
        feat, depth, wsum = ag.RenderFn.apply(planes, self, c, jitter, u, depth_range, kw, False)
        img = ag.SuperresFn.apply(feat, styles_flat, self, b, tap)
        self._premod = None
        return {'image': img.permute(0, 3, 1, 2), 'image_raw': feat[..., :3].permute(0, 3, 1, 2),
                'image_depth': depth.view(b, 1, res, res)}

    def forward(self, *a, **k):
        raise HfagpError('HFA-GP drives the generator through .synthesis(ws, c=..., noise_mode=...) only '
                         '(headnerf.py:112); the mapping network is not part of this path')


def make_generator(cfg: Optional[GeneratorConfig] = None, seed: int = 0, noise_strength: float = 0.0,
                   device='cuda') -> TriPlaneGenerator:
    """Seeded random-init generator for synthetic runs (no EG3D pickle is available offline)."""
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        g = TriPlaneGenerator(cfg)
    if noise_strength:
        with torch.no_grad():
            for name, p in g.named_parameters():
                if name.endswith('noise_strength') and name.startswith('backbone'):
                    p.fill_(noise_strength)
    return g.eval().requires_grad_(False).to(device)
