"""CPU fp32 ORACLE for the EG3D tri-plane generator that HFA-GP renders with.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``hfa_gp_b200``); only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and only
as the checker / the timed CPU arm.

PARITY UNPINNED.  The arithmetic of this path is *not* in the reference repo:
``/root/reference/code/networks/headnerf.py:6-7`` imports ``dnnlib``/``legacy`` and
``headnerf.py:31-38`` un-pickles NVlabs/eg3d's ``TriPlaneGenerator`` (dependency is
un-vendored and un-pinned: no submodule, no requirements file, no commit hash), every
render being ``generator.synthesis(latent, c=label, noise_mode='const')['image']``
(``headnerf.py:112,118,133,207,218,267,277``).  The reference holds no tests, golden
vectors or fixtures for it.  This file therefore restates the *published* algorithm
of NVlabs/eg3d (``training/triplane.py``, ``training/networks_stylegan2.py``,
``training/superresolution.py``, ``training/volumetric_rendering/*``) for the
``ffhqrebalanced512-128`` configuration named at ``headnerf.py:31``, as specified in
SURVEY.md Appendix A, and is anchored on the reference's own call sites:
``ws`` is ``[B,14,512]`` (``headnerf.py:55,100``), ``c`` is ``[B,25]``
(``trainer_rgb.py:30-32``), the result dict is indexed with ``'image'``
(``headnerf.py:112``) and is a 512x512 image in about [-1,1]
(``run_recon_video_rgb.py:233-234``).  Self-consistency identities (App. A.10) are
checked in ``tests/test_oracle_generator.py``.

Differences from upstream that are deliberate: the two random draws of the renderer
(stratified jitter, importance ``u``) are explicit arguments so the CUDA path can be
fed the very same numbers (upstream draws them with ``torch.rand_like`` /
``torch.rand`` in this order and with these shapes); everything runs in fp32 with the
super-resolution clamp (what upstream executes on a CPU device).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn


# ----------------------------------------------------------------------------- config

@dataclass
class GeneratorConfig:
    """Hyper-parameters carried by the EG3D pickle."""
    w_dim: int = 512
    c_dim: int = 25
    plane_res: int = 256            # backbone img_resolution
    plane_channels: int = 32        # per plane, x3 = backbone img_channels 96
    channel_base: int = 32768
    channel_max: int = 512
    nrr: int = 128                  # neural_rendering_resolution
    img_resolution: int = 512
    sr_channels: tuple = (256, 128)  # SuperresolutionHybrid8XDC block0/1 out channels
    sr_clamp: float = 256.0
    decoder_hidden: int = 64
    depth_res: int = 48
    depth_res_importance: int = 48
    ray_start: float = 2.25
    ray_end: float = 3.3
    box_warp: float = 1.0
    mapping_layers: int = 2

    @property
    def block_resolutions(self):
        return [2 ** i for i in range(2, int(math.log2(self.plane_res)) + 1)]

    def channels(self, res: int) -> int:
        return min(self.channel_base // res, self.channel_max)

    @property
    def num_ws(self) -> int:
        return 2 * len(self.block_resolutions)      # 1 + 2*(n-1) convs + last torgb


def tiny_config() -> GeneratorConfig:
    """A reduced generator (same topology) so CPU tests run in seconds."""
    return GeneratorConfig(plane_res=32, channel_base=1024, channel_max=64, nrr=16,
                           img_resolution=64, sr_channels=(64, 32), depth_res=12,
                           depth_res_importance=12)


def small14_config() -> GeneratorConfig:
    """Full depth (7 backbone blocks -> num_ws == 14, what get_latent produces) but thin channels."""
    return GeneratorConfig(plane_res=256, channel_base=2048, channel_max=64, nrr=16, img_resolution=64,
                           sr_channels=(32, 16), depth_res=16, depth_res_importance=16)


# ----------------------------------------------------------------------------- StyleGAN2 ops

def setup_filter():
    """[1,3,3,1] (x) [1,3,3,1] / 64 — the resample filter of every block."""
    f = torch.tensor([1.0, 3.0, 3.0, 1.0])
    f = torch.outer(f, f)
    return f / f.sum()


def upfirdn2d_ref(x, f, up=1, pad=(0, 0, 0, 0), gain=1.0):
    """Zero-insert by ``up``, pad [x0,x1,y0,y1], true-convolve with ``f*gain``."""
    n, c, h, w = x.shape
    if up > 1:
        z = x.new_zeros(n, c, h, up, w, up)
        z[:, :, :, 0, :, 0] = x
        x = z.reshape(n, c, h * up, w * up)
    x = F.pad(x, [pad[0], pad[1], pad[2], pad[3]])
    k = (f * gain).flip([0, 1])[None, None].repeat(c, 1, 1, 1)
    return F.conv2d(x, k, groups=c)


def upsample2d_ref(x, f):
    """Skip-image x2 upsample: pad [2,1,2,1], gain 4 (App. A.4)."""
    return upfirdn2d_ref(x, f, up=2, pad=(2, 1, 2, 1), gain=4.0)


def modulated_conv2d_ref(x, weight, styles, noise=None, up=1, demodulate=True, f=None,
                         fused=True):
    """Style-modulated convolution (App. A.4).

    fused:      grouped conv with per-sample weights  w' * d
    non-fused:  (x*s) -> shared-weight conv -> *d         (same math, different rounding)
    up=2:       conv_transpose2d(stride 2, weight not flipped) -> 4x4 FIR pad 1 gain 4
    Noise is added after demodulation, before bias/activation.
    """
    b, cin, h, w = x.shape
    cout, _, kh, kw = weight.shape
    pad = kh // 2
    dcoef = None
    if demodulate:
        wmod = weight[None] * styles[:, None, :, None, None]            # [B,O,I,k,k]
        dcoef = (wmod.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()        # [B,O]

    def conv(inp, wt, groups):
        if up == 1:
            return F.conv2d(inp, wt, padding=pad, groups=groups)
        # transposed path: weight [O,I,k,k] -> [I,O,k,k] per group, no flip
        if groups == 1:
            wt_t = wt.transpose(0, 1)
        else:
            wt_t = wt.reshape(groups, cout, cin, kh, kw).transpose(1, 2).reshape(groups * cin, cout, kh, kw)
        y = F.conv_transpose2d(inp, wt_t, stride=2, padding=0, groups=groups)
        return upfirdn2d_ref(y, f, up=1, pad=(1, 1, 1, 1), gain=4.0)

    if fused:
        wmod = weight[None] * styles[:, None, :, None, None]
        if demodulate:
            wmod = wmod * dcoef[:, :, None, None, None]
        y = conv(x.reshape(1, b * cin, h, w), wmod.reshape(b * cout, cin, kh, kw), b)
        y = y.reshape(b, cout, y.shape[-2], y.shape[-1])
        if noise is not None:
            y = y + noise
    else:
        y = conv(x * styles[:, :, None, None], weight, 1)
        if demodulate:
            y = y * dcoef[:, :, None, None]
        if noise is not None:
            y = y + noise
    return y


def bias_act_ref(x, b=None, act='linear', gain=None, clamp=None):
    """+b -> act -> *gain -> clamp(+-c) (App. A.4)."""
    if b is not None:
        x = x + b.reshape(1, -1, *([1] * (x.ndim - 2)))
    if act == 'lrelu':
        x = F.leaky_relu(x, 0.2)
        g = math.sqrt(2.0) if gain is None else gain
    else:
        g = 1.0 if gain is None else gain
    if g != 1.0:
        x = x * g
    if clamp is not None:
        x = x.clamp(-clamp, clamp)
    return x


class FullyConnectedRef(nn.Module):
    def __init__(self, cin, cout, bias_init=0.0, lr_mul=1.0):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(cout, cin) / lr_mul)
        self.bias = nn.Parameter(torch.full([cout], float(bias_init)))
        self.weight_gain = lr_mul / math.sqrt(cin)
        self.bias_gain = lr_mul

    def forward(self, x):
        return torch.addmm((self.bias * self.bias_gain)[None], x, (self.weight * self.weight_gain).t())


class SynthesisLayerRef(nn.Module):
    def __init__(self, cin, cout, w_dim, res, up=1, use_noise=True, clamp=None):
        super().__init__()
        self.up, self.res, self.use_noise, self.clamp = up, res, use_noise, clamp
        self.affine = FullyConnectedRef(w_dim, cin, bias_init=1.0)
        self.weight = nn.Parameter(torch.randn(cout, cin, 3, 3))
        self.register_buffer('resample_filter', setup_filter())
        self.register_buffer('noise_const', torch.randn(res, res))
        self.noise_strength = nn.Parameter(torch.zeros([]))
        self.bias = nn.Parameter(torch.zeros(cout))

    def forward(self, x, w, noise_mode='const', fused=True, gain=1.0):
        styles = self.affine(w)
        noise = None
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        elif self.use_noise and noise_mode == 'random':
            noise = torch.randn(x.shape[0], 1, self.res, self.res) * self.noise_strength
        x = modulated_conv2d_ref(x, self.weight, styles, noise=noise, up=self.up,
                                 f=self.resample_filter, fused=fused)
        clamp = self.clamp * gain if self.clamp is not None else None
        return bias_act_ref(x, self.bias, act='lrelu', gain=math.sqrt(2.0) * gain, clamp=clamp)


class ToRGBLayerRef(nn.Module):
    def __init__(self, cin, cout, w_dim, clamp=None):
        super().__init__()
        self.clamp = clamp
        self.affine = FullyConnectedRef(w_dim, cin, bias_init=1.0)
        self.weight = nn.Parameter(torch.randn(cout, cin, 1, 1))
        self.bias = nn.Parameter(torch.zeros(cout))
        self.weight_gain = 1.0 / math.sqrt(cin)

    def forward(self, x, w, fused=True):
        styles = self.affine(w) * self.weight_gain
        x = modulated_conv2d_ref(x, self.weight, styles, demodulate=False, fused=fused)
        return bias_act_ref(x, self.bias, clamp=self.clamp)


class SynthesisBlockRef(nn.Module):
    """'skip' architecture block: [const|conv0(up2)] -> conv1 -> img = up(img) + torgb(x)."""

    def __init__(self, cin, cout, w_dim, res, img_channels, clamp=None, use_noise=True):
        super().__init__()
        self.cin, self.res = cin, res
        self.register_buffer('resample_filter', setup_filter())
        if cin == 0:
            self.const = nn.Parameter(torch.randn(cout, res, res))
        else:
            self.conv0 = SynthesisLayerRef(cin, cout, w_dim, res, up=2, use_noise=use_noise, clamp=clamp)
        self.conv1 = SynthesisLayerRef(cout, cout, w_dim, res, use_noise=use_noise, clamp=clamp)
        self.torgb = ToRGBLayerRef(cout, img_channels, w_dim, clamp=clamp)
        self.num_conv = 1 if cin == 0 else 2

    def forward(self, x, img, ws, noise_mode='const', fused=True, tap=None, name=''):
        it = iter(ws.unbind(dim=1))
        if self.cin == 0:
            x = self.const[None].repeat(ws.shape[0], 1, 1, 1)
        else:
            x = self.conv0(x, next(it), noise_mode=noise_mode, fused=fused)
            if tap is not None:
                tap[name + '.conv0'] = x
        x = self.conv1(x, next(it), noise_mode=noise_mode, fused=fused)
        if tap is not None:
            tap[name + '.conv1'] = x
        if img is not None:
            img = upsample2d_ref(img, self.resample_filter)
        y = self.torgb(x, next(it), fused=fused)
        img = y if img is None else img + y
        if tap is not None:
            tap[name + '.img'] = img
        return x, img


class SynthesisNetworkRef(nn.Module):
    def __init__(self, cfg: GeneratorConfig):
        super().__init__()
        self.cfg = cfg
        self.block_resolutions = cfg.block_resolutions
        for res in self.block_resolutions:
            cin = cfg.channels(res // 2) if res > 4 else 0
            setattr(self, f'b{res}', SynthesisBlockRef(cin, cfg.channels(res), cfg.w_dim, res,
                                                       3 * cfg.plane_channels))

    def forward(self, ws, noise_mode='const', fused=True, tap=None):
        x = img = None
        w_idx = 0
        for res in self.block_resolutions:
            blk = getattr(self, f'b{res}')
            cur = ws.narrow(1, w_idx, blk.num_conv + 1)       # torgb shares w with next conv0
            w_idx += blk.num_conv
            x, img = blk(x, img, cur, noise_mode=noise_mode, fused=fused, tap=tap, name=f'b{res}')
        return img


class MappingNetworkRef(nn.Module):
    """Never executed by HFA-GP (ws come from get_latent, headnerf.py:81-102); its
    tensors only ride along in state_dict / checkpoints (trainer_rgb.py:146)."""

    def __init__(self, cfg: GeneratorConfig):
        super().__init__()
        self.embed = FullyConnectedRef(cfg.c_dim, cfg.w_dim)
        self.fc0 = FullyConnectedRef(2 * cfg.w_dim, cfg.w_dim, lr_mul=0.01)
        self.fc1 = FullyConnectedRef(cfg.w_dim, cfg.w_dim, lr_mul=0.01)
        self.register_buffer('w_avg', torch.zeros(cfg.w_dim))


class BackboneRef(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.synthesis = SynthesisNetworkRef(cfg)
        self.mapping = MappingNetworkRef(cfg)


class SuperresolutionRef(nn.Module):
    """SuperresolutionHybrid8XDC (App. A.6): two 'skip' blocks, all layers driven by ws[:, -1]."""

    def __init__(self, cfg: GeneratorConfig):
        super().__init__()
        c0, c1 = cfg.sr_channels
        r = cfg.nrr
        self.block0 = SynthesisBlockRef(cfg.plane_channels, c0, cfg.w_dim, 2 * r, 3, clamp=cfg.sr_clamp)
        self.block1 = SynthesisBlockRef(c0, c1, cfg.w_dim, 4 * r, 3, clamp=cfg.sr_clamp)

    def forward(self, rgb, x, ws, fused=True, tap=None):
        ws = ws[:, -1:, :].repeat(1, 3, 1)
        x, rgb = self.block0(x, rgb, ws, noise_mode='none', fused=fused, tap=tap, name='sr0')
        x, rgb = self.block1(x, rgb, ws, noise_mode='none', fused=fused, tap=tap, name='sr1')
        return rgb


# ----------------------------------------------------------------------------- renderer

PLANE_AXES = torch.tensor([[[1, 0, 0], [0, 1, 0], [0, 0, 1]],
                           [[1, 0, 0], [0, 0, 1], [0, 1, 0]],
                           [[0, 0, 1], [1, 0, 0], [0, 1, 0]]], dtype=torch.float32)


def ray_sampler_ref(cam2world, intrinsics, res):
    """Pinhole rays through pixel centres, x fastest (App. A.5)."""
    n = cam2world.shape[0]
    cam_pos = cam2world[:, :3, 3]
    fx, fy = intrinsics[:, 0, 0], intrinsics[:, 1, 1]
    cx, cy, sk = intrinsics[:, 0, 2], intrinsics[:, 1, 2], intrinsics[:, 0, 1]
    ar = torch.arange(res, dtype=torch.float32)
    uv = torch.stack(torch.meshgrid(ar, ar, indexing='ij')) * (1.0 / res) + (0.5 / res)
    uv = uv.flip(0).reshape(2, -1).transpose(1, 0)[None].repeat(n, 1, 1)
    x_cam, y_cam = uv[:, :, 0], uv[:, :, 1]
    z_cam = torch.ones(n, res * res)
    x_lift = (x_cam - cx[:, None] + cy[:, None] * sk[:, None] / fy[:, None]
              - sk[:, None] * y_cam / fy[:, None]) / fx[:, None] * z_cam
    y_lift = (y_cam - cy[:, None]) / fy[:, None] * z_cam
    pts = torch.stack((x_lift, y_lift, z_cam, torch.ones_like(z_cam)), dim=-1)
    world = torch.bmm(cam2world, pts.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :3]
    dirs = F.normalize(world - cam_pos[:, None, :], dim=2)
    origins = cam_pos[:, None, :].repeat(1, dirs.shape[1], 1)
    return origins, dirs


def sample_from_planes_ref(planes, coords, box_warp):
    """planes [N,3,C,H,W], coords [N,M,3] -> [N,3,M,C]; bilinear, zeros, align_corners=False."""
    n, p, c, h, w = planes.shape
    m = coords.shape[1]
    coords = (2.0 / box_warp) * coords
    cexp = coords[:, None].expand(-1, p, -1, -1).reshape(n * p, m, 3)
    inv = torch.linalg.inv(PLANE_AXES)[None].expand(n, -1, -1, -1).reshape(n * p, 3, 3)
    proj = torch.bmm(cexp, inv)[..., :2]
    out = F.grid_sample(planes.reshape(n * p, c, h, w), proj[:, None].float(), mode='bilinear',
                        padding_mode='zeros', align_corners=False)
    return out.permute(0, 3, 2, 1).reshape(n, p, m, c)


class OSGDecoderRef(nn.Module):
    def __init__(self, cfg: GeneratorConfig):
        super().__init__()
        self.net = nn.Sequential(FullyConnectedRef(cfg.plane_channels, cfg.decoder_hidden), nn.Softplus(),
                                 FullyConnectedRef(cfg.decoder_hidden, 1 + cfg.plane_channels))

    def forward(self, feats):
        x = feats.mean(1)
        n, m, c = x.shape
        x = self.net(x.reshape(n * m, c)).reshape(n, m, -1)
        rgb = torch.sigmoid(x[..., 1:]) * (1 + 2 * 0.001) - 0.001
        return rgb, x[..., 0:1]


def ray_march_ref(colors, densities, depths):
    """MipRayMarcher2 (clamp_mode='softplus', white_back=False)."""
    deltas = depths[:, :, 1:] - depths[:, :, :-1]
    colors_mid = (colors[:, :, :-1] + colors[:, :, 1:]) / 2
    dens_mid = (densities[:, :, :-1] + densities[:, :, 1:]) / 2
    depths_mid = (depths[:, :, :-1] + depths[:, :, 1:]) / 2
    dens_mid = F.softplus(dens_mid - 1)
    alpha = 1 - torch.exp(-(dens_mid * deltas))
    shifted = torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-10], -2)
    weights = alpha * torch.cumprod(shifted, -2)[:, :, :-1]
    rgb = torch.sum(weights * colors_mid, -2)
    wtot = weights.sum(2)
    depth = torch.sum(weights * depths_mid, -2) / wtot
    depth = torch.nan_to_num(depth, float('inf'))
    depth = torch.clamp(depth, torch.min(depths), torch.max(depths))
    return rgb * 2 - 1, depth, weights


def smooth_weights_ref(weights):
    """max_pool1d(2,1,pad 1) -> avg_pool1d(2,1) -> +0.01 on the per-interval weights."""
    w = F.max_pool1d(weights[:, None].float(), 2, 1, padding=1)
    w = F.avg_pool1d(w, 2, 1).squeeze(1)
    return w + 0.01


def sample_pdf_ref(bins, weights, u, eps=1e-5):
    """Inverse-CDF sampling with upstream's index conventions; returns samples + integer bookkeeping."""
    n_samples_ = weights.shape[1]
    weights = weights + eps
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n_samples_)
    idx = torch.stack([below, above], -1).view(u.shape[0], 2 * u.shape[1])
    cdf_g = torch.gather(cdf, 1, idx).view(u.shape[0], u.shape[1], 2)
    bins_g = torch.gather(bins, 1, idx).view(u.shape[0], u.shape[1], 2)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom[denom < eps] = 1
    samples = bins_g[..., 0] + (u - cdf_g[..., 0]) / denom * (bins_g[..., 1] - bins_g[..., 0])
    return samples, dict(cdf=cdf, inds=inds, below=below, above=above)


def render_ref(planes, decoder, ray_o, ray_d, cfg: GeneratorConfig, jitter_coarse=None, u_fine=None,
               tap: Optional[Dict] = None):
    """ImportanceRenderer.forward (App. A.5).  planes [N,3,C,H,W]; rays [N,M,3].

    jitter_coarse [N,M,S,1] and u_fine [N*M,S_imp] are upstream's two random draws
    (``torch.rand_like(depths_coarse)`` then ``torch.rand(N_rays, N_importance)``).
    """
    n, m, _ = ray_o.shape
    s = cfg.depth_res
    if jitter_coarse is None:
        jitter_coarse = torch.rand(n, m, s, 1)
    depths_c = torch.linspace(cfg.ray_start, cfg.ray_end, s).reshape(1, 1, s, 1).repeat(n, m, 1, 1)
    delta = (cfg.ray_end - cfg.ray_start) / (s - 1)
    depths_c = depths_c + jitter_coarse * delta

    def run(depths):
        k = depths.shape[2]
        pts = (ray_o[:, :, None] + depths * ray_d[:, :, None]).reshape(n, -1, 3)
        feats = sample_from_planes_ref(planes, pts, cfg.box_warp)
        rgb, sigma = decoder(feats)
        return rgb.reshape(n, m, k, -1), sigma.reshape(n, m, k, 1)

    col_c, den_c = run(depths_c)
    if tap is not None:
        tap.update(depths_coarse=depths_c, colors_coarse=col_c, densities_coarse=den_c)
    s_imp = cfg.depth_res_importance
    if s_imp > 0:
        _, _, w_c = ray_march_ref(col_c, den_c, depths_c)
        with torch.no_grad():
            z = depths_c.reshape(n * m, s)
            w = smooth_weights_ref(w_c.reshape(n * m, -1))
            z_mid = 0.5 * (z[:, :-1] + z[:, 1:])
            if u_fine is None:
                u_fine = torch.rand(n * m, s_imp)
            z_f, book = sample_pdf_ref(z_mid, w[:, 1:-1], u_fine)
            depths_f = z_f.detach().reshape(n, m, s_imp, 1)
        col_f, den_f = run(depths_f)
        all_d = torch.cat([depths_c, depths_f], dim=-2)
        all_c = torch.cat([col_c, col_f], dim=-2)
        all_s = torch.cat([den_c, den_f], dim=-2)
        _, order = torch.sort(all_d, dim=-2)
        all_d = torch.gather(all_d, -2, order)
        all_c = torch.gather(all_c, -2, order.expand(-1, -1, -1, all_c.shape[-1]))
        all_s = torch.gather(all_s, -2, order)
        if tap is not None:
            tap.update(weights_coarse=w_c, depths_fine=depths_f, sort_idx=order, depths_sorted=all_d,
                       colors_fine=col_f, densities_fine=den_f, **book)
        rgb, depth, weights = ray_march_ref(all_c, all_s, all_d)
    else:
        rgb, depth, weights = ray_march_ref(col_c, den_c, depths_c)
    return rgb, depth, weights.sum(2)


# ----------------------------------------------------------------------------- generator

class TriPlaneGeneratorRef(nn.Module):
    """state_dict keys follow SURVEY.md App. A.9 (``backbone.synthesis.b4.const`` ...)."""

    def __init__(self, cfg: Optional[GeneratorConfig] = None):
        super().__init__()
        self.cfg = cfg or GeneratorConfig()
        self.backbone = BackboneRef(self.cfg)
        self.superresolution = SuperresolutionRef(self.cfg)
        self.decoder = OSGDecoderRef(self.cfg)
        self.neural_rendering_resolution = self.cfg.nrr

    @property
    def num_ws(self):
        return self.cfg.num_ws

    def synthesis(self, ws, c, noise_mode='const', jitter_coarse=None, u_fine=None, fused=True,
                  tap: Optional[Dict] = None, **_unused):
        cfg = self.cfg
        n = ws.shape[0]
        cam2world = c[:, :16].view(-1, 4, 4)
        intrinsics = c[:, 16:25].view(-1, 3, 3)
        ray_o, ray_d = ray_sampler_ref(cam2world, intrinsics, cfg.nrr)
        planes = self.backbone.synthesis(ws, noise_mode=noise_mode, fused=fused, tap=tap)
        if tap is not None:
            tap.update(planes=planes, ray_origins=ray_o, ray_dirs=ray_d)
        planes = planes.view(n, 3, cfg.plane_channels, planes.shape[-2], planes.shape[-1])
        feat, depth, wsum = render_ref(planes, self.decoder, ray_o, ray_d, cfg, jitter_coarse, u_fine, tap)
        feature_image = feat.permute(0, 2, 1).reshape(n, feat.shape[-1], cfg.nrr, cfg.nrr).contiguous()
        depth_image = depth.permute(0, 2, 1).reshape(n, 1, cfg.nrr, cfg.nrr)
        rgb_image = feature_image[:, :3]
        if tap is not None:
            tap.update(feature_image=feature_image, weight_sum=wsum)
        sr_image = self.superresolution(rgb_image, feature_image, ws, fused=fused, tap=tap)
        return {'image': sr_image, 'image_raw': rgb_image, 'image_depth': depth_image}


def make_generator(cfg: Optional[GeneratorConfig] = None, seed: int = 0, noise_strength: float = 0.0):
    """Seeded random-init generator (App. A.8): weights N(0,1), biases 0, affine bias 1."""
    g = torch.Generator().manual_seed(seed)
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        net = TriPlaneGeneratorRef(cfg)
    if noise_strength:
        with torch.no_grad():
            for name, p in net.named_parameters():
                if name.endswith('noise_strength') and name.startswith('backbone'):
                    p.fill_(noise_strength)
    del g
    return net.eval().requires_grad_(False)
