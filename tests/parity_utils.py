"""Helpers shared by the parity tests: build an oracle / product pair with identical weights, run both
with the same random draws and compare stage by stage."""
import math

import torch

from oracle import eg3d_ref, hfagp_ref

# Tolerance of the north star: 1e-3 relative fp32 per pixel.  "Relative" is taken against
# max(|reference pixel|, rms(reference tensor)) so that zero crossings do not blow the ratio up.
REL_TOL = 1e-3


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = torch.clamp(b.abs(), min=float(b.pow(2).mean().sqrt()) + 1e-12)
    return float(((a - b).abs() / scale).max())


def pixel_rel_stats(a: torch.Tensor, b: torch.Tensor) -> dict:
    """The PURE per-pixel relative error |a-b|/|b| (no rms floor), reported next to rel_err: its maximum is dominated by
    zero crossings of b (the ratio is unbounded where b -> 0), so the distribution is what is informative."""
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    r = (a - b).abs() / b.abs().clamp_min(1e-30)
    rms = float(b.pow(2).mean().sqrt())
    over = r > REL_TOL
    return dict(p50=float(r.median()), p999=float(r.kthvalue(int(0.999 * r.numel())).values), max=float(r.max()),
                frac_over=float(over.float().mean()),
                max_abs_ref_where_over=float((b.abs()[over].max() / rms) if bool(over.any()) else 0.0),
                max_over_away_from_zero=float(r[b.abs() > 0.05 * rms].max()))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b|| / ||b||.  Used for GRADIENTS on the tensor-core path: two forward passes that differ by 1e-5 flip
    the leaky-ReLU branch of the few activations sitting at the kink, which changes those elements' gradient by
    O(1) and makes a per-element max meaningless; the exact-fp32 kernels are held to the per-element tolerance."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def product_config(cfg):
    from hfa_gp_b200.generator import GeneratorConfig
    return GeneratorConfig(w_dim=cfg.w_dim, c_dim=cfg.c_dim, plane_res=cfg.plane_res,
                           plane_channels=cfg.plane_channels, channel_base=cfg.channel_base,
                           channel_max=cfg.channel_max, nrr=cfg.nrr, img_resolution=cfg.img_resolution,
                           sr_channels=cfg.sr_channels, sr_clamp=cfg.sr_clamp, decoder_hidden=cfg.decoder_hidden,
                           depth_res=cfg.depth_res, depth_res_importance=cfg.depth_res_importance,
                           ray_start=cfg.ray_start, ray_end=cfg.ray_end, box_warp=cfg.box_warp)


def make_pair(cfg, seed=0, noise_strength=0.1, bias_std=0.2):
    """Oracle generator (CPU) + product generator (cuda) holding the same tensors.  Biases and noise
    strengths are perturbed away from their zero init so those code paths are exercised."""
    from hfa_gp_b200.generator import TriPlaneGenerator
    ref = eg3d_ref.make_generator(cfg, seed=seed, noise_strength=noise_strength)
    g = torch.Generator().manual_seed(seed + 1000)
    with torch.no_grad():
        for name, p in ref.named_parameters():
            if name.endswith('.bias') and 'affine' not in name and 'mapping' not in name:
                p.copy_(torch.randn(p.shape, generator=g) * bias_std)
    prod = TriPlaneGenerator(product_config(cfg))
    missing = prod.load_state_dict(ref.state_dict(), strict=True)
    prod = prod.eval().requires_grad_(False).cuda()
    return ref, prod


def make_inputs(cfg, batch, seed=0):
    g = torch.Generator().manual_seed(seed + 7)
    ws = torch.randn(batch, cfg.num_ws, cfg.w_dim, generator=g)
    c = hfagp_ref.synthetic_labels(batch, seed=seed)
    hfagp_ref.flip_label_(c)                                  # what get_image does before synthesis
    rays = cfg.nrr ** 2
    jitter = torch.rand(batch, rays, cfg.depth_res, 1, generator=g)
    u = torch.rand(batch * rays, max(cfg.depth_res_importance, 1), generator=g)
    return ws, c, jitter, u


STAGES_NHWC = ['conv0', 'conv1', 'img']


def compare_taps(tap_ref, tap_gpu):
    """Yield (name, rel_err) for every float intermediate both sides recorded."""
    out = []
    for k, v in tap_ref.items():
        if k not in tap_gpu or not torch.is_floating_point(v):
            continue
        g = tap_gpu[k]
        if g.dim() == 4 and v.dim() == 4:
            g = to_nchw(g)                      # every 4-D tap of the CUDA path is channels-last
        g = g.reshape(v.shape) if g.numel() == v.numel() else g
        out.append((k, rel_err(g, v)))
    return out
