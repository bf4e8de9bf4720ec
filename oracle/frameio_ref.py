"""CPU ORACLE of the frame egress / ingress conversions.

TEST INFRASTRUCTURE ONLY (see ``oracle/eg3d_ref.py`` header for who may import this).

Restates, with plain torch CPU ops in the reference's order:
  save_image_uint8   torchvision.utils.save_image(img, normalize=True, range=(-1,1))  (run_recon_video_rgb.py:233-234):
                     make_grid's norm_ip  clamp_(low, high).sub_(low).div_(max(high-low, 1e-5))  then
                     mul(255).add_(0.5).clamp_(0,255).permute(1,2,0).to(uint8)
  layout_grid_uint8  (img * 127.5 + 128).clamp(0, 255).to(torch.uint8)  + CHW -> HWC  (run_recon_video_rgb.py:26-40)
  to_tensor_normalize  transforms.ToTensor() + Normalize([0.5]*3, [0.5]*3)  (train_rgb.py:78-81)

PINNED against the real torchvision functions the reference calls (tests/test_oracle_frameio.py; torchvision is in
this image, the reference's own scripts are not importable — missing lpips/imageio/... — so its call sites are
transcribed with their arguments).
"""
import torch


def save_image_uint8(img: torch.Tensor) -> torch.Tensor:
    """[N,3,H,W] float -> [N,H,W,3] uint8."""
    v = img.clone().float().clamp_(min=-1, max=1)
    v = v.sub_(-1).div_(max(1 - (-1), 1e-5))
    return v.mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8).contiguous()


def layout_grid_uint8(img: torch.Tensor) -> torch.Tensor:
    """[N,3,H,W] float -> [N,H,W,3] uint8."""
    return (img.float() * 127.5 + 128).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def to_tensor_normalize(u8: torch.Tensor) -> torch.Tensor:
    """[N,H,W,3] uint8 -> [N,3,H,W] float in [-1,1]."""
    v = u8.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
    mean = torch.tensor([0.5, 0.5, 0.5]).view(1, 3, 1, 1)
    std = torch.tensor([0.5, 0.5, 0.5]).view(1, 3, 1, 1)
    return v.sub_(mean).div_(std)


# ---------------------------------------------------------------- transforms.Resize on the decoded frame (PIL bilinear)
# The reference's ingress is  Resize(args.size) -> ToTensor -> Normalize  on a PIL image (run_recon_video_3dmm.py:258-261,
# run_recon_video_audio.py:258-261, train_rgb.py:78-81).  torchvision hands a PIL image to Image.resize(BILINEAR), i.e.
# Pillow's two-pass fixed-point resampler (src/libImaging/Resample.c, Pillow 12.2 in this image): coefficients in double
# (precompute_coeffs), rounded to 22-bit fixed point (normalize_coeffs_8bpc), a horizontal pass to a uint8 intermediate
# and a vertical pass, each `clip8((1 << 21) + sum(pixel * k)) >> 22`.  PINNED against Image.resize itself in
# tests/test_oracle_frameio.py.
PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0), box = the whole axis.
    Returns (ksize, bounds [out_size][2] = (xmin, count), coeffs [out_size][ksize] int32)."""
    import math
    import numpy as np
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = []
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0 else a
            w.append(1.0 - a if a < 1.0 else 0.0)
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(0.5 + v * (1 << PRECISION_BITS)) if v >= 0 else int(-0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis(img, out_size, axis):
    """One Pillow pass along ``axis`` of a uint8 array [..]: int32 accumulation, (1 << 21) rounding offset, clip8."""
    import numpy as np
    in_size = img.shape[axis]
    ksize, bounds, kk = pil_bilinear_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_uint8(u8: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """[N,H,W,C] uint8 -> [N,out_h,out_w,C] uint8 exactly as PIL.Image.resize((out_w, out_h), BILINEAR) per frame:
    horizontal pass first (only when the width changes), then the vertical pass (only when the height changes)."""
    a = u8.numpy()
    if a.shape[2] != out_w:
        a = _resample_axis(a, out_w, 2)
    if a.shape[1] != out_h:
        a = _resample_axis(a, out_h, 1)
    return torch.from_numpy(a.copy())


def resize_output_size(h: int, w: int, size: int):
    """transforms.Resize(int): the SHORTER side becomes ``size``, the other keeps the aspect ratio (int() truncation)."""
    if w <= h:
        return int(size * h / w), size
    return size, int(size * w / h)
