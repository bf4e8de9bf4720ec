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
