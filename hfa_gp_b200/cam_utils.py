"""Camera helpers of the reference (code/cam_utils.py:12-80, trainer_rgb.py:27-33): look-at cameras on the
r = 2.7 sphere -> 25-d labels (cam2world 4x4 row-major + intrinsics).  Host-side, tiny, plain PyTorch."""
from __future__ import annotations

import math

import torch

INTRINSICS = (4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1)


def normalize_vecs(v: torch.Tensor) -> torch.Tensor:
    return v / torch.norm(v, dim=-1, keepdim=True)


def sample_camera_positions(device, n=1, r=1, horizontal_stddev=1, vertical_stddev=1, horizontal_mean=math.pi * 0.5,
                            vertical_mean=math.pi * 0.5, mode='normal', generator=None):
    """Points on a sphere of radius r; theta = yaw, phi = pitch.  Modes used by HFA-GP: 'gaussian'/'normal',
    'uniform' and None (the mean itself)."""
    if mode == 'uniform':
        theta = (torch.rand((n, 1), device=device, generator=generator) - 0.5) * 2 * horizontal_stddev + horizontal_mean
        phi = (torch.rand((n, 1), device=device, generator=generator) - 0.5) * 2 * vertical_stddev + vertical_mean
    elif mode in ('normal', 'gaussian'):
        theta = torch.randn((n, 1), device=device, generator=generator) * horizontal_stddev + horizontal_mean
        phi = torch.randn((n, 1), device=device, generator=generator) * vertical_stddev + vertical_mean
    elif mode is None:
        theta = torch.full((n, 1), float(horizontal_mean), device=device)
        phi = torch.full((n, 1), float(vertical_mean), device=device)
    else:
        raise ValueError(f'camera sampling mode {mode!r} is not used by the HFA-GP path')
    phi = torch.clamp(phi, 1e-5, math.pi - 1e-5)
    pts = torch.zeros((n, 3), device=device)
    pts[:, 0:1] = r * torch.sin(phi) * torch.cos(theta)
    pts[:, 2:3] = r * torch.sin(phi) * torch.sin(theta)
    pts[:, 1:2] = r * torch.cos(phi)
    return pts, phi, theta


def create_cam2world_matrix(forward_vector, origin, device=None):
    fwd = normalize_vecs(forward_vector)
    up = torch.tensor([0, 1, 0], dtype=torch.float, device=device).expand_as(fwd)
    left = normalize_vecs(torch.cross(up, fwd, dim=-1))
    up = normalize_vecs(torch.cross(fwd, left, dim=-1))
    n = fwd.shape[0]
    rot = torch.eye(4, device=device).unsqueeze(0).repeat(n, 1, 1)
    rot[:, :3, :3] = torch.stack((-left, up, -fwd), dim=-1)
    tr = torch.eye(4, device=device).unsqueeze(0).repeat(n, 1, 1)
    tr[:, :3, 3] = origin
    return tr @ rot


def cam_sampler(batch, device, generator=None, horizontal_stddev=0.3, vertical_stddev=0.155):
    """trainer_rgb.py:27-33 — random look-at labels [batch, 25]."""
    pts, _, _ = sample_camera_positions(device, n=batch, r=2.7, horizontal_mean=0.5 * math.pi,
                                        vertical_mean=0.5 * math.pi, horizontal_stddev=horizontal_stddev,
                                        vertical_stddev=vertical_stddev, mode='gaussian', generator=generator)
    c = create_cam2world_matrix(-pts, pts, device=device).reshape(batch, -1)
    intr = torch.tensor(INTRINSICS, dtype=torch.float32, device=device).reshape(1, -1).repeat(batch, 1)
    return torch.cat((c, intr), -1)
