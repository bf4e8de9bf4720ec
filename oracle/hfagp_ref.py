"""CPU fp32 ORACLE for the parts of the hot path the reference itself owns.

TEST INFRASTRUCTURE ONLY (see ``oracle/eg3d_ref.py`` header for who may import this).

PARITY PINNED against the reference's own code: ``tests/test_oracle_reference.py``
imports ``/root/reference/code/networks/{encoder3d,headnerf}.py`` (with ``dnnlib`` /
``legacy`` stubbed, ``oracle/ref_bridge.py``) and asserts these functions reproduce it on
seeded inputs; ``oracle/make_golden.py`` freezes reference outputs into
``tests/golden/encoder_*.npz`` so the pin travels to machines without ``/root/reference``.

Everything is a pure function of a ``state_dict`` whose keys are the reference's
(``net_app.convs.N...``, ``fc.N.weight`` ... — SURVEY.md App. A.9 / C).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


def blur_kernel():
    """make_kernel([1,3,3,1]) — encoder3d.py:47-56."""
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = k[None, :] * k[:, None]
    return k / k.sum()


def blur_ref(x, kernel, pad0, pad1):
    """Blur.forward = upfirdn2d(up=1, down=1, pad) — encoder3d.py:23-41,59-75.
    Zero-pad, then true convolution (flipped kernel) applied per channel."""
    c = x.shape[1]
    x = F.pad(x, [pad0, pad1, pad0, pad1])
    k = torch.flip(kernel, [0, 1])[None, None].repeat(c, 1, 1, 1)
    return F.conv2d(x, k, groups=c)


def equal_conv_ref(x, weight, bias=None, stride=1, padding=0):
    """EqualConv2d.forward — encoder3d.py:86-103 (scale 1/sqrt(Cin*k^2) folded per call)."""
    scale = 1.0 / math.sqrt(weight.shape[1] * weight.shape[2] ** 2)
    return F.conv2d(x, weight * scale, bias=bias, stride=stride, padding=padding)


def flrelu_ref(x, bias):
    """fused_leaky_relu — encoder3d.py:7-8."""
    return F.leaky_relu(x + bias, 0.2) * SQRT2


def conv_layer_ref(sd: Dict, prefix: str, x, k: int, downsample=False, activate=True):
    """ConvLayer — encoder3d.py:142-179.  Sub-module indices shift by one when a Blur leads."""
    i = 0
    if downsample:
        p = (4 - 2) + (k - 1)
        x = blur_ref(x, sd[f'{prefix}.0.kernel'], (p + 1) // 2, p // 2)
        i = 1
        x = equal_conv_ref(x, sd[f'{prefix}.{i}.weight'], None, stride=2, padding=0)
    else:
        x = equal_conv_ref(x, sd[f'{prefix}.{i}.weight'], None, stride=1, padding=k // 2)
    if activate:
        x = flrelu_ref(x, sd[f'{prefix}.{i + 1}.bias'])
    return x


def resblock_ref(sd, prefix, x):
    """ResBlock.forward — encoder3d.py:182-198."""
    out = conv_layer_ref(sd, f'{prefix}.conv1', x, 3)
    out = conv_layer_ref(sd, f'{prefix}.conv2', out, 3, downsample=True)
    skip = conv_layer_ref(sd, f'{prefix}.skip', x, 1, downsample=True, activate=False)
    return (out + skip) / SQRT2


def equal_linear_ref(x, weight, bias, lr_mul=1.0):
    """EqualLinear.forward with activation=None — encoder3d.py:112-136."""
    scale = (1.0 / math.sqrt(weight.shape[1])) * lr_mul
    return F.linear(x, weight * scale, bias=bias * lr_mul)


def fc_stack_ref(sd, prefix, x):
    i = 0
    while f'{prefix}.{i}.weight' in sd:
        x = equal_linear_ref(x, sd[f'{prefix}.{i}.weight'], sd[f'{prefix}.{i}.bias'])
        i += 1
    return x


def encoder_app_ref(sd, x, prefix='net_app.convs'):
    """EncoderApp.forward — encoder3d.py:201-239."""
    size = x.shape[-1]
    n_res = int(math.log2(size)) - 2
    h = conv_layer_ref(sd, f'{prefix}.0', x, 1)
    for i in range(1, n_res + 1):
        h = resblock_ref(sd, f'{prefix}.{i}', h)
    h = equal_conv_ref(h, sd[f'{prefix}.{n_res + 1}.weight'])
    return h.squeeze(-1).squeeze(-1)


def encoder_ref(sd, x, use_softmax=False, out_pose=False):
    """Encoder.get_weights / forward — encoder3d.py:242-298."""
    h = encoder_app_ref(sd, x)
    w = fc_stack_ref(sd, 'fc', h)
    if use_softmax:
        w = torch.softmax(w, dim=1)
    if out_pose:
        return w, fc_stack_ref(sd, 'pose', h)
    return w


def get_latent_ref(bases, delta, weights, dim=512):
    """HeadNeRF_*.get_latent — headnerf.py:81-102 (dup :182-195, :242-255).
    Thin QR of (bases+1e-8).T, then sum_j w_j Q[:, j] + delta, viewed [B, 14, dim]."""
    b = weights.shape[0]
    q, _ = torch.linalg.qr((bases + 1e-8).T, mode='reduced')
    out = torch.matmul(torch.diag_embed(weights), q.T).sum(dim=1)
    return out.view(b, -1, dim) + delta.view(-1, dim)


def flip_label_(label):
    """In-place GL flip done by get_image/forward — headnerf.py:108,132."""
    label[:, [1, 2, 5, 6, 9, 10]] *= -1
    return label


def weights_3dmm_ref(sd, params, use_softmax=False):
    """Weights_3DMM.forward — headnerf.py:138-158."""
    w = fc_stack_ref(sd, 'fc', params)
    return torch.softmax(w, dim=1) if use_softmax else w


def audionet_ref(sd, x, win_size=16):
    """AudioNet.forward — headnerf.py:319-349.  x [n,16,29] -> [n,dim_aud]."""
    half = win_size // 2
    x = x[:, 8 - half:8 + half, :].permute(0, 2, 1)
    for i in (0, 2, 4, 6):
        x = F.leaky_relu(F.conv1d(x, sd[f'encoder_conv.{i}.weight'], sd[f'encoder_conv.{i}.bias'],
                                  stride=2, padding=1), 0.02)
    x = x.squeeze(-1)
    x = F.leaky_relu(F.linear(x, sd['encoder_fc1.0.weight'], sd['encoder_fc1.0.bias']), 0.02)
    return F.linear(x, sd['encoder_fc1.2.weight'], sd['encoder_fc1.2.bias']).squeeze()


def audioattnet_ref(sd, x, dim_aud=32, seq_len=8):
    """AudioAttNet.forward — headnerf.py:284-314.  x [seq_len, D] -> [D]."""
    y = x[..., :dim_aud].permute(1, 0).unsqueeze(0)
    for i in (0, 2, 4, 6, 8):
        y = F.leaky_relu(F.conv1d(y, sd[f'attentionConvNet.{i}.weight'], sd[f'attentionConvNet.{i}.bias'],
                                  padding=1), 0.02)
    y = F.linear(y.view(1, seq_len), sd['attentionNet.0.weight'], sd['attentionNet.0.bias'])
    y = torch.softmax(y, dim=1).view(seq_len, 1)
    return torch.sum(y * x, dim=0)


# ------------------------------------------------------------------ synthetic weights / cameras

def make_encoder_state(size=256, dim=512, dim_motion=50, out_pose=False, seed=0, channels=None):
    """Random-init state_dict with the reference Encoder's keys/shapes/init (randn weights, 0 bias)."""
    channels = channels or {4: 512, 8: 512, 16: 512, 32: 512, 64: 256, 128: 128, 256: 64, 512: 32, 1024: 16}
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    sd = {}
    p = 'net_app.convs'
    cin = channels[size]
    sd[f'{p}.0.0.weight'] = rn(cin, 3, 1, 1)
    sd[f'{p}.0.1.bias'] = torch.zeros(1, cin, 1, 1)
    log_size = int(math.log2(size))
    idx = 1
    for i in range(log_size, 2, -1):
        cout = channels[2 ** (i - 1)]
        sd[f'{p}.{idx}.conv1.0.weight'] = rn(cin, cin, 3, 3)
        sd[f'{p}.{idx}.conv1.1.bias'] = torch.zeros(1, cin, 1, 1)
        sd[f'{p}.{idx}.conv2.0.kernel'] = blur_kernel()
        sd[f'{p}.{idx}.conv2.1.weight'] = rn(cout, cin, 3, 3)
        sd[f'{p}.{idx}.conv2.2.bias'] = torch.zeros(1, cout, 1, 1)
        sd[f'{p}.{idx}.skip.0.kernel'] = blur_kernel()
        sd[f'{p}.{idx}.skip.1.weight'] = rn(cout, cin, 1, 1)
        cin = cout
        idx += 1
    sd[f'{p}.{idx}.weight'] = rn(dim, cin, 4, 4)
    dims = [dim] * 5 + [dim_motion]
    for j in range(5):
        sd[f'fc.{j}.weight'] = rn(dims[j + 1], dims[j])
        sd[f'fc.{j}.bias'] = torch.zeros(dims[j + 1])
    if out_pose:
        dims = [dim] * 5 + [25]
        for j in range(5):
            sd[f'pose.{j}.weight'] = rn(dims[j + 1], dims[j])
            sd[f'pose.{j}.bias'] = torch.zeros(dims[j + 1])
    return sd


INTRINSICS = [4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1]


def lookat_label(theta, phi, r=2.7):
    """sample_camera_positions + create_cam2world_matrix(-p, p) + intrinsics — cam_utils.py:12-80,
    trainer_rgb.py:27-33.  theta/phi: [n] tensors (yaw, pitch in radians). Returns c [n,25]."""
    theta = torch.as_tensor(theta, dtype=torch.float32).reshape(-1, 1)
    phi = torch.as_tensor(phi, dtype=torch.float32).reshape(-1, 1).clamp(1e-5, math.pi - 1e-5)
    n = theta.shape[0]
    pts = torch.zeros(n, 3)
    pts[:, 0:1] = r * torch.sin(phi) * torch.cos(theta)
    pts[:, 2:3] = r * torch.sin(phi) * torch.sin(theta)
    pts[:, 1:2] = r * torch.cos(phi)
    nrm = lambda v: v / torch.norm(v, dim=-1, keepdim=True)
    fwd = nrm(-pts)
    up = torch.tensor([0.0, 1.0, 0.0]).expand_as(fwd)
    left = nrm(torch.cross(up, fwd, dim=-1))
    up = nrm(torch.cross(fwd, left, dim=-1))
    rot = torch.eye(4)[None].repeat(n, 1, 1)
    rot[:, :3, :3] = torch.stack((-left, up, -fwd), dim=-1)
    tr = torch.eye(4)[None].repeat(n, 1, 1)
    tr[:, :3, 3] = pts
    c2w = (tr @ rot).reshape(n, 16)
    return torch.cat([c2w, torch.tensor(INTRINSICS)[None].repeat(n, 1)], dim=-1)


def synthetic_labels(n, seed=0):
    """cam_sampler (trainer_rgb.py:27-33): yaw ~ N(pi/2, 0.3), pitch ~ N(pi/2, 0.155), seeded."""
    g = torch.Generator().manual_seed(seed)
    theta = torch.randn(n, generator=g) * 0.3 + 0.5 * math.pi
    phi = torch.randn(n, generator=g) * 0.155 + 0.5 * math.pi
    return lookat_label(theta, phi)
