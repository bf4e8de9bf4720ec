"""CPU: pin oracle/frameio_ref.py against the torchvision functions the reference calls for frame egress / ingress
(run_recon_video_rgb.py:233-234, :26-40; train_rgb.py:78-81), bit for bit."""
import io

import numpy as np
import pytest
import torch

from oracle import frameio_ref

tv = pytest.importorskip('torchvision')
PIL = pytest.importorskip('PIL.Image')


def _img(seed, n=1, h=37, w=53):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, h, w, generator=g) * 0.8            # some values beyond [-1, 1]
    x.view(-1)[:8] = torch.tensor([-1.0, 1.0, 0.0, -1.5, 1.5, 0.999999, -0.999999, 1e-8])
    return x


def test_save_image_convention_equals_torchvision():
    x = _img(0)
    buf = io.BytesIO()
    tv.utils.save_image(x.clone(), buf, format='png', normalize=True, value_range=(-1, 1))   # `range=` in the reference's
    got = np.array(PIL.open(io.BytesIO(buf.getvalue())).convert('RGB'))                      # older torchvision
    want = frameio_ref.save_image_uint8(x)[0].numpy()
    assert got.shape == want.shape and np.array_equal(got, want)


def test_layout_grid_convention_is_the_reference_expression():
    x = _img(1, n=2)
    want = (x * 127.5 + 128).clamp(0, 255).to(torch.uint8)          # run_recon_video_rgb.py:34, verbatim
    assert torch.equal(frameio_ref.layout_grid_uint8(x), want.permute(0, 2, 3, 1))


def test_ingress_equals_torchvision_transforms():
    g = torch.Generator().manual_seed(2)
    u8 = torch.randint(0, 256, (1, 41, 29, 3), generator=g, dtype=torch.uint8)
    u8.view(-1)[:3] = torch.tensor([0, 255, 128], dtype=torch.uint8)
    pil = PIL.fromarray(u8[0].numpy(), 'RGB')
    t = tv.transforms.Compose([tv.transforms.ToTensor(), tv.transforms.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])
    assert torch.equal(frameio_ref.to_tensor_normalize(u8)[0], t(pil))
