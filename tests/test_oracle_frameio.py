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


@pytest.mark.parametrize('h,w,size', [(512, 512, 256), (500, 500, 256), (300, 420, 256), (200, 200, 256), (256, 256, 256),
                                      (37, 53, 16), (512, 384, 128), (513, 513, 256)])
def test_resize_equals_pil_bilinear(h, w, size):
    """transforms.Resize(size) on a PIL image (the reference's ingress, run_recon_video_3dmm.py:258-261,
    train_rgb.py:78-81) = Pillow's two-pass fixed-point bilinear resampler: the oracle's integer restatement is bit-equal
    to PIL itself for down-scaling, up-scaling, identity and non-square frames; and so is the whole transform."""
    g = torch.Generator().manual_seed(h * 7 + w)
    u8 = torch.randint(0, 256, (1, h, w, 3), generator=g, dtype=torch.uint8)
    pil = PIL.fromarray(u8[0].numpy(), 'RGB')
    want = np.asarray(tv.transforms.Resize(size)(pil))
    oh, ow = frameio_ref.resize_output_size(h, w, size)
    got = frameio_ref.resize_uint8(u8, oh, ow)[0].numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    t = tv.transforms.Compose([tv.transforms.Resize(size), tv.transforms.ToTensor(),
                               tv.transforms.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])
    assert torch.equal(frameio_ref.to_tensor_normalize(frameio_ref.resize_uint8(u8, oh, ow))[0], t(pil))


def test_product_coefficient_table_equals_the_oracle():
    """hfa_gp_b200.frameio.pil_bilinear_table (host side of hfagp_frame_resize_u8) against the oracle's transcription of
    Pillow's precompute_coeffs + normalize_coeffs_8bpc."""
    from hfa_gp_b200 import frameio
    for a, b in ((512, 256), (500, 256), (200, 256), (256, 256), (53, 22), (420, 358), (1, 1), (3, 7)):
        ks, bounds, kk = frameio.pil_bilinear_table(a, b)
        ks_r, bounds_r, kk_r = frameio_ref.pil_bilinear_coeffs(a, b)
        assert ks == ks_r and np.array_equal(bounds.numpy(), bounds_r) and np.array_equal(kk.numpy(), kk_r)
        assert frameio.resize_output_size(300, 420, 256) == frameio_ref.resize_output_size(300, 420, 256) == (256, 358)
