"""CPU: self-consistency identities of the EG3D oracle restatement — the
substitute for upstream golden vectors, which do not exist (parity unpinned)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import eg3d_ref as E
from oracle import hfagp_ref as H


@pytest.fixture(scope='module')
def tiny():
    cfg = E.tiny_config()
    g = E.make_generator(cfg, seed=0, noise_strength=0.1)
    gen = torch.Generator().manual_seed(3)
    ws = torch.randn(2, cfg.num_ws, cfg.w_dim, generator=gen)
    c = H.flip_label_(H.synthetic_labels(2, seed=1))
    jit = torch.rand(2, cfg.nrr ** 2, cfg.depth_res, 1, generator=gen)
    u = torch.rand(2 * cfg.nrr ** 2, cfg.depth_res_importance, generator=gen)
    tap = {}
    out = g.synthesis(ws, c, jitter_coarse=jit, u_fine=u, tap=tap)
    return cfg, g, ws, c, jit, u, out, tap


def test_num_ws_and_shapes(tiny):
    cfg, g, ws, c, jit, u, out, tap = tiny
    assert E.GeneratorConfig().num_ws == 14            # headnerf.py:55 bases are [K, 14*dim]
    assert out['image'].shape == (2, 3, cfg.img_resolution, cfg.img_resolution)
    assert out['image_raw'].shape == (2, 3, cfg.nrr, cfg.nrr)
    assert out['image_depth'].shape == (2, 1, cfg.nrr, cfg.nrr)
    assert tap['planes'].shape == (2, 96, cfg.plane_res, cfg.plane_res)


def test_state_dict_keys_follow_eg3d():
    g = E.TriPlaneGeneratorRef(E.tiny_config())
    keys = set(g.state_dict().keys())
    for k in ('backbone.synthesis.b4.const', 'backbone.synthesis.b4.resample_filter',
              'backbone.synthesis.b8.conv0.affine.weight', 'backbone.synthesis.b8.conv0.noise_const',
              'backbone.synthesis.b8.conv1.noise_strength', 'backbone.synthesis.b8.torgb.affine.bias',
              'backbone.mapping.fc0.weight', 'backbone.mapping.w_avg', 'superresolution.block0.conv0.weight',
              'superresolution.block1.torgb.bias', 'superresolution.block1.resample_filter',
              'decoder.net.0.weight', 'decoder.net.2.bias'):
        assert k in keys, k
    assert not any('b4.conv0' in k for k in keys)
    full = E.TriPlaneGeneratorRef.__new__(E.TriPlaneGeneratorRef)   # shapes of the real config, no alloc
    cfg = E.GeneratorConfig()
    assert [cfg.channels(r) for r in cfg.block_resolutions] == [512, 512, 512, 512, 512, 256, 128]


def test_fused_equals_nonfused(tiny):
    cfg, g, ws, c, jit, u, out, tap = tiny
    out2 = g.synthesis(ws, c, jitter_coarse=jit, u_fine=u, fused=False)
    assert (out['image'] - out2['image']).abs().max() < 1e-4 * out['image'].abs().max()


def test_modconv_identity_styles_is_plain_conv():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 9, 9, generator=g)
    w = torch.randn(5, 8, 3, 3, generator=g)
    y = E.modulated_conv2d_ref(x, w, torch.ones(2, 8), demodulate=False)
    assert torch.allclose(y, F.conv2d(x, w, padding=1), atol=1e-5)


def test_upconv_equals_zero_insert_conv_fir():
    """up=2 path == zero-insert -> pad 2 -> convolution with flipped kernel -> FIR (built independently)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, 6, 6, generator=g)
    w = torch.randn(3, 4, 3, 3, generator=g)
    f = E.setup_filter()
    y = E.modulated_conv2d_ref(x, w, torch.ones(1, 4), up=2, demodulate=False, f=f)
    z = x.new_zeros(1, 4, 11, 11)
    z[:, :, ::2, ::2] = x                                   # zero-inserted (last zero row/col dropped)
    t = F.conv2d(F.pad(z, [2, 2, 2, 2]), w.flip([2, 3]))     # full true convolution -> 13x13 = 2H+1
    ref = E.upfirdn2d_ref(t, f, pad=(1, 1, 1, 1), gain=4.0)
    assert y.shape == (1, 3, 12, 12)
    assert torch.allclose(y, ref, atol=1e-4)


def test_upsample2d_is_polyphase_quarter_threequarter():
    x = torch.arange(16.0).reshape(1, 1, 4, 4)
    y = E.upsample2d_ref(x, E.setup_filter())
    assert y.shape == (1, 1, 8, 8)
    row = x[0, 0, 1]
    # interior even/odd phases along x at an interior row pair
    assert torch.allclose(y[0, 0, 2, 2], (0.25 * x[0, 0, 0, 0] + 0.75 * x[0, 0, 0, 1]) * 0.25
                          + (0.25 * x[0, 0, 1, 0] + 0.75 * x[0, 0, 1, 1]) * 0.75)
    del row


def test_render_invariants(tiny):
    cfg, g, ws, c, jit, u, out, tap = tiny
    s = cfg.depth_res
    assert float(tap['weight_sum'].max()) <= 1 + 1e-5 and float(tap['weight_sum'].min()) >= 0
    ds = tap['depths_sorted']
    assert bool((ds[:, :, 1:] >= ds[:, :, :-1]).all())
    assert int(tap['inds'].min()) >= 1 and int(tap['inds'].max()) <= s - 2
    assert int(tap['above'].max()) <= s - 3 and int(tap['below'].min()) >= 0
    assert bool((tap['above'] - tap['below'] <= 1).all())
    assert tap['cdf'].shape[1] == s - 2
    assert float(out['image_raw'].abs().max()) <= 1.002 + 1e-5
    # fine depths fall inside the coarse mid-point range
    zc = tap['depths_coarse']
    assert float(tap['depths_fine'].min()) >= float(zc.min()) and float(tap['depths_fine'].max()) <= float(zc.max())
    # sort indices are a permutation of 0..T-1
    si = tap['sort_idx'].squeeze(-1)
    assert bool((si.sort(dim=-1).values == torch.arange(si.shape[-1])).all())


def test_sr_uses_last_ws_only(tiny):
    cfg, g, ws, c, jit, u, out, tap = tiny
    ws2 = ws.clone()
    feat = tap['feature_image']
    a = g.superresolution(feat[:, :3], feat, ws)
    ws2[:, :-1] = 0
    b = g.superresolution(feat[:, :3], feat, ws2)
    assert torch.equal(a, b)


def test_ray_sampler_centre_ray_points_at_origin():
    c = H.flip_label_(H.lookat_label([0.5 * math.pi], [0.5 * math.pi]))
    ro, rd = E.ray_sampler_ref(c[:, :16].view(1, 4, 4), c[:, 16:].view(1, 3, 3), 128)
    assert torch.allclose(ro[0, 0], torch.tensor([0.0, 0.0, 2.7]), atol=1e-5)
    centre = rd[0].mean(0)
    assert centre[2] < -0.99 and abs(float(centre[0])) < 1e-3     # looks down -z after the GL flip
    assert torch.allclose(rd.norm(dim=-1), torch.ones(1, 128 * 128), atol=1e-5)
