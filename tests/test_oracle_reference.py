"""CPU: pin the oracle's restatement of the reference-owned parts against (a) the reference itself when
/root/reference is on this machine and (b) golden vectors frozen from it (tests/golden/*.npz)."""
import math
import os
import types
import warnings

import numpy as np
import pytest
import torch

from oracle import hfagp_ref as H
from oracle import ref_bridge

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
needs_ref = pytest.mark.skipif(not ref_bridge.available(), reason='/root/reference not on this machine')


@needs_ref
@pytest.mark.parametrize('size,pose', [(64, True), (128, False)])
def test_encoder_restatement_equals_reference(size, pose):
    enc, head, cam = ref_bridge.load()
    torch.manual_seed(0)
    e = enc.Encoder(size, 512, 50, False, pose).eval()
    with torch.no_grad():
        for n, p in e.named_parameters():
            if 'bias' in n:
                p.normal_(0, 0.3)
    sd = {k: v.detach() for k, v in e.state_dict().items()}
    x = torch.rand(2, 3, size, size) * 2 - 1
    with torch.no_grad():
        a = e(x)
    b = H.encoder_ref(sd, x, out_pose=pose)
    a, b = (a, b) if pose else ((a,), (b,))
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    # synthetic state builder matches the reference's key set and shapes
    sd2 = H.make_encoder_state(size, out_pose=pose)
    assert set(sd2) == set(sd) and all(sd2[k].shape == sd[k].shape for k in sd)


@needs_ref
def test_latent_heads_and_cameras_equal_reference():
    enc, head, cam = ref_bridge.load()
    warnings.simplefilter('ignore')
    torch.manual_seed(1)
    self = types.SimpleNamespace(bases=torch.randn(50, 14 * 512), delta=torch.randn(14 * 512), dim=512, args=None)
    w = torch.randn(3, 50)
    assert torch.equal(head.HeadNeRF_3DMM.get_latent(self, w), H.get_latent_ref(self.bases, self.delta, w))
    m = head.Weights_3DMM(76, 512, 50).eval()
    p = torch.randn(4, 76)
    assert torch.equal(m(p), H.weights_3dmm_ref(m.state_dict(), p))
    a1 = head.AudioNet(64, 16).eval()
    xa = torch.randn(8, 16, 29)
    assert torch.equal(a1(xa), H.audionet_ref(a1.state_dict(), xa))
    a2 = head.AudioAttNet().eval()
    assert torch.equal(a2(a1(xa)), H.audioattnet_ref(a2.state_dict(), a1(xa)))
    pts, _, _ = cam.sample_camera_positions('cpu', n=1, r=2.7, horizontal_mean=0.5 * math.pi,
                                            vertical_mean=0.5 * math.pi, mode=None)
    c = cam.create_cam2world_matrix(-pts, pts, device='cpu').reshape(1, -1)
    assert torch.equal(c, H.lookat_label([0.5 * math.pi], [0.5 * math.pi])[:, :16])


def test_golden_encoder_vectors():
    """Outputs of the REFERENCE's Encoder / get_latent / heads frozen by oracle/make_golden.py."""
    z = np.load(os.path.join(GOLD, 'reference_encoder.npz'))
    for size, pose in ((64, True), (128, False)):
        sd = H.make_encoder_state(size, 512, 50, out_pose=pose, seed=int(z['seed']))
        x = torch.from_numpy(z[f'x{size}'])
        out = H.encoder_ref(sd, x, out_pose=pose)
        w = out[0] if pose else out
        ref = torch.from_numpy(z[f'w{size}'])
        assert (w - ref).abs().max() <= 1e-4 * ref.abs().max()
        if pose:
            ref = torch.from_numpy(z[f'pose{size}'])
            assert (out[1] - ref).abs().max() <= 1e-4 * ref.abs().max()
    g = torch.Generator().manual_seed(int(z['seed']))
    bases = torch.randn(50, 14 * 512, generator=g)
    delta = bases.mean(0)
    wts = torch.from_numpy(z['latent_w'])
    lat = H.get_latent_ref(bases, delta, wts)
    ref = torch.from_numpy(z['latent_out'])
    assert (lat - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert torch.allclose(H.lookat_label(z['cam_theta'], z['cam_phi']), torch.from_numpy(z['cam_label']), atol=1e-6)


def test_golden_generator_vectors():
    """Oracle-minted vectors (parity UNPINNED upstream): guards the oracle against silent edits."""
    from oracle import eg3d_ref as E
    z = np.load(os.path.join(GOLD, 'oracle_generator_tiny.npz'))
    cfg = E.tiny_config()
    g = E.make_generator(cfg, seed=int(z['seed']), noise_strength=0.1)
    tap = {}
    out = g.synthesis(torch.from_numpy(z['ws']), torch.from_numpy(z['c']), jitter_coarse=torch.from_numpy(z['jitter']),
                      u_fine=torch.from_numpy(z['u']), tap=tap)
    img = torch.from_numpy(z['image'])
    assert (out['image'] - img).abs().max() <= 2e-4 * img.abs().max()
    cdf = tap['cdf']
    near = (torch.from_numpy(z['u'])[:, :, None] - cdf[:, None, :]).abs().min(-1).values < 1e-6
    for k in ('inds', 'below', 'above'):
        assert bool(((tap[k] == torch.from_numpy(z[k]).long()) | near).all()), k
    ds_ref = torch.from_numpy(z['depths_sorted'])
    assert (tap['depths_sorted'].squeeze(-1) - ds_ref).abs().max() < 1e-5


@pytest.mark.parametrize('k,m', [(50, 14 * 512), (20, 300), (7, 7), (1, 33)])
def test_qr_restatement_equals_torch_qr(k, m):
    """oracle/qr_ref.py (CholeskyQR2 + Householder sign reconstruction — the algorithm of csrc/qr.cu) against the
    call the reference makes, torch.qr / torch.linalg.qr on (bases + 1e-8).T (headnerf.py:92): same Q including the
    column signs, same R, same gradient through Q."""
    from oracle import qr_ref
    g = torch.Generator().manual_seed(100 * k + m)
    bases = torch.randn(k, m, generator=g, dtype=torch.float64)
    a = (bases + 1e-8).T
    q_ref, r_ref = torch.linalg.qr(a, mode='reduced')
    q, r = qr_ref.cholqr2_signed(a)
    assert float((q - q_ref).abs().max()) < 1e-12 and float((r - r_ref).abs().max()) < 1e-10 * float(r_ref.abs().max())
    assert torch.equal(torch.sign(torch.diagonal(r)), torch.sign(torch.diagonal(r_ref)))
    # fp32 input, as the product sees it: LAPACK's fp32 factor to its own rounding level
    q32, _ = qr_ref.cholqr2_signed(a.float())
    q32_ref, _ = torch.linalg.qr(a.float(), mode='reduced')
    assert float((q32 - q32_ref).abs().max()) < 5e-6
    # gradient arriving at Q only
    gq = torch.randn(m, k, generator=g, dtype=torch.float64)
    leaf = a.clone().requires_grad_(True)
    qq, _ = torch.linalg.qr(leaf, mode='reduced')
    (qq * gq).sum().backward()
    ga = qr_ref.qr_backward_q_only(gq, q, r)
    assert float((ga - leaf.grad).abs().max()) < 1e-9 * max(1.0, float(leaf.grad.abs().max()))
