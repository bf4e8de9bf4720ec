"""CPU: the PRODUCT camera helpers (hfa_gp_b200/cam_utils.py, trainer_rgb.cam_sampler*) against the reference's
own code/cam_utils.py:12-80 and code/trainer_rgb.py:27-42 — (a) bit-equal against golden vectors frozen from the
reference (tests/golden/reference_cameras.npz, minted by oracle/make_golden.py; travels to the GPU box) and
(b) live against /root/reference when it is on this machine."""
import math
import os

import numpy as np
import pytest
import torch

from hfa_gp_b200 import cam_utils as P
from oracle import ref_bridge

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_cameras.npz'))
needs_ref = pytest.mark.skipif(not ref_bridge.available(), reason='/root/reference not on this machine')
KW = dict(n=5, r=2.7, horizontal_stddev=0.3, vertical_stddev=0.155, horizontal_mean=0.5 * math.pi, vertical_mean=0.5 * math.pi)


@pytest.mark.parametrize('mode', ['gaussian', 'uniform', None])
def test_sample_and_lookat_equal_golden(mode):
    torch.manual_seed(int(GOLD['seed']))
    pts, phi, theta = P.sample_camera_positions('cpu', mode=mode, **KW)
    assert np.array_equal(pts.numpy(), GOLD[f'pts_{mode}'])
    assert np.array_equal(phi.numpy(), GOLD[f'phi_{mode}'])
    assert np.array_equal(theta.numpy(), GOLD[f'theta_{mode}'])
    assert np.array_equal(P.create_cam2world_matrix(-pts, pts, device='cpu').numpy(), GOLD[f'c2w_{mode}'])


def test_trainer_cam_samplers_equal_golden():
    from hfa_gp_b200 import trainer_rgb as T
    seed = int(GOLD['seed'])
    torch.manual_seed(seed + 1)
    assert np.array_equal(T.cam_sampler(4, 'cpu').numpy(), GOLD['cam_sampler'])
    torch.manual_seed(seed + 1)
    assert np.array_equal(P.cam_sampler(4, 'cpu').numpy(), GOLD['cam_sampler'])
    torch.manual_seed(seed + 2)
    assert np.array_equal(T.cam_sampler_pose(4, 0.4, 0.55, 'cpu').numpy(), GOLD['cam_sampler_pose'])


def test_unsupported_mode_raises():
    with pytest.raises(ValueError):
        P.sample_camera_positions('cpu', mode='hybrid')


@needs_ref
@pytest.mark.parametrize('mode', ['gaussian', 'normal', 'uniform', None])
def test_equals_reference_live(mode):
    cam = ref_bridge.load()[2]
    for seed in (0, 3):
        torch.manual_seed(seed)
        a = cam.sample_camera_positions('cpu', mode=mode, **KW)
        torch.manual_seed(seed)
        b = P.sample_camera_positions('cpu', mode=mode, **KW)
        for u, v in zip(a, b):
            assert torch.equal(u, v)
        assert torch.equal(cam.create_cam2world_matrix(-a[0], a[0], device='cpu'),
                           P.create_cam2world_matrix(-b[0], b[0], device='cpu'))
