"""Mint the golden fixtures under tests/golden/ (run in the authoring container: python -m oracle.make_golden).

reference_encoder.npz     outputs of the REFERENCE's own Encoder / get_latent / cam utils
                          (/root/reference/code/networks, imported through oracle/ref_bridge.py) on
                          seeded inputs.  Weights are not stored: oracle.hfagp_ref.make_encoder_state(seed)
                          regenerates them bit-identically (CPU torch.Generator).
oracle_generator_tiny.npz outputs of the ORACLE generator (parity unpinned upstream — there is no
                          reference implementation of it on this machine) incl. the integer bookkeeping.
TEST INFRASTRUCTURE ONLY.
"""
import math
import os
import types
import warnings

import numpy as np
import torch

from oracle import eg3d_ref as E
from oracle import hfagp_ref as H
from oracle import ref_bridge

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def reference_encoder(seed=5):
    mods = ref_bridge.load()
    assert mods is not None, 'needs /root/reference'
    enc, head, cam = mods
    warnings.simplefilter('ignore')
    z = dict(seed=np.int64(seed))
    g = torch.Generator().manual_seed(seed + 1)
    for size, pose in ((64, True), (128, False)):
        sd = H.make_encoder_state(size, 512, 50, out_pose=pose, seed=seed)
        e = enc.Encoder(size, 512, 50, False, pose).eval()
        e.load_state_dict(sd, strict=True)
        x = torch.rand(2, 3, size, size, generator=g) * 2 - 1
        with torch.no_grad():
            out = e(x)
        z[f'x{size}'] = x.numpy()
        if pose:
            z[f'w{size}'], z[f'pose{size}'] = out[0].numpy(), out[1].numpy()
        else:
            z[f'w{size}'] = out.numpy()
    gb = torch.Generator().manual_seed(seed)
    bases = torch.randn(50, 14 * 512, generator=gb)
    fake = types.SimpleNamespace(bases=bases, delta=bases.mean(0), dim=512, args=None)
    w = torch.randn(2, 50, generator=g)
    z['latent_w'] = w.numpy()
    z['latent_out'] = head.HeadNeRF_3DMM.get_latent(fake, w).numpy()
    theta = torch.tensor([0.5 * math.pi, 1.2, 1.9])
    phi = torch.tensor([0.5 * math.pi, 1.4, 1.7])
    labels = []
    for t, p in zip(theta, phi):
        pts, _, _ = cam.sample_camera_positions('cpu', n=1, r=2.7, horizontal_mean=float(t), vertical_mean=float(p),
                                                mode=None)
        c = cam.create_cam2world_matrix(-pts, pts, device='cpu').reshape(1, -1)
        labels.append(torch.cat([c, torch.tensor(H.INTRINSICS)[None]], -1))
    z['cam_theta'], z['cam_phi'], z['cam_label'] = theta.numpy(), phi.numpy(), torch.cat(labels).numpy()
    np.savez_compressed(os.path.join(OUT, 'reference_encoder.npz'), **z)


def reference_cameras(seed=11):
    """Seeded outputs of the REFERENCE's cam_utils.py:12-80 in the three sampling modes HFA-GP uses and of the
    trainer_rgb.py:27-42 samplers (restated on the reference's own functions: trainer_rgb.py itself needs lpips)."""
    mods = ref_bridge.load()
    assert mods is not None, 'needs /root/reference'
    cam = mods[2]
    z = dict(seed=np.int64(seed))
    intr = torch.tensor([4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1])
    for mode in ('gaussian', 'uniform', None):
        torch.manual_seed(seed)
        pts, phi, theta = cam.sample_camera_positions('cpu', n=5, r=2.7, horizontal_stddev=0.3, vertical_stddev=0.155,
                                                      horizontal_mean=0.5 * math.pi, vertical_mean=0.5 * math.pi, mode=mode)
        z[f'pts_{mode}'], z[f'phi_{mode}'], z[f'theta_{mode}'] = pts.numpy(), phi.numpy(), theta.numpy()
        z[f'c2w_{mode}'] = cam.create_cam2world_matrix(-pts, pts, device='cpu').numpy()

    def sampler(batch, hm, vm, hs):
        pts, _, _ = cam.sample_camera_positions('cpu', n=batch, r=2.7, horizontal_mean=hm * math.pi, vertical_mean=vm * math.pi,
                                                horizontal_stddev=hs, vertical_stddev=0.155, mode='gaussian')
        c = cam.create_cam2world_matrix(-pts, pts, device='cpu').reshape(batch, -1)
        return torch.cat((c, intr.reshape(1, -1).repeat(batch, 1).to(c)), -1)

    torch.manual_seed(seed + 1)
    z['cam_sampler'] = sampler(4, 0.5, 0.5, 0.3).numpy()                 # trainer_rgb.py:27-33
    torch.manual_seed(seed + 2)
    z['cam_sampler_pose'] = sampler(4, 0.4, 0.55, 0.15).numpy()          # trainer_rgb.py:36-42 (0.4, 0.55)
    np.savez_compressed(os.path.join(OUT, 'reference_cameras.npz'), **z)


def train_step_full(kind):
    """ONE full-size training step of the oracle (configs[2] 'rgb' / configs[3] '3dmm', batch 1, lr 3e-4): losses, probe
    pixels of the pooled image, gradients of delta / bases (every 16th column) / the first and last head weights and the
    updated delta.  ~1 min of CPU; the GPU test re-creates the inputs from oracle.train_ref.full_step_case."""
    from oracle import train_ref
    torch.set_num_threads(8)
    c = train_ref.full_step_case(kind)
    oracle = train_ref.TrainStepRef(c['sd'], c['bases'], c['delta'], c['generator'], c['size'], 3e-4, lpips=c['lpips'],
                                    head=c['head'])
    l2, lp, img = oracle.step(c['real'], c['label'], c['jitter'], c['u'], params=c['params'])
    first, last = ('net_app.convs.0.0.weight', 'fc.4.weight') if kind == 'rgb' else ('fc.0.weight', 'fc.6.weight')
    z = dict(l2=l2.numpy(), lpips=lp.numpy(), image_probe=img[:, :, ::8, ::8].numpy(), image_mean=img.mean().numpy(),
             image_abs_mean=img.abs().mean().numpy(), d_delta=oracle.delta.grad.numpy(),
             d_bases_sub=oracle.bases.grad[:, ::16].contiguous().numpy(), d_first=oracle.sd[first].grad.numpy(),
             d_last=oracle.sd[last].grad.numpy(), delta_new=oracle.delta.detach().numpy())
    np.savez_compressed(os.path.join(OUT, f'train_step_full_{kind}.npz'), **z)
    print(kind, 'l2', float(l2), 'lpips', float(lp))


def oracle_generator_tiny(seed=0):
    cfg = E.tiny_config()
    gen = E.make_generator(cfg, seed=seed, noise_strength=0.1)
    g = torch.Generator().manual_seed(seed + 3)
    ws = torch.randn(1, cfg.num_ws, cfg.w_dim, generator=g)
    c = H.flip_label_(H.synthetic_labels(1, seed=seed))
    jit = torch.rand(1, cfg.nrr ** 2, cfg.depth_res, 1, generator=g)
    u = torch.rand(cfg.nrr ** 2, cfg.depth_res_importance, generator=g)
    tap = {}
    out = gen.synthesis(ws, c, jitter_coarse=jit, u_fine=u, tap=tap)
    np.savez_compressed(os.path.join(OUT, 'oracle_generator_tiny.npz'), seed=np.int64(seed), ws=ws.numpy(),
                        c=c.numpy(), jitter=jit.numpy(), u=u.numpy(), image=out['image'].numpy(),
                        image_raw=out['image_raw'].numpy(), inds=tap['inds'].numpy().astype(np.int32),
                        below=tap['below'].numpy().astype(np.int32), above=tap['above'].numpy().astype(np.int32),
                        sort_idx=tap['sort_idx'].squeeze(-1).numpy().astype(np.int32),
                        depths_sorted=tap['depths_sorted'].squeeze(-1).numpy())


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    reference_encoder()
    reference_cameras()
    oracle_generator_tiny()
    train_step_full('rgb')
    train_step_full('3dmm')
    print(sorted(os.listdir(OUT)), [os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))])
