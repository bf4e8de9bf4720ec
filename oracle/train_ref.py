"""CPU fp32 ORACLE of one training step of the reference (``Trainer.gen_update``,
``/root/reference/code/trainer_rgb.py:73-98``; 3DMM variant ``trainer_3dmm.py:43-67``).

TEST INFRASTRUCTURE ONLY (see ``oracle/eg3d_ref.py`` header for who may import this).

Restates the step with plain PyTorch autograd and ``torch.optim.Adam`` — what the reference executes:

    g_optim.zero_grad()                                                  :75
    weights = encoder(real) ; latent = get_latent(weights)              :77-80  (oracle/hfagp_ref.py, reference-pinned)
    image   = generator.synthesis(latent, c=flip(label))['image']       :81     (oracle/eg3d_ref.py, parity unpinned)
    image   = AdaptiveAvgPool2d(size)(image)                            :63,84
    l2      = MSELoss(mean)(real, image)                                :15,85
    lpips   = squeeze(LPIPS(real, image)).mean()                        :86-87  (module supplied by the caller: the pip
                                                                                 package and its weights are absent offline)
    (l2 + lpips).backward() ; g_optim.step()                            :91-95

PINNING: the encoder / latent pieces are pinned against the reference's own code (tests/test_oracle_reference.py);
MSELoss, AdaptiveAvgPool2d, autograd and Adam are PyTorch's own implementations, i.e. the very code the reference
runs; the generator part inherits "parity unpinned" from oracle/eg3d_ref.py.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

from . import hfagp_ref


class TrainStepRef:
    """Trainable tensors: every encoder (or Weights_3DMM) weight/bias in ``sd``, ``bases`` and ``delta``; the
    generator is frozen (trainer_rgb.py:59-60) unless ``tune`` (after ``tune_generator()``, :69-71)."""

    def __init__(self, sd: Dict[str, torch.Tensor], bases, delta, generator, size: int, lr: float,
                 lpips: Optional[Callable] = None, head: str = 'encoder', dim: int = 512, tune: bool = False):
        self.sd = {k: v.clone() for k, v in sd.items()}
        self.names = [k for k in self.sd if not k.endswith('.kernel')]
        for k in self.names:
            self.sd[k].requires_grad_(True)
        self.bases = bases.clone().requires_grad_(True)
        self.delta = delta.clone().requires_grad_(True)
        # tune=True: the state after Trainer.tune_generator() (trainer_rgb.py:69-71) — every generator parameter
        # requires grad and is already in the optimiser (it was built over gen.parameters(), :57)
        self.generator = generator.requires_grad_(tune)
        self.size, self.lpips, self.head, self.dim = size, lpips, head, dim
        self.opt = torch.optim.Adam([self.sd[k] for k in self.names] + [self.bases, self.delta]
                                    + list(self.generator.parameters()), lr=lr)

    def forward(self, real, label, jitter_coarse=None, u_fine=None, params=None):
        if self.head == 'encoder':
            weights = hfagp_ref.encoder_ref(self.sd, real)
        else:
            weights = hfagp_ref.weights_3dmm_ref(self.sd, params)
        latent = hfagp_ref.get_latent_ref(self.bases, self.delta, weights, dim=self.dim)
        c = hfagp_ref.flip_label_(label.clone())
        image = self.generator.synthesis(latent, c, jitter_coarse=jitter_coarse, u_fine=u_fine)['image']
        return F.adaptive_avg_pool2d(image, (self.size, self.size))

    def step(self, real, label, jitter_coarse=None, u_fine=None, params=None):
        self.opt.zero_grad()
        image = self.forward(real, label, jitter_coarse, u_fine, params)
        l2 = F.mse_loss(real, image, reduction='mean')
        lp = torch.squeeze(self.lpips(real, image)).mean() if self.lpips is not None else image.new_zeros(())
        (l2 + lp).backward()
        self.opt.step()
        return l2.detach(), lp.detach(), image.detach()


def full_step_case(kind: str = 'rgb', seed: int = 0, batch: int = 1):
    """Everything one FULL-SIZE training step needs (BASELINE.json configs[2] 'rgb': encoder 256 -> 512x512 render pooled
    to 256; configs[3] '3dmm': Weights_3DMM on [B,76] coefficients, one frame per rank), built from CPU generators only so
    that the GPU box re-creates the identical tensors without running the oracle: head weights, bases / delta, the EG3D
    generator (seeded random init, noise strength 0.1), LPIPS-alex weights (seeded), inputs and the renderer's two draws."""
    from . import eg3d_ref, lpips_ref
    cfg = eg3d_ref.GeneratorConfig()
    g = torch.Generator().manual_seed(1000 + seed)
    k, size = 50, 256
    if kind == 'rgb':
        sd = hfagp_ref.make_encoder_state(size=size, dim_motion=k, seed=seed + 4)
        for key in sd:
            if key.endswith('.bias'):
                sd[key] = torch.randn(sd[key].shape, generator=g) * 0.2
        head = 'encoder'
    else:
        dims = [76] + [512] * 6 + [k]
        sd = {}
        for j in range(7):
            sd[f'fc.{j}.weight'] = torch.randn(dims[j + 1], dims[j], generator=g)
            sd[f'fc.{j}.bias'] = torch.randn(dims[j + 1], generator=g) * 0.2
        head = 'weights_3dmm'
    bases = torch.randn(k, 14 * 512, generator=g)
    delta = bases.mean(dim=0)
    gen = eg3d_ref.make_generator(cfg, seed=seed, noise_strength=0.1)
    lp = lpips_ref.LPIPS(net='alex', seed=seed + 3).eval()
    rays = cfg.nrr ** 2
    real = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    label = hfagp_ref.synthetic_labels(batch, seed=seed + 7)
    params = torch.randn(batch, 76, generator=g)
    jitter = torch.rand(batch, rays, cfg.depth_res, 1, generator=g)
    u = torch.rand(batch * rays, cfg.depth_res_importance, generator=g)
    return dict(cfg=cfg, size=size, k=k, head=head, sd=sd, bases=bases, delta=delta, generator=gen, lpips=lp, real=real,
                label=label, params=params, jitter=jitter, u=u)
