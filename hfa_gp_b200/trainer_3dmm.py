"""Drop-in for ``code/trainer_3dmm.py`` of the reference (3DMM-coefficient driving; ``train_3dmm.py:117-160``).
Same step as trainer_rgb with ``Weights_3DMM`` (7 linear EqualLinear, ``headnerf.py:138-158``) in place of the
image encoder; gradients are averaged over ranks every step (the reference relies on DDP for this trainer)."""
from __future__ import annotations

import torch

from .networks.headnerf import HeadNeRF_3DMM
from .trainer_rgb import _TrainerBase, requires_grad  # noqa: F401


class Trainer(_TrainerBase):
    optim_key = 'w_optim'
    bases_weight = 5                                   # trainer_3dmm.py:89

    def __init__(self, args, device, rank):
        super().__init__()
        gen = HeadNeRF_3DMM(args, args.size, device, args.latent_dim_style, args.latent_dim_shape, args.run_id,
                            args.emb_dir)
        self.w_optim = self._setup(args, device, gen)

    def gen_update(self, real_image, label, params, person_2=False):
        self.gen.train()

        def fwd_bwd(real_image, label, params):
            generated_image = self.gen(params, label, person_2)
            l2_loss, loss_lpips, generated_image = self._losses(real_image, generated_image)
            l2_loss_3dmm = torch.zeros(1, device=self.device)
            g_loss = l2_loss_3dmm + l2_loss + loss_lpips
            g_loss.backward()
            return l2_loss_3dmm, l2_loss, loss_lpips, generated_image

        return self._run_step(fwd_bwd, [real_image, label, params], 1, variant=(bool(person_2),))

    def sample(self, real_image, label, params, person_2=False):
        with torch.no_grad():
            self.gen.eval()
            return self.gen(params, label, person_2)
