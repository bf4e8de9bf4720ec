"""Drop-in for ``code/trainer_audio.py`` of the reference (audio-driven reenactment, ``train_audio.py``).
``AudioNet`` / ``AudioAttNet`` are tiny Conv1d stacks and stay in PyTorch; everything after the
64-d audio feature is the same path as trainer_3dmm.  The 8-frame smoothing window needs only neighbouring rows
of ``auds`` (``trainer_audio.py:68-84``), which every rank holds, so frame sharding needs no exchange."""
from __future__ import annotations

import torch

from .networks.headnerf import AudioAttNet, AudioNet, HeadNeRF_Audio
from .optim import DataParallelShard, FlatAdam
from .trainer_rgb import _TrainerBase, requires_grad  # noqa: F401


class Trainer(_TrainerBase):
    optim_key = 'w_optim'
    bases_weight = 5

    def __init__(self, auds, i_train, args, device, rank):
        super().__init__()
        gen = HeadNeRF_Audio(args, args.size, device, args.latent_dim_style, args.latent_dim_shape, args.run_id,
                             args.emb_dir)
        self.w_optim = self._setup(args, device, gen)
        self.AudNet = DataParallelShard(AudioNet(args.dim_aud, args.win_size).to(device))
        self.AudAttNet = DataParallelShard(AudioAttNet().to(device))
        self.optimizer_Aud = FlatAdam(self.AudNet.parameters(), lr=args.lr, betas=(0.9, 0.999))
        self.optimizer_AudAtt = FlatAdam(self.AudAttNet.parameters(), lr=args.lr, betas=(0.9, 0.999))
        self.auds = torch.as_tensor(auds).to(device).float()
        self.i_train = i_train

    def _optims(self):
        return {'w_optim': self.w_optim, 'optimizer_Aud': self.optimizer_Aud, 'optimizer_AudAtt': self.optimizer_AudAtt}

    def resume(self, resume_ckpt):
        ckpt = torch.load(resume_ckpt, map_location=self.device, weights_only=False)
        self.AudNet.module.load_state_dict(ckpt['AudNet'])
        self.AudAttNet.module.load_state_dict(ckpt['AudAttNet'])
        return super().resume(resume_ckpt)

    def save(self, idx, checkpoint_path):
        d = {'gen': self.gen.module.state_dict(), 'AudAttNet': self.AudAttNet.module.state_dict(),
             'AudNet': self.AudNet.module.state_dict(), 'args': self.args}
        d.update({k: o.state_dict() for k, o in self._optims().items()})
        torch.save(d, f'{checkpoint_path}/{str(idx).zfill(6)}.pt')

    def _audio_feature(self, global_step, img_i, limit):
        """trainer_audio.py:66-93 (train) / :118-146 (sample): window the features, zero-pad at the ends."""
        aud = self.auds[img_i]
        if global_step >= self.args.nosmo_iters:
            half = int(self.args.smo_size / 2)
            left_i, right_i = img_i - half, img_i + half
            pad_left = pad_right = 0
            if left_i < 0:
                pad_left, left_i = -left_i, 0
            if right_i > limit:
                pad_right, right_i = right_i - limit, limit
            win = self.auds[left_i:right_i]
            if pad_left > 0:
                win = torch.cat((torch.zeros_like(win)[:pad_left], win), dim=0)
            if pad_right > 0:
                win = torch.cat((win, torch.zeros_like(win)[:pad_right]), dim=0)
            win = self.AudNet(win)
            aud_smo = self.AudAttNet(win)
            return aud_smo.unsqueeze(0) if aud_smo.dim() == 1 else aud_smo
        aud = self.AudNet(aud.squeeze(1))
        return aud.unsqueeze(0) if aud.dim() == 1 else aud

    def gen_update(self, real_image, label, params, global_step, img_i, person_2=False):
        self.gen.train()
        self.AudNet.train()
        self.AudAttNet.train()
        for o in self._optims().values():
            o.zero_grad()
        feat = self._audio_feature(global_step, img_i, self.i_train)
        generated_image = self.gen(feat, label, person_2)
        l2_loss, loss_lpips, generated_image = self._losses(real_image, generated_image)
        l2_loss_3dmm = torch.zeros(1, device=self.device)
        g_loss = l2_loss_3dmm + l2_loss + loss_lpips
        g_loss.backward()
        self.w_optim.step()
        self.optimizer_Aud.step()
        if global_step >= self.args.nosmo_iters:
            self.optimizer_AudAtt.step()
        return l2_loss_3dmm, l2_loss, loss_lpips, generated_image

    def sample(self, real_image, label, params, global_step, img_i, person_2=False):
        with torch.no_grad():
            self.gen.eval()
            self.AudNet.eval()
            self.AudAttNet.eval()
            return self.gen(self._audio_feature(global_step, img_i, self.auds.shape[0]), label, person_2)
