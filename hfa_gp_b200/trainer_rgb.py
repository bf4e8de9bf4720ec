"""Drop-in for ``code/trainer_rgb.py`` of the reference: same ``Trainer`` surface (``gen_update``, ``sample``,
``sample_bases``, ``tune_generator``, ``resume``, ``save``) and the module-level ``cam_sampler`` helpers, so
``train_rgb.py`` drives it unchanged.

One training step (``/root/reference/code/trainer_rgb.py:73-98``)::

    weights = gen.get_weights(real)           encoder convolutions + EqualLinear head   (tcgen05 / SIMT kernels)
    latent  = gen.get_latent(weights)         thin QR (hfagp_basis_qr_fwd) + hfagp_latent_fwd
    image   = gen.get_image(latent, label)    backbone -> renderer -> super-resolution   (frozen generator)
    image   = face_pool(image)                hfagp_facepool_fwd  (AdaptiveAvgPool2d(size), :63,84)
    loss    = MSE(real, image) + LPIPS        hfagp_mse_fwd + hfa_gp_b200.lpips
    loss.backward()                           hfa_gp_b200/autograd.py: every stage walks its tape with the C-ABI kernels
    g_optim.step()                            one flat gradient all-reduce + one hfagp_adam_step launch

Differences kept on purpose: gradients are averaged over ranks every step (the reference's RGB trainer bypasses its
DDP wrapper and lets replicas drift); ``tune_generator()`` flips ``requires_grad`` as upstream but
every generator parameter then receives its gradient (hfa_gp_b200/autograd.py) and joins the flat Adam update.
"""
from __future__ import annotations

import math
import os

import torch
from torch import nn

from . import ops
from .autograd import FacePoolFn, MseFn
from .cam_utils import INTRINSICS, create_cam2world_matrix, sample_camera_positions
from .lpips import LPIPS
from .networks.headnerf import HeadNeRF_final
from .optim import DataParallelShard, FlatAdam


def requires_grad(net, flag=True):
    for p in net.parameters():
        p.requires_grad = flag


def _with_intrinsics(c, batch):
    return torch.cat((c, torch.tensor(INTRINSICS).reshape(1, -1).repeat(batch, 1).to(c)), -1)


def cam_sampler(batch, device):
    """trainer_rgb.py:27-33"""
    pts, _, _ = sample_camera_positions(device, n=batch, r=2.7, horizontal_mean=0.5 * math.pi,
                                        vertical_mean=0.5 * math.pi, horizontal_stddev=0.3, vertical_stddev=0.155,
                                        mode='gaussian')
    return _with_intrinsics(create_cam2world_matrix(-pts, pts, device=device).reshape(batch, -1), batch)


def cam_sampler_pose(batch, horizontal_mean, vertical_mean, device):
    """trainer_rgb.py:36-42"""
    pts, _, _ = sample_camera_positions(device, n=batch, r=2.7, horizontal_mean=horizontal_mean * math.pi,
                                        vertical_mean=vertical_mean * math.pi, horizontal_stddev=0.15,
                                        vertical_stddev=0.155, mode='gaussian')
    return _with_intrinsics(create_cam2world_matrix(-pts, pts, device=device).reshape(batch, -1), batch)


class FacePool(nn.Module):
    """``torch.nn.AdaptiveAvgPool2d((size, size))`` on the generator's output (trainer_rgb.py:63)."""

    def __init__(self, size):
        super().__init__()
        self.size = size

    def forward(self, image):
        nhwc = image.permute(0, 2, 3, 1)            # synthesis() hands out an NCHW view of its channels-last image
        if not nhwc.is_contiguous():
            nhwc = ops.nchw_to_nhwc(image.contiguous()) if not image.requires_grad else nhwc.contiguous()
        return FacePoolFn.apply(nhwc, self.size)


class _TrainerBase(nn.Module):
    """What the three reference trainers share (trainer_rgb.py / trainer_3dmm.py / trainer_audio.py)."""

    optim_key = 'g_optim'
    bases_weight = 10

    def _setup(self, args, device, gen):
        self.args = args
        self.batch_size = args.batch_size
        self.device = device
        self.gen = DataParallelShard(gen.to(device))
        gen_ids = {id(p) for p in gen.generator.parameters()}
        # Adam over every parameter of gen, as the reference builds it BEFORE freezing the generator (:57-60);
        # the generator's share joins the flat update only after tune_generator()
        optim = FlatAdam(self.gen.parameters(), lr=args.lr, live_first=lambda p: id(p) not in gen_ids)
        for p in gen.generator.parameters():
            p.requires_grad = False
        # pretrained LPIPS-alex weights: args.lpips_weights / HFAGP_LPIPS_WEIGHTS; the seeded random init only when the run
        # is declared synthetic (args.synthetic_lpips, args.synthetic_generator or HFAGP_SYNTHETIC_LPIPS=1) — else it raises
        synthetic = getattr(args, 'synthetic_lpips', None)
        if synthetic is None and getattr(args, 'synthetic_generator', False):
            synthetic = True
        self.lpips_loss = LPIPS(net='alex', weights=getattr(args, 'lpips_weights', None), synthetic=synthetic).to(device).eval()
        self.face_pool = FacePool(args.size)
        return optim

    def l2_loss(self, real_images, generated_images):
        return MseFn.apply(real_images, generated_images)

    # ------------------------------------------------------------------ the step as ONE CUDA graph (opt-in)
    def enable_step_graph(self, warmup: int = 3):
        """Replay ``gen_update`` from a CUDA graph instead of driving its ~900 launches from Python (the eager step is
        host-bound, above all with 8 ranks sharing one host).  After ``warmup`` eager steps with unchanged input shapes the
        whole step — gradient memset, encoder / head -> QR -> latent -> generator -> face_pool -> MSE + LPIPS, backward,
        the flat gradient all-reduce (NCCL, captured) and the Adam kernels — is captured once; later calls copy their
        inputs into static buffers, refresh Adam's two step-dependent scalars on the device and replay.  Results are
        those of the eager step (same kernels, same order); the returned losses / image are static buffers overwritten by
        the next call.  While ``tune_generator()`` has the generator's own weights live the step stays eager (their
        re-packing reads scalars back to the host)."""
        self._graph_warmup = warmup
        self._graph = None
        self._graph_key = None
        self._graph_eager_left = warmup

    def _run_step(self, fwd_bwd, tensors, label_pos, variant=()):
        """zero_grad -> fwd_bwd(*tensors) -> step of every optimiser, eagerly or from the captured graph.
        ``tensors[label_pos]`` is the label the forward flips in place (headnerf.py:108)."""
        optims = list(self._optims().values())
        if getattr(self, '_graph_warmup', None) is not None:
            key = tuple((tuple(t.shape), t.dtype) for t in tensors) + tuple(variant)
            if any(o._late_live() for o in optims):
                key = None                              # generator weights are training: eager
            if key is not None and key == self._graph_key and self._graph is not None:
                return self._replay(tensors, label_pos, optims)
            if key is not None and key == self._graph_key and self._graph_eager_left <= 0:
                self._capture(fwd_bwd, tensors, optims)
                return self._replay(tensors, label_pos, optims)
            if key != self._graph_key:
                self._graph, self._graph_key, self._graph_eager_left = None, key, self._graph_warmup
                self.__dict__['_zero_arena'] = ops.ZeroArena(self.device)      # (a captured graph keeps its own)
            self._graph_eager_left -= 1
            arena = self.__dict__.setdefault('_zero_arena', ops.ZeroArena(self.device))
            arena.begin()                               # no buffer yet: counts the step's small zeroed scratch
            try:
                return self._eager_step(fwd_bwd, tensors, optims)
            finally:
                arena.end()
        return self._eager_step(fwd_bwd, tensors, optims)

    @staticmethod
    def _eager_step(fwd_bwd, tensors, optims):
        for o in optims:
            o.zero_grad()
        out = fwd_bwd(*tensors)
        for o in optims:
            o.step()
        return out

    def _capture(self, fwd_bwd, tensors, optims):
        self._static_in = [t.detach().clone() for t in tensors]
        ops.param_epoch[0] += 1          # every parameter-derived cache misses inside the capture: packing is captured too
        arena = self.__dict__.setdefault('_zero_arena', ops.ZeroArena(self.device))
        arena.materialise()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(g):
            arena.begin()                               # ONE memset for the step's ~60 small zeroed tensors
            try:
                for o in optims:
                    o.zero_grad()
                self._static_out = fwd_bwd(*self._static_in)
                for o in optims:
                    o.step_launch()
            finally:
                arena.end()
        self.graph_launches = ops.launch_count() - n0   # kernels of libhfagp_sm100.so inside one replay
        self._graph = g

    def _replay(self, tensors, label_pos, optims):
        for st, t in zip(self._static_in, tensors):
            st.copy_(t, non_blocking=True)
        for o in optims:
            o.step_prepare()
        self._graph.replay()
        if label_pos is not None and tensors[label_pos].is_cuda:
            tensors[label_pos].copy_(self._static_in[label_pos], non_blocking=True)    # the in-place GL flip stays visible
        return self._static_out

    def tune_generator(self):
        for p in self.gen.module.generator.parameters():
            p.requires_grad = True

    def _losses(self, real_image, generated_image):
        generated_image = self.face_pool(generated_image)
        l2_loss = self.l2_loss(real_image, generated_image)
        loss_lpips = torch.squeeze(self.lpips_loss(real_image, generated_image)).mean()
        return l2_loss, loss_lpips, generated_image

    def _neutral_label(self):
        pts, _, _ = sample_camera_positions(device=self.device, n=1, r=2.7, horizontal_mean=0.5 * math.pi,
                                            vertical_mean=0.5 * math.pi, mode=None)
        return _with_intrinsics(create_cam2world_matrix(-pts, pts, device=self.device).reshape(1, -1), 1)

    def sample_bases(self, person_2=False):
        """One render per basis direction (trainer_rgb.py:108-127).  NB the reference passes the SAME label tensor
        to every get_image call, which flips it in place each time; kept."""
        imgs = []
        with torch.no_grad():
            label = self._neutral_label()
            self.gen.eval()
            for base_id in range(self.args.latent_dim_shape):
                weights = torch.zeros(self.args.latent_dim_shape, device=self.device)
                weights[base_id] = self.bases_weight
                latent = self.gen.module.get_latent(weights.unsqueeze(0), person_2)
                imgs.append(self.gen.module.get_image(latent, label))
        return imgs

    def _optims(self):
        return {self.optim_key: getattr(self, self.optim_key)}

    def resume(self, resume_ckpt):
        print('load model:', resume_ckpt)
        ckpt = torch.load(resume_ckpt, map_location=self.device, weights_only=False)
        start_iter = int(os.path.splitext(os.path.basename(resume_ckpt))[0])
        self.gen.module.load_state_dict(ckpt['gen'])
        for k, o in self._optims().items():
            o.load_state_dict(ckpt[k])
        ops.param_epoch[0] += 1
        return start_iter

    def save(self, idx, checkpoint_path):
        d = {'gen': self.gen.module.state_dict(), 'args': self.args}
        d.update({k: o.state_dict() for k, o in self._optims().items()})
        torch.save(d, f'{checkpoint_path}/{str(idx).zfill(6)}.pt')


class Trainer(_TrainerBase):
    def __init__(self, args, device, rank):
        super().__init__()
        gen = HeadNeRF_final(args, args.size, device, args.latent_dim_style, args.latent_dim_shape, args.run_id,
                             args.emb_dir)
        self.g_optim = self._setup(args, device, gen)

    def gen_update(self, real_image, label, person_2=False, mask=None):
        self.gen.train()

        def fwd_bwd(real_image, label):
            m = self.gen.module
            m.prefetch_basis(m.bases if not person_2 or m.args.same_bases else m.bases_2)   # QR next to the encoder forward
            weights_i = self.gen.module.get_weights(real_image)
            if isinstance(weights_i, tuple):
                weights_i = weights_i[0]
            latent_i = self.gen.module.get_latent(weights_i, person_2)
            generated_image = self.gen.module.get_image(latent_i, label)
            l2_loss, loss_lpips, generated_image = self._losses(real_image, generated_image)
            g_loss = l2_loss + loss_lpips
            g_loss.backward()
            return l2_loss, loss_lpips, generated_image

        return self._run_step(fwd_bwd, [real_image, label], 1, variant=(bool(person_2),))

    def sample(self, real_image, label, person_2=False):
        with torch.no_grad():
            self.gen.eval()
            return self.gen(real_image, label, person_2)
