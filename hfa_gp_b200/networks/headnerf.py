"""Drop-in for ``code/networks/headnerf.py`` of the reference (same classes, ctor arguments, methods,
attributes and ``state_dict`` keys), running on the sm_100a library instead of torch/eg3d ops.

Mirrored behaviour (file:line in /root/reference/code/networks/headnerf.py):
  load_G_official(args, device, path) -> frozen generator ............... :31-38
  HeadNeRF_final: encoder, bases [K,14*dim], delta, generator .......... :44-73
  get_latent: thin QR of (bases+1e-8).T, sum_j w_j Q[:,j] + delta ....... :81-102
  forward / get_image flip label[:, [1,2,5,6,9,10]] IN PLACE, then
    generator.synthesis(latent, c=label, noise_mode='const')['image'] ... :106-134
  Weights_3DMM (7 linear EqualLinear), HeadNeRF_3DMM, HeadNeRF_Audio ... :138-279
  AudioAttNet / AudioNet (tiny Conv1d stacks; left in PyTorch) .......... :284-349

What differs on purpose: the QR factor is cached while ``bases`` is unchanged and gradients are off
(the reference recomputes ``torch.qr`` every call); ``load_G_official`` understands plain
``state_dict`` checkpoints and, when asked for explicitly, builds a seeded random-init generator
because the EG3D pickle cannot be present offline.
"""
from __future__ import annotations

import copy
import os

import torch
from torch import nn

from .. import ops
from .._cabi import HfagpError
from ..generator import GeneratorConfig, TriPlaneGenerator, make_generator
from .encoder3d import Encoder, EqualLinear

FLIP_IDX = [1, 2, 5, 6, 9, 10]
_flip_sign_cache = {}


def flip_label_(label):
    """``label[:, [1,2,5,6,9,10]] *= -1`` in place (headnerf.py:108,132), as one multiply by a cached per-device
    sign vector: the fancy-index form builds its index tensor on the host every call, which costs an H2D copy
    per frame and cannot be captured into a CUDA graph."""
    key = (label.device, label.dtype)
    sign = _flip_sign_cache.get(key)
    if sign is None:
        sign = torch.ones(25, dtype=label.dtype)
        sign[FLIP_IDX] = -1
        sign = _flip_sign_cache[key] = sign.to(label.device)
    return label.mul_(sign)


def toogle_grad(model, flag=True):
    for p in model.parameters():
        p.requires_grad = flag


def load_G_official(args, device, eg3d_ffhq='./code/pretrained_models/eg3d/ffhqrebalanced512-128.pkl'):
    """Returns the frozen generator object HeadNeRF_* store as ``self.generator``.

    * a file at ``eg3d_ffhq`` holding a ``state_dict`` (``torch.save``) of EG3D names -> loaded;
    * ``args.synthetic_generator`` truthy or ``HFAGP_SYNTHETIC_GENERATOR=1`` -> seeded random init
      (seed ``args.generator_seed`` / env ``HFAGP_GENERATOR_SEED``, default 0) — the synthetic-benchmark case;
    * otherwise FileNotFoundError, as the reference's ``dnnlib.util.open_url`` would raise.
    """
    cfg = getattr(args, 'generator_config', None) or GeneratorConfig()
    if os.path.isfile(eg3d_ffhq):
        sd = torch.load(eg3d_ffhq, map_location='cpu', weights_only=True)
        sd = sd.get('G_ema', sd) if isinstance(sd, dict) else sd
        g = TriPlaneGenerator(cfg)
        g.load_state_dict(sd)
    elif getattr(args, 'synthetic_generator', False) or os.environ.get('HFAGP_SYNTHETIC_GENERATOR') == '1':
        seed = int(getattr(args, 'generator_seed', os.environ.get('HFAGP_GENERATOR_SEED', 0)))
        g = make_generator(cfg, seed=seed, device='cpu')
    else:
        raise FileNotFoundError(eg3d_ffhq)
    g = copy.deepcopy(g.requires_grad_(False).to(device)).requires_grad_(False).to(device)
    return g


class _LatentSubspace:
    """get_latent shared by the three avatar classes (headnerf.py:81-102, :182-195, :242-255)."""

    def _q_factor(self, bases: torch.Tensor) -> torch.Tensor:
        key = (bases._version, bases.data_ptr(), bases.device, ops.param_epoch[0])
        cache = self.__dict__.get('_q_cache')
        if cache is not None and cache[0] == key and not torch.is_grad_enabled():
            return cache[1]
        q, _ = ops.basis_qr(bases.detach(), eps=1e-8, check_info=not torch.cuda.is_current_stream_capturing())        # [14*dim, K], LAPACK's Q without LAPACK
        self.__dict__['_q_cache'] = (key, q)
        return q

    def prefetch_basis(self, bases):
        """Training: start the QR factorisation of ``bases`` (independent of the frame) on a side stream so that it
        runs next to the encoder forward; ``_latent_from`` picks the result up.  Autograd runs the factorisation's
        backward on the same side stream and joins it at the end of ``backward()``."""
        if not (torch.is_grad_enabled() and bases.requires_grad and bases.is_cuda):
            return
        cur = torch.cuda.current_stream(bases.device)
        side = self.__dict__.get('_qr_stream')
        if side is None or side.device != bases.device:
            side = self.__dict__['_qr_stream'] = torch.cuda.Stream(device=bases.device)
        side.wait_stream(cur)
        from ..autograd import BasisQRFn
        with torch.cuda.stream(side):
            q = BasisQRFn.apply(bases)
        self.__dict__['_q_prefetched'] = (bases, q, side)

    def _latent_from(self, weights, bases, delta):
        if weights is None:
            return self._q_factor(bases)
        b = weights.shape[0]
        if torch.is_grad_enabled() and (weights.requires_grad or bases.requires_grad or delta.requires_grad):
            # training (trainer_rgb.py:79-80): the QR factorisation is recomputed under autograd, as the
            # reference does every call (headnerf.py:92), so d(Q) reaches ``bases`` through the factorisation's backward
            from ..autograd import BasisQRFn, LatentFn
            pre = self.__dict__.pop('_q_prefetched', None)
            if pre is not None and pre[0] is bases:
                q, side = pre[1], pre[2]
                cur = torch.cuda.current_stream(bases.device)
                cur.wait_stream(side)
                q.record_stream(cur)
            else:
                q = BasisQRFn.apply(bases)
            return LatentFn.apply(weights, q, delta).view(b, -1, self.dim)
        q = self._q_factor(bases)
        out = ops.latent(weights.detach().float().contiguous(), q, delta.detach().contiguous(), q.shape[0])
        return out.view(b, -1, self.dim)


class HeadNeRF_final(nn.Module, _LatentSubspace):
    def __init__(self, args, size, device, dim=512, dim_shape=20, run_id='nerface2', emb_dir='./PTI/embeddings/',
                 use_softmax=False):
        super().__init__()
        self.base_dir = emb_dir + run_id
        self.device = device
        self.encoder = Encoder(size, dim, dim_shape, use_softmax, args.out_pose)
        self.args = args
        self.out_pose = args.out_pose
        self.dim, self.dim_shape = dim, dim_shape
        bases = torch.randn(self.dim_shape, 14 * self.dim).to(device)
        self.bases = nn.Parameter(bases, requires_grad=True)
        self.delta = nn.Parameter(bases.mean(dim=0), requires_grad=True)
        if getattr(args, 'person_2', False):
            if getattr(args, 'init', False):
                raise HfagpError('--person_2 --init (PTI embedding files) is outside the hot-path scope')
            bases_2 = torch.randn(self.dim_shape, 14 * self.dim).to(device)
            if not args.same_bases:
                self.bases_2 = nn.Parameter(bases_2, requires_grad=True)
            self.delta_2 = nn.Parameter(bases_2.mean(dim=0), requires_grad=True)
        self.generator = load_G_official(args, self.device)

    def get_delta(self, person_2=False):
        return (self.delta_2 if person_2 else self.delta).view(-1, self.dim)

    def get_latent(self, weights, person_2=False):
        if not person_2:
            return self._latent_from(weights, self.bases, self.delta)
        bases = self.bases if self.args.same_bases else self.bases_2
        return self._latent_from(weights, bases, self.delta_2)

    def get_weights(self, image):
        return self.encoder(image)                      # (weights, pose) when out_pose

    def get_image(self, latent, label):
        flip_label_(label)                              # in place, as the reference does
        return self.generator.synthesis(latent, c=label, noise_mode='const')['image']

    def forward(self, image, label, person_2=False):
        flip_label_(label)
        if self.out_pose:
            weights, pose = self.encoder(image)
            latent = self.get_latent(weights, person_2)
            return self.generator.synthesis(latent, c=label, noise_mode='const')['image'], pose
        latent = self.get_latent(self.encoder(image), person_2)
        return self.generator.synthesis(latent, c=label, noise_mode='const')['image']


class Weights_3DMM(nn.Module):
    def __init__(self, input_dim=76, dim=512, dim_shape=50, use_softmax=False):
        super().__init__()
        fc = [EqualLinear(input_dim, dim)] + [EqualLinear(dim, dim) for _ in range(5)] + [EqualLinear(dim, dim_shape)]
        self.fc = nn.Sequential(*fc)
        self.softmax = nn.Softmax(dim=1)
        self.use_softmax = use_softmax

    def forward(self, input):
        weights = self.fc(input)
        return self.softmax(weights) if self.use_softmax else weights


class _DrivenAvatar(nn.Module, _LatentSubspace):
    """Common body of HeadNeRF_3DMM / HeadNeRF_Audio (identical in the reference, :162-279)."""

    def __init__(self, args, size, device, dim=512, dim_shape=20, run_id='nerface2', emb_dir='./PTI/embeddings/',
                 use_softmax=False):
        super().__init__()
        self.base_dir = emb_dir + run_id
        self.device = device
        self.weights_3dmm = Weights_3DMM(input_dim=args.params_len, dim=dim, dim_shape=dim_shape,
                                         use_softmax=use_softmax)
        self.dim, self.dim_shape = dim, dim_shape
        bases = torch.randn(self.dim_shape, 14 * self.dim).to(device)
        self.bases = nn.Parameter(bases, requires_grad=True)
        self.delta = nn.Parameter(bases.mean(dim=0), requires_grad=True)
        self.generator = load_G_official(args, self.device)

    def get_latent(self, weights, person_2=False):
        return self._latent_from(weights, self.bases, self.delta)

    def get_weights(self, params):
        return self.weights_3dmm(params)

    def get_image(self, latent, label):
        flip_label_(label)
        return self.generator.synthesis(latent, c=label, noise_mode='const')['image']

    def forward(self, params, label, person_2=False):
        flip_label_(label)
        latent = self.get_latent(self.weights_3dmm(params), person_2)
        return self.generator.synthesis(latent, c=label, noise_mode='const')['image']


class HeadNeRF_3DMM(_DrivenAvatar):
    pass


class HeadNeRF_Audio(_DrivenAvatar):
    pass


class _ExactConv1d(nn.Conv1d):
    """nn.Conv1d (same parameters / state_dict keys) evaluated as unfold + fp32 matmul: cuDNN runs these tiny
    convolutions in TF32 by default and picks its backward algorithm per call, which put run-to-run noise of up to 6e-2
    (relative L2) on the gradient of AudioAttNet's first layer; the reference's fp32 arithmetic is what the path promises."""

    def forward(self, x):
        k, s, p = self.kernel_size[0], self.stride[0], self.padding[0]
        cols = torch.nn.functional.pad(x, (p, p)).unfold(2, k, s)           # [N, C, L_out, k]
        n, c, lo, _ = cols.shape
        y = cols.permute(0, 2, 1, 3).reshape(n * lo, c * k) @ self.weight.reshape(self.out_channels, c * k).t()
        if self.bias is not None:
            y = y + self.bias
        return y.view(n, lo, self.out_channels).permute(0, 2, 1)


class AudioAttNet(nn.Module):
    """8-frame attention smoothing of audio features (tiny; stays in PyTorch per SURVEY §2 #4)."""

    def __init__(self, dim_aud=32, seq_len=8):
        super().__init__()
        self.seq_len, self.dim_aud = seq_len, dim_aud
        chans = [dim_aud, 16, 8, 4, 2, 1]
        layers = []
        for a, b in zip(chans[:-1], chans[1:]):
            layers += [_ExactConv1d(a, b, kernel_size=3, stride=1, padding=1, bias=True), nn.LeakyReLU(0.02, True)]
        self.attentionConvNet = nn.Sequential(*layers)
        self.attentionNet = nn.Sequential(nn.Linear(seq_len, seq_len, bias=True), nn.Softmax(dim=1))

    def forward(self, x):
        y = x[..., :self.dim_aud].permute(1, 0).unsqueeze(0)
        y = self.attentionConvNet(y)
        y = self.attentionNet(y.view(1, self.seq_len)).view(self.seq_len, 1)
        return torch.sum(y * x, dim=0)


class AudioNet(nn.Module):
    """DeepSpeech window [n,16,29] -> [n,dim_aud] (tiny; stays in PyTorch per SURVEY §2 #4)."""

    def __init__(self, dim_aud=76, win_size=16):
        super().__init__()
        self.win_size, self.dim_aud = win_size, dim_aud
        chans = [29, 32, 32, 64, 64]
        layers = []
        for a, b in zip(chans[:-1], chans[1:]):
            layers += [_ExactConv1d(a, b, kernel_size=3, stride=2, padding=1, bias=True), nn.LeakyReLU(0.02, True)]
        self.encoder_conv = nn.Sequential(*layers)
        self.encoder_fc1 = nn.Sequential(nn.Linear(64, 64), nn.LeakyReLU(0.02, True), nn.Linear(64, dim_aud))

    def forward(self, x):
        half_w = self.win_size // 2
        x = x[:, 8 - half_w:8 + half_w, :].permute(0, 2, 1)
        x = self.encoder_conv(x).squeeze(-1)
        return self.encoder_fc1(x).squeeze()
