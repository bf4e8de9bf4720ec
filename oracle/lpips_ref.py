"""CPU fp32 ORACLE of LPIPS(net='alex') as the trainers use it (``/root/reference/code/trainer_rgb.py:10,62,86-87``):
``loss = LPIPS(net='alex').to(device).eval()(real, generated)`` -> ``[B,1,1,1]``.

TEST INFRASTRUCTURE ONLY (see ``oracle/eg3d_ref.py`` header for who may import this).

The reference imports the un-vendored, un-pinned pip package ``lpips`` (richzhang/PerceptualSimilarity), absent
offline together with its AlexNet / linear-head weights.  This module restates its published structure in plain
PyTorch with the package's ``state_dict`` names (``net.sliceK.*``, ``linK.model.1.weight``, ``scaling_layer.*``): ScalingLayer
-> torchvision-AlexNet ``features`` cut after each ReLU -> channel-unit-normalise -> squared difference -> 1x1 ``lin``
(Dropout is identity in eval) -> spatial mean -> sum over the five layers.

PARITY UNPINNED: neither the package nor any golden value for it is on this machine; the restatement is anchored on
the reference's call site (two [-1,1] images in, ``[B,1,1,1]`` out, ``squeeze().mean()`` taken by the trainer) and on
the package's published layer table.  Weights: seeded random init (SURVEY.md 8d cfg 3), ``lin`` heads non-negative.
"""
from __future__ import annotations

import torch
from torch import nn


class _ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer('shift', torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('scale', torch.tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, x):
        return (x - self.shift) / self.scale


class _Alex(nn.Module):
    """torchvision alexnet().features cut into the five LPIPS slices (module indices kept)."""

    def __init__(self):
        super().__init__()
        self.slice1, self.slice2, self.slice3 = nn.Sequential(), nn.Sequential(), nn.Sequential()
        self.slice4, self.slice5 = nn.Sequential(), nn.Sequential()
        self.slice1.add_module('0', nn.Conv2d(3, 64, 11, 4, 2))
        self.slice1.add_module('1', nn.ReLU())
        self.slice2.add_module('2', nn.MaxPool2d(3, 2))
        self.slice2.add_module('3', nn.Conv2d(64, 192, 5, 1, 2))
        self.slice2.add_module('4', nn.ReLU())
        self.slice3.add_module('5', nn.MaxPool2d(3, 2))
        self.slice3.add_module('6', nn.Conv2d(192, 384, 3, 1, 1))
        self.slice3.add_module('7', nn.ReLU())
        self.slice4.add_module('8', nn.Conv2d(384, 256, 3, 1, 1))
        self.slice4.add_module('9', nn.ReLU())
        self.slice5.add_module('10', nn.Conv2d(256, 256, 3, 1, 1))
        self.slice5.add_module('11', nn.ReLU())

    def forward(self, x):
        outs = []
        for s in (self.slice1, self.slice2, self.slice3, self.slice4, self.slice5):
            x = s(x)
            outs.append(x)
        return outs


class _NetLinLayer(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.model = nn.Sequential(nn.Dropout(), nn.Conv2d(cin, 1, 1, 1, 0, bias=False))

    def forward(self, x):
        return self.model(x)


def _normalize(x, eps=1e-10):
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


class LPIPS(nn.Module):
    CHNS = (64, 192, 384, 256, 256)

    def __init__(self, net='alex', seed=0, verbose=False):
        super().__init__()
        if net != 'alex':
            raise ValueError("only LPIPS(net='alex') is on the HFA-GP path (trainer_rgb.py:62)")
        with torch.random.fork_rng():
            torch.manual_seed(seed)
            self.scaling_layer = _ScalingLayer()
            self.net = _Alex()
            self.lins = nn.ModuleList(_NetLinLayer(c) for c in self.CHNS)
            for lin in self.lins:                      # the trained heads are non-negative
                lin.model[1].weight.data.abs_()
        for k, lin in enumerate(self.lins):
            setattr(self, f'lin{k}', lin)
        self.requires_grad_(False)

    def forward(self, in0, in1, retPerLayer=False, normalize=False):
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        f0, f1 = self.net(self.scaling_layer(in0)), self.net(self.scaling_layer(in1))
        val = 0
        for k in range(len(self.CHNS)):
            d = (_normalize(f0[k]) - _normalize(f1[k])) ** 2
            val = val + self.lins[k](d).mean([2, 3], keepdim=True)
        return val
