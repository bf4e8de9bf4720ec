"""Optimiser and data-parallel gradient exchange of the training step.

The reference trains with ``torch.optim.Adam(self.gen.parameters(), lr=args.lr)`` over a
``DistributedDataParallel`` wrapper (``/root/reference/code/trainer_rgb.py:55-57``, ``trainer_3dmm.py:29-30``,
``trainer_audio.py:28-40``).  Here both live on ONE flat fp32 buffer per optimiser:

  * every parameter's ``.data`` / ``.grad`` is re-pointed at a slice of ``flat_param`` / ``flat_grad`` (64-byte
    aligned slices), so ``zero_grad`` is one memset, the gradient mean over ranks is ONE all-reduce of the live
    prefix (NCCL over NVLink on the GPU box, gloo in the CPU tests) and the update is ONE ``hfagp_adam_step``
    launch per group with the 1/world_size folded in;
  * parameters are ordered so those that can receive gradients while the generator is frozen come first: only that
    prefix is exchanged and stepped until ``tune_generator()`` (``trainer_rgb.py:69-71``) makes the rest live —
    torch's Adam likewise skips parameters whose ``.grad`` is None, and starts their ``step`` count at the first
    gradient they see;
  * ``state_dict()`` / ``load_state_dict()`` use torch.optim.Adam's layout (per-parameter ``step``, ``exp_avg``,
    ``exp_avg_sq`` keyed by position in ``gen.parameters()``), so checkpoints interchange with the reference's
    ``"g_optim"`` / ``"w_optim"`` entries (``trainer_rgb.py:130-151``).

Known divergence from ``torch.optim.Adam``: torch skips a parameter whose ``.grad`` is None on a step (and counts steps
per parameter); here every parameter of a live group is stepped with the group's count and a zero gradient, so a parameter
that stops receiving gradients (``bases_2`` / ``delta_2`` while ``person_2`` alternates) keeps moving on its decaying
momentum for a few steps instead of freezing.  The trainers' default paths (one person) never hit this.

Reference quirk handled on purpose: ``trainer_rgb.gen_update`` calls ``self.gen.module.*`` and so
bypasses DDP's reducer — replicas silently diverge.  We implement the intended synchronous mean for all three
trainers.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops
from ._cabi import HfagpError

_ALIGN = 16          # elements (64 bytes): keeps every slice 16-byte aligned for the float4 kernels


def _round_up(n: int, a: int = _ALIGN) -> int:
    return (n + a - 1) // a * a


class FlatAdam:
    """Adam over flat buffers.  ``params``: every parameter the reference hands to Adam, in ``parameters()`` order;
    ``live_first``: predicate choosing the parameters that train from step 1 (the rest follow in the buffer and
    join once they require gradients)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, live_first=None, process_group=None):
        self.params: List[torch.nn.Parameter] = list(params)
        if not self.params:
            raise HfagpError('FlatAdam got an empty parameter list')
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.lr, self.betas, self.eps, self.weight_decay = lr, tuple(betas), eps, weight_decay
        self.process_group = process_group
        dev = self.params[0].device
        live = [i for i, p in enumerate(self.params) if (live_first(p) if live_first else p.requires_grad)]
        late = [i for i in range(len(self.params)) if i not in set(live)]
        self._order = live + late                      # buffer order (indices into self.params)
        self._n_live_first = len(live)
        offs, total = {}, 0
        for i in self._order:
            offs[i] = total
            total += _round_up(self.params[i].numel())
            if i == (live[-1] if live else -1):
                self._live_elems = total
        if not live:
            self._live_elems = 0
        self._offs, self._total = offs, total
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        for i, p in enumerate(self.params):
            if p.dtype != torch.float32:
                raise HfagpError('FlatAdam handles fp32 parameters only (the reference trains in fp32)')
            v = self._view(self.flat_param, i)
            v.copy_(p.data)
            p.data = v
            p.grad = self._view(self.flat_grad, i)
        self.steps = [0, 0]                            # Adam step count of the early / late group
        # step-dependent scalars (lr / (1 - b1^t), sqrt(1 - b2^t)) of both groups: pinned host pair + device copy, read by
        # hfagp_adam_step_dev so that a captured step can be replayed (step_prepare() refreshes them each step)
        self._sched_host = torch.zeros(4, dtype=torch.float32).pin_memory() if dev.type == 'cuda' else torch.zeros(4)
        self._sched_dev = torch.zeros(4, device=dev, dtype=torch.float32)
        self.world = 1
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(process_group)

    # ------------------------------------------------------------------ views
    def _view(self, flat, i):
        p = self.params[i]
        o = self._offs[i]
        return flat[o:o + p.numel()].view(p.shape)

    def _late_live(self) -> bool:
        return any(self.params[i].requires_grad for i in self._order[self._n_live_first:])

    def live_elements(self) -> int:
        return self._total if self._late_live() else self._live_elems

    # ------------------------------------------------------------------ torch.optim surface used by the trainers
    def zero_grad(self, set_to_none: bool = False):
        """One memset; ``.grad`` stays a view of the flat buffer (autograd accumulates into it in place)."""
        self.flat_grad.zero_()
        for i, p in enumerate(self.params):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * self._offs[i]:
                p.grad = self._view(self.flat_grad, i)

    def sync_gradients(self):
        """SUM all-reduce of the live gradient prefix (the mean's 1/world is folded into the Adam kernel)."""
        if self.world > 1:
            n = self.live_elements()
            if n:
                dist.all_reduce(self.flat_grad[:n], op=dist.ReduceOp.SUM, group=self.process_group)

    def _spans(self):
        spans = [(0, self._live_elems, 0)]
        if self._late_live():
            spans.append((self._live_elems, self._total, 1))
        return [(lo, hi, g) for lo, hi, g in spans if hi > lo]

    def step(self):
        self.sync_gradients()
        b1, b2 = self.betas
        kw = dict(lr=self.lr, beta1=b1, beta2=b2, eps=self.eps, weight_decay=self.weight_decay,
                  grad_scale=1.0 / self.world)
        for lo, hi, g in self._spans():
            self.steps[g] += 1
            ops.adam_step(self.flat_param[lo:hi], self.flat_grad[lo:hi], self.exp_avg[lo:hi],
                          self.exp_avg_sq[lo:hi], step=self.steps[g], **kw)
        ops.param_epoch[0] += 1                        # packed-weight caches must be rebuilt

    # ---- the same step in two halves, for a training step replayed from a CUDA graph: the host half advances the step
    # counts and refreshes the device copy of the step-dependent scalars (outside the graph, every step); the device half
    # (captured once) is the all-reduce + the Adam kernels reading those scalars from device memory
    def step_prepare(self):
        b1, b2 = self.betas
        for _, _, g in self._spans():
            self.steps[g] += 1
            ops.adam_sched(self.lr, b1, b2, self.steps[g], self._sched_host, 2 * g)
        self._sched_dev.copy_(self._sched_host, non_blocking=True)
        ops.param_epoch[0] += 1

    def step_launch(self):
        self.sync_gradients()
        b1, b2 = self.betas
        for lo, hi, g in self._spans():
            ops.adam_step_dev(self.flat_param[lo:hi], self.flat_grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi],
                              self._sched_dev[2 * g:2 * g + 2], beta1=b1, beta2=b2, eps=self.eps,
                              weight_decay=self.weight_decay, grad_scale=1.0 / self.world)

    # ------------------------------------------------------------------ checkpoints (torch.optim.Adam layout)
    def _group_of(self, i) -> int:
        return 0 if i in self._order[:self._n_live_first] else 1

    def state_dict(self):
        state = {}
        for i in range(len(self.params)):
            st = self.steps[self._group_of(i)]
            if st == 0:
                continue
            state[i] = {'step': torch.tensor(float(st)), 'exp_avg': self._view(self.exp_avg, i).clone(),
                        'exp_avg_sq': self._view(self.exp_avg_sq, i).clone()}
        group = dict(lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, amsgrad=False,
                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     decoupled_weight_decay=False, params=list(range(len(self.params))))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        g = sd['param_groups'][0]
        if len(g['params']) != len(self.params):
            raise HfagpError(f"optimizer state has {len(g['params'])} parameters, this model {len(self.params)}")
        self.lr, self.betas, self.eps = g['lr'], tuple(g['betas']), g['eps']
        self.weight_decay = g.get('weight_decay', 0.0)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = [set(), set()]
        for k, st in sd['state'].items():
            i = g['params'].index(k) if k in g['params'] else int(k)
            self._view(self.exp_avg, i).copy_(st['exp_avg'])
            self._view(self.exp_avg_sq, i).copy_(st['exp_avg_sq'])
            steps[self._group_of(i)].add(int(float(st['step'])))
        for gi in (0, 1):
            if len(steps[gi]) > 1:
                # torch.optim.Adam counts steps per parameter and skips parameters whose .grad is None (e.g. bases_2 /
                # delta_2 on the steps person_2 is off); the flat optimiser steps a whole group with ONE count and a zero
                # gradient for such parameters (their momentum keeps decaying).  Loading a reference checkpoint with mixed
                # counts keeps the largest one: bias correction of the stragglers is then slightly ahead of torch's.
                import warnings
                warnings.warn(f'FlatAdam: per-parameter step counts {sorted(steps[gi])} differ inside one group; '
                              f'using {max(steps[gi])} for all of them')
            self.steps[gi] = max(steps[gi]) if steps[gi] else 0


class DataParallelShard(torch.nn.Module):
    """Stands where the reference puts ``DistributedDataParallel`` (``trainer_rgb.py:55``): exposes ``.module``,
    forwards calls, and on construction broadcasts rank 0's parameters so replicas start identical.  Gradient
    exchange is FlatAdam.sync_gradients (one flat all-reduce per step) instead of DDP's bucketed hooks."""

    def __init__(self, module: torch.nn.Module, device_ids: Optional[Sequence[int]] = None, process_group=None,
                 **_ddp_kwargs):
        super().__init__()
        self.module = module
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)

    def forward(self, *a, **k):
        return self.module(*a, **k)
