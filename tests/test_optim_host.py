"""CPU: host logic of the flat optimiser / data-parallel exchange (hfa_gp_b200/optim.py) — buffer layout, the live
prefix, zero_grad views, torch.optim.Adam state_dict layout, and the world_size-2 gradient all-reduce + parameter
broadcast over gloo.  The Adam arithmetic itself is a CUDA kernel and is tested on the GPU
(tests/test_gpu_training.py); no compute call through the C ABI happens here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hfa_gp_b200.optim import DataParallelShard, FlatAdam


def _params():
    g = torch.Generator().manual_seed(0)
    shapes = [(5,), (3, 7), (2, 2, 3, 3), (9,)]
    return [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]


def test_layout_live_prefix_and_views():
    ps = _params()
    frozen = ps[1]
    before = [p.detach().clone() for p in ps]
    opt = FlatAdam(ps, lr=1e-3, live_first=lambda p: p is not frozen)
    frozen.requires_grad = False
    for p, b in zip(ps, before):
        assert torch.equal(p.detach(), b)                                  # values survive the re-pointing
        assert p.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0      # float4-aligned slices
    # live parameters first, the frozen one last; only the prefix is exchanged / stepped
    assert opt._order == [0, 2, 3, 1]
    assert opt.live_elements() == 16 + 48 + 16
    frozen.requires_grad = True
    assert opt.live_elements() == opt._total == 16 + 48 + 16 + 32
    # gradients accumulate into the flat buffer through autograd, zero_grad is one memset that keeps the views
    (ps[0].sum() * 2 + ps[2].sum()).backward()
    assert torch.equal(opt.flat_grad[:5], torch.full((5,), 2.0)) and float(opt.flat_grad.sum()) == 10 + 36
    ptr = ps[0].grad.data_ptr()
    opt.zero_grad()
    assert float(opt.flat_grad.abs().sum()) == 0 and ps[0].grad.data_ptr() == ptr
    ps[3].grad = None                                                      # e.g. someone called zero_grad(set_to_none=True)
    opt.zero_grad()
    assert ps[3].grad is not None and ps[3].grad.data_ptr() == opt.flat_grad.data_ptr() + 4 * opt._offs[3]


def test_state_dict_uses_torch_adam_layout():
    ps = _params()
    opt = FlatAdam(ps, lr=2e-3)
    ref = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=2e-3)
    for p in ref.param_groups[0]['params']:
        p.grad = torch.ones_like(p)
    ref.step()
    opt.load_state_dict(ref.state_dict())
    assert opt.steps == [1, 0]
    sd = opt.state_dict()
    assert sd['param_groups'][0]['params'] == [0, 1, 2, 3] and sd['param_groups'][0]['lr'] == 2e-3
    for k, st in ref.state_dict()['state'].items():
        assert torch.equal(sd['state'][k]['exp_avg'], st['exp_avg'])
        assert torch.equal(sd['state'][k]['exp_avg_sq'], st['exp_avg_sq'])
        assert float(sd['state'][k]['step']) == float(st['step'])
    # and torch's own Adam accepts what we write
    ref2 = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1.0)
    ref2.load_state_dict(sd)
    assert ref2.param_groups[0]['lr'] == 2e-3


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                                      # replicas start DIFFERENT on purpose
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        shard = DataParallelShard(net)                                     # broadcast from rank 0
        frozen = net[1].weight
        opt = FlatAdam(shard.parameters(), lr=1e-3, live_first=lambda p: p is not frozen)
        frozen.requires_grad = False
        assert opt.world == world
        start = opt.flat_param.clone()
        opt.zero_grad()
        x = torch.full((1, 4), float(rank + 1))
        shard(x).sum().backward()
        local = opt.flat_grad.clone()
        opt.sync_gradients()
        out[rank] = (start, local, opt.flat_grad.clone(), opt.live_elements())
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_and_broadcast_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = {k: v for k, v in out.items()}
    (s0, l0, g0, n0), (s1, l1, g1, n1) = res[0], res[1]
    assert torch.equal(s0, s1)                                             # parameters broadcast from rank 0
    assert n0 == n1 and 0 < n0 < s0.numel()
    assert not torch.equal(l0[:n0], l1[:n0])                               # ranks saw different data
    assert torch.equal(g0, g1)                                             # ... and hold the same SUM afterwards
    assert torch.allclose(g0[:n0], l0[:n0] + l1[:n0])
    assert torch.equal(g0[n0:], l0[n0:])                                   # the frozen tail is not exchanged
