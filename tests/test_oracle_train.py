"""CPU: the training-step oracle (oracle/train_ref.py).  Pinned, where /root/reference is present, against a literal
run of the reference's own modules (Encoder + HeadNeRF_final.get_latent + torch Adam) around the oracle generator;
always checked for determinism and that the step actually trains."""
import types
import warnings

import pytest
import torch
import torch.nn.functional as F

from oracle import eg3d_ref, hfagp_ref, ref_bridge, train_ref

needs_ref = pytest.mark.skipif(not ref_bridge.available(), reason='/root/reference not on this machine')


def _inputs(cfg, b, size, seed):
    g = torch.Generator().manual_seed(seed)
    real = torch.rand(b, 3, size, size, generator=g) * 2 - 1
    label = hfagp_ref.synthetic_labels(b, seed=seed)
    jit = torch.rand(b, cfg.nrr ** 2, cfg.depth_res, 1, generator=g)
    u = torch.rand(b * cfg.nrr ** 2, cfg.depth_res_importance, generator=g)
    return real, label, jit, u


def _setup(seed=0, size=16, k=6):
    cfg = eg3d_ref.tiny_config()
    gen = eg3d_ref.make_generator(cfg, seed=seed, noise_strength=0.1)
    sd = hfagp_ref.make_encoder_state(size=size, dim_motion=k, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    bases = torch.randn(k, cfg.num_ws * 512, generator=g)
    return cfg, gen, sd, bases, bases.mean(0)


def test_train_step_is_deterministic_and_trains():
    cfg, gen, sd, bases, delta = _setup()
    a = train_ref.TrainStepRef(sd, bases, delta, gen, 16, 1e-3)
    b = train_ref.TrainStepRef(sd, bases, delta, gen, 16, 1e-3)
    real, label, jit, u = _inputs(cfg, 2, 16, 0)
    first = None
    for it in range(4):
        la, _, ia = a.step(real, label, jit, u)
        lb, _, ib = b.step(real, label, jit, u)
        assert torch.equal(la, lb) and torch.equal(ia, ib)
        first = first if first is not None else float(la)
    assert float(la) < first                       # same batch four times: the loss must go down
    assert not torch.equal(a.bases.detach(), bases) and ia.shape == (2, 3, 16, 16)
    assert all(p.grad is None for p in gen.parameters())      # generator frozen (trainer_rgb.py:59-60)


@needs_ref
def test_train_step_equals_reference_modules():
    """gen_update transcribed onto the reference's OWN Encoder / get_latent (trainer_rgb.py:73-98) gives the same
    gradients and the same parameters after two Adam steps as the oracle's functional restatement."""
    warnings.simplefilter('ignore')
    enc_mod, head, _ = ref_bridge.load()
    size, k = 16, 6
    cfg, gen, sd, bases, delta = _setup(seed=3, size=size, k=k)
    oracle = train_ref.TrainStepRef(sd, bases, delta, gen, size, 3e-4)

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = enc_mod.Encoder(size, 512, k)
            self.encoder.load_state_dict(sd)
            self.bases = torch.nn.Parameter(bases.clone())
            self.delta = torch.nn.Parameter(delta.clone())
            self.dim, self.args = 512, None
    m = Model()
    opt = torch.optim.Adam(m.parameters(), lr=3e-4)
    for it in range(2):
        real, label, jit, u = _inputs(cfg, 2, size, it)
        l2_o, _, img_o = oracle.step(real, label, jit, u)
        m.train()
        opt.zero_grad()
        w = m.encoder(real)
        latent = head.HeadNeRF_final.get_latent(m, w, False)
        lab = label.clone()
        lab[:, [1, 2, 5, 6, 9, 10]] *= -1                                        # headnerf.py:132
        img = gen.synthesis(latent, lab, jitter_coarse=jit, u_fine=u)['image']  # headnerf.py:133
        img = torch.nn.AdaptiveAvgPool2d((size, size))(img)                     # trainer_rgb.py:63,84
        l2 = torch.nn.MSELoss(reduction='mean')(real, img)                      # trainer_rgb.py:15,85
        l2.backward()
        opt.step()
        assert torch.equal(l2.detach(), l2_o) and torch.equal(img.detach(), img_o)
        for n, p in m.encoder.named_parameters():
            assert torch.equal(p.grad, oracle.sd[n].grad), n
        assert torch.equal(m.bases.grad, oracle.bases.grad)
    for n, p in m.encoder.named_parameters():
        assert torch.equal(p.detach(), oracle.sd[n].detach()), n
    assert torch.equal(m.delta.detach(), oracle.delta.detach())
