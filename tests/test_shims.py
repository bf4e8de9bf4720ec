"""CPU: the unmodified reference ``code/networks/headnerf.py`` constructs on top of the B200 generator through
the ``dnnlib`` / ``legacy`` shims (skipped where /root/reference is absent)."""
import argparse
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/code'


@pytest.mark.skipif(not os.path.isdir(REF), reason='/root/reference not on this machine')
def test_reference_headnerf_runs_on_shimmed_generator(monkeypatch):
    from hfa_gp_b200.generator import GeneratorConfig, TriPlaneGenerator
    monkeypatch.setenv('HFAGP_SYNTHETIC_GENERATOR', '1')
    monkeypatch.syspath_prepend(os.path.join(ROOT, 'hfa_gp_b200', 'shims'))
    monkeypatch.syspath_prepend(REF)
    for k in ('dnnlib', 'dnnlib.util', 'legacy', 'networks', 'networks.headnerf', 'networks.encoder3d'):
        sys.modules.pop(k, None)
    try:
        head = importlib.import_module('networks.headnerf')
        args = argparse.Namespace(out_pose=False, person_2=False)
        model = head.HeadNeRF_final(args, 64, 'cpu', 512, 50, 'x', './')
        assert isinstance(model.generator, TriPlaneGenerator)
        assert not any(p.requires_grad for p in model.generator.parameters())       # headnerf.py:34-36
        keys = model.state_dict().keys()
        assert 'generator.backbone.synthesis.b4.const' in keys and 'bases' in keys and 'encoder.fc.4.weight' in keys
        lat = model.get_latent(torch.randn(2, 50))
        assert lat.shape == (2, GeneratorConfig().num_ws, 512)
        with pytest.raises(RuntimeError):       # the generator refuses CPU tensors instead of falling back
            model.get_image(lat, torch.zeros(2, 25))
    finally:
        for k in ('dnnlib', 'dnnlib.util', 'legacy', 'networks', 'networks.headnerf', 'networks.encoder3d'):
            sys.modules.pop(k, None)


def test_state_dict_matches_oracle_key_for_key():
    from oracle import eg3d_ref
    from hfa_gp_b200.generator import GeneratorConfig, TriPlaneGenerator
    ref = eg3d_ref.TriPlaneGeneratorRef(eg3d_ref.tiny_config())
    cfg = eg3d_ref.tiny_config()
    prod = TriPlaneGenerator(GeneratorConfig(**{k: getattr(cfg, k) for k in GeneratorConfig.__dataclass_fields__}))
    a, b = ref.state_dict(), prod.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    prod.load_state_dict(a, strict=True)


def test_encoder_dropin_state_dict_matches_reference_keys():
    from oracle import hfagp_ref
    from hfa_gp_b200.networks.encoder3d import Encoder
    for size, pose in ((64, True), (256, False)):
        sd = hfagp_ref.make_encoder_state(size, 512, 50, out_pose=pose)
        enc = Encoder(size, 512, 50, False, pose)
        assert set(enc.state_dict().keys()) == set(sd.keys())
        enc.load_state_dict(sd, strict=True)
