"""Stand-in for NVlabs/eg3d's ``legacy`` — ``load_network_pkl(f)['G_ema']`` (``code/networks/headnerf.py:34``)
returns the B200-native ``TriPlaneGenerator``.  Accepted stream contents: a ``torch.save``-d ``state_dict`` with
EG3D names (optionally wrapped as ``{'G_ema': state_dict}``), or the synthetic marker of ``dnnlib.util.open_url``."""
import io
import os

import torch

from dnnlib.util import SYNTHETIC_MAGIC


def load_network_pkl(f, force_fp16=False):
    from hfa_gp_b200.generator import TriPlaneGenerator, make_generator
    data = f.read()
    if data == SYNTHETIC_MAGIC:
        g = make_generator(seed=int(os.environ.get('HFAGP_GENERATOR_SEED', 0)), device='cpu')
    else:
        sd = torch.load(io.BytesIO(data), map_location='cpu', weights_only=True)
        sd = sd.get('G_ema', sd)
        g = TriPlaneGenerator()
        g.load_state_dict(sd)
    return {'G_ema': g, 'G': g}
