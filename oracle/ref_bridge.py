"""Import the reference's own ``networks/{encoder3d,headnerf}.py`` when it is on this machine.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the authoring container but not
on the GPU box, so everything here degrades to ``None`` and callers skip.

``headnerf.py:6-7`` does ``import dnnlib`` / ``import legacy`` (NVlabs/eg3d, un-vendored);
both are stubbed in ``sys.modules`` just long enough for the import to succeed.  The
generator those modules would un-pickle is *not* available, so only the encoder, the
latent map and the driving heads of the reference can be executed here.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_CODE = os.environ.get('HFAGP_REFERENCE_CODE', '/root/reference/code')


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_CODE, 'networks', 'headnerf.py'))


def load():
    """Returns (encoder3d_module, headnerf_module, cam_utils_module) or None."""
    if not available():
        return None
    saved_path = list(sys.path)
    saved = {k: sys.modules.get(k) for k in ('dnnlib', 'legacy', 'networks', 'networks.encoder3d',
                                             'networks.headnerf', 'cam_utils')}
    try:
        sys.path.insert(0, REFERENCE_CODE)
        for k in ('networks', 'networks.encoder3d', 'networks.headnerf', 'cam_utils'):
            sys.modules.pop(k, None)
        sys.modules['dnnlib'] = types.ModuleType('dnnlib')
        sys.modules['legacy'] = types.ModuleType('legacy')
        enc = importlib.import_module('networks.encoder3d')
        head = importlib.import_module('networks.headnerf')
        cam = importlib.import_module('cam_utils')
        return enc, head, cam
    finally:
        sys.path[:] = saved_path
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
