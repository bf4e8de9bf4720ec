"""hfa_gp_b200 — B200-native (sm_100a) hot path of HFA-GP's per-frame render.

Layout: ``csrc/`` CUDA kernels + C ABI (``include/hfagp.h``), ``_cabi.py``/``ops.py`` the ctypes
binding, ``generator.py`` the EG3D tri-plane generator protocol, ``networks/`` the drop-in mirror of
the reference's ``code/networks`` package, ``shims/`` the ``dnnlib``/``legacy`` stand-ins that let the
unmodified reference ``headnerf.py`` run on this generator.
"""
__version__ = '0.1.0'
