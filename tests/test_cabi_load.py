"""CPU: the C-ABI library loads and exports every symbol include/hfagp.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'hfagp.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(hfagp_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from hfa_gp_b200 import _cabi
    if not os.path.isfile(_cabi.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 13
    for s in declared:
        assert hasattr(lib, s), f'{s} declared in include/hfagp.h but not exported'
    assert sorted(_cabi.SYMBOLS) == declared, 'python binding list out of sync with the header'
    lib.hfagp_abi_version.restype = ctypes.c_int
    assert lib.hfagp_abi_version() == 1


def test_invalid_arguments_fail_without_touching_the_gpu():
    from hfa_gp_b200 import _cabi
    lib = _cabi.lib()
    d = _cabi.ConvDesc()
    rc = lib.hfagp_conv2d_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None, None)
    assert rc == -1
    assert b'null pointer' in lib.hfagp_last_error()
    with pytest.raises(_cabi.HfagpError):
        _cabi.check(rc, 'hfagp_conv2d_fwd')


def test_workspace_queries_are_pure_functions_of_the_shapes():
    """hfagp_*_workspace_bytes (caller-provided scratch): no device is touched."""
    from hfa_gp_b200 import _cabi
    lib = _cabi.lib()
    d = _cabi.ConvDesc()
    assert lib.hfagp_conv2d_tc_acc_workspace_bytes(ctypes.byref(d)) == 0
    d.batch, d.out_h, d.out_w, d.cout = 2, 16, 16, 512
    assert lib.hfagp_conv2d_tc_acc_workspace_bytes(ctypes.byref(d)) == 2 * 16 * 16 * 512 * 4
    r = _cabi.RenderDesc(2, 128, 256, 256, 48, 48, 0.02, 2.0)
    fb, db = ctypes.c_size_t(), ctypes.c_size_t()
    tot = lib.hfagp_render_bwd_dec_workspace_bytes(ctypes.byref(r), ctypes.byref(fb), ctypes.byref(db))
    samples = 2 * 128 * 128 * 96
    assert fb.value == samples * 32 * 4 and db.value == samples * 33 * 4 and tot == fb.value + db.value
    assert lib.hfagp_device_sm_count() > 0          # 148 when no device can be asked


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'hfa_gp_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, os.path.join(dp, f)


def test_hot_path_refuses_cpu_tensors():
    import torch
    from hfa_gp_b200 import _cabi
    with pytest.raises(_cabi.HfagpError):
        _cabi.ptr(torch.zeros(4))


def test_workspace_queries_are_host_only_and_aligned():
    """The *_workspace_bytes queries run without a GPU (no compute call): sizes are 256-byte multiples, grow with the
    problem and reject nonsense."""
    from hfa_gp_b200 import _cabi
    lib = _cabi.lib()
    a = lib.hfagp_basis_qr_workspace_bytes(50, 14 * 512)
    b = lib.hfagp_basis_qr_workspace_bytes(64, 14 * 512)
    assert (a - 256) % 256 == 0 and b > a >= 50 * 14 * 512 * 4
    assert lib.hfagp_basis_qr_workspace_bytes(0, 100) == 0 and lib.hfagp_basis_qr_workspace_bytes(8, 0) == 0
