"""Debug: per-role cycle breakdown of conv_tc_kernel on one layer (needs libhfagp_sm100_dbg.so, built with -DHFAGP_TC_TIMING)."""
import ctypes as C, sys, math, os
sys.path.insert(0, '.')
import torch
from hfa_gp_b200 import _cabi
_cabi.LIB_PATH = os.path.join(os.path.dirname(_cabi.LIB_PATH), 'libhfagp_sm100_dbg.so')
from hfa_gp_b200 import ops
res, cin, cout, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
x = ops.split(torch.randn(1, res, res, cin, device='cuda'))
w = ops.split(torch.randn(1, 9, cout, cin, device='cuda'))
bias = torch.randn(cout, device='cuda')
for _ in range(3):
    if kind == 'conv3':
        ops.conv2d_tc(x, w, ops.TAPS_3X3, cout, oh=res, ow=res, w_batched=True, split_out=True, bias=bias, act=1, act_gain=math.sqrt(2))
    else:
        ops.conv_transpose_s2_tc(x, w, cout, w_batched=True)
torch.cuda.synchronize()
buf = (C.c_ulonglong * (256 * 8))()
assert _cabi.lib().hfagp_debug_tc_timing(buf) == 0
import numpy as np
a = np.array(buf[:], dtype=np.float64).reshape(256, 8)[:148]
names = ['mma loop', 'wait acc_empty', 'wait a_full', 'wait b_full', 'epi total', 'epi wait acc_full', 'tiles']
for i, nm in enumerate(names):
    print(f'{nm:20s} mean {a[:, i].mean():12.0f}  min {a[:, i].min():12.0f} max {a[:, i].max():12.0f}')
