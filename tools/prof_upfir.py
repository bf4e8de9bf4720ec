import sys, math
sys.path.insert(0, '.')
import torch
from hfa_gp_b200 import ops
for res, c in ((512, 128), (256, 256), (256, 128)):
    ts = [torch.randn(1, res + 1, res + 1, c, device='cuda') for _ in range(3)]
    bias = torch.randn(c, device='cuda'); d = torch.rand(1, c, device='cuda') + 0.5
    for i in range(3): ops.upfir_act(ts[i % 3], dcoef=d, bias=bias, clamp=256.0, split_out=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): ops.upfir_act(ts[i % 3], dcoef=d, bias=bias, clamp=256.0, split_out=True)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    mb = ((res + 1) ** 2 * c * 4 + res * res * c * 4) / 1e6
    print(f'upfir {res}^2 x{c}: {us:7.1f} us  {mb / us * 1e3 / 1e3:6.2f} TB/s')
