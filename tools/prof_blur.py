"""Time the encoder's blur launches (configs[1] shapes, batch 1) in isolation: stride-1 blur in front of the 3x3 stride-2
conv (pad 2,2) and stride-2 blur in front of the 1x1 skip conv (pad 1,1).  HFAGP_BLUR_BIG=0|1 forces the 1-row / 4-row form."""
import sys
sys.path.insert(0, '.')
import torch
from hfa_gp_b200 import ops
tot = 0.0
for res, c in ((256, 64), (128, 128), (64, 256), (32, 512), (16, 512), (8, 512)):
    for stride, pad in ((1, 2), (2, 1)):
        xs = [ops.split(torch.randn(1, res, res, c, device='cuda')) for _ in range(3)]
        for i in range(3): ops.blur(xs[i % 3], pad, pad, stride=stride, split_out=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30): ops.blur(xs[i % 3], pad, pad, stride=stride, split_out=True)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 30 * 1e3
        tot += us
        print(f'blur {res}^2 x{c} stride {stride}: {us:7.1f} us')
print(f'total {tot:.1f} us')
