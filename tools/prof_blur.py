"""Time the encoder's blur launches in isolation at the bench's batch (4 frames): stride-1 blur in front of the 3x3 stride-2
conv (pad 2,2) and stride-2 blur in front of the 1x1 skip conv (pad 1,1).  The calls are captured in a CUDA graph (20 per replay,
inputs rotated) so the number is GPU time, not Python launch overhead.  HFAGP_BLUR_BIG=0|1 forces the 1-row / 4-row form."""
import sys
sys.path.insert(0, '.')
import torch
from hfa_gp_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tot = 0.0
for res, c in ((256, 64), (128, 128), (64, 256), (32, 512), (16, 512), (8, 512)):
    for stride, pad in ((1, 2), (2, 1)):
        nb = 4
        xs = [ops.split(torch.randn(B, res, res, c, device='cuda')) for _ in range(nb)]
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for i in range(3): ops.blur(xs[i % nb], pad, pad, stride=stride, split_out=True)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                outs = [ops.blur(xs[i % nb], pad, pad, stride=stride, split_out=True) for i in range(20)]
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5): g.replay()
            e1.record(s); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 100 * 1e3
        oh = (res + 2 * pad - 4) // stride + 1
        mb = (res * res + oh * oh) * c * 4 * B / 1e6
        tot += us
        print(f'blur {B}x{res}^2 x{c} stride {stride}: {us:7.1f} us  {mb / us:6.2f} TB/s')
        del g, outs
print(f'total {tot:.1f} us per {B}-frame step')
