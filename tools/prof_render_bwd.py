"""Run the full-size tri-plane render BACKWARD (configs[2]: batch 2, 128x128 rays, 48+48 samples, 256x256x96 planes) a few times
so that `ncu -k regex:render_kernel` can capture it in isolation, and print its CUDA-event time."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hfa_gp_b200 import ops, cam_utils

torch.manual_seed(0)
dev = 'cuda'
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = 2
planes = torch.randn(B, 256, 256, 96, device=dev)
c = cam_utils.cam_sampler(B, dev)
c[:, [1, 2, 5, 6, 9, 10]] *= -1
mlp = torch.cat([torch.randn(64 * 32) / math.sqrt(32), torch.zeros(64), torch.randn(33 * 64) / 8, torch.zeros(33)]).to(dev)
lin = torch.linspace(2.25, 3.3, 48, device=dev)
jit = torch.rand(B, 128 * 128, 48, device=dev)
u = torch.rand(B * 128 * 128, 48, device=dev)
delta = (3.3 - 2.25) / 47
dfeat = torch.randn(B, 128, 128, 32, device=dev)
ts = []
for i in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.render_bwd(planes, c, mlp, lin, jit, u, dfeat, res=128, s_coarse=48, s_fine=48, delta=delta, box_scale=2.0)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print('render_bwd ms (incl. the dplanes memset):', ['%.3f' % t for t in ts])
