"""Run N eager frames of the configs[1] inference loop (no CUDA graph, single stream order) so that ncu can list or
capture individual launches.  usage: python tools/one_frame.py [steps] [--serial] [--batch B]  (B frames per step, bench.py's --frames-per-step)"""
import argparse, sys
sys.path.insert(0, '.')
import torch
from hfa_gp_b200.networks.headnerf import HeadNeRF_final
from hfa_gp_b200 import cam_utils

frames = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 3
dev = torch.device('cuda')
ns = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='', synthetic_generator=True, generator_seed=0)
torch.manual_seed(0)
model = HeadNeRF_final(ns, 256, dev, 512, 50, 'bench', './').to(dev).eval().requires_grad_(False)
if '--serial' in sys.argv:
    model.generator.overlap_streams = False
B = int(sys.argv[sys.argv.index('--batch') + 1]) if '--batch' in sys.argv else 1
img = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
lab = cam_utils.cam_sampler(B, 'cpu').to(dev)
for _ in range(frames):
    with torch.no_grad():
        out = model.get_image(model.get_latent(model.get_weights(img)), lab.clone())
torch.cuda.synchronize()
print('ok', tuple(out.shape), float(out.abs().mean()))
