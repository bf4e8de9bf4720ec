"""Per-kernel GPU time of one trainer_rgb.gen_update step (torch.profiler / CUPTI)."""
import argparse, collections, sys
sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from hfa_gp_b200 import trainer_rgb
dev = torch.device('cuda')
ns = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='', synthetic_generator=True,
                        generator_seed=0, batch_size=2, size=256, latent_dim_style=512, latent_dim_shape=50, run_id='b',
                        emb_dir='./', lr=3e-4)
torch.manual_seed(0)
tr = trainer_rgb.Trainer(ns, dev, 0)
if '--tune' in sys.argv:
    tr.tune_generator()
real = (torch.rand(2, 3, 256, 256, device=dev) * 2 - 1)
def step():
    tr.gen_update(real, trainer_rgb.cam_sampler(2, dev))
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
def short(n):
    n = n.replace('hfagp::', '').replace('void ', '')
    return (n[:n.index('(')] if '(' in n else n)[:56]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    agg[short(e.name)][0] += 1
    agg[short(e.name)][1] += e.time_range.elapsed_us()
tot = sum(v[1] for v in agg.values())
print(f'--- one step: {len(evs)} kernels, busy {tot:.1f} us')
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    print(f'{k:58s} n={v[0]:4d} {v[1]:9.1f} us {100 * v[1] / tot:5.1f}%')
