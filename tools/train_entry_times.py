"""Per C-ABI entry point GPU time of one trainer_rgb.gen_update step (eager, every call bracketed by CUDA events on the
launching stream, hfa_gp_b200._cabi.start_timing): total per entry point and the slowest single calls."""
import argparse, collections, sys
sys.path.insert(0, '.')
import torch
from hfa_gp_b200 import trainer_rgb, _cabi
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
calls = _cabi.start_timing()
torch.cuda._sleep(int(6e7))
step()
torch.cuda.synchronize()
_cabi.stop_timing()
agg = collections.defaultdict(list)
for name, a, b in calls:
    agg[name].append(a.elapsed_time(b) * 1e3)
tot = sum(sum(v) for v in agg.values())
print(f'{len(calls)} C-ABI calls, {tot:.0f} us between their event brackets')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:24]:
    top = ', '.join(f'{x:.0f}' for x in sorted(v, reverse=True)[:5])
    print(f'{k:34s} n={len(v):3d} {sum(v):8.0f} us   slowest: {top}')
