"""Which torch (ATen) ops still launch glue kernels inside one trainer_rgb.gen_update step, by op and input shapes."""
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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True)
rows = [(e.self_device_time_total, e.count, e.key, str(e.input_shapes)) for e in ka if e.self_device_time_total > 0 and e.key.startswith('aten::')]
rows.sort(key=lambda r: -r[0])
print('aten self device time total us', sum(r[0] for r in rows))
for t, c, k, sh in rows[:45]:
    print(f'{t:8.1f} us n={c:3d} {k:24s} {sh[:120]}')
