"""Which torch (ATen) ops still launch glue kernels inside one gen_update step (eager), by op and input shapes.
usage: python tools/timeline_train_ops.py [--tune] [--3dmm] [--batch B]"""
import argparse, collections, os, sys
sys.path.insert(0, '.')
os.environ.setdefault('HFAGP_SYNTHETIC_LPIPS', '1')
import torch
from torch.profiler import profile, ProfilerActivity
from hfa_gp_b200 import trainer_3dmm, trainer_rgb
dev = torch.device('cuda')
B = int(sys.argv[sys.argv.index('--batch') + 1]) if '--batch' in sys.argv else 2
ns = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='', synthetic_generator=True,
                        generator_seed=0, batch_size=B, size=256, latent_dim_style=512, latent_dim_shape=50, run_id='b',
                        emb_dir='./', lr=3e-4, params_len=76)
torch.manual_seed(0)
tr = (trainer_3dmm if '--3dmm' in sys.argv else trainer_rgb).Trainer(ns, dev, 0)
if '--tune' in sys.argv:
    tr.tune_generator()
real = (torch.rand(B, 3, 256, 256, device=dev) * 2 - 1)
params = torch.randn(B, 76, device=dev)
lab0 = trainer_rgb.cam_sampler(B, dev)
def step():
    if '--3dmm' in sys.argv:
        tr.gen_update(real, lab0.clone(), params)
    else:
        tr.gen_update(real, lab0.clone())
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
for t, c, k, sh in rows[:70]:
    print(f'{t:8.1f} us n={c:3d} {k:24s} {sh[:120]}')
