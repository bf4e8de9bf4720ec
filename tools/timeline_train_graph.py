"""Warm in-graph timeline of one training step replayed from its CUDA graph (the product path of bench.py's `train` /
`train_rgb` records): span, time covered by >= 1 kernel, idle gaps, per-kernel totals and (--seq) the launch sequence.
usage: python tools/timeline_train_graph.py [--trainer 3dmm|rgb] [--batch B] [--seq]"""
import argparse, collections, os, sys
sys.path.insert(0, '.')
os.environ.setdefault('HFAGP_SYNTHETIC_LPIPS', '1')
import torch
from torch.profiler import profile, ProfilerActivity
from hfa_gp_b200 import trainer_3dmm, trainer_rgb

ap = argparse.ArgumentParser()
ap.add_argument('--trainer', default='3dmm', choices=['3dmm', 'rgb'])
ap.add_argument('--batch', type=int, default=1)
ap.add_argument('--seq', action='store_true')
args = ap.parse_args()
dev = torch.device('cuda')
bs = args.batch
ns = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='', synthetic_generator=True,
                        generator_seed=0, batch_size=bs, size=256, latent_dim_style=512, latent_dim_shape=50, run_id='b',
                        emb_dir='./', lr=3e-4, params_len=76)
torch.manual_seed(0)
tr = (trainer_3dmm if args.trainer == '3dmm' else trainer_rgb).Trainer(ns, dev, 0)
tr.enable_step_graph(warmup=2)
real = torch.rand(bs, 3, 256, 256, device=dev) * 2 - 1
params = torch.randn(bs, 76, device=dev)


def step():
    lab = trainer_rgb.cam_sampler(bs, dev)
    return tr.gen_update(real, lab, params) if args.trainer == '3dmm' else tr.gen_update(real, lab)


for _ in range(6):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)


def short(n):
    n = n.replace('hfagp::', '').replace('void ', '')
    return (n[:n.index('(')] if '(' in n else n)[:56]


last = evs[len(evs) // 2:]
iv = sorted((e.time_range.start, e.time_range.end) for e in last)
covered, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
for s_, e_ in iv[1:]:
    if s_ > cur_e:
        covered += cur_e - cur_s
        gaps.append((s_ - cur_e, cur_e - iv[0][0]))
        cur_s, cur_e = s_, e_
    else:
        cur_e = max(cur_e, e_)
covered += cur_e - cur_s
span = iv[-1][1] - iv[0][0]
print(f'{args.trainer} batch {bs}: {len(last)} kernels, span {span:.1f} us, covered by >= 1 kernel {covered:.1f} us, idle gaps {span - covered:.1f} us '
      f'({len(gaps)} gaps, largest {sorted(gaps)[-3:] if gaps else []})')
if args.seq:
    t0 = last[0].time_range.start
    for e in last:
        print(f'{(e.time_range.start - t0):9.1f} {short(e.name):58s} {e.time_range.elapsed_us():8.1f} us')
agg = collections.defaultdict(lambda: [0, 0.0])
for e in last:
    agg[short(e.name)][0] += 1
    agg[short(e.name)][1] += e.time_range.elapsed_us()
tot = sum(v[1] for v in agg.values())
print(f'busy {tot:.1f} us')
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f'{k:58s} n={v[0]:4d} {v[1]:9.1f} us {100 * v[1] / tot:5.1f}%')
