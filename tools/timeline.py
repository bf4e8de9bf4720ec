"""Real (warm-cache, back-to-back) per-kernel GPU durations of one frame of the inference loop, from torch.profiler
(CUPTI activity records; no replay, no cache flush — unlike ncu).  usage: python tools/timeline.py [--seq] [--train]"""
import argparse, collections, sys
sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from hfa_gp_b200.networks.headnerf import HeadNeRF_final
from hfa_gp_b200 import cam_utils

ap = argparse.ArgumentParser()
ap.add_argument('--seq', action='store_true')
ap.add_argument('--frames', type=int, default=3)
ap.add_argument('--graph', action='store_true', help='profile CUDA-graph replays of the frame loop (the product path)')
ap.add_argument('--batch', type=int, default=1, help='frames per replay (bench.py --frames-per-step)')
args = ap.parse_args()
dev = torch.device('cuda')
ns = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='', synthetic_generator=True, generator_seed=0)
torch.manual_seed(0)
model = HeadNeRF_final(ns, 256, dev, 512, 50, 'bench', './').to(dev).eval().requires_grad_(False)
img = torch.rand(args.batch, 3, 256, 256, device=dev) * 2 - 1
lab = cam_utils.cam_sampler(args.batch, 'cpu').to(dev)

def frame():
    with torch.no_grad():
        return model.get_image(model.get_latent(model.get_weights(img)), lab.clone())
if args.graph:
    from hfa_gp_b200.frame_loop import FrameLoop
    loop = FrameLoop(model, batch=args.batch, size=256, device=dev)
    def frame():
        return loop(img, lab, mutate_label=False)
for _ in range(4):
    frame()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.frames):
        frame()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
def short(n):
    n = n.replace('hfagp::', '').replace('void ', '')
    return (n[:n.index('(')] if '(' in n else n)[:52]
per = len(evs) // args.frames
last = evs[-per:]
if args.graph:
    # gaps on the critical path: time not covered by any kernel between the first and the last kernel of the frame
    iv = sorted((e.time_range.start, e.time_range.end) for e in last)
    covered, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    for s_, e_ in iv[1:]:
        if s_ > cur_e:
            covered += cur_e - cur_s
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    covered += cur_e - cur_s
    print(f'graph frame: span {iv[-1][1] - iv[0][0]:.1f} us, covered by >= 1 kernel {covered:.1f} us, idle gaps {iv[-1][1] - iv[0][0] - covered:.1f} us')
if args.seq:
    t0 = last[0].time_range.start
    for e in last:
        print(f'{(e.time_range.start - t0):9.1f} {short(e.name):54s} {e.time_range.elapsed_us():8.1f} us')
agg = collections.defaultdict(lambda: [0, 0.0])
for e in last:
    agg[short(e.name)][0] += 1
    agg[short(e.name)][1] += e.time_range.elapsed_us()
tot = sum(v[1] for v in agg.values())
span = last[-1].time_range.end - last[0].time_range.start
print(f'--- last frame: {per} kernels, busy {tot:.1f} us, span {span:.1f} us')
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
    print(f'{k:54s} n={v[0]:4d} {v[1]:9.1f} us {100 * v[1] / tot:5.1f}%')
