"""Time the tcgen05 convolution on the generator's big layer shapes in isolation (CUDA events, L2-flushing rotation of
inputs), and report achieved TFLOP/s (bf16-equivalent: 3 MMAs per K step) and operand GB/s.
usage: python tools/prof_conv.py [--reps 20] [--only NAME]"""
import argparse, sys, math
sys.path.insert(0, '.')
import torch
from hfa_gp_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument('--reps', type=int, default=20)
ap.add_argument('--only', default='')
args = ap.parse_args()
dev = 'cuda'
SHAPES = [  # name, res, cin, cout, kind
    ('sr1.conv1 512^2 128->128 3x3', 512, 128, 128, 'conv3'),
    ('sr0.conv1 256^2 256->256 3x3', 256, 256, 256, 'conv3'),
    ('sr1.conv0 256^2->512^2 256->128 up', 256, 256, 128, 'up'),
    ('sr0.conv0 128^2->256^2 32->256 up', 128, 32, 256, 'up'),
    ('b256.conv1 256^2 128->128 3x3', 256, 128, 128, 'conv3'),
    ('b128.conv1 128^2 256->256 3x3', 128, 256, 256, 'conv3'),
    ('b64.conv1 64^2 512->512 3x3', 64, 512, 512, 'conv3'),
    ('b32.conv1 32^2 512->512 3x3', 32, 512, 512, 'conv3'),
    ('b16.conv1 16^2 512->512 3x3', 16, 512, 512, 'conv3'),
    ('b256.torgb 256^2 128->96 1x1', 256, 128, 96, 'conv1'),
]
for name, res, cin, cout, kind in SHAPES:
    if args.only and args.only not in name:
        continue
    nbuf = 3
    xs = [ops.split(torch.randn(1, res, res, cin, device=dev)) for _ in range(nbuf)]
    taps = 9 if kind != 'conv1' else 1
    w = ops.split(torch.randn(1, taps, cout, cin, device=dev))
    bias = torch.randn(cout, device=dev)
    def run(i):
        x = xs[i % nbuf]
        if kind == 'conv3':
            return ops.conv2d_tc(x, w, ops.TAPS_3X3, cout, oh=res, ow=res, w_batched=True, split_out=True, bias=bias, act=1, act_gain=math.sqrt(2))
        if kind == 'conv1':
            return ops.conv2d_tc(x, w, ops.TAPS_1X1, cout, oh=res, ow=res, w_batched=True, bias=bias)
        return ops.conv_transpose_s2_tc(x, w, cout, w_batched=True)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / args.reps * 1e3
    flop = 2.0 * res * res * cin * cout * taps
    print(f'{name:40s} {us:8.1f} us   {flop / us / 1e6:7.1f} TFLOP/s fp32-equiv   {3 * flop / us / 1e6:7.1f} TFLOP/s bf16 (x3)')
