"""VERDICT r1 item 9 — measure the 3x MMA tax instead of assuming it (CPU emulation, no GPU needed).

Every tensor-core convolution of the generator computes x (*) w with operands cut to what a given MMA scheme can carry;
accumulation stays fp32 (TMEM accumulators are fp32).  The oracle generator (oracle/eg3d_ref.py, full 512^2 config) is run
with torch.nn.functional.conv2d / conv_transpose2d wrapped so that the activations and the style-modulated weights are
rounded exactly as each scheme would see them, and the final image is compared with the un-rounded fp32 oracle under the
parity tests' own criterion  max |a-b| / max(|b|, rms(b))  (tests/parity_utils.py) plus the rms of that ratio.

schemes (cost in bf16-MMA equivalents per algorithmic FMA):
  bf16x3   x_hi*w_hi + x_lo*w_hi + x_hi*w_lo   (shipped; 3)          == x*w - x_lo*w_lo
  bf16x2a  (x_hi + x_lo) * w_hi                 (2)                   weights cut to bf16
  bf16x2b  x_hi * (w_hi + w_lo)                 (2)                   activations cut to bf16
  tf32rn   rn_tf32(x) * rn_tf32(w)              (kind::tf32, 2)       operands rounded to 10 mantissa bits by the producer
  tf32tr   trunc_tf32(x) * trunc_tf32(w)        (kind::tf32, 2)       raw fp32 fed to the MMA (hardware drops 13 bits)
  tf32x2   (x_hi + x_lo)*w_hi in tf32 pieces    (4)                   listed for completeness: costlier than bf16x3
  fp16x2   (x_hi + x_lo) * w_hi in fp16         (kind::f16, 2)        11-bit weights; range-limited
usage: python tools/precision_study.py [--small]   (writes a table to stdout; kept in profiles/r2_precision_study.md)
"""
import sys, math
sys.path.insert(0, '.')
import torch
import torch.nn.functional as F
from oracle import eg3d_ref as E, hfagp_ref as H

torch.set_num_threads(8)
_conv2d, _convT = F.conv2d, F.conv_transpose2d


def bf16_hi(t):
    return t.bfloat16().float()


def tf32_rn(t):
    i = t.view(torch.int32)
    r = ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF)
    return r.view(torch.float32)


def tf32_tr(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def fp16_hi(t):
    return t.half().float()


SCHEMES = {
    'bf16x3': lambda conv, x, w: conv(x, w) - conv(x - bf16_hi(x), w - bf16_hi(w)),
    'bf16x2a': lambda conv, x, w: conv(bf16_hi(x) + bf16_hi(x - bf16_hi(x)), bf16_hi(w)),
    'bf16x2b': lambda conv, x, w: conv(bf16_hi(x), bf16_hi(w) + bf16_hi(w - bf16_hi(w))),
    'tf32rn': lambda conv, x, w: conv(tf32_rn(x), tf32_rn(w)),
    'tf32tr': lambda conv, x, w: conv(tf32_tr(x), tf32_tr(w)),
    'tf32x2': lambda conv, x, w: conv(tf32_rn(x) + tf32_rn(x - tf32_rn(x)), tf32_rn(w)),
    'fp16x2': lambda conv, x, w: conv(fp16_hi(x) + fp16_hi(x - fp16_hi(x)), fp16_hi(w)),
}
COST = {'bf16x3': 3, 'bf16x2a': 2, 'bf16x2b': 2, 'tf32rn': 2, 'tf32tr': 2, 'tf32x2': 4, 'fp16x2': 2}
mode = [None]


def is_fir(inp, w, groups):
    return w.shape[1] == 1 and groups == inp.shape[1] and groups > 1 and w.shape[0] == groups


def conv2d(inp, w, bias=None, stride=1, padding=0, dilation=1, groups=1):
    if mode[0] is None or is_fir(inp, w, groups) or inp.shape[1] // groups % 8:
        return _conv2d(inp, w, bias, stride, padding, dilation, groups)
    y = SCHEMES[mode[0]](lambda a, b: _conv2d(a, b, None, stride, padding, dilation, groups), inp, w)
    return y if bias is None else y + bias.view(1, -1, 1, 1)


def convT(inp, w, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    if mode[0] is None or w.shape[0] // 1 % 8:
        return _convT(inp, w, bias, stride, padding, output_padding, groups, dilation)
    return SCHEMES[mode[0]](lambda a, b: _convT(a, b, None, stride, padding, output_padding, groups, dilation), inp, w)


F.conv2d, F.conv_transpose2d = conv2d, convT
E.F.conv2d, E.F.conv_transpose2d = conv2d, convT

small = '--small' in sys.argv
cfg = E.small14_config() if small else E.GeneratorConfig()
gen = E.make_generator(cfg, seed=0, noise_strength=0.1)
g = torch.Generator().manual_seed(0)
ws = torch.randn(1, cfg.num_ws, cfg.w_dim, generator=g)
c = H.flip_label_(H.synthetic_labels(1, seed=0))
jit = torch.rand(1, cfg.nrr ** 2, cfg.depth_res, 1, generator=g)
u = torch.rand(cfg.nrr ** 2, cfg.depth_res_importance, generator=g)


def run():
    tap = {}
    with torch.no_grad():
        out = gen.synthesis(ws, c, jitter_coarse=jit, u_fine=u, tap=tap)
    return out['image'], tap.get('planes'), tap.get('feature_image')


def crit(a, b):
    den = torch.maximum(b.abs(), b.square().mean().sqrt())
    r = (a - b).abs() / den
    return float(r.max()), float(r.square().mean().sqrt())


ref = run()
print(f'config: {"small14" if small else "full 512^2 (configs[1])"}; criterion max|a-b|/max(|b|,rms(b)); tolerance 1e-3')
print(f'| scheme | MMAs/FMA | image max | image rms | planes max | feature-image max | holds 1e-3 |')
print(f'|---|---|---|---|---|---|---|')
for name in SCHEMES:
    mode[0] = name
    got = run()
    mode[0] = None
    im = crit(got[0], ref[0])
    pl = crit(got[1], ref[1]) if ref[1] is not None else (float('nan'),) * 2
    ft = crit(got[2], ref[2]) if ref[2] is not None else (float('nan'),) * 2
    print(f'| {name} | {COST[name]} | {im[0]:.2e} | {im[1]:.2e} | {pl[0]:.2e} | {ft[0]:.2e} | {"yes" if im[0] < 1e-3 else "NO"} |', flush=True)
