import sys, torch, math
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch.nn.functional as F
import parity_utils as pu
from oracle import hfagp_ref
from hfa_gp_b200 import ops
from hfa_gp_b200.networks.encoder3d import Encoder

_ct, _blur, _conv, _bu, _act = ops.conv_transpose_s2, ops.blur, ops.conv2d, ops.blur_up, ops.act_bwd
K1 = torch.tensor([1., 3., 3., 1.], dtype=torch.float64, device='cuda') / 8

def ct(x, w, cout, wbs, out=None):
    t = _ct(x, w, cout, wbs, out)
    wd = w.double().view(3, 3, cout, -1).permute(3, 2, 0, 1)     # [cin_of_x, cout, 3,3]: conv_transpose weight layout (in, out, kh, kw)
    tr = F.conv_transpose2d(x.double().permute(0, 3, 1, 2), wd, stride=2)
    print('  conv_transpose_s2', tuple(x.shape), 'err %.2e' % pu.rel_err(pu.to_nchw(t), tr))
    return t

def blur(x, pad0, pad1, stride=1, split_out=False, gain=1.0):
    y = _blur(x, pad0, pad1, stride, split_out, gain)
    if not isinstance(x, ops.Split) and not split_out:
        c = x.shape[-1]
        k = torch.outer(K1, K1)[None, None].repeat(c, 1, 1, 1) * gain
        xr = F.pad(x.double().permute(0, 3, 1, 2), [pad0, pad1, pad0, pad1])
        yr = F.conv2d(xr, k, groups=c, stride=stride)
        print('  blur', tuple(x.shape), pad0, pad1, stride, 'err %.2e' % pu.rel_err(pu.to_nchw(y), yr))
    return y

def act(y, **kw):
    dz = _act(y, **kw)
    g = kw['g0'].double()
    if kw.get('g1') is not None:
        g = g + kw['g1'].double()
    yy = y.double()
    rs = kw.get('residual_scale', 1.0)
    if kw.get('residual') is not None:
        av = yy / rs - kw['residual'].double()
    else:
        av = yy
    slope = torch.where(av > 0, 1.0, 0.2) * kw.get('act_gain', math.sqrt(2.0)) if kw.get('act', 1) == 1 else torch.ones_like(av) * kw.get('act_gain', 1.0)
    ref = g * kw.get('post_scale', 1.0) * slope
    print('  act_bwd', tuple(y.shape), 'err %.2e' % pu.rel_err(dz, ref), 'min|av| %.2e' % float(av.abs().min()))
    return dz

ops.conv_transpose_s2, ops.blur, ops.act_bwd = ct, blur, act
size, b = 32, 2
sd = hfagp_ref.make_encoder_state(size=size, dim_motion=10, seed=0)
g = torch.Generator().manual_seed(5)
for k in sd:
    if k.endswith('.bias'):
        sd[k] = torch.randn(sd[k].shape, generator=g) * 0.2
g = torch.Generator().manual_seed(3)
x = torch.rand(b, 3, size, size, generator=g) * 2 - 1
gout = torch.randn(b, 10, generator=g)
enc = Encoder(size, 512, 10); enc.load_state_dict(sd); enc = enc.cuda(); enc.net_app.precision = 'fp32'
out = enc(x.cuda())
print('--- backward')
(out * gout.cuda()).sum().backward()
