import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from oracle import eg3d_ref
import parity_utils as pu
from hfa_gp_b200 import autograd as ag

precision = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
cfg = eg3d_ref.small14_config()
ref, prod = pu.make_pair(cfg, seed=0)
prod.precision = precision
g = torch.Generator().manual_seed(4)
b = 2
ws = torch.randn(b, cfg.num_ws, cfg.w_dim, generator=g)
ws_r = ws.clone().requires_grad_(True)
saved = {}
def hook(name):
    def f(mod, inp, out):
        out.retain_grad(); saved[name] = out
    return f
for name, m in ref.backbone.synthesis.named_modules():
    if name.endswith('affine'):
        m.register_forward_hook(hook(name))
planes_r = ref.backbone.synthesis(ws_r, noise_mode='const')
gp = torch.randn(planes_r.shape, generator=g)
(planes_r * gp).sum().backward()

ws_g = ws.clone().cuda().requires_grad_(True)
flat = ag.StylesFn.apply(ws_g, prod)
flat.retain_grad()
planes = ag.BackboneFn.apply(flat, prod, 'const', b, None)
(planes * gp.permute(0, 2, 3, 1).contiguous().cuda()).sum().backward()
pk = prod._ensure_packed()
views = ag._dviews(pk, flat.grad, b)
names = {id(m): n for n, m in prod.backbone.synthesis.named_modules()}
for i, (kind, m, widx) in enumerate(pk['order']):
    n = names.get(id(m))
    if n is None:
        continue
    r = saved[n + '.affine'].grad
    if kind == 'torgb':
        r = r / (1.0 / (m.cin ** 0.5))    # oracle hook is before the weight_gain multiply; product styles include it
    print(f'{n:14s} {kind:6s} relerr {pu.rel_err(views[i], r):.3e}  |ref| {float(r.abs().mean()):.3e}')
print('dws', pu.rel_err(ws_g.grad, ws_r.grad))
