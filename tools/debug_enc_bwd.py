import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import parity_utils as pu
from oracle import hfagp_ref
from hfa_gp_b200.networks.encoder3d import Encoder
size, b = 32, 2
sd = hfagp_ref.make_encoder_state(size=size, dim_motion=10, seed=0)
g = torch.Generator().manual_seed(5)
for k in sd:
    if k.endswith('.bias'):
        sd[k] = torch.randn(sd[k].shape, generator=g) * 0.2
g = torch.Generator().manual_seed(3)
x = torch.rand(b, 3, size, size, generator=g) * 2 - 1
gout = torch.randn(b, 10, generator=g)
s64 = {k: v.double().clone().requires_grad_(not k.endswith('.kernel')) for k, v in sd.items()}
o64 = hfagp_ref.encoder_ref(s64, x.double())
(o64 * gout.double()).sum().backward()
for prec in ('fp32', 'tc'):
    enc = Encoder(size, 512, 10); enc.load_state_dict(sd); enc = enc.cuda(); enc.net_app.precision = prec
    out = enc(x.cuda())
    print(prec, 'fwd', pu.rel_err(out, o64))
    (out * gout.cuda()).sum().backward()
    for n, p in enc.named_parameters():
        print(prec, n, 'max %.3e l2 %.3e' % (pu.rel_err(p.grad, s64[n].grad), pu.rel_l2(p.grad, s64[n].grad)))
