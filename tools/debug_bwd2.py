import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from oracle import eg3d_ref
import parity_utils as pu
from hfa_gp_b200 import autograd as ag, ops

cfg = eg3d_ref.small14_config()
ref, prod = pu.make_pair(cfg, seed=0)
g = torch.Generator().manual_seed(4)
b = 2
ws = torch.randn(b, cfg.num_ws, cfg.w_dim, generator=g)
gp = torch.randn(b, 96, cfg.plane_res, cfg.plane_res, generator=g)
logs = {}
orig_act, orig_dgrad, orig_demod = ops.act_bwd, ag._dgrad, ops.demod_bwd
def run(precision):
    log = []
    def act(*a, **k):
        dd = k.get('ddcoef')
        r = orig_act(*a, **k)
        log.append(('act_dz', None if r is None else (r.float() if isinstance(r, ops.Split) else r).clone()))
        if dd is not None: log.append(('ddc', dd.clone()))
        if k.get('ds0') is not None: log.append(('ds0', k['ds0'].clone()))
        return r
    def dg(gen, dz, pl, taps, **k):
        r = orig_dgrad(gen, dz, pl, taps, **k)
        log.append(('dxu', r.clone()))
        return r
    def dm(w2, styles, dcoef, ddcoef, dstyles):
        orig_demod(w2, styles, dcoef, ddcoef, dstyles)
        log.append(('dsty_after_demod', dstyles.clone()))
    ops.act_bwd, ag._dgrad, ops.demod_bwd = act, dg, dm
    prod.precision = precision
    ws_g = ws.clone().cuda().requires_grad_(True)
    planes = ag.BackboneFn.apply(ag.StylesFn.apply(ws_g, prod), prod, 'const', b, None)
    (planes * gp.permute(0, 2, 3, 1).contiguous().cuda()).sum().backward()
    return log
la, lb = run('fp32'), run('tc')
for (na, ta), (nb, tb) in zip(la, lb):
    if ta is None: continue
    print(f'{na:18s} shape {tuple(ta.shape)} relerr {pu.rel_err(tb, ta):.3e}')
