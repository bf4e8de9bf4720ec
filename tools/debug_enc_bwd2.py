import sys, torch, math
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch.nn.functional as F
import parity_utils as pu
from oracle import hfagp_ref
from hfa_gp_b200 import ops
from hfa_gp_b200.networks.encoder3d import ConvLayer
torch.manual_seed(0)
for (c, h, b) in [(512, 16, 2), (512, 8, 2), (64, 16, 1), (64, 16, 2), (64, 32, 2)]:
    layer = ConvLayer(c, c, 3, downsample=True).cuda()
    with torch.no_grad():
        layer[2].bias.normal_(0, 0.2)
    x = torch.randn(b, c, h, h)
    gy = torch.randn(b, c, h // 2, h // 2)
    sd = {'p.0.kernel': hfagp_ref.blur_kernel(), 'p.1.weight': layer[1].weight.detach().cpu().double(), 'p.2.bias': layer[2].bias.detach().cpu().double()}
    xr = x.double().requires_grad_(True)
    yr = hfagp_ref.conv_layer_ref({k: v.double() for k, v in sd.items()}, 'p', xr, 3, downsample=True)
    (yr * gy.double()).sum().backward()
    rec, grads = {}, {}
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    y = layer.run(xn, rec=rec)
    dx = layer.backward(rec, gy.permute(0, 2, 3, 1).contiguous().cuda(), grads)
    print(c, h, b, 'fwd %.2e dx max %.2e l2 %.2e' % (pu.rel_err(pu.to_nchw(y), yr), pu.rel_err(pu.to_nchw(dx), xr.grad), pu.rel_l2(pu.to_nchw(dx), xr.grad)))
    # pieces
    dz = ops.act_bwd(rec['y'], g0=gy.permute(0, 2, 3, 1).contiguous().cuda(), act=1, act_gain=math.sqrt(2.0), out='f32')
    wt = layer[1].packed_T(False)
    t = ops.conv_transpose_s2(dz, wt, c, 0)
    # torch reference for t: conv_transpose2d(dz, W*scale, stride=2)
    w = (layer[1].weight.detach() * layer[1].scale).double().cpu()
    tr = F.conv_transpose2d(pu.to_nchw(dz).double().cpu(), w, stride=2)
    print('   t shape', tuple(t.shape), 'err %.2e' % pu.rel_err(pu.to_nchw(t), tr))
    e = (pu.to_nchw(t).double().cpu() - tr).abs().amax(dim=(0, 1))
    print('   err map rows', [('%.1e' % v) for v in e.amax(dim=1).tolist()])
