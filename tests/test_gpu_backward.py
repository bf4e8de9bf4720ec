"""GPU gradient-parity tests (``-m gpu``): the backward kernels, called through the C ABI, against PyTorch
autograd of the CPU oracle on the same seeded inputs.  Tolerance: 1e-3 relative (parity_utils.rel_err)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import eg3d_ref, hfagp_ref
import parity_utils as pu

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-3          # relative L2, exact-fp32 kernels (precision='fp32'); per element only up to 2e-2: one
                         # activation within 1e-6 of the leaky-ReLU kink takes the other branch than on the CPU and
                         # moves individual near-cancelling gradient elements by ~1e-3 (see tests/test_gpu_training.py)
GRAD_TOL_TC_L2 = 5e-3    # relative L2, tensor-core path (see parity_utils.rel_l2)


def _check_grad(got, want, precision, what):
    e_max, e_l2 = pu.rel_err(got, want), pu.rel_l2(got, want)
    print(f'{what} [{precision}]: max-rel {e_max:.3e}  rel-L2 {e_l2:.3e}')
    if precision == 'fp32':
        assert e_l2 < GRAD_TOL and e_max < 2e-2, (what, e_max, e_l2)
    else:
        assert e_l2 < GRAD_TOL_TC_L2, (what, e_l2)


def _ops():
    from hfa_gp_b200 import ops
    return ops


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().cuda()


# ------------------------------------------------------------------ unit kernels

def test_act_bwd_matches_autograd():
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    n, c, h = 2, 24, 9
    z = torch.randn(n, c, h, h, generator=g, requires_grad=True)
    dco = (torch.rand(n, c, generator=g) + 0.5).requires_grad_(True)
    noise = torch.randn(h, h, generator=g)
    bias = torch.randn(c, generator=g, requires_grad=True)
    s0 = torch.rand(n, c, generator=g) + 0.5
    s1 = torch.rand(n, c, generator=g) + 0.5
    g0 = torch.randn(n, c, h, h, generator=g)
    g1 = torch.randn(n, c, h, h, generator=g)
    pre = z * dco[:, :, None, None] + noise * 0.3 + bias[None, :, None, None]
    y = (F.leaky_relu(pre, 0.2) * 1.3).clamp(-1.0, 1.0)
    gtot = g0 * s0[:, :, None, None] + g1 * s1[:, :, None, None]
    (y * gtot).sum().backward()
    yc = y.detach()
    ds0 = torch.zeros(n, c).cuda(); ds1 = torch.zeros(n, c).cuda(); db = torch.zeros(c).cuda(); ddc = torch.zeros(n, c).cuda()
    for form in ('f32', 'split'):
        ds0.zero_(); ds1.zero_(); db.zero_(); ddc.zero_()
        yin = nhwc(yc) if form == 'f32' else ops.split(nhwc(yc))
        dz = ops.act_bwd(yin, g0=nhwc(g0), s0=s0.cuda(), g1=nhwc(g1), s1=s1.cuda(), dcoef=dco.detach().cuda(),
                         noise=noise.cuda(), noise_gain=0.3, bias=bias.detach().cuda(), act=1, act_gain=1.3, clamp=1.0,
                         out=form, ds0=ds0, ds1=ds1, dbias=db, ddcoef=ddc)
        dz = dz.float() if form == 'split' else dz
        assert pu.rel_err(pu.to_nchw(dz), z.grad) < 1e-4
        assert pu.rel_err(db, bias.grad) < 1e-4
        assert pu.rel_err(ddc, dco.grad) < 1e-3
        assert pu.rel_err(ds0, (g0 * yc).sum(dim=[2, 3])) < 1e-4
        assert pu.rel_err(ds1, (g1 * yc).sum(dim=[2, 3])) < 1e-4


def test_act_bwd_residual_merge_and_small_torgb():
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    n, c, h = 1, 16, 8
    pre = torch.randn(n, c, h, h, generator=g, requires_grad=True)
    skip = torch.randn(n, c, h, h, generator=g)
    y = (F.leaky_relu(pre, 0.2) * math.sqrt(2) + skip) / math.sqrt(2)
    dimg = torch.randn(n, 3, h, h, generator=g)
    wrgb = torch.randn(3, c, generator=g)
    srgb = torch.rand(n, c, generator=g) + 0.5
    gy = torch.einsum('nohw,oc->nchw', dimg, wrgb) * srgb[:, :, None, None]
    (y * gy).sum().backward()
    dsr = torch.zeros(n, c).cuda()
    dz = ops.act_bwd(nhwc(y.detach()), dimg=nhwc(dimg), wrgb=wrgb.cuda(), srgb=srgb.cuda(), residual=nhwc(skip),
                     residual_scale=1 / math.sqrt(2), post_scale=1 / math.sqrt(2), act=1, act_gain=math.sqrt(2),
                     out='f32', dsrgb=dsr)
    assert pu.rel_err(pu.to_nchw(dz), pre.grad) < 1e-4
    ref = (torch.einsum('nohw,oc->nchw', dimg, wrgb) * y.detach()).sum(dim=[2, 3])
    assert pu.rel_err(dsr, ref) < 1e-4


@pytest.mark.parametrize('c', [8, 3])
def test_blur_transposes(c):
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    k = hfagp_ref.blur_kernel()
    # stride-2 skip-path blur: transpose via blur_up
    x = torch.randn(2, c, 12, 12, generator=g, requires_grad=True)
    y = hfagp_ref.blur_ref(x, k, 1, 1)[:, :, ::2, ::2]
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    dx = ops.blur_up(nhwc(gy), 12, 12, 1, 1, 2)
    assert pu.rel_err(pu.to_nchw(dx), x.grad) < 1e-5
    # stride-1 blur (pad 2,2): its transpose is the same FIR with pads (1,1)
    x2 = torch.randn(1, c, 9, 9, generator=g, requires_grad=True)
    y2 = hfagp_ref.blur_ref(x2, k, 2, 2)
    gy2 = torch.randn(y2.shape, generator=g)
    (y2 * gy2).sum().backward()
    assert pu.rel_err(pu.to_nchw(ops.blur(nhwc(gy2), 1, 1)), x2.grad) < 1e-5
    # upsample2d transpose = stride-2 FIR with pads (1,1), gain 4
    lo = torch.randn(1, c, 6, 6, generator=g, requires_grad=True)
    up = eg3d_ref.upsample2d_ref(lo, eg3d_ref.setup_filter())
    gu = torch.randn(up.shape, generator=g)
    (up * gu).sum().backward()
    assert pu.rel_err(pu.to_nchw(ops.blur(nhwc(gu), 1, 1, stride=2, gain=4.0)), lo.grad) < 1e-5


def test_linear_and_wgrad():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 40, generator=g, requires_grad=True)
    w = torch.randn(24, 40, generator=g, requires_grad=True)
    b = torch.randn(24, generator=g, requires_grad=True)
    y = hfagp_ref.equal_linear_ref(x, w, b)
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    dw = torch.zeros(24, 40).cuda(); db = torch.zeros(24).cuda()
    dx = ops.linear_bwd(gy.cuda(), x.detach().cuda(), w.detach().cuda(), 1 / math.sqrt(40), 1.0, dw=dw, db=db)
    assert pu.rel_err(dx, x.grad) < 1e-5 and pu.rel_err(dw, w.grad) < 1e-5 and pu.rel_err(db, b.grad) < 1e-5
    # conv weight gradient: 3x3 stride 1 pad 1, and 3x3 stride 2 pad 0 (encoder down-conv)
    for stride, pad, hh in ((1, 1, 10), (2, 0, 11)):
        xi = torch.randn(2, 16, hh, hh, generator=g)
        wt = torch.randn(24, 16, 3, 3, generator=g, requires_grad=True)
        yo = F.conv2d(xi, wt * 0.5, stride=stride, padding=pad)
        gz = torch.randn(yo.shape, generator=g)
        (yo * gz).sum().backward()
        taps = tuple((ky - pad, kx - pad, ky * 3 + kx) for ky in range(3) for kx in range(3))
        dwp = torch.zeros(9, 24, 16).cuda()
        ops.conv2d_wgrad(nhwc(xi), nhwc(gz), taps, dwp, oh=yo.shape[2], ow=yo.shape[3], in_stride=stride, scale=0.5)
        got = dwp.view(3, 3, 24, 16).permute(2, 3, 0, 1)
        assert pu.rel_err(got, wt.grad) < 1e-4
        dwp.zero_()
        ops.conv2d_wgrad(ops.split(nhwc(xi)), ops.split(nhwc(gz)), taps, dwp, oh=yo.shape[2], ow=yo.shape[3],
                         in_stride=stride, scale=0.5)
        assert pu.rel_err(dwp.view(3, 3, 24, 16).permute(2, 3, 0, 1), wt.grad) < 1e-4


# ------------------------------------------------------------------ generator stages

def _pair(cfg, precision):
    ref, prod = pu.make_pair(cfg, seed=0)
    prod.precision = precision
    return ref, prod


@pytest.mark.parametrize('precision', ['tc', 'fp32'])
def test_backbone_backward(precision):
    from hfa_gp_b200 import autograd as ag
    cfg = eg3d_ref.small14_config()
    ref, prod = _pair(cfg, precision)
    g = torch.Generator().manual_seed(4)
    b = 2
    ws = torch.randn(b, cfg.num_ws, cfg.w_dim, generator=g)
    ws_r = ws.clone().requires_grad_(True)
    planes_r = ref.backbone.synthesis(ws_r, noise_mode='const')
    gp = torch.randn(planes_r.shape, generator=g)
    (planes_r * gp).sum().backward()
    ws_g = ws.clone().cuda().requires_grad_(True)
    planes = ag.BackboneFn.apply(ag.StylesFn.apply(ws_g, prod), prod, 'const', b, None)
    assert pu.rel_err(pu.to_nchw(planes), planes_r) < pu.REL_TOL
    (planes * nhwc(gp)).sum().backward()
    _check_grad(ws_g.grad, ws_r.grad, precision, 'backbone dws')


@pytest.mark.parametrize('precision', ['tc', 'fp32'])
def test_superres_backward(precision):
    from hfa_gp_b200 import autograd as ag
    cfg = eg3d_ref.small14_config()
    ref, prod = _pair(cfg, precision)
    g = torch.Generator().manual_seed(5)
    b = 2
    ws = torch.randn(b, cfg.num_ws, cfg.w_dim, generator=g)
    feat = torch.randn(b, 32, cfg.nrr, cfg.nrr, generator=g) * 0.5
    ws_r, feat_r = ws.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    img_r = ref.superresolution(feat_r[:, :3], feat_r, ws_r)
    gi = torch.randn(img_r.shape, generator=g)
    (img_r * gi).sum().backward()
    ws_g = ws.clone().cuda().requires_grad_(True)
    feat_g = nhwc(feat).requires_grad_(True)
    img = ag.SuperresFn.apply(feat_g, ag.StylesFn.apply(ws_g, prod), prod, b, None)
    assert pu.rel_err(pu.to_nchw(img), img_r) < pu.REL_TOL
    (img * nhwc(gi)).sum().backward()
    _check_grad(pu.to_nchw(feat_g.grad), feat_r.grad, precision, 'superres dfeat')
    _check_grad(ws_g.grad, ws_r.grad, precision, 'superres dws')


# ------------------------------------------------------------------ renderer backward

@pytest.mark.parametrize('res,s,sf,batch', [(8, 16, 0, 1), (8, 16, 16, 2), (8, 20, 9, 1), (16, 48, 48, 1)])
def test_render_backward(res, s, sf, batch):
    """d(feat) -> d(planes) through composite, march, decoder MLP and bilinear gather vs autograd of the oracle."""
    ops = _ops()
    plane_res = 32
    cfg = eg3d_ref.GeneratorConfig(nrr=res, depth_res=s, depth_res_importance=sf, plane_res=plane_res)
    g = torch.Generator().manual_seed(res + s)
    planes = torch.randn(batch, 3, 32, plane_res, plane_res, generator=g).requires_grad_(True)
    with torch.random.fork_rng():
        torch.manual_seed(1)
        dec = eg3d_ref.OSGDecoderRef(cfg)
        with torch.no_grad():
            dec.net[0].bias.normal_(0, 0.5)
            dec.net[2].bias.normal_(0, 0.5)
    dec.requires_grad_(False)
    c = hfagp_ref.flip_label_(hfagp_ref.synthetic_labels(batch, seed=3))
    rays = res * res
    jitter = torch.rand(batch, rays, s, 1, generator=g)
    u = torch.rand(batch * rays, max(sf, 1), generator=g)
    ro, rd = eg3d_ref.ray_sampler_ref(c[:, :16].view(-1, 4, 4), c[:, 16:25].view(-1, 3, 3), res)
    feat_ref, _, _ = eg3d_ref.render_ref(planes, dec, ro, rd, cfg, jitter, u if sf > 0 else None)
    gf = torch.randn(feat_ref.shape, generator=g)
    (feat_ref * gf).sum().backward()
    d0, d2 = dec.net[0], dec.net[2]
    mlp = torch.cat([(d0.weight * d0.weight_gain).reshape(-1), d0.bias * d0.bias_gain,
                     (d2.weight * d2.weight_gain).reshape(-1), d2.bias * d2.bias_gain]).detach().cuda()
    pl = planes.detach().reshape(batch, 96, plane_res, plane_res).permute(0, 2, 3, 1).contiguous().cuda()
    lin = torch.linspace(cfg.ray_start, cfg.ray_end, s)
    delta = (cfg.ray_end - cfg.ray_start) / (s - 1)
    dpl = ops.render_bwd(pl, c.cuda(), mlp, lin.cuda(), jitter.reshape(batch, -1, s).contiguous().cuda(),
                         u.cuda() if sf > 0 else None, gf.reshape(batch, res, res, 32).contiguous().cuda(),
                         res=res, s_coarse=s, s_fine=sf, delta=delta, box_scale=2.0 / cfg.box_warp)
    want = planes.grad.reshape(batch, 96, plane_res, plane_res)
    e_max, e_l2 = pu.rel_err(pu.to_nchw(dpl), want), pu.rel_l2(pu.to_nchw(dpl), want)
    print(f'render dplanes: max-rel {e_max:.3e} rel-L2 {e_l2:.3e}')
    assert e_l2 < 1e-3 and e_max < 5e-3, (e_max, e_l2)


@pytest.mark.parametrize('n,h,cin,split', [(2, 64, 128, True), (1, 37, 256, True), (3, 16, 24, False), (2, 48, 512, True)])
def test_wgrad_narrow_1x1_output(n, h, cin, split):
    """The 3-channel ToRGB's weight gradient (cout padded to 4, 1x1, per-sample style factors) takes the streaming
    wgrad_cout4_kernel: against fp64 einsum, fp32 and split-bf16 layer inputs, channel counts beyond one 32-quad sweep."""
    ops = _ops()
    g = torch.Generator().manual_seed(n * 100 + h + cin)
    x = torch.randn(n, h, h, cin, generator=g)
    dz = torch.randn(n, h, h, 4, generator=g)
    st = torch.randn(n, cin, generator=g)
    want = torch.einsum('nhwo,nhwi,ni->oi', dz.double(), x.double(), st.double()).float() * 0.5
    dw = torch.zeros(1, 4, cin, device='cuda')
    xin = ops.split(x.cuda()) if split else x.cuda()
    ops.conv2d_wgrad(xin, dz.cuda(), ops.TAPS_1X1, dw, oh=h, ow=h, scale=0.5, xscale=st.cuda())
    e = pu.rel_l2(dw[0], want)
    print(f'wgrad 1x1 cout 4 [{n},{h},{h},{cin}] split={split}: rel-L2 {e:.3e}')
    assert e < (2e-5 if split else 1e-5)


@pytest.mark.parametrize('taps,cout,cin,batch,transposed', [(9, 64, 96, 2, False), (9, 40, 24, 3, True), (1, 512, 512, 9, False),
                                                            (9, 128, 256, 2, True)])
def test_modconv_wgrad_finish(taps, cout, cin, batch, transposed):
    """hfagp_modconv_wgrad_finish: unpack(dw) - w * sum_n ddcoef dcoef^3 styles^2, added to grad[O][I][k][k]; packed and
    role-swapped ([tap][I][O]) inputs, ragged 32 x 32 tiles, more samples than one staging round."""
    ops = _ops()
    g = torch.Generator().manual_seed(taps * 1000 + cout + cin)
    k = int(math.isqrt(taps))
    dw = torch.randn(taps, cout, cin, generator=g)
    w = torch.randn(taps, cout, cin, generator=g)
    ddc, dco = torch.randn(batch, cout, generator=g), torch.rand(batch, cout, generator=g) + 0.5
    st = torch.randn(batch, cin, generator=g)
    grad0 = torch.randn(cout, cin, k, k, generator=g)
    coef = torch.einsum('no,ni->oi', (ddc * dco ** 3).double(), (st * st).double())
    want = grad0.double() + (dw.double() - w.double() * coef[None]).view(k, k, cout, cin).permute(2, 3, 0, 1)
    grad = grad0.clone().cuda()
    dw_in = dw.transpose(1, 2).contiguous() if transposed else dw
    ops.modconv_wgrad_finish(dw_in.cuda(), w.cuda(), grad, transposed=transposed, ddcoef=ddc.cuda(), dcoef=dco.cuda(), styles=st.cuda())
    assert pu.rel_err(grad, want.float()) < 1e-5
    # without the demodulation term (ToRGB-style layers): plain unpack + accumulate
    grad = grad0.clone().cuda()
    ops.modconv_wgrad_finish(dw_in.cuda(), w.cuda(), grad, transposed=transposed)
    assert pu.rel_err(grad, (grad0.double() + dw.double().view(k, k, cout, cin).permute(2, 3, 0, 1)).float()) < 1e-6


@pytest.mark.parametrize('samples', [64, 1000, 20011])
def test_decoder_weight_gradient_kernel(samples):
    """hfagp_decoder_wgrad (hidden layer recomputed, four reductions over the samples) against fp64 autograd of the same
    32 -> 64 softplus -> 33 MLP on the same per-sample operands; ragged sample counts exercise the zero-filled tail."""
    ops = _ops()
    g = torch.Generator().manual_seed(samples)
    f = torch.randn(samples, 32, generator=g)
    do = torch.randn(samples, 33, generator=g) * 0.1
    w0 = (torch.randn(64, 32, generator=g) / math.sqrt(32)).requires_grad_()
    b0 = (torch.randn(64, generator=g) * 0.1).requires_grad_()
    w1 = (torch.randn(33, 64, generator=g) / 8).requires_grad_()
    b1 = torch.zeros(33, requires_grad=True)
    out = F.softplus(f.double() @ w0.double().t() + b0.double()) @ w1.double().t() + b1.double()
    (out * do.double()).sum().backward()
    want = torch.cat([w0.grad.reshape(-1), b0.grad, w1.grad.reshape(-1), b1.grad]).float()
    mlp = torch.cat([w0.detach().reshape(-1), b0.detach(), w1.detach().reshape(-1), b1.detach()]).cuda()
    got = ops.decoder_wgrad(f.cuda(), do.cuda(), mlp).cpu()
    for name, lo, hi in (('dW0', 0, 2048), ('db0', 2048, 2112), ('dW1', 2112, 2112 + 2112), ('db1', 4224, 4257)):
        e = pu.rel_l2(got[lo:hi], want[lo:hi])
        print(f'decoder_wgrad S={samples} {name}: rel-L2 {e:.3e}')
        assert e < 1e-4, (name, e)


@pytest.mark.parametrize('precision', ['tc', 'fp32'])
def test_generator_backward_to_ws(precision):
    """loss(image) -> d(ws) through super-resolution, renderer and backbone (generator frozen), vs the oracle."""
    cfg = eg3d_ref.small14_config()
    ref, prod = _pair(cfg, precision)
    b = 2
    ws, c, jitter, u = pu.make_inputs(cfg, b, seed=2)
    g = torch.Generator().manual_seed(11)
    ws_r = ws.clone().requires_grad_(True)
    img_r = ref.synthesis(ws_r, c, jitter_coarse=jitter, u_fine=u)['image']
    gi = torch.randn(img_r.shape, generator=g)
    (img_r * gi).sum().backward()
    ws_g = ws.clone().cuda().requires_grad_(True)
    img = prod.synthesis(ws_g, c.cuda(), noise_mode='const', jitter_coarse=jitter.cuda(), u_fine=u.cuda())['image']
    assert pu.rel_err(img, img_r) < pu.REL_TOL
    (img * gi.cuda()).sum().backward()
    _check_grad(ws_g.grad, ws_r.grad, precision, 'generator dws')


@pytest.mark.parametrize('precision', ['fp32', 'tc'])
def test_generator_weight_gradients(precision):
    """The post-tune_iter regime (train_rgb.py:132-134, trainer_rgb.py:69-71): gradients of the generator's own
    parameters — modulated-conv weights (style-scaled wgrad + demodulation term), biases, noise strengths, affine
    layers, ToRGB layers, the learned constant, the decoder MLP — against autograd of the oracle."""
    cfg = eg3d_ref.small14_config()
    ref, prod = _pair(cfg, precision)
    b = 2
    ws, c, jitter, u = pu.make_inputs(cfg, b, seed=4)
    g = torch.Generator().manual_seed(13)
    skip = ('backbone.mapping.',)
    for n, p in ref.named_parameters():
        p.requires_grad_(not n.startswith(skip))
    for n, p in prod.named_parameters():
        p.requires_grad_(not n.startswith(skip))
    img_r = ref.synthesis(ws, c, jitter_coarse=jitter, u_fine=u)['image']
    gi = torch.randn(img_r.shape, generator=g)
    (img_r * gi).sum().backward()
    img = prod.synthesis(ws.cuda(), c.cuda(), noise_mode='const', jitter_coarse=jitter.cuda(), u_fine=u.cuda())['image']
    assert pu.rel_err(img, img_r) < pu.REL_TOL
    (img * gi.cuda()).sum().backward()
    want = dict(ref.named_parameters())
    checked = 0
    for n, p in prod.named_parameters():
        if n.startswith(skip):
            assert p.grad is None
            continue
        wr = want[n].grad
        if wr is None:                       # e.g. noise strengths of the SR blocks (noise_mode 'none' there)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        assert p.grad is not None, n
        e_max, e_l2 = pu.rel_err(p.grad, wr), pu.rel_l2(p.grad, wr)
        print(f'{n}: max-rel {e_max:.2e} rel-L2 {e_l2:.2e}')
        assert e_l2 < (GRAD_TOL_TC_L2 if precision == 'tc' else GRAD_TOL) * 2, (n, e_l2)
        checked += 1
    assert checked > 100


@pytest.mark.parametrize('n,cin,cout,h,k,stride,pad', [
    (2, 64, 64, 32, 3, 1, 1),        # ResBlock conv1 (patch groups of 3 taps, N = 64)
    (1, 128, 256, 34, 3, 2, 0),      # down-sampling conv2 after the blur: stride 2, single-tap groups, partial tiles
    (2, 256, 128, 16, 1, 1, 0),      # 1x1 (skip / ToRGB-like), N = 128, two cout... one co tile
    (1, 128, 128, 24, 3, 1, 1),      # N = 128: three accumulators of 128 columns
])
def test_wgrad_tensor_core_matches_autograd(n, cin, cout, h, k, stride, pad):
    """hfagp_conv2d_wgrad on split-bf16 operands takes the tcgen05 kernel (csrc/wgrad_tc.cu: MN-major operands straight from
    the channels-last activations, three split terms, K split over CTAs) and must reproduce autograd's weight gradient of
    F.conv2d to fp32-class accuracy; the same call on fp32 operands (SIMT kernel) is checked beside it."""
    from hfa_gp_b200 import ops
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, cin, h, h, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g, requires_grad=True)
    y = F.conv2d(x, w, stride=stride, padding=pad)
    dz = torch.randn(y.shape, generator=g)
    (y * dz).sum().backward()
    want = w.grad.permute(2, 3, 0, 1).reshape(k * k, cout, cin)
    oh = y.shape[2]
    taps = tuple((ky - pad, kx - pad, ky * k + kx) for ky in range(k) for kx in range(k))
    xn, dzn = x.permute(0, 2, 3, 1).contiguous().cuda(), dz.permute(0, 2, 3, 1).contiguous().cuda()
    for split in (True, False):
        dw = torch.zeros(k * k, cout, cin, device='cuda')
        ops.conv2d_wgrad(ops.split(xn) if split else xn, ops.split(dzn) if split else dzn, taps, dw, oh=oh, ow=oh,
                         in_stride=stride, scale=0.5)
        assert pu.rel_err(dw, 0.5 * want) < 2e-4, ('split' if split else 'fp32')


@pytest.mark.parametrize('n,cin,cout,h,k,stride,pad', [(2, 128, 128, 16, 3, 1, 1), (2, 64, 96, 32, 1, 1, 0), (3, 64, 128, 18, 3, 2, 0)])
def test_wgrad_tensor_core_with_style_factors(n, cin, cout, h, k, stride, pad):
    """hfagp_conv2d_wgrad_mod on the tcgen05 kernel: per-sample style factors on either operand (a modulated convolution's
    weight gradient, code/train_rgb.py:132-134 regime) are applied to the per-sample accumulators in the epilogue."""
    from hfa_gp_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, cin, h, h, generator=g)
    xs, = (torch.rand(n, cin, generator=g) + 0.5,)
    w = torch.randn(cout, cin, k, k, generator=g, requires_grad=True)
    y = F.conv2d(x * xs[:, :, None, None], w, stride=stride, padding=pad)
    dz = torch.randn(y.shape, generator=g)
    dzs = torch.rand(n, cout, generator=g) + 0.5
    (y * dz * dzs[:, :, None, None]).sum().backward()
    want = w.grad.permute(2, 3, 0, 1).reshape(k * k, cout, cin)
    oh = y.shape[2]
    taps = tuple((ky - pad, kx - pad, ky * k + kx) for ky in range(k) for kx in range(k))
    xn, dzn = x.permute(0, 2, 3, 1).contiguous().cuda(), dz.permute(0, 2, 3, 1).contiguous().cuda()
    for split in (True, False):
        dw = torch.zeros(k * k, cout, cin, device='cuda')
        ops.conv2d_wgrad(ops.split(xn) if split else xn, ops.split(dzn) if split else dzn, taps, dw, oh=oh, ow=oh,
                         in_stride=stride, xscale=xs.cuda(), dzscale=dzs.cuda())
        assert pu.rel_err(dw, want) < 2e-4, ('split' if split else 'fp32')
