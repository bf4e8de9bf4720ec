"""GPU parity tests (``-m gpu``): every CUDA entry point, called through the C ABI, against the CPU
oracle on the same seeded inputs.  fp32 tolerance: REL_TOL = 1e-3 relative per element (north star);
integer bookkeeping bit-exact given identical float inputs (ties within 1e-6 of a cdf knot are masked,
SURVEY.md §7 hard part 6)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import eg3d_ref, hfagp_ref
import parity_utils as pu

pytestmark = pytest.mark.gpu


def _ops():
    from hfa_gp_b200 import ops
    return ops


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def pack(w):
    o, i, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, o, i).contiguous().cuda()


# ------------------------------------------------------------------ unit ops

@pytest.mark.parametrize('n,cin,cout,h,w,k,stride,pad', [
    (2, 64, 64, 16, 16, 3, 1, 1),
    (1, 3, 64, 32, 32, 1, 1, 0),        # first encoder layer: cin = 3 (scalar loads)
    (2, 32, 96, 8, 8, 1, 1, 0),         # ToRGB-like, cout not a tile multiple
    (1, 128, 256, 19, 19, 3, 2, 0),     # stride-2 after blur, odd extent
    (1, 512, 512, 4, 4, 4, 1, 0),       # final 4x4 conv -> 1x1
    (1, 128, 128, 64, 64, 3, 1, 1),     # big-tile path (>=148 CTAs at 128x128)
    (1, 40, 20, 7, 5, 3, 1, 1),         # ragged everything
])
def test_conv2d_matches_torch(n, cin, cout, h, w, k, stride, pad):
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g)
    ref = F.leaky_relu(F.conv2d(x, wt, b, stride=stride, padding=pad), 0.2) * math.sqrt(2)
    taps = tuple((ky - pad, kx - pad, ky * k + kx) for ky in range(k) for kx in range(k))
    oh, ow = ref.shape[2], ref.shape[3]
    y = ops.conv2d(nhwc(x), pack(wt), taps, cout, oh=oh, ow=ow, in_stride=stride, bias=b.cuda(),
                   act=1, act_gain=math.sqrt(2))
    assert pu.rel_err(pu.to_nchw(y), ref) < 1e-4


def test_conv2d_epilogue_residual_noise_clamp_upimg():
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    n, cin, cout, h = 2, 32, 48, 16
    x = torch.randn(n, cin, h, h, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / 10
    d = torch.rand(n, cout, generator=g) + 0.5
    noise = torch.randn(h, h, generator=g)
    b = torch.randn(cout, generator=g)
    res = torch.randn(n, cout, h, h, generator=g)
    up = torch.randn(n, cout, h // 2, h // 2, generator=g)
    v = F.conv2d(x, wt, padding=1) * d[:, :, None, None] + noise * 0.3 + b[None, :, None, None]
    v = (F.leaky_relu(v, 0.2) * 1.3).clamp(-1.0, 1.0)
    v = (v + res) * 0.7
    v = v + eg3d_ref.upsample2d_ref(up, eg3d_ref.setup_filter())
    y = ops.conv2d(nhwc(x), pack(wt), ops.TAPS_3X3, cout, oh=h, ow=h, dcoef=d.cuda(), noise=noise.cuda(),
                   noise_gain=0.3, bias=b.cuda(), act=1, act_gain=1.3, clamp=1.0, residual=nhwc(res),
                   residual_scale=0.7, up_img=nhwc(up))
    assert pu.rel_err(pu.to_nchw(y), v) < 1e-4


@pytest.mark.parametrize('n,cin,cout,h', [(2, 32, 64, 8), (1, 64, 32, 17)])
def test_upconv_matches_oracle(n, cin, cout, h):
    """modulated up-sampling layer = transposed conv (4 parity classes) + FIR/act kernel."""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, cin, h, h, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g)
    s = torch.randn(n, cin, generator=g) + 1
    noise = torch.randn(2 * h, 2 * h, generator=g)
    b = torch.randn(cout, generator=g) * 0.1
    ref = eg3d_ref.modulated_conv2d_ref(x, wt, s, noise=noise * 0.2, up=2, f=eg3d_ref.setup_filter())
    ref = eg3d_ref.bias_act_ref(ref, b, act='lrelu', clamp=2.0)
    wmod, dcoef = ops.modulate(pack(wt), s.cuda(), True)
    t = ops.conv_transpose_s2(nhwc(x), wmod, cout, wmod.stride(0))
    y = ops.upfir_act(t, dcoef=dcoef, noise=noise.cuda(), noise_gain=0.2, bias=b.cuda(), clamp=2.0)
    assert pu.rel_err(pu.to_nchw(y), ref) < 1e-4


@pytest.mark.parametrize('n,h2,c', [(1, 256, 128), (1, 128, 256), (2, 64, 512), (1, 256, 24), (3, 9, 16)])
def test_upfir_register_tiled_forms(n, h2, c):
    """hfagp_upfir_act_fwd at sizes that take the 4-rows-per-thread kernels (channel count as template parameter: 128 / 256 /
    512, generic otherwise) and the 1-row form, fp32 and split-bf16 outputs: FIR [1,3,3,1]^2/16 with one pixel of zero padding
    on the (2H+1)^2 intermediate, then demodulation, noise, bias, leaky ReLU * sqrt(2), clamp (eg3d upfirdn2d + bias_act)."""
    ops = _ops()
    g = torch.Generator().manual_seed(n * 1000 + h2 + c)
    t = torch.randn(n, c, h2 + 1, h2 + 1, generator=g)
    dcoef = torch.rand(n, c, generator=g) + 0.5
    noise = torch.randn(h2, h2, generator=g)
    b = torch.randn(c, generator=g) * 0.1
    f = eg3d_ref.setup_filter() * 4
    y = F.conv2d(F.pad(t, [1, 1, 1, 1]), f[None, None].repeat(c, 1, 1, 1), groups=c)
    y = y * dcoef[:, :, None, None] + noise[None, None] * 0.3
    ref = eg3d_ref.bias_act_ref(y, b, act='lrelu', clamp=1.5)
    tg = nhwc(t)
    got = ops.upfir_act(tg, dcoef=dcoef.cuda(), noise=noise.cuda(), noise_gain=0.3, bias=b.cuda(), clamp=1.5)
    assert pu.rel_err(pu.to_nchw(got), ref) < 1e-5
    if c % 8 == 0:
        gs = ops.upfir_act(tg, dcoef=dcoef.cuda(), noise=noise.cuda(), noise_gain=0.3, bias=b.cuda(), clamp=1.5, split_out=True)
        assert pu.rel_err(pu.to_nchw(gs.float()), ref) < 1e-4


@pytest.mark.parametrize('n,h,c,stride,pad', [(4, 128, 128, 1, 2), (8, 128, 128, 2, 1), (1, 256, 64, 1, 2), (1, 256, 64, 2, 1),
                                              (2, 15, 8, 1, 2), (2, 14, 8, 2, 1)])
def test_blur_register_tiled_forms(n, h, c, stride, pad):
    """hfagp_blur_fwd in its 4-/2-rows-per-thread forms (batched encoder sizes) and the 1-row form, fp32 and split-bf16
    I/O, against the reference's Blur (upfirdn2d_native, encoder3d.py:23-75) followed by the stride the next conv applies."""
    ops = _ops()
    g = torch.Generator().manual_seed(h * 10 + c + stride)
    x = torch.randn(n, c, h, h, generator=g)
    ref = hfagp_ref.blur_ref(x, hfagp_ref.blur_kernel(), pad, pad)[:, :, ::stride, ::stride]
    xg = nhwc(x)
    assert pu.rel_err(pu.to_nchw(ops.blur(xg, pad, pad, stride=stride)), ref) < 1e-5
    ys = ops.blur(ops.split(xg), pad, pad, stride=stride, split_out=True)
    assert pu.rel_err(pu.to_nchw(ys.float()), ref) < 1e-4


def test_modulate_and_styles():
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    b, nws, wd = 3, 6, 512
    ws = torch.randn(b, nws, wd, generator=g)
    layers = []
    refs = []
    for cin, widx, gain in ((64, 0, 1.0), (32, 5, 1 / math.sqrt(32)), (40, 3, 1.0)):
        a = torch.randn(cin, wd, generator=g)
        bb = torch.randn(cin, generator=g)
        layers.append((a.cuda(), bb.cuda(), cin, widx, gain))
        refs.append((torch.addmm(bb[None], ws[:, widx], (a / math.sqrt(wd)).t())) * gain)
    outs = ops.StyleTable(layers).run(ws.cuda())
    for o, r in zip(outs, refs):
        assert pu.rel_err(o, r) < 1e-5
    wt = torch.randn(24, 64, 3, 3, generator=g)
    wmod, dcoef = ops.modulate(pack(wt), refs[0].cuda(), True)
    wm_ref = wt[None] * refs[0][:, None, :, None, None]
    d_ref = (wm_ref.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    assert pu.rel_err(dcoef, d_ref) < 1e-5
    assert pu.rel_err(wmod.view(b, 3, 3, 24, 64).permute(0, 3, 4, 1, 2), wm_ref) < 1e-6


def test_small_ops():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 76, generator=g)
    w = torch.randn(50, 76, generator=g)
    b = torch.randn(50, generator=g)
    ref = hfagp_ref.equal_linear_ref(x, w, b)
    assert pu.rel_err(ops.linear(x.cuda(), w.cuda(), b.cuda(), 1 / math.sqrt(76), 1.0), ref) < 1e-5
    # blur, both pads / strides
    xi = torch.randn(2, 8, 13, 13, generator=g)
    k = hfagp_ref.blur_kernel()
    assert pu.rel_err(pu.to_nchw(ops.blur(nhwc(xi), 2, 2)), hfagp_ref.blur_ref(xi, k, 2, 2)) < 1e-5
    assert pu.rel_err(pu.to_nchw(ops.blur(nhwc(xi[..., :12, :12]), 1, 1, stride=2)),
                      hfagp_ref.blur_ref(xi[..., :12, :12], k, 1, 1)[:, :, ::2, ::2]) < 1e-5
    # blur with split-bf16 input and output (tensor-core encoder path)
    ys = ops.blur(ops.split(nhwc(xi)), 2, 2, split_out=True)
    assert pu.rel_err(pu.to_nchw(ys.float()), hfagp_ref.blur_ref(xi, k, 2, 2)) < 1e-4
    # small-N ToRGB + skip upsample
    xr = torch.randn(2, 64, 8, 8, generator=g)
    wr = torch.randn(2, 3, 64, generator=g)
    br = torch.randn(3, generator=g)
    up = torch.randn(2, 3, 4, 4, generator=g)
    ref = torch.einsum('nchw,noc->nohw', xr, wr) + br[None, :, None, None]
    ref = ref.clamp(-3, 3) + eg3d_ref.upsample2d_ref(up, eg3d_ref.setup_filter())
    y = ops.torgb_small(nhwc(xr), wr.cuda().contiguous(), br.cuda(), 3.0, nhwc(up), 3)
    assert pu.rel_err(pu.to_nchw(y), ref) < 1e-5
    # layout helpers
    assert torch.equal(ops.nhwc_to_nchw(ops.nchw_to_nhwc(xi.cuda())).cpu(), xi)
    # latent map
    bases = torch.randn(50, 14 * 512, generator=g)
    delta = bases.mean(0)
    wts = torch.randn(2, 50, generator=g)
    q, _ = torch.linalg.qr((bases + 1e-8).T)
    ref = hfagp_ref.get_latent_ref(bases, delta, wts).reshape(2, -1)
    assert pu.rel_err(ops.latent(wts.cuda(), q.contiguous().cuda(), delta.cuda(), 14 * 512), ref) < 1e-5


# ------------------------------------------------------------------ renderer

def _render_case(res, s, sf, batch=1, seed=0, plane_res=64):
    cfg = eg3d_ref.GeneratorConfig(nrr=res, depth_res=s, depth_res_importance=sf, plane_res=plane_res)
    g = torch.Generator().manual_seed(seed)
    planes = torch.randn(batch, 3, 32, plane_res, plane_res, generator=g)
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        dec = eg3d_ref.OSGDecoderRef(cfg)
        with torch.no_grad():
            dec.net[0].bias.normal_(0, 0.5)
            dec.net[2].bias.normal_(0, 0.5)
    c = hfagp_ref.flip_label_(hfagp_ref.synthetic_labels(batch, seed=seed))
    rays = res * res
    jitter = torch.rand(batch, rays, s, 1, generator=g)
    u = torch.rand(batch * rays, max(sf, 1), generator=g)
    return cfg, planes, dec, c, jitter, u


def _run_render_gpu(cfg, planes, dec, c, jitter, u, book=True, simt=False):
    ops = _ops()
    d0, d2 = dec.net[0], dec.net[2]
    mlp = torch.cat([(d0.weight * d0.weight_gain).reshape(-1), d0.bias * d0.bias_gain,
                     (d2.weight * d2.weight_gain).reshape(-1), d2.bias * d2.bias_gain]).detach().cuda()
    b = planes.shape[0]
    pl = planes.reshape(b, 96, planes.shape[-2], planes.shape[-1]).permute(0, 2, 3, 1).contiguous().cuda()
    s, sf = cfg.depth_res, cfg.depth_res_importance
    lin = torch.linspace(cfg.ray_start, cfg.ray_end, s)
    delta = (cfg.ray_end - cfg.ray_start) / (s - 1)
    jit = jitter.reshape(b, -1, s)
    rng = torch.stack([lin[0] + jit[:, :, 0].min() * delta, lin[-1] + jit[:, :, -1].max() * delta])
    return ops.render(pl, c.cuda(), mlp, lin.cuda(), jit.contiguous().cuda(), u.cuda() if sf > 0 else None,
                      rng.cuda(), res=cfg.nrr, s_coarse=s, s_fine=sf, delta=delta, box_scale=2.0 / cfg.box_warp,
                      bookkeeping=book, simt=simt)


@pytest.mark.parametrize('res,s,sf,batch', [(16, 48, 48, 2), (8, 33, 20, 1), (24, 16, 16, 1), (16, 40, 0, 1)])
def test_render_tc_equals_simt(res, s, sf, batch):
    """The tcgen05 renderer (csrc/render_tc.cu, what hfagp_render_fwd dispatches to) against the legacy mma.sync
    kernel (hfagp_render_fwd_simt) on identical inputs: two independent implementations of the same fp32-class
    arithmetic, so they agree far inside the oracle tolerance; the integer bookkeeping is identical except where a
    last-bit difference of a density moves u across a cdf knot."""
    cfg, planes, dec, c, jitter, u = _render_case(res, s, sf, batch, seed=3)
    a = _run_render_gpu(cfg, planes, dec, c, jitter, u)
    b = _run_render_gpu(cfg, planes, dec, c, jitter, u, simt=True)
    for x, y in zip(a[:3], b[:3]):
        assert pu.rel_err(x, y.cpu()) < 1e-4
    if sf > 0:
        for k in ('inds', 'below', 'above', 'sort_idx'):
            same = (a[3][k] == b[3][k]).float().mean()
            assert float(same) > 0.999, f'{k}: only {float(same):.5f} equal'
        assert pu.rel_err(a[3]['depths_sorted'], b[3]['depths_sorted'].cpu()) < 1e-5


@pytest.mark.parametrize('res,s,sf,batch', [(16, 48, 0, 1), (16, 48, 48, 2), (8, 12, 12, 1), (8, 33, 20, 1),
                                            (8, 64, 64, 1)])
def test_render_matches_oracle(res, s, sf, batch):
    cfg, planes, dec, c, jitter, u = _render_case(res, s, sf, batch)
    tap = {}
    cam = c[:, :16].view(-1, 4, 4)
    intr = c[:, 16:25].view(-1, 3, 3)
    ro, rd = eg3d_ref.ray_sampler_ref(cam, intr, res)
    with torch.no_grad():
        feat_ref, depth_ref, wsum_ref = eg3d_ref.render_ref(planes, dec, ro, rd, cfg, jitter, u if sf > 0 else None, tap)
    feat, depth, wsum, book = _run_render_gpu(cfg, planes, dec, c, jitter, u)
    assert pu.rel_err(feat.reshape(batch, -1, 32), feat_ref) < pu.REL_TOL
    assert pu.rel_err(depth, depth_ref.reshape(batch, -1)) < pu.REL_TOL
    assert pu.rel_err(wsum, wsum_ref.reshape(batch, -1)) < pu.REL_TOL
    if sf > 0:
        # integer bookkeeping: bit-exact except where u sits within 1e-6 of a cdf knot
        cdf = tap['cdf']
        uu = u[:, :sf]
        near = (uu[:, :, None] - cdf[:, None, :]).abs().min(dim=-1).values < 1e-6
        for k in ('inds', 'below', 'above'):
            got = book[k].cpu().long()
            bad = (got != tap[k]) & ~near
            assert int(bad.sum()) == 0, f'{k}: {int(bad.sum())} mismatches outside ties'
        assert float(near.float().mean()) < 1e-3
        ds = book['depths_sorted'].cpu()
        assert pu.rel_err(ds, tap['depths_sorted'].reshape(ds.shape)) < 1e-5
        assert bool((ds[..., 1:] >= ds[..., :-1]).all()), 'merged depths must be sorted'
        # the permutation must reproduce the oracle's wherever depths are well separated
        order_ref = tap['sort_idx'].reshape(ds.shape)
        gap = (tap['depths_sorted'].reshape(ds.shape)[..., 1:] - tap['depths_sorted'].reshape(ds.shape)[..., :-1]) > 1e-5
        ok = torch.ones_like(order_ref, dtype=torch.bool)
        ok[..., 1:] &= gap
        ok[..., :-1] &= gap
        assert bool((book['sort_idx'].cpu().long()[ok] == order_ref[ok]).all())


def test_render_cfg1_micro():
    """BASELINE.json configs[0]: 128x128 frame, 32 rays x 48 samples, random-init tri-plane MLP."""
    res = 128
    cfg, planes, dec, c, jitter, u = _render_case(res, 48, 48, 1, seed=0, plane_res=256)
    idx = torch.arange(32) * 512 + 64
    cam = c[:, :16].view(-1, 4, 4)
    intr = c[:, 16:25].view(-1, 3, 3)
    ro, rd = eg3d_ref.ray_sampler_ref(cam, intr, res)
    with torch.no_grad():
        feat_ref, depth_ref, wsum_ref = eg3d_ref.render_ref(planes, dec, ro[:, idx], rd[:, idx], cfg,
                                                           jitter[:, idx], u[idx])
    feat, depth, wsum, _ = _run_render_gpu(cfg, planes, dec, c, jitter, u, book=False)
    assert pu.rel_err(feat.reshape(1, -1, 32)[:, idx.cuda()], feat_ref) < pu.REL_TOL
    assert pu.rel_err(wsum[:, idx.cuda()], wsum_ref.reshape(1, -1)) < pu.REL_TOL
    # depth: the oracle clamps to the min/max of the rays it saw; compare where no clamp is active
    assert pu.rel_err(depth[:, idx.cuda()], depth_ref.reshape(1, -1)) < pu.REL_TOL


# ------------------------------------------------------------------ whole generator

def _generator_case(cfg, batch, seed=0, precision='tc'):
    ref, prod = pu.make_pair(cfg, seed=seed)
    prod.precision = precision
    ws, c, jitter, u = pu.make_inputs(cfg, batch, seed=seed)
    tap_r, tap_g = {}, {}
    with torch.no_grad():
        out_r = ref.synthesis(ws, c.clone(), noise_mode='const', jitter_coarse=jitter, u_fine=u, tap=tap_r)
        out_g = prod.synthesis(ws.cuda(), c.clone().cuda(), noise_mode='const', jitter_coarse=jitter.cuda(),
                               u_fine=u.cuda(), tap=tap_g)
    return out_r, out_g, tap_r, tap_g


@pytest.mark.parametrize('precision', ['tc', 'fp32'])
def test_generator_tiny_stagewise(precision):
    cfg = eg3d_ref.tiny_config()
    out_r, out_g, tap_r, tap_g = _generator_case(cfg, batch=2, precision=precision)
    errs = pu.compare_taps(tap_r, tap_g)
    report = ', '.join(f'{k}={e:.2e}' for k, e in errs)
    for k, e in errs:
        assert e < pu.REL_TOL, f'stage {k}: rel err {e:.3e}  [{report}]'
    for k in ('image', 'image_raw', 'image_depth'):
        assert out_g[k].shape == out_r[k].shape
        assert pu.rel_err(out_g[k], out_r[k]) < pu.REL_TOL, k


@pytest.mark.parametrize('precision', ['tc', 'fp32'])
def test_generator_full_512_frame(precision):
    """BASELINE.json configs[1]: 512x512, 96 samples/ray, random-init EG3D generator.
    'tc' = tcgen05 split-bf16 convolutions (the shipped path), 'fp32' = SIMT kernels only."""
    cfg = eg3d_ref.GeneratorConfig()
    out_r, out_g, tap_r, tap_g = _generator_case(cfg, batch=1, precision=precision)
    print('stage errors', precision, sorted(pu.compare_taps(tap_r, tap_g), key=lambda kv: -kv[1])[:6])
    assert out_g['image'].shape == (1, 3, 512, 512)
    errs = dict(pu.compare_taps(tap_r, tap_g))
    worst = max(errs.items(), key=lambda kv: kv[1])
    assert worst[1] < pu.REL_TOL, f'worst stage {worst}'
    assert pu.rel_err(out_g['image'], out_r['image']) < pu.REL_TOL
    # north_star words the tolerance "per pixel": the pure ratio |a-b|/|b| beside the rms-floored one.  It exceeds 1e-3
    # only at zero crossings of the reference image (|b| below 5 % of the image rms); away from them it holds outright.
    st = pu.pixel_rel_stats(out_g['image'], out_r['image'])
    print('per-pixel relative error', precision, st)
    assert st['max_over_away_from_zero'] < pu.REL_TOL
    assert st['frac_over'] < 2e-3 and st['max_abs_ref_where_over'] < 0.05
    # size-independent properties at full size
    assert bool(torch.isfinite(out_g['image']).all())
    ws_sum = tap_g['weight_sum']
    assert float(ws_sum.max()) <= 1.0 + 1e-5 and float(ws_sum.min()) >= 0.0
    ds = tap_g['depths_sorted']
    assert bool((ds[..., 1:] >= ds[..., :-1]).all())
    assert int(tap_g['inds'].min()) >= 1 and int(tap_g['inds'].max()) <= cfg.depth_res - 2


# ------------------------------------------------------------------ encoder + avatar drop-in

@pytest.mark.parametrize('precision', ['tc', 'fp32'])
@pytest.mark.parametrize('size,out_pose', [(64, True), (256, False)])
def test_encoder_matches_reference_restatement(size, out_pose, precision):
    from hfa_gp_b200.networks.encoder3d import Encoder
    sd = hfagp_ref.make_encoder_state(size, 512, 50, out_pose=out_pose, seed=3)
    g = torch.Generator().manual_seed(11)
    for k in sd:
        if k.endswith('bias'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.2
    enc = Encoder(size, 512, 50, False, out_pose)
    enc.load_state_dict(sd, strict=True)
    enc = enc.eval().requires_grad_(False).cuda()
    enc.net_app.precision = precision
    x = torch.rand(2, 3, size, size, generator=g) * 2 - 1
    ref = hfagp_ref.encoder_ref(sd, x, out_pose=out_pose)
    with torch.no_grad():
        got = enc(x.cuda())
    if out_pose:
        assert pu.rel_err(got[0], ref[0]) < pu.REL_TOL and pu.rel_err(got[1], ref[1]) < pu.REL_TOL
    else:
        assert pu.rel_err(got, ref) < pu.REL_TOL


def test_headnerf_dropin_frame_loop():
    """run_recon_video_rgb.py:216-236 per-frame call sequence on the drop-in classes vs the oracle chain."""
    import argparse
    from hfa_gp_b200.networks.headnerf import HeadNeRF_final
    cfg = eg3d_ref.small14_config()
    args = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='',
                              synthetic_generator=True, generator_seed=0, generator_config=pu.product_config(cfg))
    torch.manual_seed(0)
    model = HeadNeRF_final(args, 64, 'cuda', 512, 50, 'x', './').cuda().eval()
    # oracle side gets the very same tensors
    ref_gen = eg3d_ref.TriPlaneGeneratorRef(cfg)
    ref_gen.load_state_dict({k: v.cpu() for k, v in model.generator.state_dict().items()})
    sd_enc = {k: v.detach().cpu() for k, v in model.encoder.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    label = hfagp_ref.synthetic_labels(1, seed=9)
    rays = cfg.nrr ** 2
    jitter = torch.rand(1, rays, cfg.depth_res, 1, generator=g)
    u = torch.rand(rays, cfg.depth_res_importance, generator=g)
    with torch.no_grad():
        w_ref = hfagp_ref.encoder_ref(sd_enc, img)
        lat_ref = hfagp_ref.get_latent_ref(model.bases.detach().cpu(), model.delta.detach().cpu(), w_ref)
        lab_ref = hfagp_ref.flip_label_(label.clone())
        img_ref = ref_gen.synthesis(lat_ref, lab_ref, jitter_coarse=jitter, u_fine=u)['image']
        lab_gpu = label.clone().cuda()
        w = model.get_weights(img.cuda())
        lat = model.get_latent(w)
        latc = lat.clone()
        lab_before = lab_gpu.clone()
        out = model.generator.synthesis(latc, c=hfagp_ref.flip_label_(lab_gpu), noise_mode='const',
                                        jitter_coarse=jitter.cuda(), u_fine=u.cuda())['image']
    assert pu.rel_err(w, w_ref) < pu.REL_TOL
    assert lat.shape == (1, 14, 512)
    # QR is torch.linalg.qr on both sides (cuSOLVER vs LAPACK): same Householder convention expected
    assert pu.rel_err(lat, lat_ref) < pu.REL_TOL
    assert pu.rel_err(out, img_ref) < pu.REL_TOL
    # in-place label flip semantics of get_image (headnerf.py:132)
    lab2 = lab_before.clone()
    with torch.no_grad():
        model.get_image(lat, lab2)
    assert torch.equal(lab2[:, [1, 2, 5, 6, 9, 10]], -lab_before[:, [1, 2, 5, 6, 9, 10]])
    assert torch.equal(lab2[:, [0, 3, 4, 7, 8, 11]], lab_before[:, [0, 3, 4, 7, 8, 11]])


# ------------------------------------------------------------------ tensor-core (tcgen05) convolution

@pytest.mark.parametrize('n,cin,cout,h,w,k,batched', [
    (1, 64, 128, 16, 8, 1, False),       # plain GEMM: one M tile, one K chunk
    (1, 128, 128, 16, 16, 1, False),     # two K chunks, two M tiles
    (2, 64, 64, 16, 16, 3, True),        # 3x3 taps (TMA zero padding), per-sample weights, bn = 64
    (1, 256, 96, 32, 32, 1, True),       # ToRGB-like: bn = 96
    (1, 128, 256, 24, 40, 3, False),     # ragged spatial extent, two N tiles
    (1, 512, 512, 8, 8, 3, True),        # low-res block: half-empty M tile, 72 K iterations
    (1, 32, 256, 16, 16, 3, True),       # first SR layer: cin = 32 = half a K chunk (TMA zero-fills the rest)
    (2, 96, 64, 16, 16, 1, False),       # 1.5 K chunks (dgrad of the 96-channel ToRGB)
    (1, 40, 24, 9, 7, 3, True),          # ragged everything, cin % 8 == 0 only
])
def test_conv2d_tc_matches_fp64(n, cin, cout, h, w, k, batched):
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(n if batched else 1, cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g)
    pad = k // 2
    ref = torch.cat([F.conv2d(x[i:i + 1].double(), wt[i if batched else 0].double(), b.double(), padding=pad)
                     for i in range(n)])
    ref = (F.leaky_relu(ref, 0.2) * math.sqrt(2)).float()
    taps = tuple((ky - pad, kx - pad, ky * k + kx) for ky in range(k) for kx in range(k))
    wp = wt.permute(0, 3, 4, 1, 2).reshape(wt.shape[0], k * k, cout, cin).contiguous().cuda()
    xs, ws = ops.split(nhwc(x)), ops.split(wp)
    y = ops.conv2d_tc(xs, ws, taps, cout, oh=h, ow=w, w_batched=batched, bias=b.cuda(), act=1, act_gain=math.sqrt(2))
    err = pu.rel_err(pu.to_nchw(y), ref)
    assert err < 1e-4, err          # split-bf16: ~2^-16 per product, K up to 4608
    ys = ops.conv2d_tc(xs, ws, taps, cout, oh=h, ow=w, w_batched=batched, bias=b.cuda(), act=1,
                       act_gain=math.sqrt(2), split_out=True)
    assert pu.rel_err(pu.to_nchw(ys.float()), ref) < 1e-4


def test_conv_transpose_tc_parity_classes():
    ops = _ops()
    g = torch.Generator().manual_seed(10)
    n, cin, cout, h = 1, 64, 128, 16
    x = torch.randn(n, cin, h, h, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    ref = F.conv_transpose2d(x.double(), wt.double().transpose(0, 1), stride=2).float()
    xs, ws = ops.split(nhwc(x)), ops.split(pack(wt)[None].contiguous())
    out = torch.zeros(n, 2 * h + 1, 2 * h + 1, cout, device='cuda')
    for a in (0, 1):
        for b in (0, 1):
            ops.conv2d_tc(xs, ws, ops._parity_taps(a, b), cout, oh=h + 1 - a, ow=h + 1 - b, out=out,
                          out_hw=(2 * h + 1, 2 * h + 1), out_stride=2, out_off=(a, b))
    assert pu.rel_err(pu.to_nchw(out), ref) < 1e-4


@pytest.mark.parametrize('n,cin,cout,h,batched', [
    (1, 512, 512, 8, True),       # low-res up layer: 4 classes x 1 tile, split-K accumulate path (hfagp_conv2d_tc_acc_fwd)
    (2, 64, 128, 40, True),       # 4 classes x several tiles x 2 samples in one persistent launch (..._multi_fwd)
    (1, 32, 48, 130, False),      # more tiles than SMs: every CTA walks several tiles of different classes
])
def test_conv_transpose_tc_merged_launch(n, cin, cout, h, batched):
    """ops.conv_transpose_s2_tc: the four output-parity classes as ONE launch vs torch's conv_transpose2d (fp64)."""
    ops = _ops()
    g = torch.Generator().manual_seed(21)
    x = torch.randn(n, cin, h, h, generator=g)
    wt = torch.randn(n if batched else 1, cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    ref = torch.cat([F.conv_transpose2d(x[i:i + 1].double(), wt[i if batched else 0].double().transpose(0, 1), stride=2)
                     for i in range(n)]).float()
    wp = wt.permute(0, 3, 4, 1, 2).reshape(wt.shape[0], 9, cout, cin).contiguous().cuda()
    t = ops.conv_transpose_s2_tc(ops.split(nhwc(x)), ops.split(wp), cout, w_batched=batched)
    assert tuple(t.shape) == (n, 2 * h + 1, 2 * h + 1, cout)
    assert pu.rel_err(pu.to_nchw(t), ref) < 1e-4


def test_conv2d_tc_persistent_many_tiles_and_epilogues():
    """More output tiles than SMs (each CTA loops: ring wrap-around across tiles, both TMEM accumulators) with every
    epilogue term: demodulation, noise, bias, leaky-ReLU, gain, clamp, skip-image upsample-add; then the residual
    merge (the generic epilogue path)."""
    ops = _ops()
    g = torch.Generator().manual_seed(22)
    n, cin, cout, h, w = 2, 64, 96, 120, 104            # 2 * ceil(120*104/128) = 196 tiles
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    d = torch.rand(n, cout, generator=g) + 0.5
    noise = torch.randn(h, w, generator=g)
    b = torch.randn(cout, generator=g) * 0.1
    up = torch.randn(n, cout, h // 2, w // 2, generator=g)
    res = torch.randn(n, cout, h, w, generator=g)
    v = F.conv2d(x.double(), wt.double(), padding=1) * d.double()[:, :, None, None] + noise.double() * 0.3
    v = F.leaky_relu(v + b.double()[None, :, None, None], 0.2) * 1.3
    v = v.clamp(-1.0, 1.0)
    want_up = (v + eg3d_ref.upsample2d_ref(up.double(), eg3d_ref.setup_filter().double())).float()
    want_res = ((v + res.double()) * 0.7).float()
    xs, ws = ops.split(nhwc(x)), ops.split(pack(wt)[None].contiguous())
    common = dict(oh=h, ow=w, dcoef=d.cuda(), noise=noise.cuda(), noise_gain=0.3, bias=b.cuda(), act=1, act_gain=1.3,
                  clamp=1.0)
    y = ops.conv2d_tc(xs, ws, ops.TAPS_3X3, cout, up_img=nhwc(up), **common)
    assert pu.rel_err(pu.to_nchw(y), want_up) < 1e-4
    y = ops.conv2d_tc(xs, ws, ops.TAPS_3X3, cout, residual=nhwc(res), residual_scale=0.7, split_out=True, **common)
    assert pu.rel_err(pu.to_nchw(y.float()), want_res) < 1e-4


def test_frame_loop_graph_matches_eager():
    """hfa_gp_b200.frame_loop.FrameLoop: the captured CUDA graph replays what the eager
    get_weights -> get_latent -> get_image sequence computes (random draws pinned)."""
    import argparse
    from hfa_gp_b200.frame_loop import FrameLoop
    from hfa_gp_b200.networks.headnerf import HeadNeRF_final
    cfg = eg3d_ref.small14_config()
    args = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='',
                              synthetic_generator=True, generator_seed=0, generator_config=pu.product_config(cfg))
    torch.manual_seed(0)
    model = HeadNeRF_final(args, 64, 'cuda', 512, 50, 'x', './').cuda().eval().requires_grad_(False)
    g = torch.Generator().manual_seed(5)
    rays = cfg.nrr ** 2
    model.generator.fixed_draws = (torch.rand(1, rays, cfg.depth_res, 1, generator=g).cuda(),
                                   torch.rand(rays, cfg.depth_res_importance, generator=g).cuda())
    loop = FrameLoop(model, batch=1, size=64)
    for seed in (1, 2):
        img = (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1).cuda()
        label = hfagp_ref.synthetic_labels(1, seed=seed).cuda()
        lab_e, lab_g = label.clone(), label.clone()
        with torch.no_grad():
            want = model.get_image(model.get_latent(model.get_weights(img)), lab_e).clone()
        got = loop(img, lab_g).clone()
        # not bit-equal: the split-K layers add fp32 partial sums with atomics, whose order varies run to run
        assert pu.rel_err(got, want) < 1e-4
        assert torch.equal(lab_g, lab_e)            # the in-place GL flip is visible to the caller


def test_frame_loop_full_size_properties_and_egress():
    """configs[1] at full size through the public frame loop (CUDA graph, side-stream overlap, uint8 egress):
    size-independent properties instead of an oracle run — shapes, finiteness, run-to-run reproducibility up to the
    split-K atomics, weight sums in [0, 1], depths inside the sampled range, and the egress bytes equal to the
    reference's save_image convention applied to the returned image."""
    import argparse
    from hfa_gp_b200 import frameio
    from hfa_gp_b200.frame_loop import FrameLoop
    from hfa_gp_b200.networks.headnerf import HeadNeRF_final
    from oracle import frameio_ref
    args = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='',
                              synthetic_generator=True, generator_seed=0)
    torch.manual_seed(0)
    model = HeadNeRF_final(args, 256, 'cuda', 512, 50, 'x', './').cuda().eval().requires_grad_(False)
    cfg = model.generator.cfg
    g = torch.Generator().manual_seed(7)
    rays = cfg.nrr ** 2
    model.generator.fixed_draws = (torch.rand(1, rays, cfg.depth_res, 1, generator=g).cuda(),
                                   torch.rand(rays, cfg.depth_res_importance, generator=g).cuda())
    loop = FrameLoop(model, batch=1, size=256, egress='save_image')
    img = (torch.rand(1, 3, 256, 256, generator=g) * 2 - 1).cuda()
    label = hfagp_ref.synthetic_labels(1, seed=4).cuda()
    a = loop(img, label.clone()).clone()
    u8 = loop.out_u8.clone()
    b = loop(img, label.clone()).clone()
    assert tuple(a.shape) == (1, 3, 512, 512) and torch.isfinite(a).all()
    assert pu.rel_err(b, a) < 5e-4          # split-K fp32 atomics: summation order varies run to run (measured ~1e-4 at full depth)
    assert tuple(u8.shape) == (1, 512, 512, 3) and u8.dtype == torch.uint8
    assert torch.equal(u8.cpu(), frameio_ref.save_image_uint8(a.cpu()))
    # the renderer's own invariants, on the same frame, through synthesis()
    with torch.no_grad():
        lat = model.get_latent(model.get_weights(img))
        lab = label.clone()
        hfagp_ref.flip_label_(lab)
        tap = {}
        out = model.generator.synthesis(lat, lab, noise_mode='const', tap=tap)
    wsum, depth = tap['weight_sum'], out['image_depth']
    assert float(wsum.min()) >= 0.0 and float(wsum.max()) <= 1.0 + 1e-5
    assert float(depth.min()) >= cfg.ray_start - 1e-4 and float(depth.max()) <= cfg.ray_end + (cfg.ray_end - cfg.ray_start) / (cfg.depth_res - 1) + 1e-4
    ds = tap['depths_sorted']
    assert bool((ds[..., 1:] >= ds[..., :-1]).all())                      # sortedness of the merged sample list
    srt = tap['sort_idx'].long()
    assert bool((srt.sort(dim=-1).values == torch.arange(srt.shape[-1], device=srt.device)).all())   # a permutation
    assert pu.rel_err(out['image'], a) < 5e-4                           # eager synthesis == graph replay (same caveat)


def test_deterministic_mode_reproduces_frames_bit_for_bit():
    """ops.set_deterministic(): no split-K atomics, no fused-ToRGB atomics — two replays of the full-size frame graph
    (and a freshly captured second graph) give identical bits, within the parity tolerance of the default path."""
    import argparse
    from hfa_gp_b200 import ops
    from hfa_gp_b200.frame_loop import FrameLoop
    from hfa_gp_b200.networks.headnerf import HeadNeRF_final
    args = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='',
                              synthetic_generator=True, generator_seed=0)
    torch.manual_seed(0)
    model = HeadNeRF_final(args, 256, 'cuda', 512, 50, 'x', './').cuda().eval().requires_grad_(False)
    cfg = model.generator.cfg
    g = torch.Generator().manual_seed(11)
    rays = cfg.nrr ** 2
    model.generator.fixed_draws = (torch.rand(1, rays, cfg.depth_res, 1, generator=g).cuda(),
                                   torch.rand(rays, cfg.depth_res_importance, generator=g).cuda())
    img = (torch.rand(1, 3, 256, 256, generator=g) * 2 - 1).cuda()
    label = hfagp_ref.synthetic_labels(1, seed=5).cuda()
    default = FrameLoop(model, batch=1, size=256)(img, label.clone()).clone()
    prev = ops.set_deterministic(True)
    try:
        loop = FrameLoop(model, batch=1, size=256)
        a = loop(img, label.clone()).clone()
        b = loop(img, label.clone()).clone()
        c = FrameLoop(model, batch=1, size=256)(img, label.clone()).clone()
    finally:
        ops.set_deterministic(prev)
    assert torch.equal(a, b) and torch.equal(a, c)
    assert pu.rel_err(a, default) < 5e-4


def test_batched_replay_equals_frame_by_frame():
    """bench.py's default feeds FOUR independent frames per graph replay (FrameLoop(batch=4)) and keeps three replays in
    flight (FramePipeline): every frame must come out as it does from the reference loop's one-frame-per-call form —
    different frames, labels, per-frame ray jitter — up to the split-K summation order (layer shapes, hence K splits and
    the CTA-pair choice, depend on the batch)."""
    import argparse
    from hfa_gp_b200.frame_loop import FrameLoop, FramePipeline
    from hfa_gp_b200.networks.headnerf import HeadNeRF_final
    args = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='',
                              synthetic_generator=True, generator_seed=0)
    torch.manual_seed(0)
    model = HeadNeRF_final(args, 256, 'cuda', 512, 50, 'x', './').cuda().eval().requires_grad_(False)
    cfg = model.generator.cfg
    g = torch.Generator().manual_seed(21)
    rays = cfg.nrr ** 2
    B = 4
    jit = torch.rand(B, rays, cfg.depth_res, 1, generator=g).cuda()
    u = torch.rand(B * rays, cfg.depth_res_importance, generator=g).cuda()
    imgs = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).cuda()
    labels = hfagp_ref.synthetic_labels(B, seed=9).cuda()
    model.generator.fixed_draws = (jit, u)
    batched = FrameLoop(model, batch=B, size=256)(imgs, labels.clone()).clone()
    pipe = FramePipeline(model, depth=3, batch=B, size=256)
    outs = [pipe.submit(imgs, labels.clone())[0] for _ in range(3)]
    pipe.join()
    torch.cuda.synchronize()
    for o in outs:
        assert pu.rel_err(o, batched) < 5e-4
    worst = 0.0
    for i in range(B):
        model.generator.fixed_draws = (jit[i:i + 1], u[i * rays:(i + 1) * rays])
        one = FrameLoop(model, batch=1, size=256)(imgs[i:i + 1], labels[i:i + 1].clone()).clone()
        worst = max(worst, pu.rel_err(batched[i:i + 1], one))
    print(f'batched vs frame-by-frame: worst rel err {worst:.3e}')
    assert worst < 5e-4


@pytest.mark.parametrize('drive', ['3dmm', 'audio'])
def test_driven_frame_loops_match_oracle(drive):
    """configs[4] / run_recon_video_{3dmm,audio}.py: the graph-captured frame loop of the driven avatars
    (FrameLoop(drive=...)) against the oracle run end to end on the CPU — AudioNet -> AudioAttNet (8-frame window,
    zero-padded at the clip ends exactly as run_recon_video_audio.py:323-339) -> Weights_3DMM -> get_latent ->
    synthesis — at the first, an interior and the last frame of a short clip."""
    import argparse
    from hfa_gp_b200.frame_loop import FrameLoop, audio_windows
    from hfa_gp_b200.networks.headnerf import AudioAttNet, AudioNet, HeadNeRF_3DMM, HeadNeRF_Audio
    cfg = eg3d_ref.small14_config()
    plen = 76 if drive == '3dmm' else 64
    args = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='', params_len=plen,
                              synthetic_generator=True, generator_seed=0, generator_config=pu.product_config(cfg))
    torch.manual_seed(0)
    cls = HeadNeRF_3DMM if drive == '3dmm' else HeadNeRF_Audio
    model = cls(args, 64, 'cuda', 512, 50, 'x', './').cuda().eval().requires_grad_(False)
    ref_gen = eg3d_ref.make_generator(cfg, seed=0)
    ref_gen.load_state_dict({k: v.cpu() for k, v in model.generator.state_dict().items()})
    g = torch.Generator().manual_seed(11)
    rays = cfg.nrr ** 2
    jit, u = torch.rand(1, rays, cfg.depth_res, 1, generator=g), torch.rand(rays, cfg.depth_res_importance, generator=g)
    model.generator.fixed_draws = (jit.cuda(), u.cuda())
    head_sd = {k: v.cpu() for k, v in model.weights_3dmm.state_dict().items()}
    bases, delta = model.bases.detach().cpu(), model.delta.detach().cpu()

    def oracle_image(weights_in, label):
        with torch.no_grad():
            w = hfagp_ref.weights_3dmm_ref(head_sd, weights_in)
            lat = hfagp_ref.get_latent_ref(bases, delta, w)
            lab = hfagp_ref.flip_label_(label.clone())
            return ref_gen.synthesis(lat, lab, noise_mode='const', jitter_coarse=jit, u_fine=u)['image']

    if drive == '3dmm':
        loop = FrameLoop(model, batch=1, size=64, drive='3dmm', params_len=plen)
        for seed in (1, 2):
            params = torch.randn(1, plen, generator=g)
            label = hfagp_ref.synthetic_labels(1, seed=seed)
            got = loop(params.cuda(), label.cuda()).clone()
            assert pu.rel_err(got, oracle_image(params, label)) < pu.REL_TOL
        return
    torch.manual_seed(1)
    aud_net, aud_att = AudioNet(64, 16).cuda().eval(), AudioAttNet().cuda().eval()
    sd_net = {k: v.cpu() for k, v in aud_net.state_dict().items()}
    sd_att = {k: v.cpu() for k, v in aud_att.state_dict().items()}
    frames = 11
    auds = torch.randn(frames, 16, 29, generator=g)
    padded = audio_windows(auds.cuda(), 8)
    loop = FrameLoop(model, batch=1, size=64, drive='audio', aud_net=aud_net, aud_att=aud_att)
    for i in (0, 5, frames - 1):
        # the reference's window, literally (run_recon_video_audio.py:323-339)
        left, right = i - 4, i + 4
        pad_l, pad_r = max(-left, 0), max(right - frames, 0)
        win = auds[max(left, 0):min(right, frames)]
        win = torch.cat((torch.zeros_like(win)[:pad_l], win), 0) if pad_l else win
        win = torch.cat((win, torch.zeros_like(win)[:pad_r]), 0) if pad_r else win
        assert torch.equal(padded[i:i + 8].cpu(), win)
        with torch.no_grad():
            feats = hfagp_ref.audionet_ref(sd_net, win)
            smo = hfagp_ref.audioattnet_ref(sd_att, feats)
        label = hfagp_ref.synthetic_labels(1, seed=20 + i)
        got = loop(padded[i:i + 8], label.cuda()).clone()
        assert pu.rel_err(got, oracle_image(smo.unsqueeze(0), label)) < pu.REL_TOL


@pytest.mark.parametrize('res,s,sf', [(16, 48, 48), (8, 33, 20), (8, 64, 64)])
def test_bookkeeping_on_oracle_floats(res, s, sf):
    """north_star: "bit-exact for ray-index bookkeeping".  The integer stage of the render kernels (the shared device
    functions searchsorted_right / stable_ranks, run by hfagp_render_bookkeeping) is fed the ORACLE's own floats — its cdf,
    its u and its unsorted merged depths — and must return torch.searchsorted(right=True), the clamped below / above and
    the torch.sort permutation EXACTLY: no tie mask, no tolerance.  (The only freedom torch leaves is the order of exactly
    equal depths under its non-stable sort; there the permuted depths must still be identical.)"""
    cfg, planes, dec, c, jitter, u = _render_case(res, s, sf, 1, seed=2)
    tap = {}
    ro, rd = eg3d_ref.ray_sampler_ref(c[:, :16].view(-1, 4, 4), c[:, 16:25].view(-1, 3, 3), res)
    with torch.no_grad():
        eg3d_ref.render_ref(planes, dec, ro, rd, cfg, jitter, u, tap)
    rays = res * res
    cdf = tap['cdf'].reshape(rays, -1).contiguous()
    depths = torch.cat([tap['depths_coarse'], tap['depths_fine']], dim=-2).reshape(rays, s + sf).contiguous()
    ops = _ops()
    inds, below, above, order = ops.render_bookkeeping(cdf.cuda(), u[:, :sf].contiguous().cuda(), depths.cuda())
    assert torch.equal(inds.cpu().long(), tap['inds'])
    assert torch.equal(below.cpu().long(), tap['below'])
    assert torch.equal(above.cpu().long(), tap['above'])
    order = order.cpu().long()
    want = tap['sort_idx'].reshape(rays, s + sf)
    assert torch.equal(torch.gather(depths, 1, order), tap['depths_sorted'].reshape(rays, s + sf))
    neq = order != want
    if bool(neq.any()):          # only inside runs of exactly equal depths
        ds = torch.gather(depths, 1, want)
        tie = torch.zeros_like(neq)
        tie[:, 1:] |= ds[:, 1:] == ds[:, :-1]
        tie[:, :-1] |= ds[:, 1:] == ds[:, :-1]
        assert bool((~neq | tie).all())
    # and a stable sort reproduces torch.sort(stable=True) bit for bit, ties included
    dq = (depths * 64).round() / 64                       # force many exact ties
    _, _, _, oq = ops.render_bookkeeping(cdf.cuda(), u[:, :sf].contiguous().cuda(), dq.cuda())
    assert torch.equal(oq.cpu().long(), torch.sort(dq, dim=1, stable=True).indices)


def test_frame_pipeline_equals_single_loop():
    """FramePipeline (several frames in flight on several streams, one captured graph each) returns, frame for frame, what the
    single FrameLoop returns — to the host buffers too — with slots reused several times over."""
    import argparse
    from hfa_gp_b200.frame_loop import FrameLoop, FramePipeline
    from hfa_gp_b200.networks.headnerf import HeadNeRF_final
    cfg = eg3d_ref.small14_config()
    args = argparse.Namespace(out_pose=False, person_2=False, init=False, same_bases=False, run_id_2='',
                              synthetic_generator=True, generator_seed=0, generator_config=pu.product_config(cfg))
    torch.manual_seed(0)
    model = HeadNeRF_final(args, 64, 'cuda', 512, 50, 'x', './').cuda().eval().requires_grad_(False)
    g = torch.Generator().manual_seed(5)
    rays = cfg.nrr ** 2
    model.generator.fixed_draws = (torch.rand(1, rays, cfg.depth_res, 1, generator=g).cuda(),
                                   torch.rand(rays, cfg.depth_res_importance, generator=g).cuda())
    single = FrameLoop(model, batch=1, size=64)
    pipe = FramePipeline(model, depth=3, batch=1, size=64)
    frames = [(torch.rand(1, 3, 64, 64, generator=g) * 2 - 1).pin_memory() for _ in range(8)]
    labels = [hfagp_ref.synthetic_labels(1, seed=s).pin_memory() for s in range(8)]
    want = [single(f, l.clone()).clone() for f, l in zip(frames, labels)]
    hosts = [torch.empty(1, 3, 64, 64).pin_memory() for _ in range(8)]
    evs = [pipe.submit(f, l, host_out=h)[1] for f, l, h in zip(frames, labels, hosts)]
    pipe.join()
    torch.cuda.synchronize()
    for w, h, e in zip(want, hosts, evs):
        assert e.query()
        assert pu.rel_err(h, w) < 1e-4
