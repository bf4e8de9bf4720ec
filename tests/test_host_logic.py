"""CPU: host-side arithmetic that decides what the kernels are asked to do — tap lists, split-K / tile heuristics,
weight re-packings and the FLOP accounting of bench.py.  No compute call through the C ABI happens here."""
import importlib.util
import math
import os

import pytest
import torch
import torch.nn.functional as F

from hfa_gp_b200 import ops
from hfa_gp_b200.generator import GeneratorConfig, TriPlaneGenerator, _pack_conv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _conv_by_taps(x, w_packed, taps, oh, ow, in_stride=1, out_stride=1, out_off=(0, 0), out_hw=None):
    """Literal evaluation of the hfagp_conv2d_fwd contract (include/hfagp.h) in torch: x [C,H,W], w [taps][O][I]."""
    cin, h, w = x.shape
    cout = w_packed.shape[1]
    out_h, out_w = out_hw or (oh, ow)
    y = torch.zeros(cout, out_h, out_w, dtype=x.dtype)
    for my in range(oh):
        for mx in range(ow):
            acc = torch.zeros(cout, dtype=x.dtype)
            for dy, dx, t in taps:
                iy, ix = my * in_stride + dy, mx * in_stride + dx
                if 0 <= iy < h and 0 <= ix < w:
                    acc += w_packed[t] @ x[:, iy, ix]
            y[:, my * out_stride + out_off[0], mx * out_stride + out_off[1]] = acc
    return y


def test_parity_taps_reproduce_the_stride2_transposed_conv():
    """The four output-parity tap lists (ops._parity_taps) cover each of the 9 taps once and, evaluated by the
    conv contract with out_stride 2, equal F.conv_transpose2d(stride=2)."""
    seen = sorted(t for a in (0, 1) for b in (0, 1) for _, _, t in ops._parity_taps(a, b))
    assert seen == list(range(9))
    g = torch.Generator().manual_seed(0)
    cin, cout, h = 3, 2, 4
    x = torch.randn(cin, h, h, generator=g, dtype=torch.float64)
    wt = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv_transpose2d(x[None], wt.transpose(0, 1), stride=2)[0]
    packed = wt.permute(2, 3, 0, 1).reshape(9, cout, cin)
    out = torch.zeros(cout, 2 * h + 1, 2 * h + 1, dtype=torch.float64)
    for a in (0, 1):
        for b in (0, 1):
            out += _conv_by_taps(x, packed, ops._parity_taps(a, b), h + 1 - a, h + 1 - b, out_stride=2, out_off=(a, b),
                                 out_hw=(2 * h + 1, 2 * h + 1))
    assert torch.allclose(out, ref, atol=1e-12)


def test_pack_conv_and_3x3_taps_equal_conv2d():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 5, 6, generator=g, dtype=torch.float64)
    wt = torch.randn(3, 4, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(x[None], wt, padding=1)[0]
    packed = wt.permute(2, 3, 0, 1).reshape(9, 3, 4)
    got = _conv_by_taps(x, packed, ops.TAPS_3X3, 5, 6)
    assert torch.allclose(got, ref, atol=1e-12)
    assert torch.equal(_pack_conv(wt.float()), packed.float())
    assert torch.equal(_pack_conv(wt.float())[5], wt.float()[:, :, 1, 2])        # tap = ky*3 + kx, [O][I]


def test_lpips_stem_packing_equals_the_11x11_stride4_conv():
    """pack_stem_weight + the space-to-depth layout of hfagp_lpips_stem_fwd reproduce AlexNet's first convolution."""
    from hfa_gp_b200 import lpips
    g = torch.Generator().manual_seed(2)
    h = w = 32
    x = torch.randn(1, 3, h, w, generator=g, dtype=torch.float64)
    w1 = torch.randn(8, 3, 11, 11, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w1, stride=4, padding=2)[0]
    # the stem's layout, restated: out[Y][X][(py*4+px)*3+c] = x[c][4Y+py-2][4X+px-2] (zero outside)
    xp = F.pad(x[0], (2, 2, 2, 2))                                               # [3, h+4, w+4]
    s2d = xp.reshape(3, (h + 4) // 4, 4, (w + 4) // 4, 4).permute(2, 4, 0, 1, 3).reshape(48, (h + 4) // 4, (w + 4) // 4)
    packed = lpips.pack_stem_weight(w1.float()).double()
    # pack_stem_weight works in fp32; rebuild in fp64 with the same index map for an exact comparison
    wd = F.pad(w1, (0, 1, 0, 1)).reshape(8, 3, 3, 4, 3, 4).permute(2, 4, 0, 3, 5, 1).reshape(9, 8, 48)
    assert torch.allclose(packed, wd, atol=1e-6)
    got = _conv_by_taps(s2d, wd, lpips.TAPS_STEM, ref.shape[1], ref.shape[2])
    assert got.shape == ref.shape and torch.allclose(got, ref, atol=1e-10)
    assert lpips._mirror(lpips.TAPS_5X5)[0] == (2, 2, 0) and len(lpips.TAPS_5X5) == 25


def test_ksplit_heuristic():
    assert ops._ksplit(4, 512, ops.TAPS_3X3, 1) == 24              # b4/b8 conv1: 8 chunks x 3 tap groups
    assert ops._ksplit(32, 512, ops.TAPS_3X3, 1) == 4              # 32 tiles -> 128 CTAs
    assert ops._ksplit(128, 512, ops.TAPS_3X3, 1) == 1             # enough tiles already
    assert ops._ksplit(1, 64, ops.TAPS_1X1, 1) == 1                # a single K unit cannot be split
    assert ops._ksplit(2, 512, ops.TAPS_3X3, 2) == 72              # stride 2: every tap is its own group
    for tiles in range(1, 40):
        k = ops._ksplit(tiles, 512, ops.TAPS_3X3, 1)
        assert k >= 1 and k * tiles <= ops.NUM_SMS


def test_deterministic_mode_disables_split_k():
    """ops.set_deterministic(): no layer may take the split-K path (fp32 atomics); the previous setting is returned, and
    the SM count used by the heuristic can be given explicitly (a part with fewer SMs splits less)."""
    assert ops.deterministic() is False
    prev = ops.set_deterministic(True)
    try:
        assert prev is False and ops.deterministic() is True
        for tiles in (1, 4, 32):
            assert ops._ksplit(tiles, 512, ops.TAPS_3X3, 1) == 1
    finally:
        ops.set_deterministic(prev)
    assert ops._ksplit(4, 512, ops.TAPS_3X3, 1) == 24
    assert ops._ksplit(4, 512, ops.TAPS_3X3, 1, sms=64) == 16
    assert ops.num_sms() == ops.NUM_SMS or torch.cuda.is_available()


def test_bench_flop_accounting_matches_the_layer_table():
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cfg = GeneratorConfig()
    gen = TriPlaneGenerator(cfg)
    total = 0.0
    for kind, m, _ in gen._layers_in_order():
        k = m.weight.shape[-1]
        if m.cin % 8:
            continue
        if kind == 'torgb':
            if m.cout > 4:
                blk_res = next(getattr(gen.backbone.synthesis, f'b{r}').res for r in cfg.block_resolutions
                               if getattr(gen.backbone.synthesis, f'b{r}').torgb is m)
                total += 2.0 * blk_res ** 2 * m.cin * m.cout
            continue
        pixels = (m.res // 2) ** 2 if m.up == 2 else m.res ** 2        # the transposed conv does 9 MACs per INPUT pixel
        total += 2.0 * pixels * m.cin * m.cout * k * k
    # + encoder ResBlocks (size 256), as bench.py counts them
    channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256, 128: 128, 256: 64}
    r, c = 256, 64
    while r > 4:
        c2 = channels[r // 2]
        total += 2.0 * r * r * c * c * 9 + 2.0 * (r // 2) ** 2 * c * c2 * 10
        r, c = r // 2, c2
    assert math.isclose(bench.tensor_core_conv_flops(cfg, 256), total, rel_tol=1e-12)
    assert 300e9 < total < 340e9                                      # 302 G generator + 31 G encoder - SIMT bits


def test_frameio_modes_and_errors():
    from hfa_gp_b200 import frameio
    assert frameio.MODES == {'save_image': 0, 'layout_grid': 1}
    with pytest.raises(Exception):
        frameio.to_uint8(torch.zeros(1, 3, 4, 4), 'png')
    with pytest.raises(Exception):
        frameio.to_uint8(torch.zeros(1, 3, 4, 4), 'save_image')      # CPU tensor: no fallback


def test_bench_counts_every_conv_tc_entry_point():
    """bench.py's roofline divides the tensor-core FLOPs by the summed time of TC_ENTRY_POINTS: the list must hold
    every C-ABI entry point of conv_tc.cu that launches conv_tc_kernel (round-1 bug: the _rgb_ entry was missing,
    which dropped the two largest launches from the denominator), and each must be declared in include/hfagp.h."""
    import re
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    src = open(os.path.join(ROOT, 'hfa_gp_b200', 'csrc', 'conv_tc.cu')).read()
    entries = set(re.findall(r'extern "C" int (hfagp_conv2d_tc\w*_fwd)\s*\(', src))
    assert entries, 'no tensor-core entry points found in conv_tc.cu'
    assert entries == set(bench.TC_ENTRY_POINTS)
    header = open(os.path.join(ROOT, 'include', 'hfagp.h')).read()
    for e in entries:
        assert e in header


def test_audio_windows_equal_the_reference_slicing():
    """frame_loop.audio_windows: out[i:i+smo] must be the window run_recon_video_audio.py:323-339 /
    trainer_audio.py:68-84 builds for frame i (slice [i-half, i+half), zero-padded where it leaves the clip)."""
    from hfa_gp_b200.frame_loop import audio_windows
    g = torch.Generator().manual_seed(0)
    for frames, smo in ((10, 8), (3, 8), (9, 4)):
        auds = torch.randn(frames, 16, 29, generator=g)
        padded = audio_windows(auds, smo)
        half = int(smo / 2)
        for i in range(frames):
            left_i, right_i = i - half, i + half
            pad_left = pad_right = 0
            if left_i < 0:
                pad_left, left_i = -left_i, 0
            if right_i > frames:
                pad_right, right_i = right_i - frames, frames
            win = auds[left_i:right_i]
            if pad_left > 0:
                win = torch.cat((torch.zeros_like(win)[:pad_left], win), dim=0)
            if pad_right > 0:
                win = torch.cat((win, torch.zeros_like(win)[:pad_right]), dim=0)
            if win.shape[0] == smo:       # (a clip shorter than the window cannot fill it upstream either)
                assert torch.equal(padded[i:i + smo], win)


def test_lpips_refuses_to_run_without_weights(monkeypatch, tmp_path):
    """a drop-in training run must not silently optimise a random-feature loss.  LPIPS() raises unless pretrained
    weights are given (argument or HFAGP_LPIPS_WEIGHTS) or the run is declared synthetic; a saved state_dict loads by name."""
    from hfa_gp_b200.lpips import LPIPS
    monkeypatch.delenv('HFAGP_LPIPS_WEIGHTS', raising=False)
    monkeypatch.delenv('HFAGP_SYNTHETIC_LPIPS', raising=False)
    with pytest.raises(FileNotFoundError):
        LPIPS(net='alex')
    a = LPIPS(net='alex', seed=7, synthetic=True)
    assert a.synthetic
    path = tmp_path / 'lpips_alex.pt'
    torch.save(a.state_dict(), path)
    b = LPIPS(net='alex', seed=1, weights=str(path))
    assert not b.synthetic
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    monkeypatch.setenv('HFAGP_SYNTHETIC_LPIPS', '1')
    assert LPIPS(net='alex').synthetic


def test_zero_arena_counts_then_serves_zeroed_views():
    """ops.ZeroArena (the one memset of a captured training step): a measuring pass only counts, the next pass serves
    aligned, zeroed, non-overlapping views; large requests and other devices keep their own fill; nothing is served
    once the arena is ended (eager steps never see it)."""
    arena = ops.ZeroArena('cpu')
    arena.begin()
    a = ops.zeros((3, 5), 'cpu')
    b = ops.zeros((7,), torch.device('cpu'))
    big = ops.zeros((ops.ZERO_ARENA_MAX_ITEM // 4 + 1,), 'cpu')
    arena.end()
    assert arena.buf is None and arena.need == 256 + 256 and big.numel() * 4 > ops.ZERO_ARENA_MAX_ITEM
    assert a.untyped_storage().data_ptr() != b.untyped_storage().data_ptr()        # measuring: plain torch.zeros
    arena.materialise()
    arena.buf.fill_(255)                                     # stale contents of the previous step
    arena.begin()
    a = ops.zeros((3, 5), 'cpu')
    b = ops.zeros((7,), 'cpu')
    c = ops.zeros((1000,), 'cpu')                            # more than was measured: falls back, still zero
    arena.end()
    assert a.shape == (3, 5) and b.shape == (7,) and a.dtype == torch.float32
    assert float(a.abs().sum()) == 0.0 and float(b.abs().sum()) == 0.0 and float(c.abs().sum()) == 0.0
    base = arena.buf.data_ptr()
    assert a.data_ptr() == base and b.data_ptr() == base + 256       # 256-byte slots (the CUDA allocator aligns the base)
    assert not (base <= c.data_ptr() < base + arena.buf.numel())
    a.fill_(1.0)
    assert float(b.abs().sum()) == 0.0                       # no overlap
    outside = ops.zeros((3, 5), 'cpu')
    assert not (base <= outside.data_ptr() < base + arena.buf.numel())


def test_exact_conv1d_equals_conv1d_and_keeps_the_state_dict():
    """The audio nets' Conv1d layers as unfold + matmul (networks/headnerf.py:_ExactConv1d) against F.conv1d, forward
    and both gradients, for the two geometries the reference uses (headnerf.py:284-349)."""
    from hfa_gp_b200.networks.headnerf import AudioAttNet, AudioNet, _ExactConv1d
    torch.manual_seed(0)
    for cin, cout, stride, length in [(29, 32, 2, 16), (32, 16, 1, 8), (2, 1, 1, 8)]:
        m = _ExactConv1d(cin, cout, kernel_size=3, stride=stride, padding=1)
        x = torch.randn(5, cin, length, requires_grad=True)
        y, yr = m(x), F.conv1d(x, m.weight, m.bias, stride=stride, padding=1)
        assert y.shape == yr.shape and float((y - yr).abs().max()) < 2e-6
        g = torch.randn_like(y)
        gx, gw, gb = torch.autograd.grad((y * g).sum(), [x, m.weight, m.bias])
        gxr, gwr, gbr = torch.autograd.grad((yr * g).sum(), [x, m.weight, m.bias])
        assert float((gx - gxr).abs().max()) < 1e-5 and float((gw - gwr).abs().max()) < 1e-5 and float((gb - gbr).abs().max()) < 1e-5
    keys = set(AudioNet(64, 16).state_dict()) | set(AudioAttNet(64, 8).state_dict())
    assert 'encoder_conv.0.weight' in keys and 'attentionConvNet.8.bias' in keys and 'attentionNet.0.weight' in keys


def test_grad_slot_only_accepts_dense_fp32_leaf_gradients():
    from hfa_gp_b200.autograd import _grad_slot
    p = torch.nn.Parameter(torch.randn(4, 3))
    assert not _grad_slot(p)                                 # no gradient storage yet: the Function returns a tensor
    flat = torch.zeros(12)
    p.grad = flat.view(4, 3)                                 # FlatAdam's binding
    assert _grad_slot(p)
    p.grad = torch.zeros(3, 4).t()
    assert not _grad_slot(p)                                 # not contiguous
    assert not _grad_slot(p * 2.0)                           # not a leaf
    assert not _grad_slot(None)
