"""Frame egress / ingress kernels against the CPU oracle (oracle/frameio_ref.py, itself pinned to torchvision):
integer outputs bit-exact, float outputs bit-exact (same separately-rounded fp32 steps)."""
import pytest
import torch

from oracle import frameio_ref

pytestmark = pytest.mark.gpu


def _img(seed, n, h, w):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, h, w, generator=g) * 0.8
    m = min(8, x.numel())
    x.view(-1)[:m] = torch.tensor([-1.0, 1.0, 0.0, -1.5, 1.5, 0.999999, -0.999999, 1e-8])[:m]
    return x


@pytest.mark.parametrize('n,h,w', [(1, 512, 512), (2, 37, 53), (1, 1, 1)])
def test_to_uint8_bit_exact(n, h, w):
    from hfa_gp_b200 import frameio
    x = _img(3, n, h, w)
    # every multiple of 1/255 and 1/127.5 and its fp32 neighbours: the rounding boundaries of both conventions
    k = torch.arange(0, 256, dtype=torch.float32)
    edges = torch.cat([k / 127.5 - 1, (k - 128) / 127.5, (k + 0.5) / 127.5 - 1])
    edges = torch.cat([edges, torch.nextafter(edges, torch.tensor(9.0)), torch.nextafter(edges, torch.tensor(-9.0))])
    flat = x.view(-1)
    m = min(flat.numel(), edges.numel())
    flat[-m:] = edges[:m]
    for mode, ref in (('save_image', frameio_ref.save_image_uint8), ('layout_grid', frameio_ref.layout_grid_uint8)):
        want = ref(x)
        got_nchw = frameio.to_uint8(x.cuda(), mode)                                   # NCHW input
        got_view = frameio.to_uint8(x.permute(0, 2, 3, 1).contiguous().cuda().permute(0, 3, 1, 2), mode)   # get_image's view
        got_nhwc = frameio.to_uint8(x.permute(0, 2, 3, 1).contiguous().cuda(), mode, layout='nhwc')
        assert torch.equal(got_nhwc.cpu(), want), mode
        assert got_nchw.dtype == torch.uint8 and tuple(got_nchw.shape) == (n, h, w, 3)
        assert torch.equal(got_nchw.cpu(), want), mode
        assert torch.equal(got_view.cpu(), want), mode


def test_from_uint8_bit_exact_all_byte_values():
    from hfa_gp_b200 import frameio
    g = torch.Generator().manual_seed(4)
    u8 = torch.randint(0, 256, (2, 31, 45, 3), generator=g, dtype=torch.uint8)
    u8.view(-1)[:256] = torch.arange(256, dtype=torch.uint8)
    got = frameio.from_uint8(u8.cuda())
    assert torch.equal(got.cpu(), frameio_ref.to_tensor_normalize(u8))


def test_sink_and_feeder_round_trip():
    """Asynchronous rings: frames come back in order, bit-equal to the synchronous conversion; the feeder's output
    drives the encoder-side layout ([N,3,H,W] fp32)."""
    from hfa_gp_b200 import frameio
    sink = frameio.FrameSink(64, 48, depth=3, mode='layout_grid')
    imgs = [_img(10 + i, 1, 64, 48) for i in range(7)]
    got = []
    for i, im in enumerate(imgs):
        if i >= 3:
            got.append(sink.pop().copy())
        sink.push(im.cuda())
    got += sink.drain()
    assert len(got) == 7 and sink.pop() is None
    for im, g_ in zip(imgs, got):
        assert (torch.from_numpy(g_) == frameio_ref.layout_grid_uint8(im)).all()
    with pytest.raises(Exception):
        for im in imgs:
            sink.push(im.cuda())
    feeder = frameio.FrameFeeder(64, 48, depth=2)
    frames = [torch.from_numpy(g_) for g_ in got[:4]]
    outs = []
    for f in frames:
        feeder.push(f.numpy())
        outs.append(feeder.next())
    assert feeder.next() is None
    for f, o in zip(frames, outs):
        assert torch.equal(o.cpu(), frameio_ref.to_tensor_normalize(f))


@pytest.mark.gpu
def test_rings_survive_a_gpu_that_lags_the_host():
    """with the GPU held busy (a long spin kernel in front) the host wraps the rings several times; no frame
    may be corrupted or duplicated, and a popped frame must not change when later pushes reuse its slot."""
    from hfa_gp_b200 import frameio
    depth, n = 2, 9
    feeder = frameio.FrameFeeder(32, 32, depth=depth)
    g = torch.Generator().manual_seed(3)
    frames = [torch.randint(0, 256, (1, 32, 32, 3), dtype=torch.uint8, generator=g) for _ in range(n)]
    torch.cuda._sleep(int(2e8))                       # ~100 ms of GPU work queued ahead of everything below
    outs = []
    for f in frames:
        feeder.push(f.numpy())
        outs.append(feeder.next())                    # enqueued behind the spin: the GPU has consumed nothing yet
    torch.cuda.synchronize()
    for f, o in zip(frames, outs):
        assert torch.equal(o.cpu(), frameio_ref.to_tensor_normalize(f))
    sink = frameio.FrameSink(32, 32, depth=depth, mode='save_image')
    imgs = [_img(40 + i, 1, 32, 32).cuda() for i in range(n)]
    torch.cuda._sleep(int(2e8))
    popped = []
    for i, im in enumerate(imgs):
        if i >= depth:
            popped.append(sink.pop())
        sink.push(im)
    first = popped[0].copy()
    popped += sink.drain()
    assert (popped[0] == first).all()                 # later pushes wrapped onto its slot: the popped frame owns its bytes
    for im, p_ in zip(imgs, popped):
        assert (torch.from_numpy(p_) == frameio_ref.save_image_uint8(im.cpu())).all()


@pytest.mark.parametrize('n,h,w,size', [(2, 512, 512, 256), (1, 500, 500, 256), (1, 300, 420, 256), (1, 200, 200, 256),
                                        (1, 256, 256, 256), (3, 37, 53, 16)])
def test_resize_bit_exact(n, h, w, size):
    """hfagp_frame_resize_u8 (Pillow's two-pass fixed-point bilinear resampler on the device) against the oracle, which
    tests/test_oracle_frameio.py pins to PIL.Image.resize itself: uint8 output and the fused ToTensor + Normalize output."""
    from hfa_gp_b200 import frameio
    g = torch.Generator().manual_seed(n * 1000 + h + w)
    u8 = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
    u8.view(-1)[:4] = torch.tensor([0, 255, 1, 254], dtype=torch.uint8)
    oh, ow = frameio.resize_output_size(h, w, size)
    want = frameio_ref.resize_uint8(u8, oh, ow)
    got = frameio.resize_uint8(u8.cuda(), oh, ow)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (n, oh, ow, 3)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(frameio.ingest(u8.cuda(), size).cpu(), frameio_ref.to_tensor_normalize(want))


def test_feeder_with_resize_returns_the_reference_transform():
    """FrameFeeder(size=...) = decoded 512^2 uint8 frame -> Resize(256) -> ToTensor -> Normalize, on the device."""
    from hfa_gp_b200 import frameio
    g = torch.Generator().manual_seed(12)
    feeder = frameio.FrameFeeder(96, 96, batch=2, depth=2, size=48)
    frames = [torch.randint(0, 256, (2, 96, 96, 3), generator=g, dtype=torch.uint8) for _ in range(3)]
    for f in frames:
        feeder.push(f.numpy())
        got = feeder.next()
        want = frameio_ref.to_tensor_normalize(frameio_ref.resize_uint8(f, 48, 48))
        assert torch.equal(got.cpu(), want)
