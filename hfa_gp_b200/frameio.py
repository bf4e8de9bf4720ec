"""Frame egress / ingress around the render path.

Reference, per frame of ``run_recon_video_rgb.py``:

  egress   ``torchvision.utils.save_image(img_recon, path, normalize=True, range=(-1,1))`` (:233-234) — clamp, min-max
           normalise, ``*255+0.5`` -> uint8, PNG-encode, all synchronously on the render stream; and
           ``layout_grid`` (:26-40, ``(img*127.5+128).clamp(0,255).to(uint8)``) for the mp4 writer
  ingress  ``Image.open -> Resize(size) -> ToTensor -> Normalize(0.5,0.5)`` (``train_rgb.py:78-81``, ``dataset.py:205-214``)

Here the float<->uint8 conversions are sm_100a kernels (bit-exact with torch, ``hfagp_frame_to_uint8`` /
``hfagp_frame_from_uint8``) and the PCIe copies are asynchronous on a side stream through a ring of pinned buffers, so
the render stream never waits for the host: 0.79 MB per 512^2 frame crosses PCIe instead of 3.1 MB.  File encoding
(PNG / mp4) and decoding stay host-side library work, outside the hot path.
"""
from __future__ import annotations

from typing import List, Optional

import math

import torch

from . import _cabi
from ._cabi import HfagpError, check, ptr, stream
from . import ops

MODES = {'save_image': 0, 'layout_grid': 1}


def to_uint8(image: torch.Tensor, mode: str = 'save_image', out: Optional[torch.Tensor] = None,
             layout: str = 'nchw') -> torch.Tensor:
    """image [N,3,H,W] (``layout='nchw'``, as ``get_image`` returns it: an NCHW tensor or an NCHW view of the
    channels-last image) or channels-last [N,H,W,3] (``layout='nhwc'``) -> uint8 [N,H,W,3] on the device."""
    if mode not in MODES:
        raise HfagpError(f'unknown uint8 convention {mode!r}; use one of {sorted(MODES)}')
    if image.dim() != 4 or layout not in ('nchw', 'nhwc'):
        raise HfagpError("to_uint8 expects a 4-D image batch and layout 'nchw' | 'nhwc'")
    if layout == 'nchw':
        nhwc = image.permute(0, 2, 3, 1)
        if not nhwc.is_contiguous() or image.shape[2] * image.shape[3] == 1:
            nhwc = ops.nchw_to_nhwc(image.detach().float().contiguous())
    else:
        nhwc = image
    nhwc = nhwc.detach().float().contiguous()
    if out is None:
        out = torch.empty(nhwc.shape, device=nhwc.device, dtype=torch.uint8)
    ops._ok(_cabi.lib().hfagp_frame_to_uint8(nhwc.numel(), ptr(nhwc), MODES[mode], ptr(out), stream()),
            'hfagp_frame_to_uint8')
    return out


def from_uint8(frames: torch.Tensor) -> torch.Tensor:
    """uint8 [N,H,W,C] (decoded RGB) on the device -> fp32 [N,C,H,W] in [-1,1] (ToTensor + Normalize(0.5,0.5))."""
    if frames.dtype != torch.uint8 or frames.dim() != 4:
        raise HfagpError('from_uint8 expects uint8 [N,H,W,C]')
    n, h, w, c = frames.shape
    out = torch.empty((n, c, h, w), device=frames.device, dtype=torch.float32)
    ops._ok(_cabi.lib().hfagp_frame_from_uint8(n, h, w, c, ptr(frames.contiguous()), ptr(out), stream()),
            'hfagp_frame_from_uint8')
    return out


_TABLES = {}


def pil_bilinear_table(in_size: int, out_size: int):
    """Pillow's coefficient table for resampling one axis from ``in_size`` to ``out_size`` with the bilinear filter:
    ``precompute_coeffs`` (double arithmetic, filter support scaled by the down-sampling factor, weights normalised to
    sum 1) followed by ``normalize_coeffs_8bpc`` (22-bit fixed point, round half away from zero) — src/libImaging/
    Resample.c.  Returns ``(ksize, bounds [out,2] int32 = (first, count), coeffs [out,ksize] int32)`` as CPU tensors."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale                                # bilinear: filter support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    bounds = torch.zeros((out_size, 2), dtype=torch.int32)
    coeffs = torch.zeros((out_size, ksize), dtype=torch.int32)
    one = 1 << 22
    for o in range(out_size):
        center = (o + 0.5) * scale
        first = max(int(center - support + 0.5), 0)
        count = min(int(center + support + 0.5), in_size) - first
        w = [max(1.0 - abs((k + first - center + 0.5) * inv), 0.0) for k in range(count)]
        tot = 0.0
        for v in w:
            tot += v
        for k in range(count):
            v = w[k] / tot if tot != 0.0 else w[k]
            coeffs[o, k] = int(0.5 + v * one) if v >= 0 else int(-0.5 + v * one)
        bounds[o, 0], bounds[o, 1] = first, count
    return ksize, bounds, coeffs


def _table(in_size, out_size, device):
    key = (in_size, out_size, str(device))
    if key not in _TABLES:
        ksize, b, c = pil_bilinear_table(in_size, out_size)
        _TABLES[key] = (ksize, b.to(device), c.to(device))
    return _TABLES[key]


def resize_output_size(h: int, w: int, size: int):
    """``transforms.Resize(int)``: the shorter side becomes ``size``, the other keeps the aspect ratio."""
    return (int(size * h / w), size) if w <= h else (size, int(size * w / h))


def resize_uint8(frames: torch.Tensor, out_h: int, out_w: int, normalized: bool = False) -> torch.Tensor:
    """uint8 [N,H,W,C] decoded frames on the device -> what ``PIL.Image.resize((out_w, out_h), BILINEAR)`` gives for each
    of them, bit for bit: uint8 [N,out_h,out_w,C], or with ``normalized`` the ToTensor + Normalize(0.5, 0.5) of that as
    fp32 [N,C,out_h,out_w] (the encoder's input) straight from the vertical pass (``hfagp_frame_resize_u8``)."""
    if frames.dtype != torch.uint8 or frames.dim() != 4 or not frames.is_cuda:
        raise HfagpError('resize_uint8 expects a CUDA uint8 tensor [N,H,W,C]')
    frames = frames.contiguous()
    n, h, w, c = frames.shape
    dev = frames.device
    kh, bh, ch_ = _table(w, out_w, dev) if out_w != w else (0, None, None)
    kv, bv, cv = _table(h, out_h, dev)
    tmp = torch.empty((n, h, out_w, c), device=dev, dtype=torch.uint8) if out_w != w else None
    if normalized:
        y8, yf = None, torch.empty((n, c, out_h, out_w), device=dev, dtype=torch.float32)
    else:
        y8, yf = torch.empty((n, out_h, out_w, c), device=dev, dtype=torch.uint8), None
    ops._ok(_cabi.lib().hfagp_frame_resize_u8(n, h, w, c, out_h, out_w, kh, ptr(bh), ptr(ch_), kv, ptr(bv), ptr(cv), ptr(frames),
                                              ptr(tmp), ptr(y8), ptr(yf), stream()), 'hfagp_frame_resize_u8')
    return yf if normalized else y8


def ingest(frames: torch.Tensor, size: int) -> torch.Tensor:
    """The reference's frame transform ``Resize(size) -> ToTensor -> Normalize([0.5]*3, [0.5]*3)`` (train_rgb.py:78-81,
    run_recon_video_3dmm.py:258-261) on decoded uint8 frames [N,H,W,3] already on the device -> fp32 [N,3,h',w']."""
    n, h, w, c = frames.shape
    oh, ow = resize_output_size(h, w, size)
    return resize_uint8(frames, oh, ow, normalized=True)


class FrameSink:
    """Asynchronous egress: ``push(image)`` converts on the render stream, then copies device->pinned host on a side
    stream; ``pop()`` hands back the oldest finished frame as a ``[N,H,W,3]`` uint8 numpy array that OWNS its bytes (a
    copy out of the pinned slot: the slot itself is reused by the next ``push`` that wraps onto it, by asynchronous DMA).
    Slot reuse is ordered on the device as well: the conversion into a device slot waits for the previous device->host
    copy out of it."""

    def __init__(self, height: int, width: int, channels: int = 3, batch: int = 1, depth: int = 4,
                 mode: str = 'save_image', device=None):
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.mode, self.depth = mode, depth
        shape = (batch, height, width, channels)
        self.dev = [torch.empty(shape, device=self.device, dtype=torch.uint8) for _ in range(depth)]
        self.host = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.done = [None] * depth                     # device->host copy of the slot's current frame
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.head = self.tail = 0

    def push(self, image: torch.Tensor):
        if self.head - self.tail >= self.depth:
            raise HfagpError('FrameSink is full: pop() finished frames before pushing more')
        s = self.head % self.depth
        if self.done[s] is not None:                   # the previous copy out of dev[s] must have read it
            torch.cuda.current_stream(self.device).wait_event(self.done[s])
        to_uint8(image, self.mode, out=self.dev[s])
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            self.host[s].copy_(self.dev[s], non_blocking=True)
            self.done[s] = torch.cuda.Event()
            self.done[s].record()
        self.head += 1

    def pop(self):
        if self.tail == self.head:
            return None
        s = self.tail % self.depth
        self.done[s].synchronize()
        frame = self.host[s].numpy().copy()            # the slot is free for the next push from here on
        self.tail += 1
        return frame

    def drain(self) -> List:
        out = []
        while self.tail < self.head:
            out.append(self.pop())
        return out


class FrameFeeder:
    """Asynchronous ingress: decoded uint8 frames ``[N,H,W,3]`` (numpy or CPU tensor) are staged in pinned memory,
    copied host->device on a side stream and (resized and) normalised on the device; ``next()`` returns fp32 ``[N,3,H,W]``.
    A slot is reused only when its previous occupants are done with it: the host write waits for the previous
    host->device copy out of the pinned buffer, and the copy into the device buffer waits (on the device) for the
    ``from_uint8`` kernel that last read it — a GPU that lags the host cannot see frames overwritten."""

    def __init__(self, height: int, width: int, channels: int = 3, batch: int = 1, depth: int = 4, device=None,
                 size: Optional[int] = None):
        """``size``: also apply the reference's ``transforms.Resize(size)`` (Pillow bilinear, bit-exact) on the device, so
        the decoded frames cross PCIe as they come out of the decoder and ``next()`` returns the encoder's input."""
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        shape = (batch, height, width, channels)
        self.depth, self.size = depth, size
        self.host = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.dev = [torch.empty(shape, device=self.device, dtype=torch.uint8) for _ in range(depth)]
        self.ready = [None] * depth                    # host->device copy of the slot's current frame
        self.consumed = [None] * depth                 # from_uint8 of the slot's previous frame (main stream)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.head = self.tail = 0

    def push(self, frames):
        if self.head - self.tail >= self.depth:
            raise HfagpError('FrameFeeder is full: consume frames with next() first')
        s = self.head % self.depth
        if self.ready[s] is not None:
            self.ready[s].synchronize()                # the DMA out of host[s] has finished: safe to overwrite it
        self.host[s].copy_(torch.as_tensor(frames).reshape(self.host[s].shape))
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[s] is not None:
                self.copy_stream.wait_event(self.consumed[s])   # the kernel that read dev[s] has finished
            self.dev[s].copy_(self.host[s], non_blocking=True)
            self.ready[s] = torch.cuda.Event()
            self.ready[s].record()
        self.head += 1

    def next(self) -> Optional[torch.Tensor]:
        if self.tail == self.head:
            return None
        s = self.tail % self.depth
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[s])
        out = from_uint8(self.dev[s]) if self.size is None else ingest(self.dev[s], self.size)
        self.consumed[s] = torch.cuda.Event()
        self.consumed[s].record(cur)
        self.tail += 1
        return out
