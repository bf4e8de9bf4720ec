"""The body of the reference's inference frame loop as one replayable CUDA graph.

Reference (``/root/reference/code/run_recon_video_rgb.py:216-236``), once per video frame::

    generated_weights = gen.get_weights(real_image)
    latent            = gen.get_latent(generated_weights)
    generated_image   = gen.get_image(latent, label)

At batch 1 this is ~130 kernel launches of 5-100 us each; driven from Python the GPU idles between them.
``FrameLoop`` captures the three calls once (``torch.cuda.graph``; every launch goes through the C ABI on
torch's current stream, so capture sees all of them) and replays the graph per frame.  The two random draws
of the renderer stay inside the graph (torch's graph-safe Philox state advances on every replay), the
label is flipped in place exactly as ``get_image`` does (``headnerf.py:132``), and the result is written to a
static output buffer that is valid until the next call.

The driven avatars replay the same way (``drive=``): ``'3dmm'`` feeds expression parameters ``[B, params_len]`` to
``HeadNeRF_3DMM`` (``run_recon_video_3dmm.py``: ``gen(params, label)``), ``'audio'`` feeds the ``smo_size``-frame window of
DeepSpeech features ``[smo_size,16,29]`` through ``AudioNet`` -> ``AudioAttNet`` -> ``gen(aud_smo, label)``
(``run_recon_video_audio.py:309-351``, the ``global_step >= nosmo_iters`` branch), all inside the captured graph;
:func:`audio_windows` builds the zero-padded windows exactly as the reference slices and pads ``auds``.
"""
from __future__ import annotations

import torch

from . import ops
from ._cabi import HfagpError


class FrameLoop:
    def __init__(self, model, batch: int = 1, size: int = 256, device=None, use_graph: bool = True, warmup: int = 3,
                 egress: str = None, drive: str = 'image', params_len: int = None, aud_net=None, aud_att=None,
                 smo_size: int = 8, win_size: int = 16):
        """``egress``: None -> the fp32 image of ``get_image``; 'save_image' | 'layout_grid' -> additionally convert to the
        uint8 [B,H,W,3] frame the reference writes out (hfa_gp_b200.frameio.to_uint8, inside the graph), in ``out_u8``.
        ``drive``: 'image' (HeadNeRF_final: frame [B,3,size,size]), '3dmm' (HeadNeRF_3DMM: params [B,params_len]) or
        'audio' (HeadNeRF_Audio + ``aud_net`` / ``aud_att``: feature window [smo_size,16,29], batch 1 as upstream)."""
        if drive not in ('image', '3dmm', 'audio'):
            raise HfagpError(f'unknown drive mode {drive!r}')
        if drive == 'audio' and (aud_net is None or aud_att is None or batch != 1):
            raise HfagpError("drive='audio' needs aud_net, aud_att and batch 1 (trainer_audio.py:66-78 indexes one frame)")
        if drive == '3dmm' and not params_len:
            raise HfagpError("drive='3dmm' needs params_len")
        self.model = model
        self.drive, self.aud_net, self.aud_att = drive, aud_net, aud_att
        self.egress = egress
        self.out_u8 = None
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        if dev.type != 'cuda':
            raise HfagpError('FrameLoop needs a CUDA model (there is no CPU fallback)')
        self.device = dev
        if drive == 'image':
            self.image = torch.zeros(batch, 3, size, size, device=dev)
        elif drive == '3dmm':
            self.image = torch.zeros(batch, params_len, device=dev)
        else:
            self.image = torch.zeros(smo_size, win_size, 29, device=dev)
        self.label = torch.zeros(batch, 25, device=dev)
        self.label[:, [0, 5, 10, 15]] = 1.0          # a valid camera for the warm-up frames
        self.label[:, 11] = 2.7
        self.label[:, 16:] = torch.tensor([4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1.0], device=dev)
        self.out = None
        self.graph = None
        self.launches_per_replay = None
        if use_graph:
            self._capture(warmup)

    def _body(self):
        m = self.model
        if self.drive == 'audio':
            feats = self.aud_net(self.image)                  # [smo_size, dim_aud]
            w = self.aud_att(feats)                           # [dim_aud] attention-smoothed feature of the centre frame
            w = m.get_weights(w.unsqueeze(0) if w.dim() == 1 else w)
        else:
            w = m.get_weights(self.image)
        if isinstance(w, tuple):          # out_pose models return (weights, pose)
            w = w[0]
        lat = m.get_latent(w)
        img = m.get_image(lat, self.label)
        if self.egress is not None:
            from . import frameio
            self.out_u8 = frameio.to_uint8(img, self.egress, out=self.out_u8)
        return img

    def _capture(self, warmup):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        keep = self.label.clone()
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):          # fills every host-side cache (packed weights, QR, TMA maps)
                self.label.copy_(keep)
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.label.copy_(keep)
        g = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.no_grad(), torch.cuda.graph(g):
            self.out = self._body()
        self.launches_per_replay = ops.launch_count() - n0     # kernels of libhfagp_sm100.so inside the graph
        self.graph = g

    def __call__(self, image: torch.Tensor, label: torch.Tensor, mutate_label: bool = True) -> torch.Tensor:
        """image [B,3,S,S] (or params [B,P] / audio window [smo,16,29], see ``drive``), label [B,25] (host-pinned or
        device) -> image [B,3,512,512] (static buffer)."""
        self.image.copy_(image, non_blocking=True)
        self.label.copy_(label, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                self.out = self._body()
        if mutate_label and label.is_cuda:
            label.copy_(self.label, non_blocking=True)   # callers see the in-place GL flip (headnerf.py:132)
        return self.out


def audio_windows(auds: torch.Tensor, smo_size: int = 8) -> torch.Tensor:
    """``auds [F,16,29]`` -> zero-padded copy ``[F + smo_size, 16, 29]`` such that ``out[i : i + smo_size]`` is the window
    the reference builds for frame ``i`` (``run_recon_video_audio.py:323-339`` / ``trainer_audio.py:68-84``:
    ``auds[i - half : i + half]`` with zeros where the range leaves ``[0, F)``)."""
    half = int(smo_size / 2)
    z = torch.zeros((half,) + tuple(auds.shape[1:]), dtype=auds.dtype, device=auds.device)
    return torch.cat((z, auds, z), dim=0)


class FramePipeline:
    """``depth`` frames in flight: ``depth`` FrameLoops (each its own captured graph and static buffers) on ``depth`` CUDA
    streams, frames dealt round-robin.  Every frame is still one batch-1 graph replay of the reference's loop body; what
    overlaps is frame i+1's host->device copy and its launch-sized kernels (the encoder and the 4^2..32^2 generator blocks
    leave most SMs idle at batch 1) with frame i's large convolutions and its device->host copy.

    ``submit(image, label, host_out=None)`` enqueues one frame and returns ``(image_out, event)``: ``image_out`` is the
    slot's static output buffer (valid until the slot is used again, ``depth`` submits later); when ``host_out`` (pinned)
    is given the result is also copied there on the slot's stream.  ``event`` completes when the frame (and its copy) is done.
    ``join()`` makes the current stream wait for everything submitted."""

    def __init__(self, model, depth: int = 2, device=None, **loop_kwargs):
        if depth < 1:
            raise HfagpError('FramePipeline needs depth >= 1')
        self.loops = [FrameLoop(model, device=device, **loop_kwargs) for _ in range(depth)]
        self.device = self.loops[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
        self.depth, self.count = depth, 0
        self.launches_per_replay = self.loops[0].launches_per_replay

    def submit(self, image, label, host_out=None):
        k = self.count % self.depth
        s = self.streams[k]
        s.wait_stream(torch.cuda.current_stream(self.device))        # inputs may have been produced on the caller's stream
        with torch.cuda.stream(s):
            out = self.loops[k](image, label, mutate_label=False)
            if host_out is not None:
                host_out.copy_(out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s)
        self.count += 1
        return out, ev

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)
