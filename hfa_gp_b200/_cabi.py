"""ctypes binding of ``libhfagp_sm100.so`` (the C ABI declared in ``include/hfagp.h``).

No fallbacks: if the shared library is missing or a call fails this raises.  Nothing here
imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# HFAGP_LIB: A/B-measure another build of the same ABI (profiling aid; the default is the in-tree library)
LIB_PATH = os.environ.get('HFAGP_LIB') or os.path.join(_HERE, 'libhfagp_sm100.so')
MAX_TAPS = 32
ACT_LINEAR, ACT_LRELU, ACT_RELU = 0, 1, 2

# every symbol include/hfagp.h declares (tests check the library exports all of them)
SYMBOLS = [
    'hfagp_abi_version', 'hfagp_last_error', 'hfagp_conv2d_fwd', 'hfagp_upfir_act_fwd',
    'hfagp_torgb_small_fwd', 'hfagp_styles_fwd', 'hfagp_modulate_fwd', 'hfagp_render_fwd',
    'hfagp_blur_fwd', 'hfagp_linear_fwd', 'hfagp_latent_fwd', 'hfagp_nchw_to_nhwc',
    'hfagp_nhwc_to_nchw', 'hfagp_conv2d_tc_fwd', 'hfagp_split_bf16', 'hfagp_modulate_split_fwd',
    'hfagp_blur_up', 'hfagp_act_bwd', 'hfagp_styles_bwd', 'hfagp_demod_bwd', 'hfagp_linear_bwd',
    'hfagp_conv2d_wgrad', 'hfagp_render_bwd', 'hfagp_latent_bwd', 'hfagp_facepool_fwd', 'hfagp_facepool_bwd',
    'hfagp_mse_fwd', 'hfagp_mse_bwd', 'hfagp_adam_step', 'hfagp_conv2d_tc_multi_fwd', 'hfagp_conv2d_tc_acc_fwd', 'hfagp_conv_epilogue_fwd', 'hfagp_frame_to_uint8', 'hfagp_frame_from_uint8', 'hfagp_frame_resize_u8', 'hfagp_stem_conv1x1_fwd',
    'hfagp_lpips_stem_fwd', 'hfagp_lpips_stem_bwd', 'hfagp_maxpool3s2_fwd', 'hfagp_maxpool3s2_bwd', 'hfagp_lpips_head_fwd',
    'hfagp_lpips_head_bwd', 'hfagp_modulate_split_multi_fwd', 'hfagp_conv2d_tc_rgb_fwd', 'hfagp_torgb_finalize_fwd', 'hfagp_conv2d_wgrad_mod', 'hfagp_render_bwd_dec',
    'hfagp_render_fwd_simt', 'hfagp_decoder_wgrad', 'hfagp_modconv_wgrad_finish', 'hfagp_set_device', 'hfagp_device_sm_count', 'hfagp_conv2d_tc_acc_workspace_bytes', 'hfagp_render_bwd_dec_workspace_bytes', 'hfagp_adam_sched', 'hfagp_adam_step_dev', 'hfagp_render_bookkeeping',
    'hfagp_torgb_small_mask_fwd', 'hfagp_pack_conv_weight', 'hfagp_unpack_conv_wgrad', 'hfagp_basis_qr_workspace_bytes', 'hfagp_basis_qr_fwd', 'hfagp_basis_qr_bwd', 'hfagp_basis_qr_info',
]


class ConvDesc(C.Structure):
    _fields_ = [
        ('batch', C.c_int32), ('in_h', C.c_int32), ('in_w', C.c_int32), ('cin', C.c_int32), ('cout', C.c_int32),
        ('oh', C.c_int32), ('ow', C.c_int32), ('in_stride', C.c_int32),
        ('out_h', C.c_int32), ('out_w', C.c_int32),
        ('out_stride', C.c_int32), ('out_off_y', C.c_int32), ('out_off_x', C.c_int32),
        ('ntaps', C.c_int32),
        ('dy', C.c_int32 * MAX_TAPS), ('dx', C.c_int32 * MAX_TAPS), ('wtap', C.c_int32 * MAX_TAPS),
        ('w_batch_stride', C.c_int64),
        ('act', C.c_int32), ('act_gain', C.c_float), ('clamp', C.c_float),
        ('noise_gain', C.c_float), ('residual_scale', C.c_float),
        ('up_h', C.c_int32), ('up_w', C.c_int32),
    ]


class RenderDesc(C.Structure):
    _fields_ = [
        ('batch', C.c_int32), ('res', C.c_int32), ('plane_h', C.c_int32), ('plane_w', C.c_int32),
        ('s_coarse', C.c_int32), ('s_fine', C.c_int32),
        ('delta', C.c_float), ('box_scale', C.c_float),
    ]


class ActBwdDesc(C.Structure):
    _fields_ = [
        ('batch', C.c_int32), ('h', C.c_int32), ('w', C.c_int32), ('c', C.c_int32),
        ('act', C.c_int32), ('act_gain', C.c_float), ('clamp', C.c_float), ('noise_gain', C.c_float),
        ('residual_scale', C.c_float), ('post_scale', C.c_float), ('rgb_k', C.c_int32),
    ]


class HfagpError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load the library once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _timed is not None:
        return _timed
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise HfagpError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                         f'(or `make -C hfa_gp_b200/csrc`). There is no CPU/PyTorch fallback for the hot path.')
    l = C.CDLL(LIB_PATH)
    l.hfagp_last_error.restype = C.c_char_p
    l.hfagp_abi_version.restype = C.c_int
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    l.hfagp_conv2d_fwd.argtypes = [C.POINTER(ConvDesc)] + [vp] * 9
    l.hfagp_upfir_act_fwd.argtypes = [i32, i32, i32, i32, vp, vp, vp, f32, vp, i32, f32, f32, vp, vp, vp, vp]
    l.hfagp_torgb_small_fwd.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, vp, vp, vp]
    l.hfagp_torgb_small_mask_fwd.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp]
    l.hfagp_styles_fwd.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    l.hfagp_modulate_fwd.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp]
    l.hfagp_render_fwd.argtypes = [C.POINTER(RenderDesc)] + [vp] * 16
    l.hfagp_render_fwd_simt.argtypes = [C.POINTER(RenderDesc)] + [vp] * 16
    l.hfagp_render_bwd.argtypes = [C.POINTER(RenderDesc)] + [vp] * 9
    l.hfagp_render_bwd_dec.argtypes = [C.POINTER(RenderDesc)] + [vp] * 11
    l.hfagp_decoder_wgrad.argtypes = [C.c_longlong] + [vp] * 5
    l.hfagp_modconv_wgrad_finish.argtypes = [i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp]
    l.hfagp_conv2d_tc_acc_workspace_bytes.argtypes = [C.POINTER(ConvDesc)]
    l.hfagp_conv2d_tc_acc_workspace_bytes.restype = C.c_size_t
    l.hfagp_render_bwd_dec_workspace_bytes.argtypes = [C.POINTER(RenderDesc), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    l.hfagp_render_bwd_dec_workspace_bytes.restype = C.c_size_t
    l.hfagp_blur_fwd.argtypes = [i32] * 7 + [f32] + [vp] * 7
    l.hfagp_pack_conv_weight.argtypes = [i32, i32, i32, vp, f32, vp, vp, vp, vp, vp, vp]
    l.hfagp_unpack_conv_wgrad.argtypes = [i32, i32, i32, i32, vp, vp, vp]
    l.hfagp_basis_qr_workspace_bytes.argtypes = [i32, i32]
    l.hfagp_basis_qr_workspace_bytes.restype = C.c_size_t
    l.hfagp_basis_qr_fwd.argtypes = [i32, i32, vp, f32, vp, vp, vp, i32, vp]
    l.hfagp_basis_qr_bwd.argtypes = [i32, i32, vp, vp, vp, vp, vp, i32, vp]
    l.hfagp_basis_qr_info.argtypes = [vp, i32, i32, C.POINTER(C.c_int), vp]
    l.hfagp_blur_up.argtypes = [i32] * 7 + [f32, vp, vp, vp]
    l.hfagp_act_bwd.argtypes = [C.POINTER(ActBwdDesc)] + [vp] * 23
    l.hfagp_styles_bwd.argtypes = [i32, i32, i32, i32] + [vp] * 8
    l.hfagp_demod_bwd.argtypes = [i32, i32, i32] + [vp] * 6
    l.hfagp_linear_bwd.argtypes = [i32, i32, i32, vp, vp, vp, f32, f32, vp, vp, vp, vp]
    l.hfagp_conv2d_wgrad.argtypes = [C.POINTER(ConvDesc)] + [vp] * 6 + [f32, vp, vp]
    l.hfagp_linear_fwd.argtypes = [i32, i32, i32, vp, vp, vp, f32, f32, vp, vp]
    l.hfagp_latent_fwd.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp]
    l.hfagp_latent_bwd.argtypes = [i32, i32, i32] + [vp] * 7
    l.hfagp_facepool_fwd.argtypes = [i32] * 5 + [vp] * 3
    l.hfagp_facepool_bwd.argtypes = [i32] * 5 + [vp] * 3
    l.hfagp_mse_fwd.argtypes = [C.c_longlong, vp, vp, f32, vp, vp]
    l.hfagp_mse_bwd.argtypes = [C.c_longlong, vp, vp, f32, vp, i32, vp, vp]
    l.hfagp_adam_step.argtypes = [C.c_longlong, vp, vp, vp, vp, f32] + [C.c_double] * 5 + [C.c_longlong, vp]
    l.hfagp_render_bookkeeping.argtypes = [i32] * 4 + [vp] * 8
    l.hfagp_adam_sched.argtypes = [C.c_double] * 3 + [C.c_longlong, vp]
    l.hfagp_adam_step_dev.argtypes = [C.c_longlong, vp, vp, vp, vp, f32] + [C.c_double] * 4 + [vp, vp]
    l.hfagp_nchw_to_nhwc.argtypes = [i32, i32, i32, i32, vp, vp, vp]
    l.hfagp_nhwc_to_nchw.argtypes = [i32, i32, i32, i32, vp, vp, vp]
    l.hfagp_conv2d_tc_fwd.argtypes = [C.POINTER(ConvDesc), vp, vp, vp, vp, i32] + [vp] * 9
    l.hfagp_conv2d_tc_multi_fwd.argtypes = [vp, i32, vp, vp, vp, vp, i32] + [vp] * 9
    l.hfagp_conv2d_tc_acc_fwd.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, vp, vp]
    l.hfagp_conv_epilogue_fwd.argtypes = [C.POINTER(ConvDesc)] + [vp] * 10
    l.hfagp_frame_to_uint8.argtypes = [C.c_longlong, vp, i32, vp, vp]
    l.hfagp_frame_from_uint8.argtypes = [i32, i32, i32, i32, vp, vp, vp]
    l.hfagp_stem_conv1x1_fwd.argtypes = [i32] * 5 + [vp, vp, vp, i32, f32, vp, vp, vp]
    l.hfagp_frame_resize_u8.argtypes = [i32] * 7 + [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    l.hfagp_lpips_stem_fwd.argtypes = [i32, i32, i32] + [vp] * 6
    l.hfagp_lpips_stem_bwd.argtypes = [i32, i32, i32] + [vp] * 4
    l.hfagp_maxpool3s2_fwd.argtypes = [i32] * 4 + [vp] * 7
    l.hfagp_maxpool3s2_bwd.argtypes = [i32] * 4 + [vp] * 6
    l.hfagp_lpips_head_fwd.argtypes = [i32, i32, i32] + [vp] * 6
    l.hfagp_lpips_head_bwd.argtypes = [i32, i32, i32] + [vp] * 7
    l.hfagp_modulate_split_multi_fwd.argtypes = [i32, i32] + [vp] * 10
    l.hfagp_conv2d_tc_rgb_fwd.argtypes = [C.POINTER(ConvDesc), vp, vp, vp, vp, i32] + [vp] * 7 + [i32, vp, vp]
    l.hfagp_torgb_finalize_fwd.argtypes = [i32, i32, i32, i32, vp, vp, f32, vp, vp, vp]
    l.hfagp_conv2d_wgrad_mod.argtypes = [C.POINTER(ConvDesc)] + [vp] * 8 + [f32, vp, vp]
    l.hfagp_split_bf16.argtypes = [C.c_longlong, vp, vp, vp, vp]
    l.hfagp_modulate_split_fwd.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    for s in SYMBOLS:
        fn = getattr(l, s)
        if s not in ('hfagp_last_error',):
            fn.restype = C.c_int
    if l.hfagp_abi_version() != 1:
        raise HfagpError(f'ABI version mismatch: library {l.hfagp_abi_version()}, binding 1')
    _lib = l
    return l


class _TimedLib:
    """Wraps the CDLL so that every entry point is bracketed by CUDA events on torch's current stream (the stream
    the kernels are enqueued on).  bench.py uses it for per-kernel durations measured live, outside any profiler."""

    def __init__(self, cdll, sink):
        self._cdll, self._sink = cdll, sink

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)
        if name in ('hfagp_last_error', 'hfagp_abi_version', 'hfagp_set_device', 'hfagp_device_sm_count',
                    'hfagp_conv2d_tc_acc_workspace_bytes', 'hfagp_render_bwd_dec_workspace_bytes') or not name.startswith('hfagp_'):
            return fn
        sink = self._sink

        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            sink.append((name, e0, e1))
            return rc
        return call


_timed: Optional[_TimedLib] = None


def start_timing() -> list:
    """Every C-ABI call from now on appends (entry point, start event, end event) to the returned list."""
    global _timed
    sink = []
    _timed = _TimedLib(lib(), sink)
    return sink


def stop_timing():
    global _timed
    _timed = None


def check(rc: int, what: str):
    if rc != 0:
        msg = (_lib or lib()).hfagp_last_error()
        raise HfagpError(f'{what} failed (rc={rc}): {msg.decode() if msg else "?"}')


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor on the CURRENT device (None passes NULL).  The element type is the
    callee's contract (fp32 unless the C signature says uint16_t / int32_t / uint8_t) and is not checked here."""
    if t is None:
        return None
    if not t.is_cuda:
        raise HfagpError('hot-path tensors must live on a CUDA device (no CPU fallback)')
    if t.device.index != torch.cuda.current_device():
        raise HfagpError(f'tensor lives on cuda:{t.device.index} but the current device (whose stream the kernels are '
                         f'enqueued on) is cuda:{torch.cuda.current_device()}')
    if not t.is_contiguous():
        raise HfagpError('hot-path tensors must be contiguous')
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream
