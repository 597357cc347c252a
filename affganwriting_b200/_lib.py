"""ctypes binding of csrc/libaffgw.so (the C ABI declared in include/affgw.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(mirroring the reference's convention of Python asserts / exceptions, blocks.py:83,95,121,134,146,189-190).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libaffgw.so")

F32, BF16 = 0, 1
ACT = {"none": 0, "relu": 1, "lrelu": 2, "tanh": 3}
PAD = {"zero": 0, "reflect": 1, "replicate": 2}
ALGO_AUTO, ALGO_SIMT, ALGO_TC = 0, 1, 2
WLAYOUT_IM2COL, WLAYOUT_SHIFT = 1, 2
FMT_BF16, FMT_F16 = 0, 1


class ConvDesc(C.Structure):
    """Mirror of `affgw_conv_desc` (include/affgw.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "N", "H", "W", "Cin", "Cout", "KH", "KW", "stride", "pad", "pad_mode", "upsample", "Ho", "Wo",
        "in_pitch", "out_pitch", "pre_act", "post_act", "x_dtype", "w_dtype", "y_dtype", "algo", "passes", "grad_dtype", "stride_w", "operand_fmt")]


class PosFrame(C.Structure):
    """Mirror of `affgw_pos_frame` (include/affgw.h)."""
    _fields_ = [("N", C.c_int32), ("Hp", C.c_int32), ("Wp", C.c_int32), ("G", C.c_int32), ("lead", C.c_int32),
                ("reserved", C.c_int32), ("QA", C.c_int64)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float
_D = C.POINTER(ConvDesc)

# name -> argtypes (restype is int unless listed in _RESTYPE); every symbol include/affgw.h declares is here
SIGNATURES = {
    "affgw_version": [],
    "affgw_last_error": [],
    "affgw_launch_count": [],
    "affgw_device_ok": [],
    "affgw_pack_weight": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_pack_weight_tc": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_pack_weight_tc_bytes": [_I, _I, _I, _I, _I, _I, _I, _I],
    "affgw_maxpool3s2_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "affgw_maxpool3s2_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "affgw_u8_to_image": [_P, _P, _L, _P],
    "affgw_maxpool3_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_maxpool3_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_resize_bilinear_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_resize_bilinear_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_add_act": [_P, _P, _P, _I, _L, _I, _P],
    "affgw_conv_thin_supported": [_D],
    "affgw_conv_thin_ws_bytes": [_D, _I],
    "affgw_conv_thin_fwd": [_P, _P, _P, _P, _P, _D, _P],
    "affgw_conv_thin_dgrad": [_P, _P, _P, _P, _D, _P],
    "affgw_conv_thin_wgrad": [_P, _P, _P, _D, _P],
    "affgw_conv_tc_layout": [_D, _I],
    "affgw_conv_tc_tile_n": [_D, _I],
    "affgw_conv_tc_tile_m": [_D, _I],
    "affgw_conv_pos_frames": [_D, C.POINTER(PosFrame), C.POINTER(PosFrame)],
    "affgw_position_planes_bytes": [C.POINTER(PosFrame), _I],
    "affgw_split_positions": [_P, _I, _P, C.POINTER(PosFrame), _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "affgw_amax_scale": [_P, _L, _P, _P, _P],
    "affgw_split_positions_fmt": [_P, _I, _P, C.POINTER(PosFrame), _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    "affgw_split_planes_fmt": [_P, _I, _P, _L, _I, _I, _I, _I, _I, _I, _P, _P],
    "affgw_pack_weight_tc_fmt": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_conv2d_dgrad_scaled": [_P, _P, _P, _P, _P, _D, _P, _P],
    "affgw_conv2d_wgrad_scaled": [_P, _P, _P, _P, _D, _P, _P],
    "affgw_conv_tc_prefer_shift": [_I],
    "affgw_operand_planes_bytes": [_L, _I, _I],
    "affgw_split_planes": [_P, _I, _P, _L, _I, _I, _I, _I, _I, _P],
    "affgw_conv_tc_supported": [_D],
    "affgw_conv2d_fwd": [_P, _P, _P, _P, _P, _D, _P],
    "affgw_conv2d_dgrad_ws_bytes": [_D],
    "affgw_conv2d_dgrad": [_P, _P, _P, _P, _P, _D, _P],
    "affgw_conv2d_wgrad_ws_bytes": [_D],
    "affgw_conv2d_wgrad": [_P, _P, _P, _P, _D, _P],
    "affgw_colsum": [_P, _I, _P, _L, _I, _I, _P],
    "affgw_norm_stats": [_P, _I, _P, _P, _P, _P, _I, _L, _I, _F, _I, _P],
    "affgw_norm_apply": [_P, _I, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _P],
    "affgw_norm_bwd": [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _I, _P],
    "affgw_bn_update_running": [_P, _P, _P, _P, _P, _I, _F, _P],
    "affgw_bn_eval_stats": [_P, _P, _P, _P, _I, _F, _P],
    "affgw_maxpool2_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "affgw_maxpool2_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "affgw_avgpool3s2_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "affgw_avgpool3s2_bwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "affgw_resize_nearest_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_resize_nearest_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_gate_fwd": [_P, _P, _P, _P, _P, _I, _I, _L, _I, _P],
    "affgw_gate_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _I, _P],
    "affgw_gap_fwd": [_P, _P, _I, _I, _L, _I, _P],
    "affgw_bcast_add": [_P, _P, _P, _I, _I, _L, _I, _F, _P],
    "affgw_add2": [_P, _P, _P, _I, _L, _P],
    "affgw_act_bwd": [_P, _P, _P, _I, _L, _I, _P],
    "affgw_embedding_fwd": [_P, _P, _P, _I, _L, _I, _I, _P, _P],
    "affgw_embedding_bwd": [_P, _P, _P, _I, _L, _I, _I, _P],
    "affgw_text_tile_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_text_tile_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "affgw_bce_logits_fwd": [_P, _I, _F, _P, _L, _P],
    "affgw_bce_logits_bwd": [_P, _I, _F, _P, _P, _L, _P],
    "affgw_softmax_ce_fwd": [_P, _I, _P, _P, _I, _I, _P, _P],
    "affgw_softmax_ce_bwd": [_P, _I, _P, _P, _P, _I, _I, _P],
    "affgw_gru_cell_fwd": [_P, _L, _P, _P, _P, _I, _I, _P],
    "affgw_gru_cell_bwd": [_P, _P, _L, _P, _P, _P, _P, _P, _I, _I, _P],
    "affgw_scale_nc": [_P, _P, _P, _I, _L, _I, _P],
    "affgw_mul2": [_P, _P, _P, _L, _P],
    "affgw_map_seq": [_P, _P, _I, _I, _I, _I, _I, _P],
    "affgw_attn_energy_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "affgw_attn_energy_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "affgw_attn_ctx_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "affgw_attn_ctx_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "affgw_layernorm_fwd": [_P, _P, _P, _P, _L, _I, _F, _P],
    "affgw_gelu_fwd": [_P, _P, _L, _P],
    "affgw_scale_residual": [_P, _P, _P, _P, _L, _I, _P],
    "affgw_attention_fwd": [_P, _P, _I, _I, _I, _I, _F, _P],
    "affgw_blur3": [_P, _P, _I, _I, _I, _I, _P],
    "affgw_pixelnorm": [_P, _P, _I, _I, _F, _P],
    "affgw_label_smooth_kl_fwd": [_P, _P, _P, _I, _I, _I, _F, _P, _P],
    "affgw_label_smooth_kl_bwd": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "affgw_nchw_to_nhwc": [_P, _P, _I, _I, _I, _L, _I, _P],
    "affgw_nhwc_to_nchw": [_P, _P, _I, _I, _I, _L, _I, _P],
    "affgw_cast": [_P, _I, _P, _I, _L, _P],
    "affgw_concat_channels": [_P, _P, _P, _I, _L, _I, _I, _P],
    "affgw_split_channels": [_P, _P, _P, _I, _L, _I, _I, _P],
    "affgw_bucket_pack": [_P, _P, _P, _I, _P, _P],
    "affgw_bucket_unpack": [_P, _P, _P, _I, _P, _F, _P],
    "affgw_adam_step": [_P, _P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _L, _F, _P],
}
_RESTYPE = {"affgw_last_error": C.c_char_p, "affgw_launch_count": _L, "affgw_pack_weight_tc_bytes": _L, "affgw_operand_planes_bytes": _L, "affgw_position_planes_bytes": _L, "affgw_conv_thin_ws_bytes": _L,
            "affgw_conv2d_dgrad_ws_bytes": _L, "affgw_conv2d_wgrad_ws_bytes": _L}

_lib = None


def lib():
    """Load (once) and return the ctypes handle.  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libaffgw.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C affganwriting_b200/csrc`.  affganwriting_b200 has no CPU or PyTorch fallback.")
        h = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(h, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = h
    return _lib


def last_error():
    return lib().affgw_last_error().decode()


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"libaffgw {what} failed ({rc}): {last_error()}")


def launch_count():
    return int(lib().affgw_launch_count())


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"libaffgw supports float32 / bfloat16 tensors, got {t.dtype}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("affganwriting_b200 runs on CUDA tensors only (there is no CPU fallback)")


PROFILE = None      # list of (entry point, start event, end event, algorithmic bytes) while bench.py's instrumented pass runs


def call(name, *args, nbytes=0):
    """Invoke an entry point; raises on a non-zero return.  `nbytes` = the ALGORITHMIC HBM traffic of the call (every tensor
    read once + written once), recorded together with CUDA events on the launching stream when profiling is switched on."""
    if PROFILE is not None and nbytes:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(lib(), name)(*args), name)
        e1.record()
        PROFILE.append((name, e0, e1, int(nbytes)))
        return
    check(getattr(lib(), name)(*args), name)
