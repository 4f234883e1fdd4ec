"""ctypes binding of libhupr_b200.so — the C ABI declared in include/hupr_b200.h.

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  Tensors are passed as raw device pointers + the current CUDA stream.
"""
import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libhupr_b200.so")
_lib = None

class ConvDesc(ctypes.Structure):
    """Mirror of ``hupr_conv_desc`` (include/hupr_b200.h) — field order must match the C struct."""
    _fields_ = [
        ("a_hi", ctypes.c_void_p), ("a_lo", ctypes.c_void_p),
        ("n", ctypes.c_int), ("d", ctypes.c_int), ("h", ctypes.c_int), ("w", ctypes.c_int), ("ca", ctypes.c_int),
        ("a_ch_off", ctypes.c_int), ("cin", ctypes.c_int),
        ("w_hi", ctypes.c_void_p), ("w_lo", ctypes.c_void_p),
        ("cout", ctypes.c_int),
        ("kd", ctypes.c_int), ("kh", ctypes.c_int), ("kw", ctypes.c_int),
        ("pd", ctypes.c_int), ("ph", ctypes.c_int), ("pw", ctypes.c_int),
        ("w_batched", ctypes.c_int),
        ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p), ("slope", ctypes.c_void_p),
        ("r_hi", ctypes.c_void_p), ("r_lo", ctypes.c_void_p), ("r_ld", ctypes.c_int), ("r_ch_off", ctypes.c_int),
        ("o_hi", ctypes.c_void_p), ("o_lo", ctypes.c_void_p), ("o_ld", ctypes.c_int), ("o_ch_off", ctypes.c_int),
        ("o_f32", ctypes.c_void_p), ("o_f32_ld", ctypes.c_int),
        ("w_ld", ctypes.c_int), ("w_ch_off", ctypes.c_int),
        ("a_n_stride", ctypes.c_longlong),
        ("k_split", ctypes.c_int), ("w_k_off", ctypes.c_int),
        ("row_vec", ctypes.c_void_p), ("row_mode", ctypes.c_int),
        ("ws", ctypes.c_void_p), ("ws_bytes", ctypes.c_size_t),
        ("no_tma_store", ctypes.c_int),
        ("nprod", ctypes.c_int),
        ("stats", ctypes.c_void_p), ("stats_ld", ctypes.c_int),
        ("a_q16", ctypes.c_void_p), ("a_q8", ctypes.c_void_p),
        ("w_q16", ctypes.c_void_p), ("w_q8", ctypes.c_void_p),
        ("o_q16", ctypes.c_void_p), ("o_q8", ctypes.c_void_p),
        ("q_sat", ctypes.c_void_p),
    ]


class AttnDesc(ctypes.Structure):
    """Mirror of ``hupr_attn_desc``."""
    _fields_ = [
        ("q_hi", ctypes.c_void_p), ("q_lo", ctypes.c_void_p), ("q_ld", ctypes.c_int), ("q_off", ctypes.c_int),
        ("k_hi", ctypes.c_void_p), ("k_lo", ctypes.c_void_p), ("k_ld", ctypes.c_int), ("k_off", ctypes.c_int),
        ("vt_hi", ctypes.c_void_p), ("vt_lo", ctypes.c_void_p),
        ("r_hi", ctypes.c_void_p), ("r_lo", ctypes.c_void_p), ("r_ld", ctypes.c_int), ("r_off", ctypes.c_int),
        ("o_hi", ctypes.c_void_p), ("o_lo", ctypes.c_void_p), ("o_ld", ctypes.c_int), ("o_off", ctypes.c_int),
        ("batch", ctypes.c_int), ("s", ctypes.c_int), ("c", ctypes.c_int),
        ("lse", ctypes.c_void_p),
    ]


class AttnBwdDesc(ctypes.Structure):
    """Mirror of ``hupr_attn_bwd_desc``."""
    _fields_ = [
        ("q_hi", ctypes.c_void_p), ("q_lo", ctypes.c_void_p), ("q_ld", ctypes.c_int), ("q_off", ctypes.c_int),
        ("k_hi", ctypes.c_void_p), ("k_lo", ctypes.c_void_p), ("k_ld", ctypes.c_int), ("k_off", ctypes.c_int),
        ("v_hi", ctypes.c_void_p), ("v_lo", ctypes.c_void_p), ("v_ld", ctypes.c_int), ("v_off", ctypes.c_int),
        ("do_hi", ctypes.c_void_p), ("do_lo", ctypes.c_void_p), ("do_ld", ctypes.c_int), ("do_off", ctypes.c_int),
        ("lse", ctypes.c_void_p), ("rowdot", ctypes.c_void_p),
        ("dq", ctypes.c_void_p), ("dq_ld", ctypes.c_int), ("dq_off", ctypes.c_int),
        ("dk", ctypes.c_void_p), ("dk_ld", ctypes.c_int), ("dk_off", ctypes.c_int),
        ("dv", ctypes.c_void_p), ("dv_ld", ctypes.c_int), ("dv_off", ctypes.c_int),
        ("batch", ctypes.c_int), ("s", ctypes.c_int), ("c", ctypes.c_int),
    ]


class WgradDesc(ctypes.Structure):
    """Mirror of ``hupr_wgrad_desc``."""
    _fields_ = [
        ("x_hi", ctypes.c_void_p), ("x_lo", ctypes.c_void_p),
        ("n", ctypes.c_int), ("d", ctypes.c_int), ("h", ctypes.c_int), ("w", ctypes.c_int), ("cx", ctypes.c_int),
        ("x_ch_off", ctypes.c_int), ("cin", ctypes.c_int),
        ("dy_hi", ctypes.c_void_p), ("dy_lo", ctypes.c_void_p), ("cy", ctypes.c_int), ("y_ch_off", ctypes.c_int), ("cout", ctypes.c_int),
        ("kd", ctypes.c_int), ("kh", ctypes.c_int), ("kw", ctypes.c_int),
        ("pd", ctypes.c_int), ("ph", ctypes.c_int), ("pw", ctypes.c_int),
        ("dw", ctypes.c_void_p), ("dw_ld", ctypes.c_int),
        ("batched", ctypes.c_int), ("dw_batch_stride", ctypes.c_longlong),
        ("nprod", ctypes.c_int),
    ]


class PackJob(ctypes.Structure):
    """Mirror of ``hupr_pack_job``."""
    _fields_ = [
        ("w", ctypes.c_void_p), ("a_hi", ctypes.c_void_p), ("a_lo", ctypes.c_void_p), ("b_hi", ctypes.c_void_p), ("b_lo", ctypes.c_void_p),
        ("cout", ctypes.c_int), ("cin", ctypes.c_int), ("taps", ctypes.c_int),
        ("cout_total", ctypes.c_int), ("cin_pad", ctypes.c_int), ("cout_off", ctypes.c_int),
        ("blocks_x", ctypes.c_int), ("block_begin", ctypes.c_int),
    ]


class TensorView(ctypes.Structure):
    """Mirror of ``hupr_tensor_view``."""
    _fields_ = [("hi", ctypes.c_void_p), ("lo", ctypes.c_void_p), ("ld", ctypes.c_int), ("ch_off", ctypes.c_int)]


_P = ctypes.c_void_p
_I = ctypes.c_int
_TV = ctypes.POINTER(TensorView)

# name -> (restype, argtypes); must list every symbol of include/hupr_b200.h
SIGNATURES = {
    "hupr_version": (ctypes.c_int, []),
    "hupr_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "hupr_launch_count": (ctypes.c_longlong, []),
    "hupr_set_pdl": (ctypes.c_int, [ctypes.c_int]),
    "hupr_fft_cascade_i16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "hupr_conv_gemm": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.c_void_p]),
    "hupr_conv_quant_eligible": (ctypes.c_int, [ctypes.POINTER(ConvDesc)]),
    "hupr_quantize_planes": (ctypes.c_int, [_P, _P, ctypes.c_longlong, _I, _I, _I, _P, _P, _I, _P, _P]),
    "hupr_conv_wgrad": (ctypes.c_int, [ctypes.POINTER(WgradDesc), ctypes.c_void_p]),
    "hupr_attention_bwd": (ctypes.c_int, [ctypes.POINTER(AttnBwdDesc), ctypes.c_void_p]),
    "hupr_attention_fwd": (ctypes.c_int, [ctypes.POINTER(AttnDesc), ctypes.c_void_p]),
    "hupr_window_normalize": (ctypes.c_int, [_P, _P, _I, _P, _P]),
    "hupr_frame_features_workspace_bytes": (ctypes.c_size_t, [_I]),
    "hupr_plane_stats": (ctypes.c_int, [_P, _I, _P, ctypes.c_size_t, _P]),
    "hupr_frame_features": (ctypes.c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "hupr_mnet_fwd": (ctypes.c_int, [_P, _P, _P, _P, _P, _I, _P]),
    "hupr_resample_linear": (ctypes.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "hupr_softmax_rows": (ctypes.c_int, [_P, _P, _P, ctypes.c_longlong, _I, _P]),
    "hupr_transpose_split": (ctypes.c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "hupr_prgcn_workspace_bytes": (ctypes.c_size_t, [_I]),
    "hupr_prgcn_fwd": (ctypes.c_int, [_P, _I, ctypes.POINTER(_P), ctypes.POINTER(_P), _P, _P, ctypes.c_size_t, _P, _P, _I, _P]),
    "hupr_gcn_nodes": (ctypes.c_int, [_P, _I, _P, _P, _P, _P, _I, _P]),
    "hupr_gcn_mix": (ctypes.c_int, [_P, _P, _P, _P, _P, _I, _P]),
    "hupr_gcn_heads": (ctypes.c_int, [_P, _I, _P, _I, _P]),
    "hupr_keypoints_argmax": (ctypes.c_int, [_P, _I, _P, _P, _P]),
    "hupr_keypoint_oks": (ctypes.c_int, [_P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "hupr_channel_sums": (ctypes.c_int, [_I, _TV, _TV, _TV, _P, _P, ctypes.c_longlong, _I, _P, _P, _P]),
    "hupr_affine_act": (ctypes.c_int, [_TV, _P, _P, _TV, _P, _P, _P, _TV, ctypes.c_longlong, _I, _P]),
    "hupr_bn_bwd_apply": (ctypes.c_int, [_TV, _TV, _TV, _P, _P, _P, _P, _P, _TV, ctypes.c_longlong, _I, _P]),
    "hupr_act_bwd": (ctypes.c_int, [_TV, _TV, _P, _TV, ctypes.c_longlong, _I, _P]),
    "hupr_bn_finalize": (ctypes.c_int, [_P, _P, ctypes.c_longlong, _P, _P, ctypes.c_float, ctypes.c_float, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "hupr_bn_bwd_finalize": (ctypes.c_int, [_P, _P, ctypes.c_longlong, _P, _P, _P, _P, _I, _P]),
    "hupr_rowdot": (ctypes.c_int, [_TV, _TV, _TV, _P, ctypes.c_longlong, _I, _P]),
    "hupr_accumulate": (ctypes.c_int, [_TV, _TV, _P, _I, _I, _TV, ctypes.c_longlong, _I, _P]),
    "hupr_resample_linear_bwd": (ctypes.c_int, [_TV, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P]),
    "hupr_softmax_bwd_rows": (ctypes.c_int, [_P, _P, _P, _P, _P, ctypes.c_longlong, _I, _P]),
    "hupr_gcn_bias_grad": (ctypes.c_int, [_P, _P, _P, _I, _P]),
    "hupr_gcn_heads_bwd": (ctypes.c_int, [_P, _P, _I, _P]),
    "hupr_gcn_nodes_bwd": (ctypes.c_int, [_P, _P, _P, _P, _I, _I, _P]),
    "hupr_mnet_bwd": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "hupr_to_kmajor": (ctypes.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, ctypes.c_longlong, _P]),
    "hupr_to_kmajor_multi": (ctypes.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, ctypes.c_longlong,
                                            ctypes.c_longlong, _P]),
    "hupr_heatmap_loss_bwd": (ctypes.c_int, [_P, _P, _P, _I, _I, ctypes.c_float, ctypes.c_float, _P, _P, _P]),
    "hupr_heatmap_bwd": (ctypes.c_int, [_P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "hupr_adam_step": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_longlong, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                      ctypes.c_float, _I, _P, _P, _P]),
    "hupr_heatmap_loss_fwd": (ctypes.c_int, [_P, _P, _P, _I, _P, ctypes.c_size_t, _P, _P, _P, _P]),
    "hupr_pack_conv_weights": (ctypes.c_int, [_P, _I, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P]),
    "hupr_unpack_wgrad": (ctypes.c_int, [_P, _I, _I, _I, _I, _P, _I, _I, _P]),
    "hupr_pack_conv_weights_multi": (ctypes.c_int, [_P, _P, _I, _I, _P]),
    "hupr_unpack_wgrad_multi": (ctypes.c_int, [_P, _P, _I, _I, _P]),
    "hupr_reduce_f64": (ctypes.c_int, [_P, _I, _I, _P, _P]),
    "hupr_broadcast_f32": (ctypes.c_int, [_P, _P, _I, _P]),
    "hupr_gcn_bias_rows": (ctypes.c_int, [_P, _I, _I, _P, _P, _P]),
    "hupr_bump_i32": (ctypes.c_int, [_P, _P]),
    "hupr_memset_zero": (ctypes.c_int, [_P, ctypes.c_size_t, _P]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "hupr_b200: %s not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU/PyTorch fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the symbol is missing -> loud failure
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().hupr_error_string(code).decode()
        raise RuntimeError("hupr_b200.%s failed: %s (code %d)" % (what, msg, code))


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def optr(t):
    """Raw pointer value of an optional tensor (None -> NULL), for Structure fields."""
    return None if t is None else t.data_ptr()
