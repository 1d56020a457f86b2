"""ctypes binding of libfov360.so (the C ABI declared in include/fov360.h).

The library is the product path: if it is missing or fails to load this module
raises — there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfov360.so")

ACT = {None: 0, "linear": 0, "tanh": 1, "relu": 2}
# arithmetic of the convolution / dense / ConvLSTM family (include/fov360.h FOV_MATH_*):
# fp32 = CUDA-core kernels; bf16 / bf16x2 / bf16x3 = tcgen05 kernels with 1 / 2 / 3 bf16 terms per
# operand and fp32 accumulation in tensor memory (bf16x2: ~16 mantissa bits, ~1e-5 max-abs forward error, inside the 1e-4 bar; bf16x3 is the fp32-grade mode)
MATH = {"fp32": 0, "bf16": 1, "bf16x2": 2, "bf16x3": 3}
REC = {"hard_sigmoid": 0, "sigmoid": 1}

c_float_p = C.c_void_p      # device pointers travel as integers


class LstmCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "B", "T_enc", "T_dec", "in_enc", "in_dec", "H", "out_dim", "teacher_forcing",
        "head_act", "rec_act", "dec_zero_init", "training", "math")]


class LstmWeights(C.Structure):
    _fields_ = [(n, c_float_p) for n in (
        "enc_kernel", "enc_recurrent", "enc_bias", "dec_kernel", "dec_recurrent", "dec_bias",
        "head_kernel", "head_bias")]


class LstmSaved(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("xh", "gates", "c", "hseq")]


class LstmIO(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("x_enc", "x_dec", "extra", "h0", "c0", "y", "hT", "cT")] + \
               [("enc", LstmSaved), ("dec", LstmSaved), ("ws", c_float_p)]


class LstmGrads(C.Structure):
    _fields_ = [(n, c_float_p) for n in (
        "dy", "dhseq_enc", "y", "dz_enc", "dz_dec", "dpre",
        "g_enc_kernel", "g_enc_recurrent", "g_enc_bias",
        "g_dec_kernel", "g_dec_recurrent", "g_dec_bias", "g_head_kernel", "g_head_bias", "ws", "dhseq_dec")] + \
               [("wgrad_stream", C.c_void_p)]


class ConvCfg(C.Structure):
    _fields_ = [("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int),
                ("kh", C.c_int), ("kw", C.c_int), ("dil_h", C.c_int), ("dil_w", C.c_int),
                ("pad_h", C.c_int), ("pad_w", C.c_int),
                ("x_img_stride", C.c_longlong), ("x_pix_stride", C.c_int),
                ("y_img_stride", C.c_longlong), ("y_pix_stride", C.c_int),
                ("act", C.c_int), ("beta", C.c_float)]


class ConvLstmCfg(C.Structure):
    _fields_ = [("B", C.c_int), ("T", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int),
                ("F", C.c_int), ("kh", C.c_int), ("kw", C.c_int), ("dil_h", C.c_int),
                ("dil_w", C.c_int), ("rec_act", C.c_int),
                ("x_b_stride", C.c_longlong), ("x_t_stride", C.c_longlong), ("x_pix_stride", C.c_int),
                ("h_b_stride", C.c_longlong), ("h_t_stride", C.c_longlong), ("h_pix_stride", C.c_int),
                ("training", C.c_int), ("math", C.c_int), ("ws_prepacked", C.c_int), ("wave_layers", C.c_int)]


class ConvLstmIO(C.Structure):
    _fields_ = [(n, c_float_p) for n in (
        "x", "kernel", "recurrent", "bias", "h0", "c0", "hseq", "gates", "cseq",
        "hT", "cT", "ws")] + [("wave_wait", C.c_void_p), ("wave_set", C.c_void_p)]


class ConvLstmGrads(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("dhseq", "dhT", "dcT", "dx", "dh0", "dc0",
                                         "g_kernel", "g_recurrent", "g_bias", "ws")] + \
               [("dx_accumulate", C.c_int), ("wgrad_stream", C.c_void_p), ("wave_wait", C.c_void_p),
                ("wave_set", C.c_void_p)]


# every symbol include/fov360.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_LL = C.c_longlong
_I = C.c_int
_F = C.c_float
SYMBOLS = {
    "fov_last_error": (C.c_char_p, []),
    "fov_version": (_I, []),
    "fov_launch_count": (C.c_ulonglong, []),
    "fov_device_is_sm100": (_I, []),
    "fov_lstm_seq2seq_fwd": (_I, [C.POINTER(LstmCfg), C.POINTER(LstmWeights), C.POINTER(LstmIO), _P]),
    "fov_lstm_seq2seq_bwd": (_I, [C.POINTER(LstmCfg), C.POINTER(LstmWeights), C.POINTER(LstmIO),
                                  C.POINTER(LstmGrads), _P]),
    "fov_lstm_bwd_ws_floats": (C.c_size_t, [C.POINTER(LstmCfg)]),
    "fov_lstm_fwd_ws_bytes": (C.c_size_t, [C.POINTER(LstmCfg)]),
    "fov_conv2d_fwd": (_I, [C.POINTER(ConvCfg), _P, _P, _P, _P, _P]),
    "fov_conv2d_bwd_data": (_I, [C.POINTER(ConvCfg), _P, _P, _P, _P, _P]),
    "fov_conv2d_bwd_weight": (_I, [C.POINTER(ConvCfg), _P, _P, _P, _P, _P]),
    "fov_conv_tc_ws_bytes": (C.c_size_t, [C.POINTER(ConvCfg), _I, _I]),
    "fov_conv2d_fwd_tc": (_I, [C.POINTER(ConvCfg), _P, _P, _P, _P, _P, _I, _P]),
    "fov_conv2d_bwd_data_tc": (_I, [C.POINTER(ConvCfg), _P, _P, _P, _P, _I, _P]),
    "fov_conv_tc_pack": (_I, [C.POINTER(ConvCfg), _P, _P, _I, _I, _P]),
    "fov_conv2d_fwd_tc_packed": (_I, [C.POINTER(ConvCfg), _P, _P, _P, _P, _I, _P]),
    "fov_conv2d_bwd_data_tc_packed": (_I, [C.POINTER(ConvCfg), _P, _P, _P, _I, _P]),
    "fov_conv2d_bwd_weight_tc": (_I, [C.POINTER(ConvCfg), _P, _P, _P, _P, _I, _P]),
    "fov_conv_wgrad_ws_bytes": (C.c_size_t, [C.POINTER(ConvCfg), _I]),
    "fov_conv2d_bwd_weight_tc_ws": (_I, [C.POINTER(ConvCfg), _P, _P, _P, _P, _P, _I, _P]),
    "fov_act_bwd": (_I, [_I, _LL, _I, _P, _LL, _P, _LL, _P, _LL, _P]),
    "fov_tapstack_reduce": (_I, [_LL, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    "fov_tapstack_expand": (_I, [_LL, _I, _I, _I, _I, _I, _P, _P, _P]),
    "fov_convlstm_fwd": (_I, [C.POINTER(ConvLstmCfg), C.POINTER(ConvLstmIO), _P]),
    "fov_convlstm_fwd_ws_bytes": (C.c_size_t, [C.POINTER(ConvLstmCfg)]),
    "fov_convlstm_wave_groups": (_I, [_P, _I, _I]),
    "fov_convlstm_fwd_persistent": (_I, [_P]),
    "fov_convlstm_bwd_ws_floats": (C.c_size_t, [C.POINTER(ConvLstmCfg)]),
    "fov_convlstm_bwd": (_I, [C.POINTER(ConvLstmCfg), C.POINTER(ConvLstmIO), C.POINTER(ConvLstmGrads), _P]),
    "fov_softmax_fwd": (_I, [_LL, _I, _P, _P, _P]),
    "fov_softmax_bwd": (_I, [_LL, _I, _P, _P, _P, _P]),
    "fov_mse_fwd_bwd": (_I, [_LL, _P, _P, _F, _P, _P, _P]),
    "fov_gauss_nll_fwd_bwd": (_I, [_I, _I, _I, _P, _P, _F, _P, _P, _P]),
    "fov_cce_fwd_bwd": (_I, [_LL, _I, _P, _P, _F, _P, _P, _P]),
    "fov_adam_step": (_I, [_LL, _P, _P, _P, _P, _I, _F, _F, _F, _F, _F, _P, _P]),
    "fov_rmsprop_step": (_I, [_LL, _P, _P, _P, _F, _F, _F, _F, _P, _P]),
    "fov_mean_var_xyz": (_I, [_LL, _P, _P, _P]),
    "fov_gauss_resample": (_I, [_LL, _I, _P, _P, _P, _P]),
    "fov_gauss_resample_bwd": (_I, [_LL, _I, _P, _P, _P, _P, _P]),
    "fov_philox_normal": (_I, [_LL, C.c_ulonglong, C.c_ulonglong, _P, _P, _P]),
    "fov_dp_unique_id_bytes": (_I, []),
    "fov_dp_get_unique_id": (_I, [_P]),
    "fov_dp_init": (_I, [_P, _I, _I]),
    "fov_dp_world": (_I, []),
    "fov_dp_rank": (_I, []),
    "fov_dp_allreduce": (_I, [_P, C.c_size_t, _P]),
    "fov_dp_broadcast": (_I, [_P, C.c_size_t, _I, _P]),
    "fov_dp_destroy": (_I, []),
    "fov_dropout_expand": (_I, [_I, _I, _I, _I, _P, _LL, _LL, _I, _P, _P, _P]),
    "fov_dropout_reduce": (_I, [_I, _I, _I, _I, _P, _P, _P, _LL, _LL, _I, _I, _P]),
    "fov_gate_kernel_expand": (_I, [_I, _I, _I, _P, _P, _P]),
    "fov_gate_kernel_reduce": (_I, [_I, _I, _I, _P, _P, _P]),
    "fov_window_count": (_I, [_I, _I, _I, _I]),
    "fov_window_stacks": (_I, [_I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "fov_whole_span": (_I, [_LL, _LL, _P, _P, _P]),
    "fov_pick_user_gather": (_I, [_I, _I, _I, _LL, _P, _P, _P, _LL, _LL, _P]),
    "fov_m3_batches": (_I, [_I, _I, _I, _I, _P, _P, _LL, _P, _P, _P, _P, _P]),
    "fov_onehot_heatmaps": (_I, [_LL, _I, _I, _P, _P, _P]),
    "fov_hit_rate": (_I, [_LL, _P, _P, _F, _F, _F, _F, _P, _P]),
    "fov_theta_phi_frames": (_I, [_LL, _P, _P, _P]),
    "fov_gaussian_fov_tiles": (_I, [_LL, _I, _I, _P, _P, _P, _P]),
    "fov_heatmap_sum": (_I, [_LL, _I, _I, _P, _P, _P]),
}

_lib = None


class FovError(RuntimeError):
    pass


def load():
    """Load libfov360.so once; fail loudly when it is absent (no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FovError(
            "libfov360.so not found at %s - build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or longterm360fov_b200/csrc/build.sh; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().fov_last_error()
        raise FovError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
