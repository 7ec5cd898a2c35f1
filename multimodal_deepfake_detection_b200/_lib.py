"""ctypes binding of libxcp_sm100.so (the C-ABI boundary declared in include/xcp.h).

Loading is lazy (DataLoader worker processes import the model package but must not touch CUDA,
SURVEY.md §8b) and LOUD: if the shared object is missing or a call returns non-zero, a RuntimeError
carrying ``xcp_last_error_string()`` is raised.  There is no CPU / eager fallback anywhere.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxcp_sm100.so")

_T = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_longlong, "f": ctypes.c_float, "d": ctypes.c_double}

# name -> argument type string (see include/xcp.h for the meaning of every argument)
SIGNATURES = {
    "xcp_version": "",
    "xcp_check_device": "i",
    "xcp_gemm_tn": "plplpliiiippiiip",
    "xcp_gemm_tn_bias": "plplpliiipipliiip",
    "xcp_gemm_stats_parts": "lii",
    "xcp_gemm_wgrad": "plplpliiiip",
    "xcp_gemm_ref": "plplpliiiiip",
    "xcp_conv3x3_gemm": "ppppiiiiiiiiip",
    "xcp_conv3x3_wgrad": "pppiiiiiip",
    "xcp_stem_conv1_parts": "iiii",
    "xcp_stem_conv1_fwd": "pipppiiiip",
    "xcp_stem_conv1_fwd_affine": "pippppiiiip",
    "xcp_stem_conv1_wgrad_ws_bytes": "iii",
    "xcp_stem_conv1_wgrad": "pipppiiiip",
    "xcp_dw3x3_fwd": "ppppipiiiiip",
    "xcp_dw3x3_bwd": "pppppipppppiiiiiip",
    "xcp_bn_finalize": "piiidppppffppppip",
    "xcp_bn_eval_affine": "ppppfppppiiip",
    "xcp_bn_act": "pppipliip",
    "xcp_gather_s2": "pppipiiiiip",
    "xcp_pool_add_fwd": "pppppppppiiiiip",
    "xcp_bn_bwd_sums": "ppppliip",
    "xcp_bn_add_fwd": "pppppppliip",
    "xcp_bn_relu_gap": "ppppiiiip",
    "xcp_bnbwd_num_parts": "",
    "xcp_bn_bwd": "ipppppppppippppppiiiiiiiip",
    "xcp_nchw_to_nhwc": "ppiiiiip",
    "xcp_nhwc_to_nchw": "ppiiiiip",
    "xcp_pack_weight": "pppiiiiip",
    "xcp_pack_weight_scaled": "pppiiiiip",
    "xcp_pack_dw": "ppiiip",
    "xcp_pack_multi": "piiip",
    "xcp_unpack_dw_grad": "ppiiip",
    "xcp_pack_conv3x3": "pppiiip",
    "xcp_unpack_conv3x3_grad": "ppiiip",
    "xcp_bilinear_up": "ppliiip",
    "xcp_cast_f32_bf16": "pplip",
    "xcp_lstm_fwd": "pppppppppiiiip",
    "xcp_lstm_bwd": "pppppppppppiiiip",
    "xcp_linear_small_fwd": "ppppfipiiiip",
    "xcp_linear_small_bwd": "ppfpppppiiiip",
    "xcp_sigmoid_fwd": "ppiip",
    "xcp_sigmoid_bwd": "pppiip",
    "xcp_bce_fwd_bwd": "ppfpppiip",
    "xcp_bce_prob_fwd_bwd": "ppppiip",
    "xcp_fusion_head_fwd": "ppiiiippppiippffipfffppfpppppppppip",
    "xcp_fusion_head_bwd": "pppiiiippiippppfffppppppppppip",
    "xcp_head_mlp_fwd": "plppppfpppipfpppiiiip",
    "xcp_head_mlp_bwd": "pppplppfppppplpiiiip",
    "xcp_arcface_loss": "pppffipfppppppiifip",
    "xcp_fusion_pool_reg": "ppppppiiifffip",
    "xcp_fusion_pool_bwd": "pppiiiip",
    "xcp_grad_sumsq": "plpiip",
    "xcp_adam_step": "pppplfffffiipffip",
    "xcp_adam_multi": "pipifffffipffpip",
    "xcp_f32_conv3x3": "pippiiiiiiip",
    "xcp_f32_dw3x3": "pppiiiiip",
    "xcp_f32_gemm": "ppppliiip",
    "xcp_f32_bn_stats_parts": "l",
    "xcp_f32_bn_stats": "ppliip",
    "xcp_f32_affine": "pppipliip",
    "xcp_f32_pool_add": "pppiiiiip",
    "xcp_f32_add": "ppplip",
    "xcp_f32_gather": "ppiiiiiip",
    "xcp_f32_gap": "ppiiiip",
    "xcp_f32_lstm_fwd": "pppppppiiiip",
    "xcp_f32_dw3x3_fused": "ppppipiiiiip",
    "xcp_f32_pool_add_fused": "pppppppppiiiiip",
    "xcp_f32_bn_bwd_sums": "pppliip",
    "xcp_f32_bn_add": "pppppppliip",
    "xcp_f32_bn_relu_gap": "ppppiiiip",
    "xcp_f32_bn_bwd": "ipppppppppippppppiiiiiiiip",
    "xcp_f32_dw3x3_bwd": "pppppipppppiiiiiip",
    "xcp_f32_gemm_wgrad": "plplplliiip",
    "xcp_f32_conv3x3_dgrad": "pppiiiiiip",
    "xcp_f32_conv3x3_wgrad": "pippiiiiiiip",
    "xcp_split3_bf16": "ppliiiip",
    "xcp_mfcc_frames": "ii",
    "xcp_mfcc": "piipiiiiiffpppip",
}
_RET_LONGLONG = {"xcp_stem_conv1_wgrad_ws_bytes"}
_NO_STATUS = {"xcp_version", "xcp_bnbwd_num_parts", "xcp_gemm_stats_parts", "xcp_stem_conv1_parts",
              "xcp_stem_conv1_wgrad_ws_bytes", "xcp_f32_bn_stats_parts", "xcp_mfcc_frames"}

# kernels launched per C-ABI call (for bench.py's `gpu_launches`; 0 = host-only query)
_LAUNCHES = {"xcp_version": 0, "xcp_check_device": 0, "xcp_bnbwd_num_parts": 0, "xcp_gemm_stats_parts": 0,
             "xcp_stem_conv1_parts": 0, "xcp_stem_conv1_wgrad_ws_bytes": 0, "xcp_f32_bn_stats_parts": 0, "xcp_mfcc_frames": 0, "xcp_mfcc": 2, "xcp_stem_conv1_wgrad": 3, "xcp_bn_bwd": 3,
             "xcp_arcface_loss": 2, "xcp_adam_multi": 2, "xcp_f32_bn_bwd": 3, "xcp_f32_dw3x3_bwd": 2, "xcp_bn_bwd_sums": 2}
_count = 0

_lock = threading.Lock()
_lib = None


def reset_launch_count():
    global _count
    _count = 0


def launch_count() -> int:
    return _count


def add_launches(n: int):
    global _count
    _count += n


class XcpError(RuntimeError):
    pass


# ---- optional per-call CUDA-event timing (bench.py's per-kernel roofline): every C-ABI call is bracketed by two events on the
# launching (= torch current) stream and logged as (entry point, argument signature, events).  Integer / float arguments are
# kept as they are (shapes, modes), pointer arguments become "non-NULL?" flags.  Off by default: one `is None` test per call.
_timer = None


def timer_start():
    global _timer
    _timer = []


def timer_stop():
    """-> {(name, signature): [total_ms, calls]} of the calls since timer_start() (synchronises the device)."""
    global _timer
    rec, _timer = _timer, None
    if not rec:
        return {}
    import torch
    torch.cuda.synchronize()
    out = {}
    for name, sig, e0, e1 in rec:
        d = out.setdefault((name, sig), [0.0, 0])
        d[0] += e0.elapsed_time(e1); d[1] += 1
    return out


def load():
    """dlopen the library (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise XcpError(
                "libxcp_sm100.so is missing at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C multimodal_deepfake_detection_b200/csrc).  There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.xcp_last_error_string.restype = ctypes.c_char_p
        lib.xcp_last_error_string.argtypes = []
        for name, sig in SIGNATURES.items():
            fn = getattr(lib, name)     # AttributeError here == header/library mismatch: fail loudly
            fn.argtypes = [_T[c] for c in sig]
            fn.restype = ctypes.c_longlong if name in _RET_LONGLONG else ctypes.c_int
        _lib = lib
    return _lib


def call(name: str, *args):
    global _count
    lib = load()
    if _timer is not None and name not in _NO_STATUS:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        _timer.append((name, tuple(a if isinstance(a, (int, float)) else (getattr(a, "value", None) is not None) for a in args), e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if name == "xcp_bn_bwd":
        # reduce pass (unless the sums are given) + finalize (folded into the apply kernel when the sums are complete and the
        # source is not the max-pool routing) + apply pass (when dy is wanted): include/xcp.h argument order
        presums, dy = bool(args[11].value), bool(args[16].value)
        fused = presums and dy and not (args[0] == 2 and args[22] <= 0)
        _count += (0 if presums else 1) + (0 if fused else 1) + (1 if dy else 0)
    else:
        _count += _LAUNCHES.get(name, 1)
    if name in _NO_STATUS:
        return rc
    if rc != 0:
        msg = lib.xcp_last_error_string()
        raise XcpError("%s failed (rc=%d): %s" % (name, rc, msg.decode(errors="replace") if msg else "?"))
    return 0
