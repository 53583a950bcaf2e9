"""ctypes binding of libmsq_b200.so (C ABI in include/msq_b200.h).

There is no CPU fallback: if the library is missing this module raises, and every
entry point refuses non-CUDA tensors.
"""
import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSQ_B200_LIB") or os.path.join(_PKG, "lib", "libmsq_b200.so")   # override: A/B builds

MAX_CLASSES = 32
MODE_MAXSQUARE = 0
MODE_IW = 1

#: every symbol include/msq_b200.h declares
SYMBOLS = ("msq_abi_version", "msq_launch_count", "msq_fused_aux_bytes", "msq_error_string", "msq_state_layout_get", "msq_prob_fwd", "msq_prob_bwd",
           "msq_fused_fwd", "msq_fused_bwd", "msq_entropy_fwd", "msq_entropy_bwd", "msq_multi_fwd", "msq_guidance_bwd", "msq_source_ce_fwd", "msq_confusion_i64", "msq_confusion_i64_multi", "msq_confusion_per_image_logits_f32", "msq_softce_fwd", "msq_softce_bwd", "msq_confusion_logits_f32", "msq_confusion_flip_f32", "msq_tune_set",
           "msq_fused_fwd_bwd", "msq_comm_unique_id", "msq_comm_create", "msq_comm_allreduce_f64", "msq_comm_allreduce_u64", "msq_comm_sum_u64_begin", "msq_comm_sum_u64_end", "msq_comm_join", "msq_comm_destroy",
           "msq_comm_box_export", "msq_comm_box_open", "msq_comm_box_enable", "msq_comm_box_active", "msq_comm_box_errors", "msq_comm_box_timeout", "msq_comm_result",
           "msq_pipe_create", "msq_pipe_shard", "msq_pipe_submit", "msq_pipe_wait", "msq_pipe_drain", "msq_pipe_destroy")


class StateLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in (
        "sumsq_off", "kept_off", "hist_off", "flags_off", "ticket_off", "ce_off", "nvalid_off", "accum_bytes",
        "sum_out_off", "kept_out_off", "loss_off", "weights_off", "hist_out_off", "stats_off",
        "ce_fix_out_off", "nvalid_out_off", "loss2_off", "ce_out_off", "out_bytes")]


_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library once.  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m maxsquareloss_b200.build` "
                "(nvcc, sm_100a).  maxsquareloss_b200 has no CPU or PyTorch fallback.")
        import torch  # noqa: F401  (loads PyTorch's libcudart.so.12 first: the library shares that runtime)
        lib = ctypes.CDLL(LIB_PATH)
        c = ctypes
        vp, i32, i64, dbl = c.c_void_p, c.c_int, c.c_int64, c.c_double
        lib.msq_abi_version.restype = i32
        lib.msq_abi_version.argtypes = []
        lib.msq_launch_count.restype = c.c_ulonglong
        lib.msq_launch_count.argtypes = []
        lib.msq_error_string.restype = c.c_char_p
        lib.msq_error_string.argtypes = [i32]
        lib.msq_state_layout_get.restype = i32
        lib.msq_state_layout_get.argtypes = [i32, i32, c.POINTER(StateLayout)]
        lib.msq_prob_fwd.restype = i32
        lib.msq_prob_fwd.argtypes = [i32, vp, i32, i32, i64, vp, dbl, i32, i32, vp, vp, vp]
        lib.msq_prob_bwd.restype = i32
        lib.msq_prob_bwd.argtypes = [i32, vp, i32, i32, i64, i32, i32, vp, vp, vp, vp]
        lib.msq_fused_fwd.restype = i32
        lib.msq_fused_fwd.argtypes = [i32, vp, i32, i32, i32, i32, i32, i32, vp, dbl, i32, vp, vp, vp, vp, vp]
        lib.msq_fused_bwd.restype = i32
        lib.msq_fused_bwd.argtypes = [i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp]
        lib.msq_entropy_fwd.restype = i32
        lib.msq_entropy_fwd.argtypes = [i32, vp, i32, i32, i32, i32, i32, i32, dbl, i32, vp, vp, vp, vp, vp]
        lib.msq_entropy_bwd.restype = i32
        lib.msq_entropy_bwd.argtypes = [i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp]
        lib.msq_multi_fwd.restype = i32
        lib.msq_multi_fwd.argtypes = [i32, vp, vp, i32, i32, i32, i32, i32, i32, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.msq_guidance_bwd.restype = i32
        lib.msq_guidance_bwd.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp]
        lib.msq_source_ce_fwd.restype = i32
        lib.msq_source_ce_fwd.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
        lib.msq_fused_aux_bytes.restype = i64
        lib.msq_fused_aux_bytes.argtypes = [i32, i32, i32]
        lib.msq_confusion_flip_f32.restype = i32
        lib.msq_confusion_flip_f32.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
        lib.msq_confusion_i64.restype = i32
        lib.msq_confusion_i64.argtypes = [vp, vp, i64, i32, vp, vp, vp]
        lib.msq_confusion_i64_multi.restype = i32
        lib.msq_confusion_i64_multi.argtypes = [vp, vp, vp, i32, i32, vp, i64, vp, vp, vp]
        lib.msq_confusion_per_image_logits_f32.restype = i32
        lib.msq_confusion_per_image_logits_f32.argtypes = [vp, vp, i32, i32, i64, vp, vp, vp]
        lib.msq_softce_fwd.restype = i32
        lib.msq_softce_fwd.argtypes = [i32, vp, vp, i32, i32, i64, dbl, i32, i32, vp, vp, vp]
        lib.msq_softce_bwd.restype = i32
        lib.msq_softce_bwd.argtypes = [i32, vp, vp, i32, i32, i64, i32, i32, vp, vp, vp, vp, vp]
        lib.msq_confusion_logits_f32.restype = i32
        lib.msq_confusion_logits_f32.argtypes = [vp, vp, i32, i32, i64, vp, vp]
        lib.msq_tune_set.restype = i32
        lib.msq_tune_set.argtypes = [c.c_char_p, i32]
        lib.msq_pipe_create.restype = i32
        lib.msq_pipe_create.argtypes = [i32, i32, i32, i32, i32, i32, i32, dbl, i32, c.POINTER(vp)]
        lib.msq_pipe_shard.restype = i32
        lib.msq_pipe_shard.argtypes = [vp, i32, vp]
        lib.msq_pipe_submit.restype = i32
        lib.msq_pipe_submit.argtypes = [vp, vp, c.c_float, vp, vp, vp, c.POINTER(i32)]
        lib.msq_pipe_wait.restype = i32
        lib.msq_pipe_wait.argtypes = [vp, i32]
        lib.msq_pipe_drain.restype = i32
        lib.msq_pipe_drain.argtypes = [vp]
        lib.msq_fused_fwd_bwd.restype = i32
        lib.msq_fused_fwd_bwd.argtypes = [i32, vp, i32, i32, i32, i32, i32, i32, dbl, i32, vp, vp, vp, vp, c.c_float, vp, vp, i32, vp]
        lib.msq_comm_unique_id.restype = i32
        lib.msq_comm_unique_id.argtypes = [vp]
        lib.msq_comm_create.restype = i32
        lib.msq_comm_create.argtypes = [vp, i32, i32, c.POINTER(vp)]
        lib.msq_comm_allreduce_f64.restype = i32
        lib.msq_comm_allreduce_f64.argtypes = [vp, vp, i32, vp]
        lib.msq_comm_allreduce_u64.restype = i32
        lib.msq_comm_allreduce_u64.argtypes = [vp, vp, i32, vp]
        lib.msq_comm_sum_u64_begin.restype = i32
        lib.msq_comm_sum_u64_begin.argtypes = [vp, vp, i32, vp]
        lib.msq_comm_sum_u64_end.restype = i32
        lib.msq_comm_sum_u64_end.argtypes = [vp, vp, i32, vp]
        lib.msq_comm_join.restype = i32
        lib.msq_comm_join.argtypes = [vp, i32, vp]
        lib.msq_comm_box_export.restype = i32
        lib.msq_comm_box_export.argtypes = [vp, vp]
        lib.msq_comm_box_open.restype = i32
        lib.msq_comm_box_open.argtypes = [vp, vp]
        lib.msq_comm_box_enable.restype = i32
        lib.msq_comm_box_enable.argtypes = [vp, i32]
        lib.msq_comm_box_active.restype = i32
        lib.msq_comm_box_active.argtypes = [vp]
        lib.msq_comm_box_errors.restype = i32
        lib.msq_comm_box_errors.argtypes = [vp, c.POINTER(c.c_uint)]
        lib.msq_comm_box_timeout.restype = i32
        lib.msq_comm_box_timeout.argtypes = [vp, dbl]
        lib.msq_comm_result.restype = i32
        lib.msq_comm_result.argtypes = [vp, i32, vp, i32, vp]
        lib.msq_comm_destroy.restype = None
        lib.msq_comm_destroy.argtypes = [vp]
        lib.msq_pipe_destroy.restype = None
        lib.msq_pipe_destroy.argtypes = [vp]
        if lib.msq_abi_version() != 6:
            raise RuntimeError("libmsq_b200.so ABI version mismatch; rebuild it")
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        msg = load().msq_error_string(code).decode()
        raise RuntimeError(f"libmsq_b200: {msg} (code {code})")


_layouts = {}


def state_layout(n, c):
    key = (n, c)
    lay = _layouts.get(key)
    if lay is None:
        lay = StateLayout()
        check(load().msq_state_layout_get(n, c, ctypes.byref(lay)))
        _layouts[key] = lay
    return lay


def tune(key, value):
    check(load().msq_tune_set(key.encode(), int(value)))
