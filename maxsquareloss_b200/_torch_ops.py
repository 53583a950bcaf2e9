"""Loader of ``lib/libmsq_torch.so``: the C++ autograd nodes (``csrc/torch_binding.cpp``) behind
``torch.ops.msq_b200.fused_loss`` / ``prob_loss``.  They call the same C ABI as ``_lib`` (ctypes); the point of a
native node is host time -- a Python ``autograd.Function`` costs ~100 us per forward+backward, four times the GPU work
of a training step of the loss."""
import os
import threading

from . import _lib

TORCH_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmsq_torch.so")
_lock = threading.Lock()
_ops = None


class _Ops:
    __slots__ = ("fused_loss", "prob_loss")


def load():
    """-> object with ``fused_loss`` and ``prob_loss`` (the ``.default`` overloads, resolved once).  Raises if the
    binding has not been built: there is no fallback to an eager PyTorch implementation."""
    global _ops
    if _ops is not None:
        return _ops
    with _lock:
        if _ops is not None:
            return _ops
        _lib.load()                                   # libmsq_b200.so first: the binding resolves its symbols from it
        if not os.path.exists(TORCH_LIB_PATH):
            raise RuntimeError(f"{TORCH_LIB_PATH} is missing: build it with `python -m maxsquareloss_b200.build`")
        import torch
        torch.ops.load_library(TORCH_LIB_PATH)
        o = _Ops()
        o.fused_loss = torch.ops.msq_b200.fused_loss.default
        o.prob_loss = torch.ops.msq_b200.prob_loss.default
        _ops = o
    return _ops
