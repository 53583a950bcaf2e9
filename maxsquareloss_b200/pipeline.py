"""HostPipeline: the fused loss step for HOST-resident tensors (C ABI ``msq_pipe_*``).

Each ``submit`` copies the head logits host->device, runs the fused forward (and backward when
a gradient buffer is given) and copies loss / histogram / dL/dlogits back, on one of ``depth``
internal CUDA streams; the copies of one step overlap the kernels of another.  It is the same
sequence the reference's trainer performs with torch ops (``tools/solve_gta5.py:366-371,199,217``),
for callers that are not inside a PyTorch autograd graph.
"""
import ctypes

import torch

from . import _lib


class HostPipeline:
    def __init__(self, mode, n, num_class, hw_in, hw_out, ratio=0.2, depth=3, device=None, comm=None, global_batch=0):
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device: there is no CPU fallback")
        if device is not None:
            torch.cuda.set_device(device)
        torch.cuda.current_stream().synchronize()      # make sure the primary context exists
        self._lib = _lib.load()
        self.mode = {"maxsquare": _lib.MODE_MAXSQUARE, "iw": _lib.MODE_IW}.get(mode, mode)
        self.shape = (n, num_class, int(hw_in[0]), int(hw_in[1]))
        self.n, self.c, self.depth = n, num_class, depth
        h = ctypes.c_void_p()
        _lib.check(self._lib.msq_pipe_create(self.mode, n, num_class, int(hw_in[0]), int(hw_in[1]), int(hw_out[0]),
                                             int(hw_out[1]), float(ratio), depth, ctypes.byref(h)))
        self._h = h
        self._comm = comm                  # dist.StatsComm: images sharded over ranks (kept alive with the pipeline)
        if comm is not None or global_batch:
            _lib.check(self._lib.msq_pipe_shard(h, int(global_batch), comm._h if comm is not None else None))
        self._ok = {}
        self._keep = [None] * depth        # keep submitted host tensors alive until their slot is waited on

    def _check(self, t, shape, dtype, name):
        key = (id(t), t.data_ptr(), t.numel(), t.dtype)
        if self._ok.get(key) is name:          # this very buffer passed before (trainers reuse their pinned buffers)
            return
        if t.device.type != "cpu" or t.dtype != dtype or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"{name} must be a contiguous CPU {dtype} tensor of shape {tuple(shape)}")
        if len(self._ok) > 256:
            self._ok.clear()
        self._ok[key] = name

    def submit(self, host_logits, host_loss, host_grad=None, host_hist=None, grad_scale=1.0):
        """Enqueue one step; returns the slot to ``wait`` on.  Tensors should be pinned."""
        self._check(host_logits, self.shape, torch.float32, "host_logits")
        self._check(host_loss, (), torch.float32, "host_loss")
        if host_grad is not None:
            self._check(host_grad, self.shape, torch.float32, "host_grad")
        if host_hist is not None:
            self._check(host_hist, (self.n, self.c), torch.int32, "host_hist")
        slot = ctypes.c_int(-1)
        _lib.check(self._lib.msq_pipe_submit(
            self._h, host_logits.data_ptr(), float(grad_scale), host_loss.data_ptr(),
            host_grad.data_ptr() if host_grad is not None else None,
            host_hist.data_ptr() if host_hist is not None else None, ctypes.byref(slot)))
        self._keep[slot.value] = (host_logits, host_loss, host_grad, host_hist)
        return slot.value

    def wait(self, slot):
        _lib.check(self._lib.msq_pipe_wait(self._h, int(slot)))
        self._keep[slot] = None

    def drain(self):
        _lib.check(self._lib.msq_pipe_drain(self._h))
        self._keep = [None] * self.depth

    def close(self):
        if getattr(self, "_h", None):
            self._lib.msq_pipe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
