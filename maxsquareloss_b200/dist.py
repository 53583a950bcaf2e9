"""Sharding by image across the GPUs of one box, and the one small all-reduce per step.

The reference is single-GPU (``nn.DataParallel(model, device_ids=[0])``,
``tools/train_source.py:133-136``); nothing here replaces reference code.  The
path shards naturally: image-wise weights are per image (``utils/loss.py:91-97``),
MaxSquare is a plain sum, confusion counts are additive (``utils/eval.py:121``).
Each rank runs the kernels on its own block of images with the loss normaliser
set to the GLOBAL batch (``crit.global_batch``), and the only exchange is ONE
all-reduce(sum) of a packed float64 vector

    [ loss_partial | class_hist[C] | confusion[C*C] ]        (8*(1+C+C*C) bytes, ~3 KB)

over NCCL/NVLink (gloo in the CPU tests).  Counts below 2**53 are exact in
fp64, so the integers come back bit-exact.  The dL/dlogits of a rank's images
depends on nothing from other ranks, so the gradient path has no collective.
"""
import os

import torch
import torch.distributed as dist

from .loss import _raw_stream


def image_shard(n_total, rank, world):
    """Contiguous block [lo, hi) of the global batch owned by ``rank`` (cfg 3:
    8 images over 2/4/8 ranks).  Remainder images go to the lowest ranks."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stats_len(num_class, with_confusion=True):
    return 1 + num_class + (num_class * num_class if with_confusion else 0)


def pack_stats(loss_partial, class_hist, confusion=None, out=None):
    """-> float64 vector [loss, hist(C), cm(C*C)?] on the inputs' device."""
    c = class_hist.numel()
    n = stats_len(c, confusion is not None)
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=class_hist.device)
    out[0] = loss_partial
    out[1:1 + c] = class_hist.reshape(-1)
    if confusion is not None:
        out[1 + c:] = confusion.reshape(-1)
    return out


def unpack_stats(buf, num_class):
    """-> (loss float64 0-dim, hist int64 (C,), cm int64 (C,C) or None)."""
    c = num_class
    loss = buf[0]
    hist = buf[1:1 + c].round().to(torch.int64)
    cm = None
    if buf.numel() >= 1 + c + c * c:
        cm = buf[1 + c:1 + c + c * c].round().to(torch.int64).view(c, c)
    return loss, hist, cm


def allreduce_stats(buf, group=None, async_op=False):
    """Sum ``buf`` over the ranks (no-op without an initialised process group).
    With ``async_op`` the NCCL work handle is returned so that the collective
    overlaps the backward kernel; call ``.wait()`` before reading ``buf``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class StatsComm:
    """The per-step all-reduce of the packed statistics vector, enqueued by libmsq_b200 on an NCCL communicator
    of its own (C ABI ``msq_comm_*``): one ``ncclAllReduce`` on a side stream per call and ~3 us of host time,
    against ~25 us for ``torch.distributed.all_reduce``.  ``torch.distributed`` is used once, to hand rank 0's
    NCCL unique id to the other ranks.  CUDA only.

    With ``peer_memory`` (default, <= 8 ranks of one NVLink box) the one-call step ``msq_fused_fwd_bwd`` does not
    call NCCL at all: an extra CTA of its backward kernel pushes the previous step's vector into every rank's mailbox over NVLink
    and reduces the one before (``include/msq_b200.h``, "Peer-memory mailboxes"); ``join()`` completes the two steps
    still in flight."""

    def __init__(self, group=None, peer_memory=True):
        import ctypes
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("StatsComm needs an initialised torch.distributed process group (for the unique id)")
        self._libmod = _lib
        self._lib = _lib.load()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        ident = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(self._lib.msq_comm_unique_id(ident))
        t = torch.frombuffer(bytearray(ident.raw), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device())
        t = t.to(dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        h = ctypes.c_void_p()
        _lib.check(self._lib.msq_comm_create(raw, self.world, self.rank, ctypes.byref(h)))
        self._h = h
        self.peer_memory = False
        if peer_memory and self.world <= 8 and os.environ.get("MSQ_PEERBOX", "1") != "0":
            self._open_mailboxes(group, dev)

    def _open_mailboxes(self, group, dev):
        """NVLink peer-memory mailboxes for the per-step statistics (C ABI ``msq_comm_box_*``): every rank exports a
        cudaIpc handle, the handles are all-gathered, every rank maps its peers.  Any failure on any rank (no peer
        access, IPC not permitted in this container) leaves the communicator on NCCL."""
        import ctypes
        mine = ctypes.create_string_buffer(64)
        ok = self._lib.msq_comm_box_export(self._h, mine) == 0
        t = torch.zeros(65, dtype=torch.uint8)
        if ok:
            t[:64] = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8)
            t[64] = 1
        t = t.to(dev)
        allh = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(allh, t, group=group)
        allh = torch.stack(allh).cpu()
        if not bool((allh[:, 64] == 1).all()):
            return                                       # every rank sees the same table: all of them return here
        blob = bytes(allh[:, :64].contiguous().numpy().tobytes())
        rc = self._lib.msq_comm_box_open(self._h, blob)
        flag = torch.tensor([1 if rc == 0 else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 1:                        # all ranks mapped all peers: switch the path on everywhere
            self._libmod.check(self._lib.msq_comm_box_enable(self._h, 1))
        self.peer_memory = bool(self._lib.msq_comm_box_active(self._h))

    def errors(self):
        """Error word of the mailbox path (0 = fine; bit 0 = a peer's vector never arrived).  Synchronises."""
        import ctypes
        v = ctypes.c_uint(0)
        self._libmod.check(self._lib.msq_comm_box_errors(self._h, ctypes.byref(v)))
        return int(v.value)

    def set_timeout(self, seconds):
        """How long a mailbox reduction waits for a peer's vector (default 600 s).  Synchronises the device."""
        self._libmod.check(self._lib.msq_comm_box_timeout(self._h, float(seconds)))

    def result(self, count, lag=0, out=None, stream=None):
        """All-reduced ``[loss | class hist]`` (``count`` = 1 + C doubles) of the ``msq_fused_fwd_bwd`` step issued ``lag``
        steps before the most recent one, as a float64 CUDA tensor (device-to-device copy on the current stream).  It exists
        once two further steps have been enqueued, or after ``join()``."""
        if out is None:
            out = torch.empty(count, dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        self._libmod.check(self._lib.msq_comm_result(self._h, int(lag), out.data_ptr(), int(count), stream))
        return out

    def allreduce(self, buf):
        """Sum the float64 CUDA vector ``buf`` over the ranks, in place, asynchronously w.r.t. the current stream."""
        if not (buf.is_cuda and buf.dtype == torch.float64 and buf.is_contiguous()):
            raise RuntimeError("StatsComm.allreduce needs a contiguous float64 CUDA tensor")
        rc = self._lib.msq_comm_allreduce_f64(self._h, buf.data_ptr(), buf.numel(), _raw_stream(buf.device.index))
        if rc:
            self._libmod.check(rc)

    def allreduce_u64(self, buf):
        """Sum the int64/uint64 CUDA vector ``buf`` over the ranks, in place (exact): the confusion counts of ``Eval``,
        the ``[ce_fix | nvalid]`` pair of the cross-entropy rows.  Asynchronous like ``allreduce``; ``join()`` orders the
        current stream after it."""
        if not (buf.is_cuda and buf.dtype == torch.int64 and buf.is_contiguous()):
            raise RuntimeError("StatsComm.allreduce_u64 needs a contiguous int64 CUDA tensor")
        rc = self._lib.msq_comm_allreduce_u64(self._h, buf.data_ptr(), buf.numel(), _raw_stream(buf.device.index))
        if rc:
            self._libmod.check(rc)

    def sum_u64_begin(self, buf):
        """Start the same-step sum over the ranks of the <= 2 int64 words of ``buf`` (mailboxes: NVLink stores from a 32-thread
        kernel; else ncclAllReduce on the side stream).  Put independent work on the stream, then ``sum_u64_end``."""
        if not (buf.is_cuda and buf.dtype == torch.int64 and buf.is_contiguous() and buf.numel() <= 2):
            raise RuntimeError("StatsComm.sum_u64_begin needs a contiguous int64 CUDA tensor of at most 2 elements")
        rc = self._lib.msq_comm_sum_u64_begin(self._h, buf.data_ptr(), buf.numel(), _raw_stream(buf.device.index))
        if rc:
            self._libmod.check(rc)

    def sum_u64_end(self, buf):
        """Finish it: after this call (in stream order) ``buf`` holds the global sums."""
        rc = self._lib.msq_comm_sum_u64_end(self._h, buf.data_ptr(), buf.numel(), _raw_stream(buf.device.index))
        if rc:
            self._libmod.check(rc)

    def allreduce_ptr(self, ptr, count, stream):
        rc = self._lib.msq_comm_allreduce_f64(self._h, ptr, count, stream)
        if rc:
            self._libmod.check(rc)

    def join(self, stream=None, lag=0):
        """Make the current stream wait for the ``allreduce`` issued ``lag`` calls before the most recent one
        (0 = the most recent); no host synchronisation."""
        if stream is None:
            stream = _raw_stream(torch.cuda.current_device())
        rc = self._lib.msq_comm_join(self._h, int(lag), stream)
        if rc:
            self._libmod.check(rc)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.msq_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
