#!/usr/bin/env python
"""Benchmark of the MaxSquare hot path on B200 (contract: see the task brief / DESIGN.md section 5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], GTA5->Cityscapes target shape): per GPU a batch of
2 images, 19 classes, DeepLabv2 head logits 65x129 upsampled to 512x1024.  A *step* is
one pass of the hot path over that batch: fused IW-MaxSquare forward (bilinear upsample +
softmax + per-image argmax histogram + image-wise weights + loss) and backward (dL/dlogits
at 65x129).  Metric: Gpixel/s = label-resolution pixels / time.

  value  device-resident inputs, kernels launched through the C ABI (ctypes), CUDA-event timed
  e2e    the same step through the REFERENCE'S API -- the IW_MaxSquareloss nn.Module + autograd, called as
         tools/solve_gta5.py:199,217 call it -- with pinned HOST logits in and the loss + dL/dlogits back on the
         host every step (e2e.c_abi_pipeline: the same through the C-ABI host pipeline, no PyTorch autograd)

Weak scaling: every rank owns its own 2 images (sharding by image); for N > 1 each step also exchanges the packed
[loss, class histogram] vector (dist.StatsComm): over NVLink peer-memory mailboxes written by an extra CTA of the step's
own backward kernel, or -- where CUDA IPC is not available -- with one ncclAllReduce.  The N > 1 runs also carry
the cfg-3 (multi-level guidance, batch 8 strong-sharded) and cfg-5 (crosscity step, 1 image per GPU) legs and the
cross-rank parity checks (stats_check, cfg3.check, cfg5.check).

``--impl reference`` times the reference's algorithm on the host CPU (oracle/loss_port.py:
F.interpolate -> softmax -> IW loss -> backward, all host threads).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "IW-MaxSquare fwd+bwd Gpixel/s"
UNIT = "Gpixel/s"
N_IMG, C, HW_LO, HW_OUT = 2, 19, (65, 129), (512, 1024)
RATIO = 0.2
LAMBDA_TARGET = 0.1                      # callers scale the loss before backward (solve_gta5.py:199)
POOL = 128                               # distinct input buffers: 128 x 1.27 MB = 163 MB > 126 MB L2
PX_PER_STEP = N_IMG * HW_OUT[0] * HW_OUT[1]
WORKLOAD = ("cfg2 GTA5->Cityscapes target shape: IW-MaxSquare fwd+bwd, batch 2/GPU, 19 classes, "
            "head logits 65x129 -> 512x1024, ratio 0.2")


def shared_config(n_gpus):
    """``config`` of the JSON line: the SAME dict from both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "images_per_gpu": N_IMG, "global_batch": N_IMG * n_gpus, "num_class": C,
            "head_logits_hw": list(HW_LO), "label_hw": list(HW_OUT), "iw_ratio": RATIO, "lambda_target": LAMBDA_TARGET}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index, interval=0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.interval = interval
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        try:
            mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            try:
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.samples.append(mhz)
            for bit, name in list(self.BAD.items()) + list(self.NOTE.items()):
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            time.sleep(self.interval)

    def start(self):
        if self.nv is None:
            return
        self._once()
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        if self._thr:
            self._thr.join()
        self._once()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_chain_time(n_img, hw_lo, hw_out, iters, warm):
    from maxsquareloss_b200 import synth
    from oracle import loss_port
    lo = synth.head_logits(n_img, C, hw_lo, 0, 5.0)
    for _ in range(warm):
        loss_port.chain_iw_maxsquare(lo, hw_out, C, RATIO, LAMBDA_TARGET)
    t0 = time.perf_counter()
    for _ in range(iters):
        loss_port.chain_iw_maxsquare(lo, hw_out, C, RATIO, LAMBDA_TARGET)
    return (time.perf_counter() - t0) / iters


def cpu_baseline(budget_s=12.0):
    """The reference's algorithm (oracle port, kind 'port') on this host's cores: a bounded
    sample of the same workload (full-size steps until ~budget_s of CPU work)."""
    torch.set_num_threads(os.cpu_count() or 1)
    t1 = cpu_chain_time(N_IMG, HW_LO, HW_OUT, 1, 1)
    iters = max(3, min(40, int(budget_s / max(t1, 1e-3))))
    t = cpu_chain_time(N_IMG, HW_LO, HW_OUT, iters, 0)
    return {"value": PX_PER_STEP / t / 1e9, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{iters} full steps of the same workload (2x19x65x129 -> 512x1024), "
                      f"F.interpolate+softmax+IW loss+backward, {t * 1e3:.0f} ms/step",
            "ms_per_step": t * 1e3}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference is pure Python and cannot travel to the GPU box) on all host threads."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    t_full = cpu_chain_time(N_IMG, HW_LO, HW_OUT, 1, 1)
    budget = 150.0
    n_img, hw_lo, hw_out = N_IMG, HW_LO, HW_OUT
    if (steps + warm) * t_full > budget:
        # shrink the per-step sample (fewer images, then fewer rows) so that the run is bounded
        frac = budget / ((steps + warm) * t_full)
        if frac < 0.5:
            n_img = 1
            frac *= 2
        if frac < 1.0:
            rows_lo = max(9, int(HW_LO[0] * frac))
            hw_lo = (rows_lo, HW_LO[1])
            hw_out = ((rows_lo - 1) * 8, HW_OUT[1])
    t = cpu_chain_time(n_img, hw_lo, hw_out, steps, warm)
    px = n_img * hw_out[0] * hw_out[1]
    val = px / t / 1e9
    sample = (f"per step {n_img}x{C}x{hw_lo[0]}x{hw_lo[1]} -> {hw_out[0]}x{hw_out[1]} "
              f"({px / PX_PER_STEP:.3f} of the workload's pixels), F.interpolate+softmax+IW loss+backward")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(args.gpus),
            "arm": "the reference's algorithm on the host CPU (oracle port of F.interpolate -> softmax -> IW_MaxSquareloss -> backward)",
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def time_loop(fn, iters, warm):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters          # ms per call


def max_over_ranks(x, dev, dist, world):
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def api_e2e_leg(msq, crit, host_in, dev, steps, depth, sync_every_step, comm):
    """The step through the reference's API, exactly as tools/solve_gta5.py:199,217 spell it -- ``loss = crit(...)``,
    ``(lambda_target * loss).backward()`` -- around the host copies the contract asks for: every step copies that
    step's head logits from pinned host memory (H2D) and copies the loss and dL/dlogits back to pinned host memory (D2H).
    ``sync_every_step``: the host waits for the step's results before it issues the next one (a trainer that logs
    ``loss.item()`` every iteration); otherwise it reads the results of step i-depth when it reuses that slot's buffers
    (a trainer that logs with a lag).  Sharded runs also all-reduce the step's [loss | hist] vector (StatsComm).
    Returns seconds per step (wall clock around ``steps`` steps, drained)."""
    shape = (N_IMG, C) + tuple(HW_LO)
    dev_in = [torch.empty(shape, device=dev) for _ in range(depth)]
    host_grad = [torch.empty(shape).pin_memory() for _ in range(depth)]
    host_loss = [torch.empty(()).pin_memory() for _ in range(depth)]
    stats = [torch.zeros(1 + C, dtype=torch.float64, device=dev) for _ in range(depth)]
    done = [torch.cuda.Event() for _ in range(depth)]
    cur = torch.cuda.current_stream()
    npool = len(host_in)
    seen = [0.0]

    def step(i):
        k = i % depth
        if i >= depth:
            done[k].synchronize()                      # step i-depth is complete: its loss / gradient are on the host
            seen[0] += float(host_loss[k])             # ... and the host reads them
        dev_in[k].copy_(host_in[i % npool], non_blocking=True)                  # H2D
        x = dev_in[k].detach().requires_grad_(True)
        loss = crit(x, out_size=HW_OUT)
        (LAMBDA_TARGET * loss).backward()
        host_grad[k].copy_(x.grad, non_blocking=True)                           # D2H
        host_loss[k].copy_(loss.detach(), non_blocking=True)
        if comm is not None:
            stats[k].copy_(crit.last_stats)
            comm.allreduce(stats[k])
            comm.join(lag=1)                           # the previous step's collective; this one overlaps the next step
        done[k].record(cur)
        if sync_every_step:
            done[k].synchronize()

    for i in range(max(3 * depth, 30)):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    if comm is not None:
        comm.join()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps, float(host_loss[(steps - 1) % depth])


def torch_floor(dev, host_in, steps):
    """PyTorch's own fixed host cost for the same call pattern with a NATIVE one-kernel loss (``x.sum()``): detach +
    loss + ``lambda * loss`` + ``.backward()`` + the same copies.  No binding of this library can go below it."""
    shape = (N_IMG, C) + tuple(HW_LO)
    dev_in = torch.empty(shape, device=dev)
    host_grad, host_loss = torch.empty(shape).pin_memory(), torch.empty(()).pin_memory()
    cur = torch.cuda.current_stream()

    def step(i):
        dev_in.copy_(host_in[i % len(host_in)], non_blocking=True)
        x = dev_in.detach().requires_grad_(True)
        loss = x.sum()
        (LAMBDA_TARGET * loss).backward()
        host_grad.copy_(x.grad, non_blocking=True)
        host_loss.copy_(loss.detach(), non_blocking=True)
        if i % 64 == 63:
            cur.synchronize()
    for i in range(64):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    import maxsquareloss_b200 as msq
    from maxsquareloss_b200 import _lib, synth
    from maxsquareloss_b200 import dist as mdist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # the only collectives of this program carry <= 3 KB: one NCCL CTA is plenty, and the one-wave kernels then
        # lose fewer SM slots to it (8 GPUs: 183.8 -> 196.2 Gpix/s together with reserve_sms=2)
        os.environ.setdefault("NCCL_MAX_CTAS", "1")
        os.environ.setdefault("NCCL_MIN_CTAS", "1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    from maxsquareloss_b200 import build as _build
    if rank == 0:
        _build.build()          # no-op when lib/libmsq_b200.so is up to date; never a fallback: load() raises if absent
    if world > 1:
        dist.barrier()
    lib = _lib.load()
    hbm_peak, peak_src = peaks()
    steps, warm = max(1, args.steps), max(3, args.warmup)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- inputs: POOL distinct batches of head logits, resident in HBM (and a pinned host copy)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_lo = N_IMG * C * HW_LO[0] * HW_LO[1]
    lo_pool = torch.randn(POOL, N_IMG, C, *HW_LO, generator=g, device=dev) * 5.0
    grad_pool = torch.empty_like(lo_pool)
    go = torch.full((), LAMBDA_TARGET, device=dev)
    lay = _lib.state_layout(N_IMG, C)
    accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
    outs = torch.empty(POOL, lay.out_bytes, dtype=torch.uint8, device=dev)
    stats_views = [outs[i, lay.stats_off:lay.stats_off + 8 * (1 + C)].view(torch.float64) for i in range(POOL)]
    n_norm = N_IMG * world
    AUX_POOL = 10                                   # 10 x 16.8 MB = 168 MB > L2
    aux_bytes = lib.msq_fused_aux_bytes(N_IMG, HW_OUT[0], HW_OUT[1])
    aux_pool = torch.empty(AUX_POOL, aux_bytes, dtype=torch.uint8, device=dev)
    aux_ptrs = [aux_pool[i].data_ptr() for i in range(AUX_POOL)]
    lo_ptrs = [lo_pool[i].data_ptr() for i in range(POOL)]
    gr_ptrs = [grad_pool[i].data_ptr() for i in range(POOL)]
    out_ptrs = [outs[i].data_ptr() for i in range(POOL)]
    acc_ptr, go_ptr = accum.data_ptr(), go.data_ptr()
    MODE = _lib.MODE_IW
    h, w = HW_LO
    H, W = HW_OUT

    def fwd(i):
        j = i % POOL
        rc = lib.msq_fused_fwd(MODE, lo_ptrs[j], N_IMG, C, h, w, H, W, None, RATIO, n_norm, acc_ptr, out_ptrs[j],
                               aux_ptrs[i % AUX_POOL], gr_ptrs[j], stream)
        if rc:
            _lib.check(rc)

    def bwd(i):
        j = i % POOL
        rc = lib.msq_fused_bwd(MODE, lo_ptrs[j], N_IMG, C, h, w, H, W, n_norm, out_ptrs[j], aux_ptrs[i % AUX_POOL],
                               go_ptr, gr_ptrs[j], 1, stream)
        if rc:
            _lib.check(rc)

    COMM_LAG = int(os.environ.get("MSQ_COMM_LAG", "0"))
    comm = mdist.StatsComm() if world > 1 else None
    if world > 1:
        # NCCL path: room for the NCCL kernel next to the one-wave grids; peer-memory mailboxes: no collective kernel at all
        _lib.tune("reserve_sms", int(os.environ.get("MSQ_RESERVE_SMS", "0" if comm.peer_memory else "2")))
    n_stats = 1 + C
    comm_h = comm._h if comm is not None else None

    def step(i):
        # ONE library call per step (C ABI msq_fused_fwd_bwd), TWO kernels: the fused forward and the fused backward, which derives
        # the image-wise weights itself and carries the finalisation -- and, when sharded, the statistics exchange over the
        # NVLink mailboxes -- in extra CTAs (without CUDA IPC: ncclAllReduce forked AFTER the backward)
        j = i % POOL
        rc = lib.msq_fused_fwd_bwd(MODE, lo_ptrs[j], N_IMG, C, h, w, H, W, RATIO, n_norm, acc_ptr, out_ptrs[j],
                                   aux_ptrs[i % AUX_POOL], go_ptr, 0.0, gr_ptrs[j], comm_h, COMM_LAG, stream)
        if rc:
            _lib.check(rc)

    # ---- warm-up: at least W steps, and enough of them (~0.2 s) for the clocks to be up.  The count is FIXED, not
    #      time-based: every rank must issue the same number of exchanges.  All of it is reported as `warmup`.
    warm_steps = warm if os.environ.get("MSQ_BENCH_MIN_WARM_S") == "0" else max(warm, 5000)      # "0": under ncu
    for i in range(warm_steps):
        step(i)
        if i % 256 == 255:
            torch.cuda.synchronize()
    if comm is not None:
        comm.join(stream)
    torch.cuda.synchronize()

    # ---- timed region: EXACTLY `steps` steps, barrier + synchronize on both sides, max over ranks
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the last warm-up steps run AFTER the barrier + synchronize, back to back with the timed ones: the first steps after
    # an idle stream carry the host's launch latency, the sampler thread's start-up and -- on 8 ranks -- tens of
    # milliseconds of skew between the ranks (measured: 48.8 us/step for the first 2000 steps after the barrier against
    # 36.7 us/step for every later 2000); they are warm-up, not steady state
    REWARM = 0 if os.environ.get("MSQ_BENCH_MIN_WARM_S") == "0" else max(256, min(steps, 4000))
    for i in range(REWARM):
        step(steps - REWARM + i)
    launches0 = int(lib.msq_launch_count())
    ev0.record()
    for i in range(steps):
        step(i)
    if comm is not None:
        comm.join(stream)
    ev1.record()
    launches = int(lib.msq_launch_count()) - launches0          # counted by the library: every kernel it launched in there
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
    ms_rank = ev0.elapsed_time(ev1)
    ms_total = max_over_ranks(ms_rank, dev, dist, world)
    ms_per_step = ms_total / steps
    value = world * PX_PER_STEP / (ms_per_step * 1e-3) / 1e9
    rank_spread = None
    if world > 1:
        allms = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allms, torch.tensor([ms_rank / steps * 1e3], device=dev, dtype=torch.float64))
        rank_spread = {"us_per_step_by_rank": [round(float(t.item()), 3) for t in allms]}

    # ---- cross-rank parity of the exchange (SCALE runs): the vector the mailboxes (or the library's ncclAllReduce)
    #      produced for the LAST timed step against (a) torch.distributed's NCCL all-reduce of the ranks' local vectors and
    #      (b) their sum in rank order: histogram bit-exact, loss <= 1e-6 relative
    stats_check = None
    if comm is not None:
        local = stats_views[(steps - 1) % POOL].clone()
        got = comm.result(n_stats, lag=0)
        ref_nccl = local.clone()
        dist.all_reduce(ref_nccl, op=dist.ReduceOp.SUM)
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        ref_sum = parts[0].clone()
        for t_ in parts[1:]:
            ref_sum += t_
        torch.cuda.synchronize()
        hist_total = float(got[1:].sum().item())
        err = comm.errors() if comm.peer_memory else 0
        rel = lambda a, b: abs(float(a) - float(b)) / max(abs(float(b)), 1e-300)          # noqa: E731
        ok_local = (torch.equal(got[1:], ref_nccl[1:]) and torch.equal(got[1:], ref_sum[1:]) and
                    rel(got[0], ref_nccl[0]) <= 1e-6 and rel(got[0], ref_sum[0]) <= 1e-6 and
                    hist_total == float(world * PX_PER_STEP) and err == 0)
        okt = torch.tensor([1 if ok_local else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        stats_check = {"exchange": "nvlink peer-memory mailboxes" if comm.peer_memory else "ncclAllReduce (library communicator)",
                       "hist_bit_exact_vs_nccl_allreduce": bool(torch.equal(got[1:], ref_nccl[1:])),
                       "hist_bit_exact_vs_rank_order_sum": bool(torch.equal(got[1:], ref_sum[1:])),
                       "loss_rel_vs_nccl_allreduce": rel(got[0], ref_nccl[0]), "loss_rel_vs_rank_order_sum": rel(got[0], ref_sum[0]),
                       "loss_bit_identical_to_rank_order_sum": bool(float(got[0]) == float(ref_sum[0])),
                       "hist_total": hist_total, "expected_hist_total": float(world * PX_PER_STEP), "mailbox_errors": err,
                       "loss_allreduced": float(got[0]), "ok": bool(int(okt.item()) == 1)}

    # ---- the same step in the "hot" regime: one buffer set every iteration (the 1.27 MB of logits and the 16.8 MB
    #      statistics cache then live in L2); reported next to the cold number above, never instead of it
    hot_ms = time_loop(lambda i: step(i & 3), min(steps, 1000), 20)     # 4 buffer sets (each step in flight owns its `out`)
    if comm is not None:
        comm.join(stream)
        torch.cuda.synchronize()
    hot_ms = max_over_ranks(hot_ms, dev, dist, world)

    # ---- e2e (headline): the reference's API.  IW_MaxSquareloss nn.Module + autograd with host copies every step
    e2e_pool = min(POOL, 16)
    host_in = [lo_pool[i].cpu().pin_memory() for i in range(e2e_pool)]
    crit = msq.IW_MaxSquareloss(-1, C, RATIO)
    crit.global_batch = n_norm
    API_DEPTH = 4
    api_steps = max(20, min(steps, 1000))
    if world > 1:
        dist.barrier()
    api_s, api_last_loss = api_e2e_leg(msq, crit, host_in, dev, api_steps, API_DEPTH, False, comm)
    api_s = max_over_ranks(api_s, dev, dist, world)
    if world > 1:
        dist.barrier()
    api_sync_s, _ = api_e2e_leg(msq, crit, host_in, dev, min(api_steps, 500), 1, True, comm)
    api_sync_s = max_over_ranks(api_sync_s, dev, dist, world)
    floor_s = max_over_ranks(torch_floor(dev, host_in, min(api_steps, 500)), dev, dist, world)
    e2e_val = world * PX_PER_STEP / api_s / 1e9

    # ---- the same through the C-ABI host-buffer pipeline (maxsquareloss_b200.HostPipeline -> msq_pipe_submit, three
    #      streams, `depth` steps in flight): no PyTorch autograd; sharded runs use the global normaliser and the exchange
    depth = int(os.environ.get("MSQ_BENCH_DEPTH", "16"))         # a step is ~100 us of latency end to end (H2D, kernels, D2H)
    pipe = msq.HostPipeline("iw", N_IMG, C, HW_LO, HW_OUT, ratio=RATIO, depth=depth, comm=comm,
                            global_batch=n_norm if world > 1 else 0)
    h_grad = [torch.empty(N_IMG, C, *HW_LO).pin_memory() for _ in range(depth)]
    h_loss = [torch.empty(()).pin_memory() for _ in range(depth)]
    pipe_steps = max(steps, 200)

    def pipe_run(nsteps):
        slots = []
        for i in range(nsteps):
            j = i % depth
            if len(slots) >= depth:
                pipe.wait(slots[i - depth])          # host reads step i-depth's loss/grad before reusing its buffers
            slots.append(pipe.submit(host_in[i % e2e_pool], h_loss[j], h_grad[j], None, LAMBDA_TARGET))
        pipe.drain()

    pipe_run(max(warm, 64))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipe_run(pipe_steps)
    pipe_s = max_over_ranks(time.perf_counter() - t0, dev, dist, world)
    pipe_val = world * PX_PER_STEP * pipe_steps / pipe_s / 1e9
    pipe_last_loss = float(h_loss[(pipe_steps - 1) % depth].item())
    pipe.close()

    # ---- per-kernel numbers of the step (rank 0 reports; every rank runs them to stay in step)
    kit = max(50, min(steps, 400))
    lo_bytes = 4.0 * n_lo
    t_fwd = time_loop(fwd, kit, 20)
    t_bwd = time_loop(bwd, kit, 20)
    # ALGORITHMIC bytes: the forward reads the low-resolution logits once (4 C h w per image); the backward
    # reads them again and writes dL/dlogits (8 C h w).  The 16 B/pixel statistics cache between the two is this
    # implementation's own traffic, NOT algorithmic: it is listed as `cache_bytes` and shows up in `traffic`.
    fwd_bytes, bwd_bytes = lo_bytes, 2 * lo_bytes
    cache_bytes = 16.0 * PX_PER_STEP
    kernels = [
        {"kernel": "fused_fwd_kernel<19,IW> + finalize_kernel", "bound": "issue/MUFU (not HBM)", "api": "msq_fused_fwd (the two-call path: the nn.Module's forward)",
         "algorithmic_bytes": fwd_bytes, "cache_bytes_written": cache_bytes, "ms": t_fwd, "achieved_GBps": fwd_bytes / t_fwd / 1e6,
         "frac_of_hbm": fwd_bytes / t_fwd / 1e6 / hbm_peak, "gpixel_per_s": PX_PER_STEP / t_fwd / 1e6},
        {"kernel": "fused_bwd_kernel<19,IW,cached>", "bound": "issue/MUFU (not HBM)", "api": "msq_fused_bwd (the two-call path: the nn.Module's backward)",
         "algorithmic_bytes": bwd_bytes, "cache_bytes_read": cache_bytes, "ms": t_bwd, "achieved_GBps": bwd_bytes / t_bwd / 1e6,
         "frac_of_hbm": bwd_bytes / t_bwd / 1e6 / hbm_peak, "gpixel_per_s": PX_PER_STEP / t_bwd / 1e6},
    ]
    extra = {}
    if not args.skip_secondary:
        c3 = cfg3_leg(lib, _lib, synth, mdist, comm, dev, stream, rank, world, dist, kit)
        ch = confusion_hist_leg(lib, _lib, synth, msq, dev, stream, rank, world, hbm_peak, dist, comm)
        c5 = crosscity_leg(dev, rank, world, dist, comm)
        if rank == 0:
            extra["cfg3_multi_level"] = c3
            extra["confusion_hist"] = ch
            extra["cfg5_crosscity"] = c5
    if rank == 0 and world == 1 and not args.skip_secondary:        # per-GPU numbers: reported by the N = 1 run only
        kernels += secondary_kernels(lib, _lib, synth, dev, stream, hbm_peak, kit)
        extra["maxsquare"] = maxsquare_variant(lib, _lib, lo_ptrs, gr_ptrs, out_ptrs, aux_ptrs, acc_ptr, go_ptr, n_norm, stream, kit)
        extra["marginal_image"] = marginal_image_cost(lib, _lib, dev, stream, kit)
        extra["next_rows"] = next_rows(lib, _lib, synth, dev, stream, kit)
        extra["torch_cuda_eager_baseline"] = torch_eager_gpu(dev)
    dom = max(kernels[:2], key=lambda k: k["ms"])
    which = "fused_fwd" if "fwd" in dom["kernel"] else "fused_bwd"
    traffic, traffic_src, tk = None, None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):       # DRAM bytes / instruction counts of one launch from the committed ncu --set full capture
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tj = json.load(f)
            tk = tj["kernels"][which]
            traffic, traffic_src = tk["dram_read_bytes"] + tk["dram_write_bytes"], tj["source"]
            break
        except Exception:
            continue
    issue = None
    if tk is not None:        # the bound that does apply: warp-instruction issue slots (SMs x 4 schedulers x SM clock)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_ginst = sms * 4 * (clocks.get("sm_mhz") or 1965) / 1e3
        issue = {"kernel": dom["kernel"], "bound": "warp-instruction issue slots", "achieved": tk["warp_instructions"] / dom["ms"] / 1e6,
                 "peak": peak_ginst, "unit": "Ginst/s", "frac": tk["warp_instructions"] / dom["ms"] / 1e6 / peak_ginst,
                 "warp_instructions_per_launch": tk["warp_instructions"],
                 "ncu_issue_active_pct": tk["issue_active_pct"],
                 "ncu_pipes_pct": {"xu": tk["xu_pipe_pct"], "fma": tk["fma_pipe_pct"], "alu": tk["alu_pipe_pct"]},
                 "source": traffic_src,
                 "note": "instruction count from the committed ncu capture; time = this run's CUDA-event time of the launch through the "
                         "two-call API (incl. the dependent finalisation launch for the forward); peak = SMs x 4 schedulers x the SM "
                         "clock sampled in this run"}
    roofline = {"bound": "hbm", "achieved": dom["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": dom["achieved_GBps"] / hbm_peak, "traffic": traffic, "kernel": dom["kernel"],
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "launch_ms": dom["ms"],
                "peak_source": peak_src, "traffic_source": traffic_src,
                "note": "SURVEY 8d: 4*C*h*w bytes per image forward, 8*C*h*w backward (3.65 B/pixel for the pair); the fused "
                        "kernels are FP32-issue/MUFU bound by design, so this fraction is ~1 % and `issue_roofline` is the binding "
                        "figure; the HBM-bound kernels of the path (>= 70 % target) are listed under 'kernels' with their own fractions"}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_baseline()
        if "confusion_hist" in extra:
            cpu["confusion_hist_port"] = cpu_eval_check(extra["confusion_hist"])

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": warm_steps + REWARM, "warmup_requested": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": shared_config(world),
                "method": {
                    "timing": "CUDA events around exactly `steps` steps on the launching stream, max over ranks; barrier + synchronize "
                              f"on both sides; {warm_steps} warm-up steps before the leading barrier and {REWARM} after it, back to "
                              "back with the timed ones",
                    "l2_policy": f"inputs rotate over {POOL} distinct logits buffers ({POOL * lo_bytes / 1e6:.0f} MB > 126 MB L2); "
                                 "outputs and statistics caches likewise",
                    "hot_regime": {"ms_per_step": hot_ms, "value": world * PX_PER_STEP / hot_ms / 1e6,
                                   "what": "4 buffer sets reused round-robin (inputs and statistics caches, 73 MB, stay in L2)"},
                    "parallelism": f"image-sharded x{world}" + ("" if world == 1 else
                        ", [loss,hist] of every step exchanged over NVLink peer-memory mailboxes by an extra CTA of the step's own "
                        "backward kernel (no NCCL call, no extra launch)" if comm.peer_memory else
                        ", 1 ncclAllReduce of [loss,hist] per step on the library's own communicator, overlapped with the next step")},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(lo_bytes),
                        "d2h_bytes_per_step": int(lo_bytes) + 4, "steps": api_steps, "ms_per_step": api_s * 1e3,
                        "how": "the reference's API as tools/solve_gta5.py:199,217 call it: IW_MaxSquareloss nn.Module (C++ autograd "
                               "node over the C ABI) -> (lambda_target * loss).backward(); per step pinned host logits H2D, loss + "
                               f"dL/dlogits D2H to pinned memory; the host reads step i-{API_DEPTH}'s results when it reuses that slot "
                               "(CUDA event), wall clock, max over ranks" + ("; each step also all-reduces [loss,hist] (StatsComm)" if world > 1 else ""),
                        "last_loss": api_last_loss,
                        "sync_every_step": {"value": world * PX_PER_STEP / api_sync_s / 1e9, "ms_per_step": api_sync_s * 1e3,
                                            "how": "same, but the host waits for each step's loss + gradient before issuing the next"},
                        "torch_floor": {"ms_per_step": floor_s * 1e3,
                                        "how": "the same loop with x.sum() as the loss (one native kernel): PyTorch's own host cost of "
                                               "detach + loss + lambda*loss + backward() + the copies, which bounds any nn.Module"},
                        "c_abi_pipeline": {"value": pipe_val, "ms_per_step": pipe_s / pipe_steps * 1e3, "steps": pipe_steps,
                                           "last_loss": pipe_last_loss,
                                           "how": "HostPipeline.submit -> C ABI msq_pipe_submit (3-stream software pipeline, no autograd): "
                                                  f"per step pinned host logits H2D, fused fwd+bwd, loss + dL/dlogits D2H; {depth} steps in "
                                                  "flight" + ("; global normaliser + statistics exchange as in the device-timed loop" if world > 1 else "")}},
                "gpu_launches": launches,
                "gpu_launches_how": "counted by the library (msq_launch_count) over the timed region: two per step -- the fused forward "
                                    "and the fused backward, which carries the finalisation in an extra CTA" +
                                    (" -- + one flush kernel at the closing join" if world > 1 else ""),
                "roofline": roofline, "issue_roofline": issue, "kernels": kernels}
        line.update(extra)
        for k in [k for k in line.get("confusion_hist", {}) if k.startswith("_")]:
            del line["confusion_hist"][k]
        if rank_spread is not None:
            line["rank_spread"] = rank_spread
        if stats_check is not None:
            line["stats_check"] = stats_check
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def confusion_hist_leg(lib, _lib, synth, msq, dev, stream, rank, world, hbm_peak, dist, comm):
    """cfg 4 (SYNTHIA->Cityscapes evaluation): Eval fast_hist + mIoU over 500 synthetic 512x1024 validation images,
    16 classes, sharded round-robin by image over the ranks (strong scaling: the 500 images are the job).  Each
    rank accumulates its images into its own device matrix; ONE exact uint64 all-reduce of the C*C counts at the end
    (the library's communicator: msq_comm_allreduce_u64); the metrics are then read once.  Timed on the device with
    CUDA events, max over ranks.
      per_image            one msq_confusion_i64 launch per image, as tools/train_source.py:429-492 calls add_batch
      per_image_eval_api   the same through Eval.add_batch (the reference's API, one call per image)
      eval_api_deferred16  Eval(defer=16).add_batch: the calls are queued and run 16 per launch (msq_confusion_i64_multi)
      batched16            a caller that stacks its validation batch: 16 images per msq_confusion_i64 call
      logits_per_image     per image from fp32 logits (1,16,512,1024): the callers' np.argmax fused into the kernel"""
    Cv, HWv, n_total, pool = 16, (512, 1024), 500, 128
    mine = list(range(rank, n_total, world))
    px_img = HWv[0] * HWv[1]
    # validation image i is pool entry i % pool on EVERY sharding, so the accumulated matrix (and mIoU) does not depend on the
    # rank count; a rank touches >= 16 distinct entries (>= 134 MB > L2) at up to 8 ranks and generates only those
    entries = sorted({i % pool for i in mine})
    gts = {e: synth.blocky_labels(1, HWv, Cv, 1000 + e).to(dev) for e in entries}
    prs = {e: synth.noisy_prediction(gts[e].cpu(), Cv, 1000 + e).to(dev) for e in entries}
    gp, pp = {e: t.data_ptr() for e, t in gts.items()}, {e: t.data_ptr() for e, t in prs.items()}
    ev = msq.Eval(Cv, device=dev)
    ev16 = msq.Eval(Cv, device=dev, defer=16)
    cm_ptr = ev._dev.data_ptr()

    def timed(fn, e=ev):
        fn()                                                 # warm (also the clocks: the caller just ran the step loop)
        e.reset()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        if comm is not None:
            comm.allreduce_u64(e.device_counts())            # uint64 counts over NCCL/NVLink: exact
            comm.join()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        ms = max_over_ranks(ms, dev, dist, world)
        e._pending = True
        return ms, e.Mean_Intersection_over_Union(), e.confusion_matrix.copy()

    def per_image():
        for i in mine:
            j = i % pool
            rc = lib.msq_confusion_i64(gp[j], pp[j], px_img, Cv, cm_ptr, cm_ptr + 8 * Cv * Cv, stream)
            if rc:
                _lib.check(rc)

    def api_per_image():
        for i in mine:
            ev.add_batch(gts[i % pool], prs[i % pool])

    def api_deferred():
        for i in mine:
            ev16.add_batch(gts[i % pool], prs[i % pool])
        ev16.flush()

    B = 16
    groups = [entries[:B], entries[B:2 * B] if len(entries) >= 2 * B else entries[:B]]
    stacks = [(torch.cat([gts[e] for e in grp]).contiguous(), torch.cat([prs[e] for e in grp]).contiguous()) for grp in groups]

    def batched():
        left, k = len(mine), 0
        while left > 0:
            nb = min(B, left)
            g_, p_ = stacks[k % 2]
            rc = lib.msq_confusion_i64(g_.data_ptr(), p_.data_ptr(), nb * px_img, Cv, cm_ptr, cm_ptr + 8 * Cv * Cv, stream)
            if rc:
                _lib.check(rc)
            left -= nb
            k += 1

    lg_pool = 6                                              # 6 x 33.5 MB = 201 MB > L2
    lgs = [torch.randn(1, Cv, *HWv, device=dev) for _ in range(lg_pool)]
    lgp = [t.data_ptr() for t in lgs]

    def logits():
        for k, i in enumerate(mine):
            rc = lib.msq_confusion_logits_f32(gp[i % pool], lgp[k % lg_pool], 1, Cv, px_img, cm_ptr, stream)
            if rc:
                _lib.check(rc)

    res = {"workload": f"cfg4 Eval over {n_total} synthetic 512x1024 val images, {Cv} classes (blocky gt, 30% noisy pred), "
                       f"images sharded round-robin over {world} rank(s), one exact uint64 all-reduce of the matrix at the end "
                       "(msq_comm_allreduce_u64, the library's NCCL communicator)",
           "images": n_total, "scaling": "strong", "unit": UNIT}
    px_total = float(n_total) * px_img
    ref_cm = None
    for name, fn, bpp, e in (("per_image", per_image, 16.0, ev), ("per_image_eval_api", api_per_image, 16.0, ev),
                             ("eval_api_deferred16", api_deferred, 16.0, ev16), ("batched16", batched, 16.0, ev),
                             ("logits_per_image", logits, 4.0 * Cv + 8, ev)):
        ms, miou, cm = timed(fn, e)
        res[name] = {"value": px_total / ms / 1e6, "ms_total": ms, "us_per_image_per_rank": ms * 1e3 / len(mine),
                     "frac_of_hbm_aggregate": bpp * px_total / ms / 1e6 / (hbm_peak * world)}
        if name == "per_image":
            ref_cm = cm
            res["miou_16_13"] = [float(miou[0]), float(miou[1])]
            res["matrix_total"] = int(cm.sum())
            res["matrix_sha1"] = __import__("hashlib").sha1(cm.astype("int64").tobytes()).hexdigest()
        elif name in ("per_image_eval_api", "eval_api_deferred16"):
            res[name]["matrix_equals_per_image"] = bool((cm == ref_cm).all())
    res["_pool"], res["_classes"], res["_hw"] = pool, Cv, list(HWv)
    return res


def cpu_eval_check(ch):
    """cpu_baseline leg (rank 0, N = 1): the cfg-4 matrix rebuilt by the NumPy port of the reference's ``Eval``
    (oracle/eval_port.py) over the same 500 images -- equality with the GPU matrix, and the port's time."""
    import hashlib
    import numpy as np
    from maxsquareloss_b200 import synth
    from oracle import eval_port
    pool, Cv, HWv, n_total = ch["_pool"], ch["_classes"], tuple(ch["_hw"]), ch["images"]
    port = eval_port.EvalPort(Cv)
    pairs = {}
    t_port = 0.0
    for i in range(n_total):
        e = i % pool
        if e not in pairs:
            gt = synth.blocky_labels(1, HWv, Cv, 1000 + e)
            pairs[e] = (gt.numpy(), synth.noisy_prediction(gt, Cv, 1000 + e).numpy())
        t0 = time.perf_counter()
        port.add_batch(*pairs[e])
        t_port += time.perf_counter() - t0
    cm = np.asarray(port.confusion_matrix)
    miou = port.Mean_Intersection_over_Union()
    return {"kind": "port", "cores": 1, "images": n_total, "value": n_total * HWv[0] * HWv[1] / t_port / 1e9, "unit": UNIT,
            "ms_total": t_port * 1e3,
            "matrix_bit_exact_vs_gpu": hashlib.sha1(cm.astype("int64").tobytes()).hexdigest() == ch["matrix_sha1"],
            "miou_bit_exact_vs_gpu": [float(miou[0]), float(miou[1])] == ch["miou_16_13"]}


def cfg3_leg(lib, _lib, synth, mdist, comm, dev, stream, rank, world, dist, kit):
    """BASELINE config 3: MaxSquare+IW+Multi (tools/solve_gta5.py:178-218 with --multi): IW-MaxSquare on head 1 +
    self-produced guidance cross-entropy on head 2, GLOBAL batch 8 strong-sharded over the ranks (8/G images each).
    Per step and rank: msq_multi_fwd (both heads, one kernel + finalisation), the exact integer sum of the
    [cross-entropy sum | valid-pixel count] pair begun right after it (the head-2 backward divides by the GLOBAL count),
    msq_fused_bwd for head 1 (the pairs cross NVLink meanwhile), the sum's end, msq_guidance_bwd for head 2; the
    [loss | hist] vector of head 1 is all-reduced as well (NCCL, side stream).  check: the all-reduced integers against ONE GPU computing all 8 images."""
    NG, thr = 8, 0.95
    if NG % world:
        return {"skipped": f"global batch {NG} does not divide over {world} ranks"}
    lo_i, hi_i = mdist.image_shard(NG, rank, world)
    nl = hi_i - lo_i
    h, w = HW_LO
    H, W = HW_OUT
    pool = 6
    full1 = [synth.head_logits(NG, C, HW_LO, 7000 + p, 5.0) for p in range(pool)]
    full2 = [synth.second_head(full1[p], 7000 + p) for p in range(pool)]
    lo1 = [t[lo_i:hi_i].contiguous().to(dev) for t in full1]
    lo2 = [t[lo_i:hi_i].contiguous().to(dev) for t in full2]
    g1, g2 = [torch.empty_like(t) for t in lo1], [torch.empty_like(t) for t in lo2]
    lay = _lib.state_layout(nl, C)
    accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
    outs = [torch.empty(lay.out_bytes, dtype=torch.uint8, device=dev) for _ in range(pool)]
    nb = lib.msq_fused_aux_bytes(nl, H, W)
    naux = max(2, min(pool, int(200e6 // (2 * nb)) + 1))
    aux1 = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(naux)]
    aux2 = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(naux)]
    go = torch.full((), LAMBDA_TARGET, device=dev)
    go2 = torch.full((), LAMBDA_TARGET * 0.1, device=dev)                 # lambda_seg * lambda_target (train_source.py:827)
    comm_h = comm._h if comm is not None else None

    def run(j, l1, l2, o, a1, a2, gr1, gr2, n, acc, sharded):
        rc = lib.msq_multi_fwd(_lib.MODE_IW, l1.data_ptr(), l2.data_ptr(), n, C, h, w, H, W, RATIO, thr, NG,
                               acc.data_ptr(), o.data_ptr(), a1.data_ptr(), a2.data_ptr(), gr1.data_ptr(), gr2.data_ptr(), None, stream)
        if rc:
            _lib.check(rc)
        if sharded:         # push this rank's [ce_fix | nvalid] into every rank's mailbox (NVLink) ...
            _lib.check(lib.msq_comm_sum_u64_begin(comm_h, o.data_ptr() + lay_of(n).ce_fix_out_off, 2, stream))
        rc = lib.msq_fused_bwd(_lib.MODE_IW, l1.data_ptr(), n, C, h, w, H, W, NG, o.data_ptr(), a1.data_ptr(), go.data_ptr(),
                               gr1.data_ptr(), 1, stream)
        if rc:
            _lib.check(rc)
        if sharded:         # ... and sum the pairs of all ranks (arrived during the head-1 backward) in place
            _lib.check(lib.msq_comm_sum_u64_end(comm_h, o.data_ptr() + lay_of(n).ce_fix_out_off, 2, stream))
        rc = lib.msq_guidance_bwd(l2.data_ptr(), n, C, h, w, H, W, o.data_ptr(), a2.data_ptr(), go2.data_ptr(), gr2.data_ptr(), 1, stream)
        if rc:
            _lib.check(rc)
        if sharded:
            _lib.check(lib.msq_comm_allreduce_f64(comm_h, o.data_ptr() + lay_of(n).stats_off, 1 + C, stream))

    def lay_of(n):
        return _lib.state_layout(n, C)

    sharded = comm is not None

    def step(i):
        j = i % pool
        run(j, lo1[j], lo2[j], outs[j], aux1[i % naux], aux2[i % naux], g1[j], g2[j], nl, accum, sharded)

    it = max(30, min(kit, 200))
    for i in range(20):
        step(i)
    if sharded:
        comm.join()
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(it):
        step(i)
    if sharded:
        comm.join()
    b.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(a.elapsed_time(b) / it, dev, dist, world)
    px = NG * H * W
    res = {"what": "cfg3 MaxSquare+IW+Multi step (tools/solve_gta5.py:178-218 --multi): IW-MaxSquare head 1 + guidance CE head 2, "
                   f"forward + both backwards, global batch {NG} strong-sharded over {world} rank(s) ({nl} images each), "
                   "65x129 -> 512x1024, thr 0.95, lambda 0.1/0.1",
           "scaling": "strong", "us_per_step": ms * 1e3, "value": px / ms / 1e6, "unit": UNIT, "launches_per_step": 4 + (2 if sharded and comm.peer_memory else 0),
           "exchange": None if not sharded else ("same-step integer sum of [ce_fix | nvalid] over the NVLink peer-memory mailboxes "
                                                 "(msq_comm_sum_u64_begin after the forward, _end before the head-2 backward: two "
                                                 "32-thread kernels)" if comm.peer_memory else "ncclAllReduce(uint64) of [ce_fix | nvalid] "
                                                 "between forward and head-2 backward") + " + fp64 ncclAllReduce of [loss | hist] on the side stream"}
    # ---- parity of the sharded step against ONE GPU doing all 8 images (pool entry 0)
    step(0)
    if sharded:
        comm.join()
    torch.cuda.synchronize()
    o = outs[0]
    pair = o[lay.ce_fix_out_off:lay.ce_fix_out_off + 16].view(torch.int64).clone()
    stats = o[lay.stats_off:lay.stats_off + 8 * (1 + C)].view(torch.float64).clone()
    grad1_mine, grad2_mine = g1[0].clone(), g2[0].clone()
    if sharded:
        layf = _lib.state_layout(NG, C)
        f1, f2 = full1[0].to(dev), full2[0].to(dev)
        fg1, fg2 = torch.empty_like(f1), torch.empty_like(f2)
        facc = torch.zeros(layf.accum_bytes, dtype=torch.uint8, device=dev)
        fo = torch.empty(layf.out_bytes, dtype=torch.uint8, device=dev)
        fnb = lib.msq_fused_aux_bytes(NG, H, W)
        fa1, fa2 = torch.empty(fnb, dtype=torch.uint8, device=dev), torch.empty(fnb, dtype=torch.uint8, device=dev)
        run(0, f1, f2, fo, fa1, fa2, fg1, fg2, NG, facc, False)
        torch.cuda.synchronize()
        rpair = fo[layf.ce_fix_out_off:layf.ce_fix_out_off + 16].view(torch.int64)
        rstats = fo[layf.stats_off:layf.stats_off + 8 * (1 + C)].view(torch.float64)
        rel = lambda x, y: abs(float(x) - float(y)) / max(abs(float(y)), 1e-300)          # noqa: E731
        gmax = lambda x, y: float((x - y).abs().max() / y.abs().max().clamp_min(1e-30))   # noqa: E731
        chk = {"label2_valid_count_bit_exact": bool(int(pair[1]) == int(rpair[1])), "label2_valid_count": int(pair[1]),
               "class_hist_bit_exact": bool(torch.equal(stats[1:], rstats[1:])),
               "ce_sum_rel": rel(pair[0], rpair[0]), "loss_head1_rel": rel(stats[0], rstats[0]),
               "grad_head1_max_rel": gmax(grad1_mine, fg1[lo_i:hi_i]), "grad_head2_max_rel": gmax(grad2_mine, fg2[lo_i:hi_i]),
               "reference": "the same 8 images on ONE GPU (this rank), same kernels, n_images_norm = 8"}
        ok = (chk["label2_valid_count_bit_exact"] and chk["class_hist_bit_exact"] and chk["ce_sum_rel"] <= 1e-6 and
              chk["loss_head1_rel"] <= 1e-6 and chk["grad_head1_max_rel"] <= 1e-4 and chk["grad_head2_max_rel"] <= 1e-4)
        okt = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        chk["ok"] = bool(int(okt.item()) == 1)
        res["check"] = chk
        del f1, f2, fg1, fg2, fa1, fa2
    else:
        res["reference_values"] = {"label2_valid_count": int(pair[1]), "ce_sum": float(pair[0]) * 2.0 ** -32,
                                   "loss_head1": float(stats[0]), "class_hist_total": float(stats[1:].sum())}
    return res


def secondary_kernels(lib, _lib, synth, dev, stream, hbm_peak, kit):
    """The HBM-bound kernels of the path, each on buffers that exceed L2."""
    out = []
    kit = min(kit, 100)
    npx = PX_PER_STEP
    # strict drop-in on full-resolution prob (2 x 19 x 512 x 1024 fp32 = 159 MB per buffer)
    probs = [torch.softmax(torch.randn(N_IMG, C, *HW_OUT, device=dev) * 3, 1) for _ in range(2)]
    grads = [torch.empty_like(p) for p in probs]
    lay = _lib.state_layout(N_IMG, C)
    accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
    o = torch.empty(lay.out_bytes, dtype=torch.uint8, device=dev)
    go = torch.ones((), device=dev)
    hw = HW_OUT[0] * HW_OUT[1]
    for mode, name in ((_lib.MODE_IW, "IW"), (_lib.MODE_MAXSQUARE, "MaxSquare")):
        f = lambda i: lib.msq_prob_fwd(mode, probs[i % 2].data_ptr(), N_IMG, C, hw, None, RATIO, -1, 0,
                                       accum.data_ptr(), o.data_ptr(), stream)
        b = lambda i: lib.msq_prob_bwd(mode, probs[i % 2].data_ptr(), N_IMG, C, hw, -1, 0, o.data_ptr(),
                                       go.data_ptr(), grads[i % 2].data_ptr(), stream)
        tf, tb = time_loop(f, kit, 5), time_loop(b, kit, 5)
        for nm, t, byt in ((f"prob_fwd_kernel<19,{name}>", tf, 4.0 * C * npx), (f"prob_bwd_kernel<19,{name}>", tb, 8.0 * C * npx)):
            out.append({"kernel": nm, "bound": "hbm", "algorithmic_bytes": byt, "ms": t, "achieved_GBps": byt / t / 1e6,
                        "frac_of_hbm": byt / t / 1e6 / hbm_peak, "gpixel_per_s": npx / t / 1e6})
    del probs, grads
    # confusion matrix, source-batch shape of cfg 2 (2 x 720 x 1280), buffers rotate past L2
    n_src, hw_src = 2, (720, 1280)
    pool = 8
    gts = [synth.blocky_labels(n_src, hw_src, C, 100 + i).to(dev) for i in range(pool)]
    prs = [synth.noisy_prediction(gts[i].cpu(), C, 100 + i).to(dev) for i in range(pool)]
    cm = torch.zeros(C * C + 1, dtype=torch.int64, device=dev)
    px_src = n_src * hw_src[0] * hw_src[1]
    f = lambda i: lib.msq_confusion_i64(gts[i % pool].data_ptr(), prs[i % pool].data_ptr(), px_src, C, cm.data_ptr(),
                                        cm.data_ptr() + 8 * C * C, stream)
    t = time_loop(f, kit, 5)
    out.append({"kernel": "confusion_i64_kernel (blocky gt, 30% noisy pred)", "bound": "hbm", "algorithmic_bytes": 16.0 * px_src,
                "ms": t, "achieved_GBps": 16.0 * px_src / t / 1e6, "frac_of_hbm": 16.0 * px_src / t / 1e6 / hbm_peak,
                "gpixel_per_s": px_src / t / 1e6})
    lgs = [torch.randn(n_src, C, *hw_src, device=dev) for _ in range(2)]
    f = lambda i: lib.msq_confusion_logits_f32(gts[i % pool].data_ptr(), lgs[i % 2].data_ptr(), n_src, C,
                                               hw_src[0] * hw_src[1], cm.data_ptr(), stream)
    t = time_loop(f, kit, 5)
    byt = (4.0 * C + 8) * px_src
    out.append({"kernel": "confusion_logits_kernel<19> (argmax fused)", "bound": "hbm", "algorithmic_bytes": byt, "ms": t,
                "achieved_GBps": byt / t / 1e6, "frac_of_hbm": byt / t / 1e6 / hbm_peak, "gpixel_per_s": px_src / t / 1e6})
    del lgs
    # flip-ensemble evaluation (tools/evaluate.py --flip): two logits tensors + labels, one pass
    fa = [torch.randn(N_IMG, C, *HW_OUT, device=dev) * 3 for _ in range(2)]
    fb = [torch.flip(a, dims=[-1]) + torch.randn_like(a) for a in fa]
    gtf = [synth.blocky_labels(N_IMG, HW_OUT, C, 300 + i).to(dev) for i in range(2)]
    f = lambda i: lib.msq_confusion_flip_f32(gtf[i % 2].data_ptr(), fa[i % 2].data_ptr(), fb[i % 2].data_ptr(), N_IMG, C,
                                             HW_OUT[0], HW_OUT[1], cm.data_ptr(), stream)
    t = time_loop(f, kit, 5)
    byt = (8.0 * C + 8) * npx
    out.append({"kernel": "confusion_flip_kernel<19> (2 softmaxes + mirrored average + argmax fused)", "bound": "hbm",
                "algorithmic_bytes": byt, "ms": t, "achieved_GBps": byt / t / 1e6, "frac_of_hbm": byt / t / 1e6 / hbm_peak,
                "gpixel_per_s": npx / t / 1e6})
    del fa, fb
    return out


def next_rows(lib, _lib, synth, dev, stream, kit):
    """Steps of the rows built after the headline path: multi-level guidance (cfg 3), the MinEnt
    losses, and the source-side CE + Eval step, each fused from low-resolution logits; us per step and Gpix/s."""
    res = {}
    kit = min(kit, 200)
    h, w = HW_LO
    H, W = HW_OUT
    pool = 16
    lay = _lib.state_layout(N_IMG, C)
    accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
    out = torch.empty(lay.out_bytes, dtype=torch.uint8, device=dev)
    go = torch.full((), LAMBDA_TARGET, device=dev)
    lo1 = torch.randn(pool, N_IMG, C, h, w, device=dev) * 5
    lo2 = lo1 + 0.5 * torch.randn_like(lo1)
    g1, g2 = torch.empty_like(lo1), torch.empty_like(lo2)
    nb = lib.msq_fused_aux_bytes(N_IMG, H, W)
    aux1 = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(4)]
    aux2 = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(4)]

    def multi(i):
        j, a = i % pool, i % 4
        lib.msq_multi_fwd(_lib.MODE_IW, lo1[j].data_ptr(), lo2[j].data_ptr(), N_IMG, C, h, w, H, W, RATIO, 0.95, 0,
                          accum.data_ptr(), out.data_ptr(), aux1[a].data_ptr(), aux2[a].data_ptr(), g1[j].data_ptr(),
                          g2[j].data_ptr(), None, stream)
        lib.msq_fused_bwd(_lib.MODE_IW, lo1[j].data_ptr(), N_IMG, C, h, w, H, W, 0, out.data_ptr(), aux1[a].data_ptr(),
                          go.data_ptr(), g1[j].data_ptr(), 1, stream)
        lib.msq_guidance_bwd(lo2[j].data_ptr(), N_IMG, C, h, w, H, W, out.data_ptr(), aux2[a].data_ptr(), go.data_ptr(),
                             g2[j].data_ptr(), 1, stream)
    t = time_loop(multi, kit, 20)
    res["multi_level_guidance"] = {"what": "cfg3 step per GPU: IW-MaxSquare on head 1 + self-produced guidance CE on head 2, "
                                           "fwd + both backwards (batch 2, 65x129 -> 512x1024)",
                                   "us_per_step": t * 1e3, "gpixel_per_s": PX_PER_STEP / t / 1e6, "launches_per_step": 4}

    def ent(i):
        j, a = i % pool, i % 4
        lib.msq_entropy_fwd(_lib.MODE_IW, lo1[j].data_ptr(), N_IMG, C, h, w, H, W, RATIO, 0, accum.data_ptr(), out.data_ptr(),
                            aux1[a].data_ptr(), g1[j].data_ptr(), stream)
        lib.msq_entropy_bwd(_lib.MODE_IW, lo1[j].data_ptr(), N_IMG, C, h, w, H, W, 0, out.data_ptr(), aux1[a].data_ptr(),
                            go.data_ptr(), g1[j].data_ptr(), 1, stream)
    t = time_loop(ent, kit, 20)
    res["iw_entropy"] = {"what": "IWsoftCrossEntropy fwd+bwd, same shape", "us_per_step": t * 1e3,
                         "gpixel_per_s": PX_PER_STEP / t / 1e6, "launches_per_step": 3}

    # source-side step of cfg 2: 2 x 19 x 91x161 -> 720x1280, CE + argmax + confusion matrix, then backward
    hs, ws, Hs, Ws = 91, 161, 720, 1280
    los = torch.randn(pool, N_IMG, C, hs, ws, device=dev) * 3
    gs = torch.empty_like(los)
    ys = [synth.blocky_labels(N_IMG, (Hs, Ws), C, 500 + i).to(dev) for i in range(4)]
    auxs = [torch.empty(lib.msq_fused_aux_bytes(N_IMG, Hs, Ws), dtype=torch.uint8, device=dev) for _ in range(4)]
    cm = torch.zeros(C * C + 1, dtype=torch.int64, device=dev)

    def src(i):
        j, a = i % pool, i % 4
        lib.msq_source_ce_fwd(los[j].data_ptr(), ys[a].data_ptr(), N_IMG, C, hs, ws, Hs, Ws, accum.data_ptr(), out.data_ptr(),
                              auxs[a].data_ptr(), gs[j].data_ptr(), cm.data_ptr(), stream)
        lib.msq_guidance_bwd(los[j].data_ptr(), N_IMG, C, hs, ws, Hs, Ws, out.data_ptr(), auxs[a].data_ptr(), go.data_ptr(),
                             gs[j].data_ptr(), 1, stream)
    t = time_loop(src, kit, 20)
    px = N_IMG * Hs * Ws
    res["source_ce_eval"] = {"what": "cfg2 source step per GPU: CrossEntropyLoss + argmax + confusion matrix fwd, CE bwd "
                                     "(batch 2, 91x161 -> 720x1280)", "us_per_step": t * 1e3, "gpixel_per_s": px / t / 1e6,
                             "launches_per_step": 3}

    # cfg 4, loss half: 16-class IW-MaxSquare at the SYNTHIA source shape (96x161 -> 760x1280), the one-call step
    c4, h4, w4, H4, W4 = 16, 96, 161, 760, 1280
    lay4 = _lib.state_layout(N_IMG, c4)
    accum4 = torch.zeros(lay4.accum_bytes, dtype=torch.uint8, device=dev)
    out4 = torch.empty(lay4.out_bytes, dtype=torch.uint8, device=dev)
    lo4 = torch.randn(pool, N_IMG, c4, h4, w4, device=dev) * 5
    g4 = torch.empty_like(lo4)
    aux4 = [torch.empty(lib.msq_fused_aux_bytes(N_IMG, H4, W4), dtype=torch.uint8, device=dev) for _ in range(4)]

    def iw16(i):
        j, a = i % pool, i % 4
        rc = lib.msq_fused_fwd_bwd(_lib.MODE_IW, lo4[j].data_ptr(), N_IMG, c4, h4, w4, H4, W4, RATIO, 0, accum4.data_ptr(),
                                   out4.data_ptr(), aux4[a].data_ptr(), go.data_ptr(), 0.0, g4[j].data_ptr(), None, 0, stream)
        if rc:
            _lib.check(rc)
    t = time_loop(iw16, kit, 20)
    px4 = N_IMG * H4 * W4
    res["iw_16class_synthia_shape"] = {"what": "cfg4, loss half: 16-class IW-MaxSquare fwd+bwd, one-call step (batch 2, 96x161 -> "
                                               "760x1280)", "us_per_step": t * 1e3, "gpixel_per_s": px4 / t / 1e6,
                                       "launches_per_step": 2}
    return res


def crosscity_leg(dev, rank, world, dist, comm, iters=8):
    """BASELINE config 5: the adaptation step of tools/solve_crosscity.py:165-249 (no --multi) around a random-init
    DeepLabv2-ResNet101 (harness/deeplabv2.py: the reference's topology, cuDNN, NOT the product), 13 classes, 512x1024,
    batch 1 per GPU, SGD step included.  'fused' = low-resolution heads into CrossEntropyLoss2d(+Eval) and
    IW_MaxSquareloss.  One GPU: also 'reference_chain' = the same step with the model's two F.interpolate calls, torch
    softmax / cross-entropy, the oracle port of IW_MaxSquareloss (per-image D2H + CPU histc + H2D) and the callers' D2H of
    the logits + np.argmax + Eval.add_batch (numpy port), and the backbone alone.  N GPUs: every rank steps its own image
    (same initial weights); the loss normaliser is the global batch, the source cross-entropy averages over the valid
    pixels of ALL ranks (exact uint64 all-reduce, StatsComm) and the confusion matrix is all-reduced at the end; the
    backbone-gradient all-reduce is DDP's job and not part of the path.  torch defaults (TF32 convolutions)."""
    import numpy as np
    import torch.nn.functional as F
    import maxsquareloss_b200 as msq
    from harness.deeplabv2 import DeepLabV2Harness
    from maxsquareloss_b200 import synth
    C5, HW5 = 13, (512, 1024)
    torch.manual_seed(12345)
    model = DeepLabV2Harness(C5).to(dev).train()
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=2.5e-4, momentum=0.9, weight_decay=5e-4)
    gs = torch.Generator().manual_seed(500 + rank)
    xs, xt = torch.randn(1, 3, *HW5, generator=gs).to(dev), torch.randn(1, 3, *HW5, generator=gs).to(dev)
    ys = synth.blocky_labels(1, HW5, C5, 5 + rank).to(dev)
    ev = msq.Eval(C5, device=dev)
    ce = msq.CrossEntropyLoss2d(ignore_index=-1, group=comm if comm is not None else False)
    iw = msq.IW_MaxSquareloss(-1, C5, 0.2)
    iw.global_batch = world

    def fused():
        lo, _ = model(xs)
        ce(lo, ys, evaluator=ev).backward()
        lt, _ = model(xt)
        (0.1 * iw(lt, out_size=HW5)).backward()
        opt.step()
        opt.zero_grad()

    def timeit(fn):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) / iters * 1e3, dev, dist, world)

    out = {"what": f"cfg5 adaptation step (source CE + Eval, target IW-MaxSquare, SGD), DeepLabv2-ResNet101 random init on cuDNN, "
                   f"batch 1 per GPU x {world} GPU(s), 13 classes, 512x1024; ms per step, max over ranks",
           "scaling": "weak", "fused_ms": timeit(fused)}
    out["images_per_s"] = world / (out["fused_ms"] * 1e-3)
    if world == 1:
        from oracle import eval_port, loss_port
        port = eval_port.EvalPort(C5)

        def ref():
            pred, _ = model(xs, upsample=True)
            F.cross_entropy(pred, ys, ignore_index=-1).backward()
            port.add_batch(ys.cpu().numpy(), np.argmax(pred.data.cpu().numpy(), axis=1))
            tp, _ = model(xt, upsample=True)
            (0.1 * loss_port.iw_maxsquare(F.softmax(tp, 1), C5, 0.2)).backward()
            opt.step()
            opt.zero_grad()

        def backbone_only():
            lo, _ = model(xs)
            lo.sum().backward()
            lt, _ = model(xt)
            lt.sum().backward()
            opt.step()
            opt.zero_grad()
        out["reference_chain_ms"] = timeit(ref)
        out["backbone_only_ms"] = timeit(backbone_only)
        out["miou_fused"] = float(ev.Mean_Intersection_over_Union())
    else:
        # the exchange of the evaluation counts: every rank's local matrix of ONE more forward, all-reduced through the
        # library (exact uint64) against the sum of the all-gathered local matrices
        ev.reset()
        with torch.no_grad():
            lo, _ = model(xs)
            ce(lo, ys, evaluator=ev)               # the fused source kernel: CE + argmax + confusion matrix of this rank's image
        local = ev.device_counts().clone()
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        expect = torch.stack(parts).sum(0)
        comm.allreduce_u64(ev.device_counts())
        comm.join()
        torch.cuda.synchronize()
        got = ev.device_counts().clone()
        ev._pending = True
        okm = bool(torch.equal(got, expect))
        okt = torch.tensor([1 if okm else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        out["check"] = {"confusion_matrix_allreduce_bit_exact": bool(int(okt.item()) == 1), "matrix_total": int(got.sum()),
                        "miou": float(ev.Mean_Intersection_over_Union()),
                        "ce_global_valid_pixels": int(ce.last_nvalid.item()),
                        "ok": bool(int(okt.item()) == 1)}
    del model, opt
    torch.cuda.empty_cache()
    return out


def torch_eager_gpu(dev, iters=20):
    """Secondary baseline: the reference's op chain executed by torch eager on this GPU (what a user of the
    reference actually runs today): F.interpolate -> softmax -> IW loss (with its per-image D2H + CPU histc + H2D)
    -> backward.  Oracle port, timed beside the CUDA path, never part of it."""
    from oracle import loss_port
    lo = torch.randn(N_IMG, C, *HW_LO, device=dev) * 5
    for _ in range(3):
        loss_port.chain_iw_maxsquare(lo, HW_OUT, C, RATIO, LAMBDA_TARGET)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        loss_port.chain_iw_maxsquare(lo, HW_OUT, C, RATIO, LAMBDA_TARGET)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / iters
    return {"value": PX_PER_STEP / t / 1e9, "unit": UNIT, "ms_per_step": t * 1e3,
            "what": "reference op chain (oracle port) on torch CUDA eager, same GPU, device-resident input"}


def marginal_image_cost(lib, _lib, dev, stream, kit):
    """us per step of the one-call fused IW step at 1, 2 and 4 images per GPU (cold inputs) and the slope between them: what
    one more image costs against the fixed part of a step (launches, one-wave ramp-up / tail, per-CTA set-up)."""
    h, w = HW_LO
    H, W = HW_OUT
    out = {}
    for n in (1, 2, 4):
        pool = max(8, min(POOL, int(170e6 // (4 * n * C * h * w)) + 1))
        lo = torch.randn(pool, n, C, h, w, device=dev) * 5.0
        gr = torch.empty_like(lo)
        lay = _lib.state_layout(n, C)
        accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
        outs = torch.empty(8, lay.out_bytes, dtype=torch.uint8, device=dev)
        nb = lib.msq_fused_aux_bytes(n, H, W)
        naux = max(2, min(10, int(170e6 // nb) + 1))
        aux = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(naux)]

        def step(i):
            j = i % pool
            rc = lib.msq_fused_fwd_bwd(_lib.MODE_IW, lo[j].data_ptr(), n, C, h, w, H, W, RATIO, 0, accum.data_ptr(),
                                       outs[i % 8].data_ptr(), aux[i % naux].data_ptr(), None, LAMBDA_TARGET, gr[j].data_ptr(),
                                       None, 0, stream)
            if rc:
                _lib.check(rc)
        out[n] = time_loop(step, max(kit, 300), 50) * 1e3
        del lo, gr, aux
    return {"us_per_step_by_images": {str(k): v for k, v in out.items()},
            "marginal_us_per_image": (out[4] - out[1]) / 3.0, "fixed_us_per_step": out[1] - (out[4] - out[1]) / 3.0,
            "what": "one-call fused IW step (forward + backward with the finalisation in an extra CTA), cold inputs; marginal = (t4 - t1) / 3, "
                    "fixed = t1 - marginal"}


def maxsquare_variant(lib, _lib, lo_ptrs, gr_ptrs, out_ptrs, aux_ptrs, acc_ptr, go_ptr, n_norm, stream, kit):
    """Same step with MaxSquareloss instead of the IW loss (cfg 2 as literally written)."""
    h, w = HW_LO
    H, W = HW_OUT

    def st(i):
        j = i % POOL
        a = aux_ptrs[i % len(aux_ptrs)]
        lib.msq_fused_fwd(_lib.MODE_MAXSQUARE, lo_ptrs[j], N_IMG, C, h, w, H, W, None, 0.0, n_norm, acc_ptr, out_ptrs[j],
                          a, gr_ptrs[j], stream)
        lib.msq_fused_bwd(_lib.MODE_MAXSQUARE, lo_ptrs[j], N_IMG, C, h, w, H, W, n_norm, out_ptrs[j], a, go_ptr,
                          gr_ptrs[j], 1, stream)
    t = time_loop(st, kit, 20)
    return {"metric": "MaxSquare fwd+bwd Gpixel/s", "value": PX_PER_STEP / t / 1e6, "ms_per_step": t}


class _QuietStdout:
    """Route everything written to fd 1 (e.g. NCCL's version banner) to stderr so that the only
    thing on stdout is the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


_OUT = None


def emit(line):
    text = json.dumps(line)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg")
    ap.add_argument("--skip-secondary", action="store_true", help="omit the secondary kernel table")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
    global _OUT
    with _QuietStdout() as q:
        _OUT = q
        run_b200(args, rank, world, local_rank)
    _OUT = None


if __name__ == "__main__":
    main()
