"""How deep the launch queue is changes the measured step time of the one-call fused step (development aid).
Times msq_fused_fwd_bwd over loops of different lengths and with the host kept at most D steps ahead of the GPU."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import _lib
lib = _lib.load()
if os.environ.get("AB_LATE"):
    _lib.tune("late_finalize", int(os.environ["AB_LATE"]))
dev = torch.device("cuda:0")
N, C, (h, w), (H, W) = int(os.environ.get("AB_N", "2")), 19, (65, 129), (512, 1024)
POOL = 128
lo = torch.randn(POOL, N, C, h, w, device=dev) * 5
gr = torch.empty_like(lo)
lay = _lib.state_layout(N, C)
accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
out = torch.empty(lay.out_bytes, dtype=torch.uint8, device=dev)
go = torch.full((), 0.1, device=dev)
auxb = [torch.empty(lib.msq_fused_aux_bytes(N, H, W), dtype=torch.uint8, device=dev) for _ in range(4)]
st = torch.cuda.current_stream().cuda_stream
lop = [lo[i].data_ptr() for i in range(POOL)]
grp = [gr[i].data_ptr() for i in range(POOL)]
axp = [a.data_ptr() for a in auxb]
ap, op, gp = accum.data_ptr(), out.data_ptr(), go.data_ptr()


def one(i):
    rc = lib.msq_fused_fwd_bwd(1, lop[i % POOL], N, C, h, w, H, W, 0.2, 0, ap, op, axp[i % 4], gp, 0.0, grp[i % POOL], None, 0, st)
    assert rc == 0, rc


def sep(i):
    lib.msq_fused_fwd(1, lop[i % POOL], N, C, h, w, H, W, None, 0.2, 0, ap, op, axp[i % 4], grp[i % POOL], st)
    lib.msq_fused_bwd(1, lop[i % POOL], N, C, h, w, H, W, 0, op, axp[i % 4], gp, grp[i % POOL], 1, st)


def timeit(fn, iters, depth=0, warm=50):
    evs = [torch.cuda.Event() for _ in range(8)]
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(iters):
        fn(i)
        if depth and i % depth == 0:
            k = (i // depth) % 8
            evs[k].record()
            evs[(k - 1) % 8].synchronize()          # at most 2 * depth steps ahead
    th = time.perf_counter() - t0
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3, th / iters * 1e6


# the one-call step must give what the two separate calls give
sep(5); torch.cuda.synchronize(); g1 = gr[5].clone(); o1 = out.clone()
one(5); torch.cuda.synchronize()
print("late_finalize", os.environ.get("AB_LATE", "default"), "one-call == separate calls: out", bool(torch.equal(o1, out)),
      "grad max rel diff", ((g1 - gr[5]).abs().max() / g1.abs().max()).item(), flush=True)
QUICK = os.environ.get("AB_QUICK") == "1"
for name, fn in (("one-call", one), ("separate", sep)):
    for iters in ((4000,) if QUICK else (100, 300, 1000, 4000)):
        t, th = timeit(fn, iters)
        print(f"{name} iters {iters:5d} unthrottled: {t:6.2f} us/step (host issue {th:5.1f} us/step)", flush=True)
    for depth in (() if QUICK else (4, 16, 64)):
        t, th = timeit(fn, 4000, depth)
        print(f"{name} iters  4000 host <= {2 * depth:3d} steps ahead: {t:6.2f} us/step (host loop {th:5.1f} us/step)", flush=True)
