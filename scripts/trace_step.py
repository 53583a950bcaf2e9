"""Timeline of one fused step from %globaltimer stamps (development aid; needs a library built with -DMSQ_TRACE=1:
``python scripts/ab_variants.py build`` makes lib/variants/libmsq_trace.so, then on the GPU box
``MSQ_B200_LIB=.../libmsq_trace.so python scripts/trace_step.py``).

Per kernel of the step (forward, finalisation, backward) it prints, relative to the first stamp of the step, the
min / median / max over the CTAs of: entry, before/after griddepcontrol.wait, set-up done, row loop done, exit."""
import ctypes, os, sys
import numpy as np
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
mode = int(os.environ.get("AB_MODE", "1"))


lop = [lo[i].data_ptr() for i in range(POOL)]
grp = [gr[i].data_ptr() for i in range(POOL)]
axp = [a.data_ptr() for a in auxb]
ap, op, gp = accum.data_ptr(), out.data_ptr(), go.data_ptr()
FAST = os.environ.get("AB_FASTHOST", "1") == "1"        # precomputed pointers: the host runs far ahead of the GPU (deep launch queue)
ONECALL = os.environ.get("AB_ONECALL", "1") == "1"      # msq_fused_fwd_bwd (weights handed over by flag) vs separate calls


def step(i):
    if FAST and ONECALL:
        rc = lib.msq_fused_fwd_bwd(mode, lop[i % POOL], N, C, h, w, H, W, 0.2, 0, ap, op, axp[i % 4], gp, 0.0, grp[i % POOL], None, 0, st)
        assert rc == 0, rc
        return
    if FAST:
        lib.msq_fused_fwd(mode, lop[i % POOL], N, C, h, w, H, W, None, 0.2, 0, ap, op, axp[i % 4], grp[i % POOL], st)
        lib.msq_fused_bwd(mode, lop[i % POOL], N, C, h, w, H, W, 0, op, axp[i % 4], gp, grp[i % POOL], 1, st)
        return
    if ONECALL:
        rc = lib.msq_fused_fwd_bwd(mode, lo[i % POOL].data_ptr(), N, C, h, w, H, W, 0.2, 0, accum.data_ptr(), out.data_ptr(),
                                   auxb[i % 4].data_ptr(), go.data_ptr(), 0.0, gr[i % POOL].data_ptr(), None, 0, st)
        assert rc == 0, rc
        return
    lib.msq_fused_fwd(mode, lo[i % POOL].data_ptr(), N, C, h, w, H, W, None, 0.2, 0, accum.data_ptr(), out.data_ptr(),
                      auxb[i % 4].data_ptr(), gr[i % POOL].data_ptr(), st)
    lib.msq_fused_bwd(mode, lo[i % POOL].data_ptr(), N, C, h, w, H, W, 0, out.data_ptr(), auxb[i % 4].data_ptr(), go.data_ptr(),
                      gr[i % POOL].data_ptr(), 1, st)


K, P = 1024, 8
for fn in ("msq_debug_trace_fused", "msq_debug_trace_finalize"):
    getattr(lib, fn).argtypes = [ctypes.c_void_p, ctypes.c_longlong]
    getattr(lib, fn).restype = ctypes.c_int
for i in range(int(os.environ.get('AB_STEPS', '200'))):
    step(i)
torch.cuda.synchronize()
fused = np.zeros(4 * K * P, dtype=np.uint64)
fin = np.zeros(4 * K * P, dtype=np.uint64)
assert lib.msq_debug_trace_fused(fused.ctypes.data, fused.size) == 0
assert lib.msq_debug_trace_finalize(fin.ctypes.data, fin.size) == 0
fused = fused.reshape(2, 2, K, P).astype(np.int64)          # [kernel][step parity][CTA][point]
fin = fin.reshape(4, K, P).astype(np.int64)
G = int((fused[0, 0, :, 0] > 0).sum())
# the two most recent steps: `old` is the parity whose forward started first
old = 0 if fused[0, 0, :G, 0].min() < fused[0, 1, :G, 0].min() else 1
t0 = fused[0, old, :G, 0].min()
names = ["entry", "before wait", "after wait", "set-up done", "rows done", "exit"]
print(f"batch {N}, grid {G}, one call {ONECALL}, fast host {FAST}; the last two steps, us relative to the first forward CTA's entry of the older one "
      f"(globaltimer); min / median / max over CTAs")
for which, par in (("older step", old), ("newer step", 1 - old)):
    for kid, title in ((0, "forward"), (1, "backward")):
        print(which, title)
        Gk = int((fused[kid, par, :, 0] > 0).sum())
        for p, nm in enumerate(names):
            v = (fused[kid, par, :Gk, p] - t0) / 1e3
            print(f"  {nm:12s} {v.min():7.2f} {np.median(v):7.2f} {v.max():7.2f}")
v = (fin[0, 0, :7] - t0) / 1e3
print(f"finalisation (newer step): entry {v[0]:.2f}  after wait {v[1]:.2f}  replicas loaded {v[3]:.2f}  weights stored {v[4]:.2f}  "
      f"after barrier {v[5]:.2f}  loss stored {v[6]:.2f}  exit {v[2]:.2f}  (sm {int(fin[0, 0, 7])})")
print("step period from the forward's wait release:", round((fused[0, 1 - old, :G, 2].min() - fused[0, old, :G, 2].min()) / 1e3, 2), "us")
if os.environ.get("AB_DUMP"):
    np.savez(os.environ["AB_DUMP"], fused=fused, fin=fin, G=G, t0=t0)
