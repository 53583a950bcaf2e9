"""A few launches of every HBM-bound kernel of the path at the BASELINE shapes, on buffers that rotate past L2 -- the
program `ncu --set full` captures for profiles/ (dram__bytes and achieved GB/s per launch).  Prints CUDA-event times when
run without a profiler."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import _lib, synth
import maxsquareloss_b200 as msq

lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
REPS = int(os.environ.get("HBM_REPS", "3"))
N, C, HW = 2, 19, (512, 1024)
hw = HW[0] * HW[1]
rows = []

flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(name, fn, nbytes):
    """every timed launch starts with a cold L2 (a 256 MB write in front of it, outside the timed region)"""
    fn(0); fn(1)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(REPS):
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(i); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    us = tot / REPS * 1e3
    rows.append((name, nbytes, us))
    print(f"{name:58s} {nbytes / 1e6:8.1f} MB  {us:7.1f} us  {nbytes / us / 1e3:7.0f} GB/s", flush=True)

def cold(fn):
    return fn

# strict drop-in
probs = [torch.softmax(torch.randn(N, C, *HW, device=dev) * 3, 1) for _ in range(2)]
grads = [torch.empty_like(p) for p in probs]
lay = _lib.state_layout(N, C)
accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
o = torch.empty(lay.out_bytes, dtype=torch.uint8, device=dev)
go = torch.ones((), device=dev)
for mode, nm in ((_lib.MODE_IW, "IW"), (_lib.MODE_MAXSQUARE, "MaxSquare")):
    run(f"prob_fwd<19,{nm}> + finalize (2x19x512x1024)", lambda i: lib.msq_prob_fwd(mode, probs[i % 2].data_ptr(), N, C, hw, None, 0.2, -1, 0, accum.data_ptr(), o.data_ptr(), st), 4.0 * C * N * hw)
    run(f"prob_bwd<19,{nm}>", lambda i: lib.msq_prob_bwd(mode, probs[i % 2].data_ptr(), N, C, hw, -1, 0, o.data_ptr(), go.data_ptr(), grads[i % 2].data_ptr(), st), 8.0 * C * N * hw)
# strict MinEnt (inputs + arbitrary target)
for mode, nm in ((_lib.MODE_IW, "IW"), (_lib.MODE_MAXSQUARE, "mean")):
    run(f"softce_fwd<19,{nm}> + finalize", lambda i: lib.msq_softce_fwd(mode, grads[i % 2].data_ptr(), probs[i % 2].data_ptr(), N, C, hw, 0.2, -1, 0, accum.data_ptr(), o.data_ptr(), st), 8.0 * C * N * hw)
gz = torch.empty_like(probs[0])
run("softce_bwd<19,IW> (d inputs only)", lambda i: lib.msq_softce_bwd(_lib.MODE_IW, grads[i % 2].data_ptr(), probs[i % 2].data_ptr(), N, C, hw, -1, 0, o.data_ptr(), go.data_ptr(), gz.data_ptr(), None, st), 12.0 * C * N * hw)
del probs, grads, gz
# confusion, cfg-2 source batch and cfg-4 single image
cm = torch.zeros(C * C + 1, dtype=torch.int64, device=dev)
for (n_, hw_, Cv, tag) in ((2, (720, 1280), 19, "cfg2 source batch 2x720x1280"), (1, (512, 1024), 16, "cfg4 one image 512x1024")):
    pool = 10
    gts = [synth.blocky_labels(n_, hw_, Cv, 100 + i).to(dev) for i in range(pool)]
    prs = [synth.noisy_prediction(gts[i].cpu(), Cv, 100 + i).to(dev) for i in range(pool)]
    px = n_ * hw_[0] * hw_[1]
    run(f"confusion_i64 ({tag})", cold(lambda i: lib.msq_confusion_i64(gts[i % pool].data_ptr(), prs[i % pool].data_ptr(), px, Cv, cm.data_ptr(), cm.data_ptr() + 8 * Cv * Cv, st)), 16.0 * px)
    if n_ == 1:
        ev = msq.Eval(Cv, device=dev, defer=pool)
        def deferred(i):
            for k in range(pool): ev.add_batch(gts[k], prs[k])
        run(f"confusion_multi: {pool} deferred add_batch, one launch ({tag})", cold(deferred), 16.0 * px * pool)
    lgs = [torch.randn(n_, Cv, *hw_, device=dev) for _ in range(3)]
    run(f"confusion_logits<{Cv}> ({tag})", lambda i: lib.msq_confusion_logits_f32(gts[i % pool].data_ptr(), lgs[i % 3].data_ptr(), n_, Cv, hw_[0] * hw_[1], cm.data_ptr(), st), (4.0 * Cv + 8) * px)
    del lgs
# flip ensemble
fa = [torch.randn(N, C, *HW, device=dev) * 3 for _ in range(2)]
fb = [torch.flip(a, dims=[-1]) + torch.randn_like(a) for a in fa]
gtf = [synth.blocky_labels(N, HW, C, 300 + i).to(dev) for i in range(2)]
run("confusion_flip2<19> (2x19x512x1024 x2)", lambda i: lib.msq_confusion_flip_f32(gtf[i % 2].data_ptr(), fa[i % 2].data_ptr(), fb[i % 2].data_ptr(), N, C, HW[0], HW[1], cm.data_ptr(), st), (8.0 * C + 8) * N * hw)
