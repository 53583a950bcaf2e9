"""int64 confusion kernel: CTA-count sweep for one 512x1024 image per call (cfg 4) and for the cfg-2 source batch.
Every CTA ends with up to C*C global atomics on the same addresses; fewer CTAs shorten that queue but also the number
of bytes in flight.  Usage: python scripts/ab_conf.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxsquareloss_b200 import _lib, synth
lib = _lib.load()
dev = torch.device("cuda")
st = torch.cuda.current_stream().cuda_stream
for name, C, n, hw in (("cfg4 image", 16, 1, (512, 1024)), ("cfg2 source batch", 19, 2, (720, 1280))):
    pool = 24
    gts = [synth.blocky_labels(n, hw, C, 1000 + i).to(dev) for i in range(pool)]
    prs = [synth.noisy_prediction(g.cpu(), C, 1000 + i).to(dev) for i, g in enumerate(gts)]
    cm = torch.zeros(C * C + 1, dtype=torch.int64, device=dev)
    npx = n * hw[0] * hw[1]
    for grid in (0, 148, 111, 74, 56, 37):
        _lib.tune("conf_grid", grid)
        f = lambda i: lib.msq_confusion_i64(gts[i % pool].data_ptr(), prs[i % pool].data_ptr(), npx, C, cm.data_ptr(), cm.data_ptr() + 8 * C * C, st)
        for i in range(30): f(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(480): f(i)
        b.record(); torch.cuda.synchronize()
        t = a.elapsed_time(b) / 480 * 1e3
        print(f"{name}: grid {grid or 'auto':>4}: {t:6.2f} us  {16.0 * npx / t / 1e3:6.0f} GB/s", flush=True)
_lib.tune("conf_grid", 0)
