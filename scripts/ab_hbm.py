"""A/B of the HBM-bound kernels in the bench's own timing loops (back to back, buffers rotate past L2):
prob_tail 0/1 (finalisation launch vs last-CTA finalisation of the strict forward) and the confusion legs."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import maxsquareloss_b200 as msq
from maxsquareloss_b200 import _lib, synth

lib = _lib.load()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
st = torch.cuda.current_stream().cuda_stream
peak, _ = bench.peaks()
for tail in (0, 1, 0, 1):
    _lib.tune("prob_tail", tail)
    rows = bench.secondary_kernels(lib, _lib, synth, dev, st, peak, 100)
    print(f"prob_tail={tail}: " + "  ".join(f"{r['kernel'].split('(')[0].strip()[:28]} {r['ms'] * 1e3:.1f}us {r['frac_of_hbm']:.3f}" for r in rows), flush=True)
_lib.tune("prob_tail", 0)
import torch.distributed as dist
ch = bench.confusion_hist_leg(lib, _lib, synth, msq, dev, st, 0, 1, peak, dist, None)
print({k: (round(v["value"], 1), round(v["frac_of_hbm_aggregate"], 3), round(v["us_per_image_per_rank"], 2)) for k, v in ch.items() if isinstance(v, dict)})
# correctness of the tail variant against the launch variant
for mode in (_lib.MODE_IW, _lib.MODE_MAXSQUARE):
    p = torch.softmax(torch.randn(2, 19, 256, 512, device=dev) * 3, 1)
    lay = _lib.state_layout(2, 19)
    acc = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
    outs = []
    for tail in (0, 1, 1):
        _lib.tune("prob_tail", tail)
        o = torch.zeros(lay.out_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.msq_prob_fwd(mode, p.data_ptr(), 2, 19, 256 * 512, None, 0.2, -1, 0, acc.data_ptr(), o.data_ptr(), st))
        torch.cuda.synchronize()
        outs.append(o)
    print("mode", mode, "tail == launch:", bool(torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])), "accum clean:", int(acc.sum()) == 0)
_lib.tune("prob_tail", 0)
