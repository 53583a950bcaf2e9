"""Flip-ensemble kernel timing (bench.py's shape); run once per MSQ_FLIP_PX setting (the knob is read at first use)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxsquareloss_b200 import _lib, synth
lib = _lib.load()
N, C, H, W = 2, 19, 512, 1024
dev = torch.device("cuda")
fa = [torch.randn(N, C, H, W, device=dev) * 3 for _ in range(2)]
fb = [torch.flip(a, dims=[-1]) + torch.randn_like(a) for a in fa]
gt = [synth.blocky_labels(N, (H, W), C, 300 + i).to(dev) for i in range(2)]
cm = torch.zeros(C * C + 1, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
f = lambda i: lib.msq_confusion_flip_f32(gt[i % 2].data_ptr(), fa[i % 2].data_ptr(), fb[i % 2].data_ptr(), N, C, H, W, cm.data_ptr(), st)
for i in range(20): f(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(200): f(i)
b.record(); torch.cuda.synchronize()
t = a.elapsed_time(b) / 200 * 1e3
byt = (8.0 * C + 8) * N * H * W
print(f"flip px={os.environ.get('MSQ_FLIP_PX', '2 (default)')}: {t:.1f} us  {byt / t / 1e3:.0f} GB/s = {byt / t / 1e3 / 6533.8 * 100:.0f}% of HBM  cm.sum={int(cm[:C*C].sum())}")
