"""Host-side cost of HostPipeline.submit/wait: a tiny problem (GPU time negligible) vs the bench shape."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maxsquareloss_b200 as msq

def run(n, C, hw, HW, depth, steps=3000):
    pipe = msq.HostPipeline("iw", n, C, hw, HW, ratio=0.2, depth=depth)
    hin = [torch.randn(n, C, *hw).pin_memory() for _ in range(16)]
    hg = [torch.empty(n, C, *hw).pin_memory() for _ in range(depth)]
    hl = [torch.empty(()).pin_memory() for _ in range(depth)]
    def loop(k):
        slots = []
        for i in range(k):
            j = i % depth
            if len(slots) >= depth:
                pipe.wait(slots[i - depth])
            slots.append(pipe.submit(hin[i % 16], hl[j], hg[j], None, 0.1))
        pipe.drain()
    loop(200)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(steps); t = (time.perf_counter() - t0) / steps
    pipe.close()
    return t * 1e6

for depth in (8, 16):
    print(f"depth {depth}: tiny (1x13x9x17 -> 64x128) {run(1, 13, (9, 17), (64, 128), depth):.1f} us/step   "
          f"bench shape (2x19x65x129 -> 512x1024) {run(2, 19, (65, 129), (512, 1024), depth):.1f} us/step", flush=True)
