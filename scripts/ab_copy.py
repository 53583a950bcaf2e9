import torch, time
n = 2 * 19 * 65 * 129
h = torch.randn(n).pin_memory(); d = torch.empty(n, device="cuda"); h2 = torch.empty(n).pin_memory()
def t(fn, it=500):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
print("H2D 1.27 MB: %.1f us" % t(lambda: d.copy_(h, non_blocking=True)))
print("D2H 1.27 MB: %.1f us" % t(lambda: h2.copy_(d, non_blocking=True)))
s2 = torch.cuda.Stream()
def both():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d, non_blocking=True)
print("H2D + D2H on two streams: %.1f us per pair" % t(both)); torch.cuda.synchronize()
