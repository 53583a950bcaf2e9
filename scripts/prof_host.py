"""Host-side cost of the drop-in module path (IW_MaxSquareloss nn.Module + autograd), stage by stage and under cProfile.
    python scripts/prof_host.py"""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maxsquareloss_b200 as msq

dev = torch.device("cuda:0")
N, C, HW_LO, HW = 2, 19, (65, 129), (512, 1024)
crit = msq.IW_MaxSquareloss(-1, C, 0.2)
host_in = [torch.randn(N, C, *HW_LO).mul_(5).pin_memory() for _ in range(8)]
host_grad = torch.empty(N, C, *HW_LO).pin_memory()
host_loss = torch.empty(()).pin_memory()
dev_in = torch.empty(N, C, *HW_LO, device=dev)
cur = torch.cuda.current_stream()
T = {}

def tick(name, t0):
    t1 = time.perf_counter()
    T[name] = T.get(name, 0.0) + (t1 - t0)
    return t1

def step(i, sync):
    t = time.perf_counter()
    dev_in.copy_(host_in[i % 8], non_blocking=True); t = tick("h2d copy_", t)
    x = dev_in.detach().requires_grad_(True); t = tick("detach/requires_grad", t)
    loss = crit(x, out_size=HW); t = tick("crit()", t)
    l2 = 0.1 * loss; t = tick("lambda*loss", t)
    l2.backward(); t = tick("backward()", t)
    host_grad.copy_(x.grad, non_blocking=True); host_loss.copy_(loss.detach(), non_blocking=True); t = tick("d2h copies", t)
    if sync:
        cur.synchronize(); t = tick("sync", t)

for sync in (False, True):
    for i in range(200): step(i, sync)
    torch.cuda.synchronize(); T.clear()
    n = 2000
    t0 = time.perf_counter()
    for i in range(n):
        step(i, sync)
        if not sync and i % 64 == 63: cur.synchronize()
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print(f"sync={sync}: {tot / n * 1e6:.1f} us/step  " + "  ".join(f"{k} {v / n * 1e6:.1f}" for k, v in T.items()), flush=True)

# PyTorch's own fixed cost for the same call pattern with a NATIVE one-kernel loss (x.sum()): what no binding can remove
def floor_step(i):
    x = dev_in.detach().requires_grad_(True)
    (0.1 * x.sum()).backward()
for i in range(200): floor_step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(2000):
    floor_step(i)
    if i % 64 == 63: cur.synchronize()
torch.cuda.synchronize()
print(f"torch floor (x.sum() as the loss: detach + sum + lambda*loss + backward): {(time.perf_counter() - t0) / 2000 * 1e6:.1f} us/step", flush=True)

pr = cProfile.Profile()
pr.enable()
for i in range(2000):
    step(i, False)
    if i % 64 == 63: cur.synchronize()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)

# Eval.add_batch host cost
ev = msq.Eval(16, device=dev)
gt = torch.randint(-1, 16, (1, 512, 1024), device=dev); pr_ = torch.randint(0, 16, (1, 512, 1024), device=dev)
for i in range(100): ev.add_batch(gt, pr_)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(2000): ev.add_batch(gt, pr_)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"Eval.add_batch host {((t1 - t0) / 2000) * 1e6:.1f} us/call, with drain {((time.perf_counter() - t0) / 2000) * 1e6:.1f}")
