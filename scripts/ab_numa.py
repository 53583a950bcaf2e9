"""Where does the host side of the e2e pipeline lose PCIe bandwidth?  NUMA placement of the pinned buffers / the
thread, write-combined pinned memory.  Prints the topology and times 1.27 MB H2D / D2H copies in each setting."""
import ctypes, glob, os, subprocess, sys, time
import torch

def sh(c):
    try: return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e: return f"<{e}>"

print(sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread'"))
print(sh("nvidia-smi topo -m | head -20"))
print("affinity now:", sorted(os.sched_getaffinity(0)))
for nd in sorted(glob.glob("/sys/devices/system/node/node*")):
    print(nd, open(nd + "/cpulist").read().strip())
import pynvml
pynvml.nvmlInit()
hdl = pynvml.nvmlDeviceGetHandleByIndex(0)
ncpu = os.cpu_count()
try:
    words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (ncpu + 63) // 64)
    local = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
except Exception as e:
    local = []; print("nvml affinity failed", e)
print("GPU0-local cpus:", local)

torch.cuda.init(); torch.zeros(1, device="cuda")
rt = ctypes.CDLL("libcudart.so.12")
n = 2 * 19 * 65 * 129
nbytes = 4 * n
d = torch.empty(n, device="cuda"); d2 = torch.empty(n, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
s2 = torch.cuda.Stream()

def host_alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    ctypes.memset(p, 1, nbytes)
    return p

def timeit(fn, it=400):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3

def run(tag, flags_in):
    hin, hout = host_alloc(flags_in), host_alloc(0)
    h2d = lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), hin, ctypes.c_size_t(nbytes), 1, ctypes.c_void_p(stream))
    d2h = lambda: rt.cudaMemcpyAsync(hout, ctypes.c_void_p(d2.data_ptr()), ctypes.c_size_t(nbytes), 2, ctypes.c_void_p(stream))
    def both():
        rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), hin, ctypes.c_size_t(nbytes), 1, ctypes.c_void_p(stream))
        rt.cudaMemcpyAsync(hout, ctypes.c_void_p(d2.data_ptr()), ctypes.c_size_t(nbytes), 2, ctypes.c_void_p(s2.cuda_stream))
    a, b, c = timeit(h2d), timeit(d2h), timeit(both)
    torch.cuda.synchronize()
    print(f"{tag:40s} H2D {a:6.1f} us ({nbytes/a/1e3:5.1f} GB/s)  D2H {b:6.1f} us ({nbytes/b/1e3:5.1f} GB/s)  both {c:6.1f} us/pair")

run("default affinity, WB pinned", 0)
run("default affinity, WC pinned input", 4)
for nd in sorted(glob.glob("/sys/devices/system/node/node*")):
    cl = open(nd + "/cpulist").read().strip()
    cpus = set()
    for part in cl.split(","):
        if "-" in part:
            lo, hi = part.split("-"); cpus |= set(range(int(lo), int(hi) + 1))
        elif part: cpus.add(int(part))
    cpus &= os.sched_getaffinity(0) if False else cpus
    try:
        os.sched_setaffinity(0, cpus)
    except Exception as e:
        print("setaffinity", nd, e); continue
    run(f"bound to {os.path.basename(nd)} ({cl}), WB", 0)
    run(f"bound to {os.path.basename(nd)} ({cl}), WC in", 4)
# large copy for the PCIe peak
big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
t = timeit(lambda: dbig.copy_(big, non_blocking=True), 20); print(f"64 MB H2D {64*1.048576/t*1e3:.1f} GB/s")
t = timeit(lambda: big.copy_(dbig, non_blocking=True), 20); print(f"64 MB D2H {64*1.048576/t*1e3:.1f} GB/s")
