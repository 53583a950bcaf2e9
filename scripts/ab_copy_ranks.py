"""Copy-only ceiling of the box for the e2e leg: N ranks (torchrun) each move the step's host traffic -- 1.27 MB of head
logits H2D and 1.27 MB of dL/dlogits D2H per step, pinned memory, both directions on their own streams -- and NOTHING
else, concurrently.  What this prints is the rate no e2e number of bench.py can exceed on this host at N GPUs.
    python -m torch.distributed.run --nproc-per-node N scripts/ab_copy_ranks.py"""
import os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
n = 2 * 19 * 65 * 129
nbytes = ctypes.c_size_t(4 * n)
hin = [torch.randn(n).pin_memory() for _ in range(8)]
hout = [torch.empty(n).pin_memory() for _ in range(8)]
din, dout = torch.empty(n, device=dev), torch.randn(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
p_in, p_out = [ctypes.c_void_p(t.data_ptr()) for t in hin], [ctypes.c_void_p(t.data_ptr()) for t in hout]
d_in, d_out, st1, st2 = ctypes.c_void_p(din.data_ptr()), ctypes.c_void_p(dout.data_ptr()), ctypes.c_void_p(s1.cuda_stream), ctypes.c_void_p(s2.cuda_stream)

def pairs(k):          # raw cudaMemcpyAsync (~2 us of host time per call): the loop is copy-engine bound, not Python bound
    for i in range(k):
        rt.cudaMemcpyAsync(d_in, p_in[i % 8], nbytes, 1, st1)
        rt.cudaMemcpyAsync(p_out[i % 8], d_out, nbytes, 2, st2)
        if i % 256 == 255:
            s1.synchronize()       # bound the queue depth

pairs(200)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
K = 4000
t0 = time.perf_counter()
pairs(K)
torch.cuda.synchronize()
us = (time.perf_counter() - t0) / K * 1e6
t = torch.tensor([us], device=dev, dtype=torch.float64)
allus = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allus, t)
else:
    allus = [t]
if rank == 0:
    v = [float(x.item()) for x in allus]
    px = 2 * 512 * 1024
    print(f"copy-only, {world} rank(s): us per (H2D 1.27 MB + D2H 1.27 MB) pair by rank: {[round(x, 1) for x in v]}  "
          f"max {max(v):.1f} us  -> ceiling {world * px / max(v) / 1e3:.1f} Gpix/s aggregate, "
          f"{2 * 4 * n / max(v) / 1e3:.1f} GB/s per GPU both directions; affinity {sorted(os.sched_getaffinity(0))[:4]}..{len(os.sched_getaffinity(0))} cpus", flush=True)
if world > 1:
    dist.destroy_process_group()
