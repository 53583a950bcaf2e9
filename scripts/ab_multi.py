"""Time the pieces of the multi-level guidance step and of the source-side step (development aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import _lib, synth

lib = _lib.load()
dev = torch.device("cuda:0")
N, C, (h, w), (H, W) = int(os.environ.get("AB_N", "2")), 19, (65, 129), (512, 1024)
POOL = 32
lo1 = torch.randn(POOL, N, C, h, w, device=dev) * 5
lo2 = lo1 + 0.5 * torch.randn_like(lo1)
g1, g2 = torch.empty_like(lo1), torch.empty_like(lo2)
lay = _lib.state_layout(N, C)
accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
out = torch.empty(lay.out_bytes, dtype=torch.uint8, device=dev)
go = torch.full((), 0.1, device=dev)
nb = lib.msq_fused_aux_bytes(N, H, W)
aux1 = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(4)]
aux2 = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(4)]
st = torch.cuda.current_stream().cuda_stream

def timeit(fn, iters=300, warm=30):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3

mf = lambda i: lib.msq_multi_fwd(1, lo1[i % POOL].data_ptr(), lo2[i % POOL].data_ptr(), N, C, h, w, H, W, 0.2, 0.95, 0, accum.data_ptr(),
                                 out.data_ptr(), aux1[i % 4].data_ptr(), aux2[i % 4].data_ptr(), g1[i % POOL].data_ptr(),
                                 g2[i % POOL].data_ptr(), None, st)
b1 = lambda i: lib.msq_fused_bwd(1, lo1[i % POOL].data_ptr(), N, C, h, w, H, W, 0, out.data_ptr(), aux1[i % 4].data_ptr(), go.data_ptr(),
                                 g1[i % POOL].data_ptr(), 1, st)
b2 = lambda i: lib.msq_guidance_bwd(lo2[i % POOL].data_ptr(), N, C, h, w, H, W, out.data_ptr(), aux2[i % 4].data_ptr(), go.data_ptr(),
                                    g2[i % POOL].data_ptr(), 1, st)
def step(i): mf(i); b1(i); b2(i)
print(f"multi N={N}: fwd {timeit(mf):.1f} bwd1 {timeit(b1):.1f} bwd2 {timeit(b2):.1f} step {timeit(step):.1f} us", flush=True)

hs, ws, Hs, Ws = 91, 161, 720, 1280
los = torch.randn(POOL, N, C, hs, ws, device=dev) * 3
gs = torch.empty_like(los)
ys = [synth.blocky_labels(N, (Hs, Ws), C, 500 + i).to(dev) for i in range(4)]
auxs = [torch.empty(lib.msq_fused_aux_bytes(N, Hs, Ws), dtype=torch.uint8, device=dev) for _ in range(4)]
cm = torch.zeros(C * C + 1, dtype=torch.int64, device=dev)
sf = lambda i: lib.msq_source_ce_fwd(los[i % POOL].data_ptr(), ys[i % 4].data_ptr(), N, C, hs, ws, Hs, Ws, accum.data_ptr(), out.data_ptr(),
                                     auxs[i % 4].data_ptr(), gs[i % POOL].data_ptr(), cm.data_ptr(), st)
sf0 = lambda i: lib.msq_source_ce_fwd(los[i % POOL].data_ptr(), ys[i % 4].data_ptr(), N, C, hs, ws, Hs, Ws, accum.data_ptr(), out.data_ptr(),
                                      auxs[i % 4].data_ptr(), gs[i % POOL].data_ptr(), None, st)
sb = lambda i: lib.msq_guidance_bwd(los[i % POOL].data_ptr(), N, C, hs, ws, Hs, Ws, out.data_ptr(), auxs[i % 4].data_ptr(), go.data_ptr(),
                                    gs[i % POOL].data_ptr(), 1, st)
print(f"source N={N}: fwd(+cm) {timeit(sf):.1f} fwd(no cm) {timeit(sf0):.1f} bwd {timeit(sb):.1f} us  ({N*Hs*Ws/1e6:.2f} Mpx)", flush=True)

# flip-ensemble evaluation kernel
Hf, Wf = 512, 1024
fa = [torch.randn(N, C, Hf, Wf, device=dev) * 3 for _ in range(2)]
fb = [torch.flip(a, dims=[-1]) + torch.randn_like(a) for a in fa]
gtf = [synth.blocky_labels(N, (Hf, Wf), C, 300 + i).to(dev) for i in range(2)]
ff = lambda i: lib.msq_confusion_flip_f32(gtf[i % 2].data_ptr(), fa[i % 2].data_ptr(), fb[i % 2].data_ptr(), N, C, Hf, Wf, cm.data_ptr(), st)
t = timeit(ff, 100, 10)
byt = (8.0 * C + 8) * N * Hf * Wf
print(f"flip N={N}: {t:.1f} us  {byt / t / 1e3:.0f} GB/s = {byt / t / 1e3 / 6533.8 * 100:.0f}% of HBM", flush=True)
