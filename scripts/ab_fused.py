"""Time the fused kernels of whichever library MSQ_B200_LIB points at (development aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import _lib, synth
import maxsquareloss_b200 as msq
from oracle import loss_port

lib = _lib.load()
if os.environ.get("AB_PDL"):
    _lib.tune("pdl_mask", int(os.environ["AB_PDL"]))
if os.environ.get("AB_ROWS"):
    _lib.tune("fused_rows", int(os.environ["AB_ROWS"]))
dev = torch.device("cuda:0")
N, C, (h, w), (H, W) = int(os.environ.get("AB_N", "2")), 19, (65, 129), (512, 1024)
POOL = 128
lo = torch.randn(POOL, N, C, h, w, device=dev) * 5
gr = torch.empty_like(lo)
lay = _lib.state_layout(N, C)
accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev)
out = torch.empty(lay.out_bytes, dtype=torch.uint8, device=dev)
go = torch.full((), 0.1, device=dev)
CACHE = os.environ.get("AB_CACHE", "1") == "1"
auxb = [torch.empty(lib.msq_fused_aux_bytes(N, H, W), dtype=torch.uint8, device=dev) for _ in range(4)]
st = torch.cuda.current_stream().cuda_stream

def timeit(fn, iters=300, warm=30):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3

# correctness spot check
x = synth.head_logits(1, C, (h, w), 1, 5.0)
rl, rg, rh = loss_port.chain_iw_maxsquare(x, (H, W), C, 0.2, 0.1)
xc = x.to(dev).requires_grad_(True)
crit = msq.IW_MaxSquareloss(-1, C, 0.2)
l = crit(xc, out_size=(H, W)); (0.1 * l).backward()
ok = bool((crit.last_hist.cpu().long() == rh).all()) and abs(l.item() - rl.item()) < 1e-5 * abs(rl.item()) and \
    (xc.grad.cpu() - rg).abs().max().item() < 1e-4 * rg.abs().max().item()
res = {}
for mode, name in ((1, "iw"), (0, "ms")):
    f = lambda i: lib.msq_fused_fwd(mode, lo[i % POOL].data_ptr(), N, C, h, w, H, W, None, 0.2, 0, accum.data_ptr(), out.data_ptr(),
                                    auxb[i % 4].data_ptr() if CACHE else None, gr[i % POOL].data_ptr() if CACHE else None, st)
    b = lambda i: lib.msq_fused_bwd(mode, lo[i % POOL].data_ptr(), N, C, h, w, H, W, 0, out.data_ptr(),
                                    auxb[i % 4].data_ptr() if CACHE else None, go.data_ptr(), gr[i % POOL].data_ptr(), 1 if CACHE else 0, st)
    def fb(i): f(i); b(i)
    def one(i):
        rc = lib.msq_fused_fwd_bwd(mode, lo[i % POOL].data_ptr(), N, C, h, w, H, W, 0.2, 0, accum.data_ptr(), out.data_ptr(),
                                   auxb[i % 4].data_ptr(), go.data_ptr(), 0.0, gr[i % POOL].data_ptr(), None, 0, st)
        assert rc == 0, rc
    res[name] = (timeit(f), timeit(b), timeit(fb), timeit(one))
    # the one-call step (weights handed over by flag) must give what the two separate calls give
    fb(7); torch.cuda.synchronize(); ga = gr[7].clone(); la = out.clone()
    one(7); torch.cuda.synchronize()
    ok = ok and bool(torch.equal(la[:lay.out_bytes - 16], out[:lay.out_bytes - 16])) and \
        (ga - gr[7]).abs().max().item() <= 1e-6 * ga.abs().max().item()
px = N * H * W
print(f"{os.path.basename(_lib.LIB_PATH):22s} pdl={os.environ.get('AB_PDL', '15')} cache={int(CACHE)} rows={os.environ.get('AB_ROWS','auto'):>4s} ok={ok} " +
      "  ".join(f"{k}: fwd {v[0]:.1f} bwd {v[1]:.1f} f+b {v[2]:.1f} one-call {v[3]:.1f} us = {px / v[3] / 1e3:.1f} Gpix/s" for k, v in res.items()), flush=True)
