"""Quick on-GPU sanity + timing sweep (development aid; the judged checks are tests/ and bench.py)."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maxsquareloss_b200 as msq
from maxsquareloss_b200 import synth, _lib
from oracle import loss_port, eval_port, loss_math

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))

def t_ms(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

# ---------------- eval
for C, shape in [(19, (2, 720, 1280)), (16, (1, 512, 1024)), (19, (1, 37, 53))]:
    gt = synth.blocky_labels(shape[0], shape[1:], C, 1)
    pr = synth.noisy_prediction(gt, C, 1)
    ref = eval_port.confusion(gt.numpy(), pr.numpy(), C)
    for agg in (0, 1, 2):
        _lib.tune("conf_agg", agg)
        ev = msq.Eval(C)
        ev.add_batch(gt.to(dev), pr.to(dev))
        got = ev.confusion_matrix.astype(np.int64)
        print("eval C", C, shape, "agg", agg, "exact:", np.array_equal(got, ref))
    lg = torch.randn(shape[0], C, *shape[1:])
    ref2 = eval_port.confusion(gt.numpy(), lg.numpy().argmax(1), C)
    ev = msq.Eval(C); ev.add_batch(gt.to(dev), lg.to(dev))
    print("eval logits exact:", np.array_equal(ev.confusion_matrix.astype(np.int64), ref2))
_lib.tune("conf_agg", 1)

# ---------------- strict loss
for C, hw, HW, N in [(19, (65, 129), (512, 1024), 2), (16, (96, 161), (760, 1280), 1), (13, (9, 17), (64, 128), 1), (5, (6, 7), (31, 45), 3)]:
    lo = synth.head_logits(N, C, hw, 0, 2.0)
    _, prob = loss_port.prologue(lo, HW)
    for kind in ("iw", "ms"):
        p = prob.clone().requires_grad_(True)
        if kind == "iw":
            lref, href, _ = loss_port.iw_maxsquare(p, C, 0.2, return_aux=True)
        else:
            lref = loss_port.maxsquare(p); href = None
        (0.1 * lref).backward()
        pg = prob.to(dev).requires_grad_(True)
        crit = msq.IW_MaxSquareloss(-1, C, 0.2) if kind == "iw" else msq.MaxSquareloss(-1, C)
        l = crit(None, pg)
        (0.1 * l).backward()
        gerr = (pg.grad.cpu() - p.grad).abs().max().item() / p.grad.abs().max().item()
        hist_ok = None if href is None else bool((crit.last_hist.cpu().long() == href).all())
        print(f"strict {kind} C{C} N{N} {HW}: loss {l.item():.9g} ref {lref.item():.9g} rel {abs(l.item()-lref.item())/abs(lref.item()):.2e} grad relmax {gerr:.2e} hist {hist_ok}")

# ---------------- fused loss
for C, hw, HW, N, scale in [(19, (65, 129), (512, 1024), 2, 1.0), (19, (65, 129), (512, 1024), 1, 5.0), (16, (96, 161), (760, 1280), 1, 1.0),
                            (13, (9, 17), (64, 128), 1, 1.0), (5, (6, 7), (31, 45), 3, 2.0), (19, (65, 129), (513, 1025), 1, 5.0)]:
    lo = synth.head_logits(N, C, hw, 0, scale, quantize=(HW == (513, 1025)))
    for kind in ("iw", "ms"):
        if kind == "iw":
            lref, gref, href = loss_port.chain_iw_maxsquare(lo, HW, C, 0.2, 0.1)
        else:
            lref, gref = loss_port.chain_maxsquare(lo, HW, 0.1); href = None
        x = lo.to(dev).requires_grad_(True)
        crit = msq.IW_MaxSquareloss(-1, C, 0.2) if kind == "iw" else msq.MaxSquareloss(-1, C)
        l = crit(x, out_size=HW)
        (0.1 * l).backward()
        gerr = (x.grad.cpu() - gref).abs().max().item() / gref.abs().max().item()
        hist_ok = None if href is None else bool((crit.last_hist.cpu().long() == href).all())
        hd = None if href is None else int((crit.last_hist.cpu().long() - href).abs().sum())
        print(f"fused {kind} C{C} N{N} {hw}->{HW} s{scale}: loss {l.item():.9g} ref {lref.item():.9g} rel {abs(l.item()-lref.item())/abs(lref.item()):.2e} grad relmax {gerr:.2e} hist {hist_ok} |dh|={hd}")

# ---------------- timings
print("--- timings (ms) ---")
C, hw, HW, N = 19, (65, 129), (512, 1024), 2
npx = N * HW[0] * HW[1]
lo = synth.head_logits(N, C, hw, 0, 5.0).to(dev)
prob = torch.softmax(torch.nn.functional.interpolate(lo, size=HW, mode="bilinear", align_corners=True), 1).contiguous()
go = torch.ones((), device=dev)
for kind in ("iw", "ms"):
    crit = msq.IW_MaxSquareloss(-1, C, 0.2) if kind == "iw" else msq.MaxSquareloss(-1, C)
    for R in (0,):
        _lib.tune("fused_rows", R)
        x = lo.clone().requires_grad_(True)
        def fwd(): return crit(x, out_size=HW)
        def fb():
            x.grad = None
            crit(x, out_size=HW).backward()
        tf = t_ms(fwd); tfb = t_ms(fb)
        print(f"fused {kind} R={R}: fwd {tf*1e3:.1f} us  fwd+bwd {tfb*1e3:.1f} us  -> {npx/tfb/1e6:.2f} Gpix/s")
    _lib.tune("fused_rows", 0)
    probs = [prob, prob.clone()]
    gradb = [torch.empty_like(prob), torch.empty_like(prob)]
    lay = _lib.state_layout(N, C)
    accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device=dev); outb = torch.empty(lay.out_bytes, dtype=torch.uint8, device=dev)
    lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
    mode = 1 if kind == "iw" else 0
    kk = [0]
    def sf():
        j = kk[0] % 2; kk[0] += 1
        lib.msq_prob_fwd(mode, probs[j].data_ptr(), N, C, HW[0] * HW[1], None, 0.2, -1, 0, accum.data_ptr(), outb.data_ptr(), st)
    def sb():
        j = kk[0] % 2; kk[0] += 1
        lib.msq_prob_bwd(mode, probs[j].data_ptr(), N, C, HW[0] * HW[1], -1, 0, outb.data_ptr(), go.data_ptr(), gradb[j].data_ptr(), st)
    tf = t_ms(sf, iters=50); tb = t_ms(sb, iters=50)
    bytes_f = 4 * C * npx; bytes_b = 8 * C * npx
    print(f"strict {kind}: fwd {tf*1e3:.1f} us ({bytes_f/tf/1e6:.0f} GB/s)  bwd {tb*1e3:.1f} us ({bytes_b/tb/1e6:.0f} GB/s) -> fwd+bwd {npx/(tf+tb)/1e6:.2f} Gpix/s")

# eval timings
for C, n, HWs in [(19, 2, (720, 1280)), (19, 16, (720, 1280)), (16, 32, (512, 1024))]:
    pool = max(1, int(200e6 // (16 * n * HWs[0] * HWs[1])))
    gts = [synth.blocky_labels(n, HWs, C, 3 + i).to(dev) for i in range(pool)]
    prs = [synth.noisy_prediction(g.cpu(), C, 3).to(dev) for g in gts]
    gtr = synth.random_labels(n, HWs, C, 4).to(dev); prr = torch.randint(0, C, gtr.shape, device=dev)
    ev = msq.Eval(C)
    cmp = ev._dev.data_ptr()
    lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
    for ctas in (1, 2):
        _lib.tune("conf_ctas", ctas)
        for agg in (0, 1):
            _lib.tune("conf_agg", agg)
            for name, G, P in (("blocky", gts, prs), ("uniform", [gtr], [prr])):
                k = [0]
                def f():
                    j = k[0] % len(G); k[0] += 1
                    lib.msq_confusion_i64(G[j].data_ptr(), P[j].data_ptr(), G[j].numel(), C, cmp, cmp + 8 * C * C, st)
                t = t_ms(f, iters=50, warm=5)
                print(f"conf_i64 C{C} n{n} {name} ctas{ctas} agg{agg}: {t*1e3:.1f} us  {G[0].numel()/t/1e6:.1f} Gpix/s  {16*G[0].numel()/t/1e6:.0f} GB/s")
    _lib.tune("conf_agg", 0); _lib.tune("conf_ctas", 1)
    lg = torch.randn(min(n, 4), C, *HWs, device=dev)
    g4 = gts[0][:min(n, 4)].contiguous()
    t = t_ms(lambda: lib.msq_confusion_logits_f32(g4.data_ptr(), lg.data_ptr(), g4.shape[0], C, HWs[0] * HWs[1], cmp, st), iters=20, warm=3)
    print(f"conf_logits C{C} n{g4.shape[0]}: {t*1e3:.1f} us  {g4.numel()/t/1e6:.1f} Gpix/s  {(4*C+8)*g4.numel()/t/1e6:.0f} GB/s")
print("done")
