import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import maxsquareloss_b200 as msq
from maxsquareloss_b200 import _lib, _torch_ops
which = sys.argv[1]
lib = _lib.load(); ops = _torch_ops.load()
lo = torch.randn(1, 19, 9, 17, device="cuda").requires_grad_(True)
if which == "c":
    lay = _lib.state_layout(1, 19)
    accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device="cuda"); out = torch.empty(lay.out_bytes, dtype=torch.uint8, device="cuda")
    aux = torch.empty(256, dtype=torch.uint8, device="cuda"); g = torch.empty_like(lo)
    print("rc", lib.msq_fused_fwd(1, lo.data_ptr(), 1, 19, 9, 17, 4, 4, None, 0.2, 0, accum.data_ptr(), out.data_ptr(), aux.data_ptr(), g.data_ptr(), torch.cuda.current_stream().cuda_stream), flush=True)
elif which == "op_nograd":
    try:
        with torch.no_grad(): ops.fused_loss(lo, None, 4, 4, 1, 0.2, 0, 0, True)
    except RuntimeError as e: print("raised:", str(e)[:60], flush=True)
elif which == "op":
    try: ops.fused_loss(lo, None, 4, 4, 1, 0.2, 0, 0, True)
    except RuntimeError as e: print("raised:", str(e)[:60], flush=True)
elif which == "op_detached":
    try: ops.fused_loss(lo.detach(), None, 4, 4, 1, 0.2, 0, 0, True)
    except RuntimeError as e: print("raised:", str(e)[:60], flush=True)
elif which == "prob_badmode":
    p = torch.softmax(torch.randn(1, 19, 8, 8, device="cuda"), 1).requires_grad_(True)
    try: ops.prob_loss(p, None, 7, 0.2, -1, 0)
    except RuntimeError as e: print("raised:", str(e)[:60], flush=True)
print("done", which, flush=True)
