"""A few small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck).
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maxsquareloss_b200 as msq
from maxsquareloss_b200 import synth

dev = "cuda"
for C, hw, HW, N in ((19, (9, 17), (64, 128), 2), (13, (5, 6), (31, 45), 1), (7, (3, 5), (7, 300), 2)):
    lo = synth.head_logits(N, C, hw, 1, 3.0).to(dev)
    lo2 = synth.second_head(lo.cpu(), 1).to(dev)
    y = synth.blocky_labels(N, HW, C, 2, grid=(4, 8)).to(dev)
    for crit in (msq.IW_MaxSquareloss(-1, C, 0.2), msq.MaxSquareloss(-1, C), msq.IWsoftCrossEntropy(-1, C, 0.2), msq.softCrossEntropy(-1)):
        x = lo.clone().requires_grad_(True)
        (0.1 * crit(x, out_size=HW)).backward()
    x = lo.clone().requires_grad_(True)
    crit = msq.IW_MaxSquareloss(-1, C, 0.2)
    crit(x, label=y, out_size=HW).backward()
    x1, x2 = lo.clone().requires_grad_(True), lo2.clone().requires_grad_(True)
    a, b = msq.MultiLevelTargetLoss(msq.IW_MaxSquareloss(-1, C, 0.2), threshold=0.9, return_label=True)((x1, x2), HW)
    (a + b).backward()
    ev = msq.Eval(C)
    x = lo.clone().requires_grad_(True)
    msq.CrossEntropyLoss2d(evaluator=ev)(x, y).backward()
    pred = torch.nn.functional.interpolate(lo, size=HW, mode="bilinear", align_corners=True)
    prob = torch.softmax(pred, 1).requires_grad_(True)
    msq.IW_MaxSquareloss(-1, C, 0.2)(pred, prob).backward()
    msq.MaxSquareloss(-1, C)(pred, prob).backward()
    ev.add_batch(y, pred.argmax(1))
    ev.add_batch(y, pred)
    ev.add_batch_flip(y, pred, torch.flip(pred, dims=[-1]))
    print(C, HW, "mIoU", ev.Mean_Intersection_over_Union())
torch.cuda.synchronize()
print("done")
