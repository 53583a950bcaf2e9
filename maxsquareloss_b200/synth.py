"""Seeded synthetic inputs of the shapes the reference's trainers feed to the
hot path.  Everything is generated on the CPU with a
fixed ``torch.Generator`` / ``numpy.random.Generator`` so that the CPU oracle
and the CUDA path see identical bits; callers copy to the device themselves.

Head sizes come from ``graphs/models/deeplab_multi.py:113-130`` (stride-8
DeepLabv2: 512x1024 -> 65x129, 720x1280 -> 91x161, 760x1280 -> 96x161,
640x1280 -> 81x161).
"""
import numpy as np
import torch

#: (name) -> (C, (h, w), (H, W)); the BASELINE.json configs
SHAPES = {
    "cityscapes_target": (19, (65, 129), (512, 1024)),     # cfg 1, 2 (target), 3
    "gta5_source": (19, (91, 161), (720, 1280)),           # cfg 2 (source side, eval)
    "synthia_source": (16, (96, 161), (760, 1280)),        # cfg 4
    "multi_readme": (19, (81, 161), (640, 1280)),          # cfg 3 variant (README.md:140)
    "dyadic": (19, (65, 129), (513, 1025)),                # exact-arithmetic suite (scale = 1/8)
    "tiny13": (13, (9, 17), (64, 128)),                    # KAT5
}


def head_logits(n, c, hw, seed, scale=1.0, class_bias=False, quantize=False):
    """Low-resolution head logits (n,c,h,w) fp32: ``randn * scale``.

    class_bias: add ``linspace(3,-3,c)`` per class so the argmax histogram is
                skewed the way street scenes are.
    quantize:   round to multiples of 2**-8 (and clamp to |x| < 64) so that every
                product and sum of a dyadic-geometry bilinear interpolation is
                exact in fp32 (the exact-arithmetic parity suite)."""
    g = torch.Generator().manual_seed(int(seed))
    x = torch.randn(n, c, hw[0], hw[1], generator=g) * scale
    if class_bias:
        x = x + torch.linspace(3, -3, c).view(1, c, 1, 1)
    if quantize:
        x = torch.clamp(torch.round(x * 256) / 256, -63.0, 63.0)
    return x.contiguous()


def random_labels(n, hw, c, seed):
    """(n,H,W) int64 uniformly in [-1, c)."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.randint(-1, c, (n, hw[0], hw[1]), generator=g, dtype=torch.int64)


def blocky_labels(n, hw, c, seed, grid=(16, 32), ignore_frac=0.1):
    """(n,H,W) int64 segmentation-like maps: a random ``grid`` of class ids,
    nearest-upsampled, with about ``ignore_frac`` of the cells set to -1."""
    g = torch.Generator().manual_seed(int(seed))
    cells = torch.randint(0, c, (n, grid[0], grid[1]), generator=g, dtype=torch.int64)
    drop = torch.rand(n, grid[0], grid[1], generator=g) < ignore_frac
    cells = torch.where(drop, torch.full_like(cells, -1), cells)
    yi = (torch.arange(hw[0]) * grid[0]) // hw[0]
    xi = (torch.arange(hw[1]) * grid[1]) // hw[1]
    return cells[:, yi][:, :, xi].contiguous()


def noisy_prediction(gt, c, seed, flip_frac=0.3):
    """Prediction map for ``gt``: equal to it except that ``flip_frac`` of the
    pixels (and every ignored pixel) are re-drawn uniformly in [0, c)."""
    g = torch.Generator().manual_seed(int(seed) + 7919)
    rnd = torch.randint(0, c, gt.shape, generator=g, dtype=torch.int64)
    flip = (torch.rand(gt.shape, generator=g) < flip_frac) | (gt < 0)
    return torch.where(flip, rnd, gt).contiguous()


def eval_pair_np(shape, c, seed):
    """KAT6/KAT7-style numpy pair: gt in [-1,c), pred in [0,c)."""
    rng = np.random.default_rng(seed)
    gt = rng.integers(-1, c, shape)
    pr = rng.integers(0, c, shape)
    return gt, pr


def second_head(head1, seed, noise=0.5):
    """Logits of the other classifier head for the multi-level guidance cases (cfg 3):
    ``head1 + noise * randn`` -- correlated with head 1, as two heads of one network are."""
    g = torch.Generator().manual_seed(int(seed) + 104729)
    return (head1 + noise * torch.randn(head1.shape, generator=g)).contiguous()


def align_logits_to_labels(lo, labels, boost=4.0):
    """Raise, in every low-resolution cell, the logit of the class the label map has at the cell's
    position (nearest sample), so that argmax(upsampled logits) mostly equals the label -- a
    prediction that looks like a trained network's (strong confusion-matrix diagonal)."""
    n, c, h, w = lo.shape
    H, W = labels.shape[-2:]
    ys = torch.linspace(0, H - 1, h).round().long()
    xs = torch.linspace(0, W - 1, w).round().long()
    cell = labels[:, ys][:, :, xs]                                   # (n,h,w)
    onehot = torch.zeros_like(lo)
    onehot.scatter_(1, cell.clamp(min=0).unsqueeze(1), 1.0)
    onehot = onehot * (cell >= 0).unsqueeze(1)
    return (lo + boost * onehot).contiguous()


def source_case(n, c, hw, HW, seed, scale, label_kind):
    """(head logits, label map) of a source-side step case (tests/golden/source_kats.json)."""
    lo = head_logits(n, c, hw, seed, scale)
    if label_kind == "random":
        y = random_labels(n, HW, c, seed)
    elif label_kind == "ignored":
        y = torch.full((n,) + tuple(HW), -1, dtype=torch.int64)
    else:
        y = blocky_labels(n, HW, c, seed, grid=(8, 16))
    if label_kind == "aligned":
        lo = align_logits_to_labels(lo, y)
    return lo, y
