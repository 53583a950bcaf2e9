"""NumPy restatement of the bilinear upsample the reference's model applies
before the loss (``graphs/models/deeplab_multi.py:124,128``:
``F.interpolate(x, size=input_size, mode='bilinear', align_corners=True)``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The arithmetic lives in PyTorch (un-vendored).  It is restated here from
``ATen/native/UpSample.h`` (``area_pixel_compute_scale``,
``area_pixel_compute_source_index``, ``guard_index_and_lambda``) and was pinned
empirically in the build container: with

    t_r  = fma(A[r, x0], lx0, A[r, x1] * lx1)          (row r = y0 or y1)
    out  = fma(t_y0, ly0, t_y1 * ly1)

this file reproduces ``F.interpolate`` on CPU **bit for bit** (torch 2.11.0;
``tests/test_oracle_golden.py::test_bilinear_bit_exact_vs_torch``), and the
sm_100 SASS of ``upsample_bilinear2d_out_frame<float,float>`` in
libtorch_cuda.so shows the same FMUL/FFMA pattern, so CPU and CUDA eager agree.
FMA is emulated exactly: the product of two float32 is exact in float64, and
the single float64 add is rounded once more to float32 (double rounding can
only matter on exact half-ulp ties of the float64 sum, which the bit-exactness
test above would expose).
"""
import numpy as np

f32 = np.float32


def source_index(in_size: int, out_size: int):
    """Per output index: (i0, i1, lambda0, lambda1) in the reference's fp32
    arithmetic for align_corners=True."""
    if out_size > 1:
        scale = f32(in_size - 1) / f32(out_size - 1)
    else:
        scale = f32(0)
    dst = np.arange(out_size, dtype=np.float32)
    src = (scale * dst).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    lam1 = np.clip((src - i0.astype(np.float32)).astype(np.float32), f32(0), f32(1))
    lam0 = (f32(1) - lam1).astype(np.float32)
    i1 = i0 + (i0 < in_size - 1)
    return i0, i1, lam0, lam1


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def upsample(lo: np.ndarray, out_hw) -> np.ndarray:
    """(N,C,h,w) float32 -> (N,C,H,W) float32, bit-exact with ATen."""
    lo = np.ascontiguousarray(lo, dtype=np.float32)
    _, _, h, w = lo.shape
    H, W = out_hw
    y0, y1, ly0, ly1 = source_index(h, H)
    x0, x1, lx0, lx1 = source_index(w, W)
    LX0 = lx0[None, None, None, :]
    LX1 = lx1[None, None, None, :]
    # horizontal pass on the low-res rows (h rows only), then vertical
    a = lo[:, :, :, x0]
    b = lo[:, :, :, x1]
    t = _fma(a, np.broadcast_to(LX0, a.shape), (b * LX1).astype(np.float32))  # (N,C,h,W)
    t0 = t[:, :, y0, :]
    t1 = t[:, :, y1, :]
    LY0 = np.broadcast_to(ly0[None, None, :, None], t0.shape)
    LY1 = ly1[None, None, :, None]
    return _fma(t0, LY0, (t1 * LY1).astype(np.float32))


def interp_matrix(in_size: int, out_size: int) -> np.ndarray:
    """Dense (out,in) float64 matrix of the 1-D interpolation weights."""
    i0, i1, l0, l1 = source_index(in_size, out_size)
    R = np.zeros((out_size, in_size), dtype=np.float64)
    rows = np.arange(out_size)
    np.add.at(R, (rows, i0), l0.astype(np.float64))
    np.add.at(R, (rows, i1), l1.astype(np.float64))
    return R


def upsample_adjoint(g: np.ndarray, in_hw) -> np.ndarray:
    """Adjoint of ``upsample`` in float64: (N,C,H,W) -> (N,C,h,w).  This is what
    autograd's ``upsample_bilinear2d_backward`` computes (up to summation
    order): out = Ry^T . g . Rx."""
    g = np.asarray(g, dtype=np.float64)
    _, _, H, W = g.shape
    h, w = in_hw
    Ry = interp_matrix(h, H)
    Rx = interp_matrix(w, W)
    return np.einsum('yh,ncyx,xw->nchw', Ry, g, Rx, optimize=True)
