"""The C-ABI library loads and exports every symbol include/msq_b200.h declares; argument
validation that needs no GPU.  CPU only (no kernel is launched)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from maxsquareloss_b200 import _lib, build
    build.build()
    return _lib.load()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "msq_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(msq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from maxsquareloss_b200 import _lib
    decl = declared_symbols()
    assert len(decl) >= 10
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/msq_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == decl, "python binding and header disagree on the symbol list"


def test_abi_version_and_error_strings(lib):
    assert lib.msq_abi_version() == 6
    assert lib.msq_error_string(0) == b"success"
    for code in (-1, -2, -3, -4):
        assert lib.msq_error_string(code).startswith(b"msq:")


def test_state_layout(lib):
    from maxsquareloss_b200 import _lib
    for n, c in [(1, 19), (2, 19), (8, 16), (3, 13), (1, 32)]:
        lay = _lib.state_layout(n, c)
        nc = n * c
        assert lay.accum_bytes % 16 == 0 and lay.out_bytes % 16 == 0
        assert lay.sumsq_off % 8 == 0 and lay.kept_off % 8 == 0 and lay.hist_off % 4 == 0
        assert lay.hist_off >= lay.sumsq_off + 8 * nc
        assert lay.accum_bytes >= lay.nvalid_off + 8 * 16 and lay.ce_off % 8 == 0
        assert lay.loss_off % 4 == 0 and lay.weights_off % 4 == 0 and lay.sum_out_off % 8 == 0
        assert lay.stats_off % 8 == 0 and lay.stats_off >= lay.hist_out_off + 4 * nc
        assert lay.out_bytes >= lay.loss2_off + 4 and lay.nvalid_out_off % 8 == 0
    bad = _lib.StateLayout()
    assert lib.msq_state_layout_get(1, 33, ctypes.byref(bad)) == -1
    assert lib.msq_state_layout_get(0, 19, ctypes.byref(bad)) == -1


def test_argument_validation_without_gpu(lib):
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.msq_confusion_i64(None, None, 16, 19, None, None, None) == -1
    assert lib.msq_confusion_i64(None, None, 16, 33, 8, None, None) == -1
    assert lib.msq_confusion_i64(None, None, 0, 19, 8, None, None) == 0          # empty batch: no-op
    assert lib.msq_confusion_logits_f32(None, None, 1, 19, 16, None, None) == -1
    assert lib.msq_prob_fwd(0, None, 1, 19, 16, None, 0.2, -1, 0, None, None, None) == -1
    assert lib.msq_prob_fwd(7, 16, 1, 19, 16, None, 0.2, -1, 0, 16, 16, None) == -1   # bad mode
    assert lib.msq_prob_bwd(1, None, 1, 19, 16, -1, 0, None, None, None, None) == -1
    assert lib.msq_fused_fwd(1, None, 1, 19, 4, 4, 8, 8, None, 0.2, 0, None, None, None, None, None) == -1
    assert lib.msq_fused_fwd(1, 16, 1, 19, 8, 8, 4, 4, None, 0.2, 0, 16, 16, None, None, None) == -2   # downsampling
    assert lib.msq_fused_bwd(1, None, 1, 19, 4, 4, 8, 8, 0, None, None, None, None, 0, None) == -1
    assert lib.msq_fused_aux_bytes(2, 512, 1024) == 16 * 2 * 512 * 1024
    assert lib.msq_pipe_create(1, 0, 19, 4, 4, 8, 8, 0.2, 2, None) == -1
    assert lib.msq_prob_fwd(0, 18, 1, 19, 16, None, 0.2, -1, 0, 16, 16, None) == -4       # misaligned prob
    assert lib.msq_tune_set(b"no_such_knob", 1) == -1
    assert lib.msq_tune_set(b"conf_agg", 1) == 0
    for knob, default in ((b"late_finalize", 1), (b"pdl_mask", 15)):                    # round-2 knobs: settable, restored
        assert lib.msq_tune_set(knob, 0) == 0 and lib.msq_tune_set(knob, default) == 0
    # the one-call step validates before it launches anything (no GPU here): null logits / accumulators, too many classes
    assert lib.msq_fused_fwd_bwd(1, None, 1, 19, 4, 4, 8, 8, 0.2, 0, 16, 16, None, None, 0.1, 16, None, 0, None) == -1
    assert lib.msq_fused_fwd_bwd(1, 16, 1, 19, 4, 4, 8, 8, 0.2, 0, None, 16, None, None, 0.1, 16, None, 0, None) == -1
    assert lib.msq_fused_fwd_bwd(1, 16, 1, 33, 4, 4, 8, 8, 0.2, 0, 16, 16, None, None, 0.1, 16, None, 0, None) == -1
    assert lib.msq_fused_fwd_bwd(1, 16, 1, 19, 4, 4, 8, 8, 0.2, 0, 16, 16, None, None, 0.1, None, None, 0, None) == -1


def test_host_side_refuses_cpu_tensors():
    import torch
    import maxsquareloss_b200 as msq
    p = torch.softmax(torch.randn(1, 19, 4, 4), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        msq.MaxSquareloss()(p, p)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        msq.IW_MaxSquareloss()(p, p)
    with pytest.raises(RuntimeError, match="out_size"):
        msq.IW_MaxSquareloss()(p)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            msq.Eval(19)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from maxsquareloss_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "maxsquareloss_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_new_rows_have_no_cpu_fallback_either():
    import torch
    import maxsquareloss_b200 as msq
    lo = torch.randn(1, 19, 4, 4)
    y = torch.zeros(1, 8, 8, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        msq.softCrossEntropy()(lo, out_size=(8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        msq.IWsoftCrossEntropy(-1, 19, 0.2)(lo, out_size=(8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        msq.CrossEntropyLoss2d()(lo, y)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        msq.MultiLevelTargetLoss(msq.IW_MaxSquareloss(-1, 19, 0.2))((lo, lo), (8, 8))
    with pytest.raises(TypeError):
        msq.MultiLevelTargetLoss(torch.nn.CrossEntropyLoss())
    with pytest.raises(RuntimeError):
        msq.CrossEntropyLoss2d(weight=torch.ones(19))


def test_comm_entry_points_reject_bad_arguments(lib):
    import ctypes
    assert lib.msq_comm_unique_id(None) == -1
    h = ctypes.c_void_p()
    assert lib.msq_comm_create(None, 2, 0, ctypes.byref(h)) == -1
    assert lib.msq_comm_create(b"\0" * 128, 2, 5, ctypes.byref(h)) == -1
    assert lib.msq_comm_allreduce_f64(None, None, 1, None) == -1
    assert lib.msq_comm_join(None, 0, None) == -1
    assert lib.msq_comm_allreduce_u64(None, None, 1, None) == -1
    assert lib.msq_comm_sum_u64_begin(None, None, 2, None) == -1
    assert lib.msq_comm_sum_u64_end(None, None, 2, None) == -1
    assert lib.msq_error_string(-5).startswith(b"msq:")


def test_peer_memory_mailbox_argument_checks(lib):
    # the mailbox entry points reject null handles before touching CUDA; a NULL communicator has no mailboxes
    buf = ctypes.create_string_buffer(64)
    assert lib.msq_comm_box_export(None, buf) == -1
    assert lib.msq_comm_box_open(None, buf) == -1
    assert lib.msq_comm_box_active(None) == 0
    assert lib.msq_comm_box_enable(None, 1) == -1
    v = ctypes.c_uint(7)
    assert lib.msq_comm_box_errors(None, ctypes.byref(v)) == -1


def test_torch_binding_loads_and_has_no_cpu_fallback(lib):
    """lib/libmsq_torch.so (the C++ autograd nodes over the C ABI) loads next to libmsq_b200.so, registers its two
    operators and refuses CPU tensors; the nn.Modules go through it."""
    import torch
    import maxsquareloss_b200 as msq
    from maxsquareloss_b200 import _torch_ops
    ops = _torch_ops.load()
    assert str(ops.fused_loss) == "msq_b200.fused_loss.default" and str(ops.prob_loss) == "msq_b200.prob_loss.default"
    lo = torch.randn(1, 19, 4, 4, requires_grad=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.fused_loss(lo, None, 8, 8, 1, 0.2, 0, 0, True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        msq.IW_MaxSquareloss(-1, 19, 0.2)(lo, out_size=(8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        msq.MaxSquareloss(-1, 19)(None, torch.softmax(lo, 1))
    with pytest.raises(ValueError):
        msq.MaxSquareloss(-1, 16)(None, torch.softmax(lo, 1))
    with pytest.raises(RuntimeError, match="out_size"):
        msq.IW_MaxSquareloss(-1, 19, 0.2)(lo)
