"""CPU-side checks of the C-ABI library: it loads, exports every symbol that
include/t5b200.h declares, its pure host functions behave, and without a GPU it
fails with an error code (never a silent fallback)."""
import ctypes
import os
import re

import pytest

import tensor_b200 as tb
from tensor_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "t5b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(t5_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 35
    L = ctypes.CDLL(_lib.SO_PATH)
    for n in names:
        assert hasattr(L, n), f"libt5b200.so does not export {n}"
    # and the binding table covers the header exactly
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_strerror():
    L = _lib.lib()
    assert b"sm_100a" in L.t5_version()
    assert L.t5_strerror(0) == b"ok"
    assert b"WrongListLength" in L.t5_strerror(_lib.ERR_SHAPE)
    assert b"no CPU fallback" in L.t5_strerror(_lib.ERR_UNSUPPORTED)


@pytest.mark.parametrize("batch", [0, 1, 255, 256, 257, 1000, 1 << 20, (1 << 24) + 3])
@pytest.mark.parametrize("ndev", [1, 2, 3, 4, 8])
def test_partition_covers_batch_contiguously(batch, ndev):
    nxt = 0
    for g in range(ndev):
        first, count = tb.partition(batch, ndev, g)
        assert first == min(nxt, batch)
        nxt = first + count
        if count and g < ndev - 1 and first + count < batch:
            assert count % 256 == 0          # shard boundaries fall on tile multiples
    assert nxt == batch
    sizes = [tb.partition(batch, ndev, g)[1] for g in range(ndev)]
    assert max(sizes) - min(s for s in sizes) <= max(sizes)  # contiguous, front-loaded
    assert max(sizes) <= (batch + ndev - 1) // ndev + 255


def test_partition_rejects_bad_arguments():
    with pytest.raises(ValueError):
        tb.partition(10, 0, 0)
    with pytest.raises(ValueError):
        tb.partition(10, 2, 2)


def test_null_handles_are_errors_not_crashes():
    L = _lib.lib()
    assert L.t5_shutdown(None) == _lib.ERR_INVALID
    assert L.t5_sync(None) == _lib.ERR_INVALID
    assert L.t5_free(None) == _lib.ERR_INVALID
    assert L.t5_matmul(None, 0, 3, 3, 3, None, None, None) == _lib.ERR_INVALID
    assert L.t5_ndev(None) == 0


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tb.DeviceError):
        tb.Context([0])


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensor_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    s = f.read()
                assert "import oracle" not in s and "from oracle" not in s and "t5oracle" not in s, fn


def test_haskell_ffi_module_binds_every_declared_symbol():
    """The reference-side binding (tensor_b200/hs, uncompilable here: no GHC) must at least name every
    entry point of the header, so that the boundary a maintainer would add is complete."""
    with open(os.path.join(ROOT, "tensor_b200", "hs", "Data", "Tensor", "Device", "FFI.hs")) as f:
        hs = f.read()
    bound = set(re.findall(r'foreign import ccall safe "&?(t5_[a-z0-9_]+)"', hs))
    missing = [n for n in declared_symbols() if n not in bound]
    assert not missing, f"FFI.hs has no foreign import for {missing}"
    # and every C entry is used by some Haskell module above the raw bindings or is a measurement helper
    used = ""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tensor_b200", "hs")):
        for fn in files:
            if fn.endswith(".hs") and fn != "FFI.hs":
                with open(os.path.join(dirpath, fn)) as f:
                    used += f.read()
    helpers = {"t5_timer_begin", "t5_timer_end", "t5_timer_per_device", "t5_flush_l2", "t5_fill_synthetic", "t5_plant_specials",
               "t5_launch_count", "t5_version", "t5_strerror", "t5_sync", "t5_free", "t5_partition", "t5_relayout",
               "t5_pinned_alloc", "t5_pinned_free", "t5_buf_batch", "t5_buf_elems", "t5_buf_dtype", "t5_buf_layout",
               "t5_dot_sum_used_nccl", "t5_structural_table", "t5_matmul_is_exact"}
    for n in declared_symbols():
        if n in helpers:
            continue
        assert "c_" + n[3:] in used, f"no Haskell caller for {n}"


def test_matmul_is_exact_reports_the_tensor_core_shape_rules():
    """Pure shape logic behind include/t5b200.h's t5_matmul contract (no GPU): which products are the bit-exact fold."""
    import tensor_b200 as tb
    ex = tb.matmul_is_exact
    for dt in (tb.F64, tb.F32, tb.I64):
        for mkn in ((3, 3, 3), (4, 4, 4), (8, 8, 8), (16, 16, 16), (4, 4, 1), (1, 64, 1), (20, 20, 20), (65, 17, 63)):
            assert ex(dt, *mkn), (dt, mkn)
    for mkn in ((128, 128, 128), (64, 64, 64), (32, 32, 32), (24, 24, 24), (48, 48, 48), (96, 96, 96), (17, 8, 24), (72, 72, 72), (32, 12, 32), (100, 100, 100), (50, 10, 26)):
        assert not ex(tb.F64, *mkn), mkn
        assert ex(tb.I64, *mkn), mkn
    for mkn in ((16, 32, 32), (32, 32, 16), (32, 13, 32), (32, 32, 25), (32, 32, 20), (17, 32, 136), (8, 64, 64)):
        assert ex(tb.F64, *mkn), mkn
    for mkn in ((128, 128, 128), (64, 64, 64), (96, 96, 96), (48, 32, 64), (32, 32, 32), (17, 40, 32), (48, 48, 48), (32, 32, 24), (28, 28, 28), (36, 32, 36)):
        assert not ex(tb.F32, *mkn), mkn
    for mkn in ((16, 32, 32), (32, 32, 16), (17, 8, 32), (24, 24, 24), (24, 24, 22), (200, 32, 32), (64, 30, 64), (32, 6, 32), (100, 20, 36)):
        assert ex(tb.F32, *mkn), mkn
    import pytest
    with pytest.raises(tb.WrongListLength):
        ex(tb.F64, 0, 3, 3)
    with pytest.raises(tb.Unsupported):
        ex(tb.U8, 3, 3, 3)
