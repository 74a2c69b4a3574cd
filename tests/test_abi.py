"""CPU: the C-ABI library loads, exports every symbol include/mergenet_b200.h declares, and refuses
to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mergenet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\(", src)
    return sorted({n for n in names if n.startswith("mn_") or n == "c_run_segmentation"})


def test_library_exports_every_declared_symbol(lib_mod):
    L = lib_mod.lib()
    declared = _declared_symbols()
    assert "c_run_segmentation" in declared and len(declared) >= 12
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(lib_mod.EXPORTS) == declared


def test_status_strings(lib_mod):
    L = lib_mod.lib()
    assert L.mn_status_string(0) == b"ok"
    assert b"CUDA" in L.mn_status_string(7)


def test_workspace_size_is_sane(lib_mod):
    L = lib_mod.lib()
    b = L.mn_workspace_bytes_per_image(1024, 2048, 9, 10)
    assert 0.5e9 < b < 2.5e9
    assert L.mn_workspace_bytes_per_image(0, 5, 9, 10) == 0


def test_wrapper_argument_errors_match_cython(lib_mod):
    from mergenet_b200 import c_segment
    a = np.zeros((2, 4, 4), np.float32)
    with pytest.raises(TypeError):
        c_segment.run_segmentation(None, a, 2, [(0, 1)], 0, 1, 0)
    with pytest.raises(ValueError):
        c_segment.run_segmentation(a.astype(np.float64), a, 2, [(0, 1)], 0, 1, 0)
    with pytest.raises(ValueError):
        c_segment.run_segmentation(a[:, :, ::2], a, 2, [(0, 1)], 0, 1, 0)
    with pytest.raises(ValueError):
        c_segment.run_segmentation(a[0], a, 2, [(0, 1)], 0, 1, 0)
    with pytest.raises(TypeError):
        c_segment.run_segmentation(a, a, 2, None, 0, 1, 0)


def test_exact_wrapper_keeps_the_same_argument_checks(lib_mod):
    from mergenet_b200 import c_segment, segmenter
    a = np.zeros((2, 4, 4), np.float32)
    with pytest.raises(TypeError):
        c_segment.run_segmentation_exact(None, a, 2, [(0, 1)], 0, 1, 0)
    with pytest.raises(ValueError):
        c_segment.run_segmentation_exact(a, a.astype(np.float64), 2, [(0, 1)], 0, 1, 0)
    with pytest.raises(TypeError):
        c_segment.run_segmentation_exact(a, a, 2, ((0, 1),), 0, 1, 0)
    with pytest.raises(ValueError):
        segmenter.ObjectSegmenter(a, a, 2, [(0, 1), (1, 0)], mode="exact")


def test_no_cpu_fallback(lib_mod):
    """Without a device the product path must fail loudly, never compute."""
    L = lib_mod.lib()
    if L.mn_device_count() > 0:
        pytest.skip("a GPU is present")
    from mergenet_b200 import c_segment, segmenter
    a = np.full((2, 4, 4), 0.5, np.float32)
    with pytest.raises(lib_mod.MergeNetError):
        c_segment.run_segmentation(a, a, 2, [(0, 1), (1, 0)], 0, 1, 0)
    with pytest.raises(lib_mod.MergeNetError):
        segmenter.BatchSegmenter(1, 4, 4, 2, [(0, 1), (1, 0)])
    with pytest.raises(lib_mod.MergeNetError):
        segmenter.ObjectSegmenter(a, a, 2, [(0, 1), (1, 0)]).run_segmentation()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mergenet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "libsegment_ref" not in text, f


def test_tie_order_switch_without_a_device_fails_loudly(lib_mod, monkeypatch, capfd):
    """MN_TIE_ORDER=reference routes the drop-in symbol to the tie-exact replay; without a device that route, like the
    hot path, leaves (0, -1) in the outputs, sets the status and says so on stderr."""
    import ctypes
    L = lib_mod.lib()
    if L.mn_device_count() > 0:
        pytest.skip("a GPU is present")
    F, I = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
    cp = np.full((2, 4, 5), 0.5, np.float32)
    sp = np.full((2, 4, 5), 0.5, np.float32)
    off = np.array([[0, 1], [1, 0]], np.int32)
    for value in ("reference", "fixed", None):
        if value is None:
            monkeypatch.delenv("MN_TIE_ORDER", raising=False)
        else:
            monkeypatch.setenv("MN_TIE_ORDER", value)
        mask = np.full((4, 5), 7, np.int32)
        ocls = np.full(20, 7, np.int32)
        L.c_run_segmentation(cp.ctypes.data_as(F), 2, sp.ctypes.data_as(F), 2, 5, 4, 2, off.ctypes.data_as(I),
                             mask.ctypes.data_as(I), ocls.ctypes.data_as(I), 0.0, 1.0, 0.0)
        assert L.mn_last_error() == 7, value
        assert (mask == 0).all() and (ocls == -1).all(), value
        assert "c_run_segmentation failed" in capfd.readouterr().err, value
