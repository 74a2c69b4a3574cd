"""Mode B (the reference's pure-Python `utils/segmenter.py::ObjectSegmenter`, the class the COCO recipe calls).

Oracle = the reference itself: tests/golden/modeb/*.npz are outputs of the imported, unmodified reference class
(tests/golden/make_golden_modeb.py).  Compared WITHOUT relabelling: the reference's labels are ascending surviving
object ids and its mask is int64, both part of the contract.

* CPU (`-m "not gpu"`): the device source mn_modeb.cuh compiled for the host (tests/emul/emul_modeb.cpp) -- the
  logic (heapq replay, dict orders, NumPy-2 promotion rules, prune) against the fixtures.
* GPU (`-m gpu`): mergenet_b200.ObjectSegmenter (default mode) through the C ABI mn_modeb_segment_host.
"""
import ctypes
import glob
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "emul")
EMUL_SO = os.path.join(EMUL_DIR, "libemul_modeb.so")
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "modeb", "*.npz")))


@pytest.fixture(scope="module")
def emul_entry():
    srcs = [os.path.join(EMUL_DIR, "emul_modeb.cpp")] + [
        os.path.join(HERE, "..", "mergenet_b200", "csrc", f) for f in ("mn_modeb.cuh", "mn_common.h")]
    if not os.path.exists(EMUL_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMUL_SO) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", EMUL_SO, srcs[0]])
    lib = ctypes.CDLL(EMUL_SO)
    V = ctypes.c_void_p
    lib.emul_modeb_segment.restype = ctypes.c_int
    lib.emul_modeb_segment.argtypes = [V, V, V, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_int), ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       V, V, V, V]

    def entry(*a):
        rc = lib.emul_modeb_segment(*a)
        return {0: 0, 1: 3, 2: 5, 3: 9}[rc]  # the library's status codes (include/mergenet_b200.h)
    return entry


def _load(f):
    g = np.load(f)
    offs = [tuple(int(v) for v in o) for o in g["offsets"]]
    return g, offs, tuple(float(v) for v in g["opts"])


def _check(seg_factory, run):
    assert FIXTURES, "tests/golden/modeb is empty"
    seen_raise = False
    for f in FIXTURES:
        g, offs, opts = _load(f)
        name = os.path.basename(f)
        # the partition prune() receives (threshold -inf prunes nothing)
        seg = seg_factory(g["class_pred"], g["adj_pred"], int(g["num_classes"]), offs, opts)
        m0, c0 = run(seg, -np.inf)
        assert m0.dtype == np.int64, name
        assert np.array_equal(m0, g["ref_mask_before_prune"]), name + " (before prune)"
        assert list(c0) == [int(v) for v in g["ref_object_class_before_prune"]], name + " (before prune)"
        # the reference's result: prune(200)
        seg = seg_factory(g["class_pred"], g["adj_pred"], int(g["num_classes"]), offs, opts)
        if int(g["raises"]):
            seen_raise = True
            with pytest.raises(UnboundLocalError):
                run(seg, 200.0)
            continue
        m1, c1 = run(seg, 200.0)
        assert np.array_equal(m1, g["ref_mask"]), name
        assert list(c1) == [int(v) for v in g["ref_object_class"]], name
    assert seen_raise, "no fixture covers the UnboundLocalError case of prune()"


def test_modeb_logic_on_host_matches_reference_fixtures(emul_entry):
    from mergenet_b200 import ObjectSegmenter, SegmenterOptions
    _check(lambda cp, sp, C, offs, opts: ObjectSegmenter(cp, sp, C, offs, SegmenterOptions(*opts)),
           lambda seg, thr: seg._run_modeb(emul_entry, thr))


@pytest.mark.gpu
def test_modeb_gpu_matches_reference_fixtures():
    from mergenet_b200 import ObjectSegmenter, SegmenterOptions
    _check(lambda cp, sp, C, offs, opts: ObjectSegmenter(cp, sp, C, offs, SegmenterOptions(*opts)),
           lambda seg, thr: seg.run_segmentation(prune_threshold=thr))


@pytest.mark.gpu
def test_modeb_default_options_and_csegment_mode_differ_in_semantics():
    """The class default is the reference CLASS's semantics; mode='csegment' is the C++ port's: on a soft map
    the two generally give different partitions (different priority denominators), and the dtypes differ."""
    import cases
    from mergenet_b200 import ObjectSegmenter
    name, cp, sp, C, offs = cases.small_cases()[0]
    mb, cb = ObjectSegmenter(cp, sp, C, offs).run_segmentation()
    ma, ca = ObjectSegmenter(cp, sp, C, offs, mode="csegment").run_segmentation()
    assert mb.dtype == np.int64 and ma.dtype == np.int32
    assert mb.shape == ma.shape
