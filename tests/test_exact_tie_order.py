"""Tie-exact replay of the reference's C++ segmenter (mergenet_b200/csrc/mn_exact.cuh, mn_stl_order.h).

The hot path breaks ties among EXACTLY equal priorities by a fixed rule; the reference's order there is an artefact
of libstdc++ (std::push_heap / std::pop_heap with a comparator that sees the priority only, segment.h:270-275, fed in
std::unordered_map iteration order, segment.cc:650-652).  `c_segment.run_segmentation_exact` replays those container
orders literally and must return the reference's RAW arrays (label numbering included).

* CPU (`-m "not gpu"`): (1) the restated containers against the real ones of this toolchain (tests/emul/
  stl_order_check.cpp); (2) the device source compiled for the host (tests/emul/emul_exact.cpp) against the fixtures
  of tests/golden/exact (raw outputs of the unmodified reference, tests/golden/make_golden_exact.py) and, where
  oracle/_ref is present, against the reference itself on further seeded inputs, also with a bucket arena so small
  that the semi-space collection runs.
* GPU (`-m gpu`): `mergenet_b200.c_segment.run_segmentation_exact` through the C ABI against the same fixtures.
"""
import ctypes
import glob
import os
import subprocess

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "emul")
CSRC = os.path.join(HERE, "..", "mergenet_b200", "csrc")
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "exact", "*.npz")))
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
_DEPS = [os.path.join(CSRC, f) for f in ("mn_exact.cuh", "mn_stl_order.h", "mn_stl_primes.h", "mn_common.h")]


def _stale(target, srcs):
    return not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in srcs)


def test_stl_order_restatement_matches_libstdcxx():
    """mn_stl_order.h == this toolchain's std::unordered_map / std::priority_queue on random operation sequences."""
    exe = os.path.join(EMUL_DIR, "stl_order_check.bin")
    src = os.path.join(EMUL_DIR, "stl_order_check.cpp")
    if _stale(exe, [src] + _DEPS):
        subprocess.check_call([CXX, "-std=c++17", "-O2", "-ffp-contract=off", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("ok "), out.stdout + out.stderr


def test_prime_table_is_this_toolchains():
    """mn_stl_primes.h (generated, committed) equals what tools/gen_stl_primes.py reads out of libstdc++ now."""
    import tempfile
    src = ('#include <cstdio>\nnamespace std { namespace __detail { extern const unsigned long __prime_list[]; } }\n'
           'int main() { for (int i = 0; i < 256; i++) printf("%lu\\n", std::__detail::__prime_list[i]); }\n')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.cpp"), "w").write(src)
        subprocess.check_call([CXX, "-O1", os.path.join(d, "p.cpp"), "-o", os.path.join(d, "p")])
        live = [int(v) for v in subprocess.check_output([os.path.join(d, "p")]).split()]
    text = open(os.path.join(CSRC, "mn_stl_primes.h")).read()
    body = text[text.index("#define MN_STL_PRIMES"):].replace("#define MN_STL_PRIMES", "").replace("\\", "")
    committed = [int(v.strip().rstrip("u")) for v in body.split(",") if v.strip()]
    assert committed == live


@pytest.fixture(scope="module")
def emul_exact():
    so = os.path.join(EMUL_DIR, "libemul_exact.so")
    src = os.path.join(EMUL_DIR, "emul_exact.cpp")
    if _stale(so, [src] + _DEPS):
        subprocess.check_call([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    lib = ctypes.CDLL(so)
    F, I = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
    lib.emul_exact_segment.restype = ctypes.c_int
    lib.emul_exact_segment.argtypes = [F, ctypes.c_int, F, ctypes.c_int, ctypes.c_int, ctypes.c_int, I, ctypes.c_float,
                                       ctypes.c_float, ctypes.c_float, I, I, I, ctypes.POINTER(ctypes.c_longlong),
                                       ctypes.c_longlong]

    def run(cp, sp, C, offs, opts, arena_half_words=0):
        """called like c_segment.pyx calls the symbol: clipped maps, int32 offsets, zeroed outputs"""
        eps = np.finfo(np.float32).eps
        cp = np.ascontiguousarray(np.asarray(cp, np.float32).clip(eps, 1.0 - eps)).copy()
        sp = np.ascontiguousarray(np.asarray(sp, np.float32).clip(eps, 1.0 - eps)).copy()
        off = np.ascontiguousarray(np.array(offs).astype(np.int32))
        k, h, w = sp.shape
        mask = np.zeros((h, w), np.int32)
        ocls = np.zeros((1, h * w), np.int32)
        n = ctypes.c_int(0)
        st = (ctypes.c_longlong * 8)()
        rc = lib.emul_exact_segment(cp.ctypes.data_as(F), cp.shape[0], sp.ctypes.data_as(F), k, w, h, off.ctypes.data_as(I),
                                    opts[0], opts[1], opts[2], mask.ctypes.data_as(I), ocls.ctypes.data_as(I),
                                    ctypes.byref(n), st, arena_half_words)
        return rc, mask, ocls, dict(zip(("pops", "merges", "pushes", "collections", "arena_used", "arena_half"), list(st)))
    return run


def _load(f):
    g = np.load(f)
    offs = [tuple(int(v) for v in o) for o in g["offsets"]]
    return g, offs, tuple(float(v) for v in g["opts"])


def test_fixtures_cover_tie_dependent_inputs():
    assert len(FIXTURES) >= 10, "tests/golden/exact is incomplete"
    assert sum(int(np.load(f)["tie_dependent"]) for f in FIXTURES) >= 5


def test_exact_replay_on_host_equals_reference_fixtures(emul_exact):
    for f in FIXTURES:
        g, offs, opts = _load(f)
        rc, mask, ocls, st = emul_exact(g["class_pred"], g["adj_pred"], int(g["num_classes"]), offs, opts)
        name = os.path.basename(f)
        assert rc == 0, name
        assert np.array_equal(mask, g["ref_mask"]), name
        assert np.array_equal(ocls, g["ref_object_class"]), name
        # a half-space not far above the live peak (~3 words per record slot; the default provides 5 E + 32 N):
        # the collection must run, many times, and change nothing
        if mask.size >= 300:
            for frac in (0.25, 0.35, 0.5, 0.7):
                rc2, mask2, ocls2, st2 = emul_exact(g["class_pred"], g["adj_pred"], int(g["num_classes"]), offs, opts,
                                                    int(st["arena_half"] * frac))
                if rc2 == 2:  # the live arrays did not fit: said so, loudly
                    continue
                assert rc2 == 0 and st2["collections"] > st["collections"], (name, rc2, st2)
                assert np.array_equal(mask2, mask) and np.array_equal(ocls2, ocls), name
                break
            else:
                pytest.fail("no arena between 0.25 and 0.7 of the default worked: " + name)


def test_exact_replay_on_host_equals_live_reference(emul_exact):
    """Further seeded inputs against the reference itself (needs oracle/_ref, i.e. the build container)."""
    import oracle
    if not oracle.have_reference():
        pytest.skip("reference .so absent")
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_exact
    todo = [("quant_%dx%d_%d" % (h, w, s),) + cases.quantized(h, w, s, C, K) + (o,)
            for (h, w, s, C, K, o) in [(33, 47, 20, 3, 6, cases.PLAIN_OPTS), (64, 96, 21, 4, 8, cases.RECIPE_OPTS),
                                       (28, 28, 22, 9, 10, cases.QUARTER_OPTS), (50, 40, 23, 2, 16, cases.PLAIN_OPTS),
                                       (128, 160, 24, 5, 10, cases.RECIPE_OPTS), (7, 90, 25, 3, 3, (0.5, 1.0, 0.0))]]
    todo += [(n, cp, sp, C, offs, cases.RECIPE_OPTS) for n, cp, sp, C, offs in cases.small_cases()]
    todo += [(n, cp, sp, C, offs, cases.PLAIN_OPTS) for n, cp, sp, C, offs in cases.pathological_cases()]
    differ = 0
    for name, cp, sp, C, offs, opts in todo:
        rmask, rocls, _ = make_golden_exact.reference_raw(cp, sp, C, offs, opts)
        rc, mask, ocls, st = emul_exact(cp, sp, C, offs, opts)
        assert rc == 0, name
        assert np.array_equal(mask, rmask), name
        assert np.array_equal(ocls, rocls), name
        if name.startswith("quant"):
            fixed = oracle.oracle_run_segmentation(cp, sp, C, offs, *opts)[:2]
            differ += 0 if cases.same_result(oracle, (rmask, oracle._trim(rocls)), fixed) else 1
    assert differ >= 3  # these inputs do depend on the tie order: the fixed rule gives another partition
    # ... and the generator of the long seeded sweep (tests/manual/soak_sweep.py): quantized / noisy / smooth / soft
    # maps, 1-16 offsets, same_different_bias
    sys.path.insert(0, os.path.join(HERE, "manual"))
    from soak_sweep import random_case
    for seed, wide, n in [(20261105, False, 150), (20261106, True, 100)]:
        rng = np.random.default_rng(seed)
        for i in range(n):
            cp, sp, C, offs, opts, quant = random_case(rng, 48, 56, wide)
            rmask, rocls, _ = make_golden_exact.reference_raw(cp, sp, C, offs, opts)
            rc, mask, ocls, st = emul_exact(cp, sp, C, offs, opts)
            assert rc == 0 and np.array_equal(mask, rmask) and np.array_equal(ocls, rocls), (seed, i, cp.shape, sp.shape, offs, opts, quant)


@pytest.mark.gpu
def test_exact_gpu_equals_reference_fixtures():
    from mergenet_b200 import ObjectSegmenter, SegmenterOptions, c_segment
    assert FIXTURES, "tests/golden/exact is empty"
    for f in FIXTURES:
        g, offs, opts = _load(f)
        name = os.path.basename(f)
        st = {}
        adj = g["adj_pred"].copy()
        mask, ocls = c_segment.run_segmentation_exact(g["class_pred"], adj, int(g["num_classes"]), offs, *opts, stats=st)
        assert mask.dtype == np.int32 and np.array_equal(mask, g["ref_mask"]), name
        ref_list = []
        for v in g["ref_object_class"][0, :-1]:
            if v == -1:
                break
            ref_list.append(int(v))
        assert [int(c) for c in ocls] == ref_list, name
        assert st["merges"] > 0 or mask.size == 1, name
    # the class facade, and the difference to the fixed tie rule of the hot path on a tie-dependent input
    g, offs, opts = _load([f for f in FIXTURES if f.endswith("quant_24x32_a.npz")][0])
    seg = ObjectSegmenter(g["class_pred"], g["adj_pred"], int(g["num_classes"]), offs, SegmenterOptions(*opts), mode="csegment-exact")
    m2, c2 = seg.run_segmentation()
    assert np.array_equal(m2, g["ref_mask"])
    m3, c3 = ObjectSegmenter(g["class_pred"], g["adj_pred"], int(g["num_classes"]), offs, SegmenterOptions(*opts),
                             mode="csegment").run_segmentation()
    import oracle
    assert not cases.same_result(oracle, (m2, list(c2)), (m3, list(c3)))


@pytest.mark.gpu
def test_exact_gpu_bucket_arena_collection():
    """A small bucket arena (test hook MN_EXACT_ARENA_WORDS): the semi-space collection runs on the device and the
    result does not change; an arena below the live peak fails loudly instead of corrupting anything."""
    from mergenet_b200 import _lib, c_segment
    g, offs, opts = _load([f for f in FIXTURES if f.endswith("quant_96x128.npz")][0])
    N = g["ref_mask"].size
    E = N * len(offs)
    old = os.environ.get("MN_EXACT_ARENA_WORDS")
    try:
        os.environ["MN_EXACT_ARENA_WORDS"] = str(int(3.6 * E))
        st = {}
        mask, ocls = c_segment.run_segmentation_exact(g["class_pred"], g["adj_pred"].copy(), int(g["num_classes"]), offs, *opts, stats=st)
        assert st["bucket_arena_collections"] > 0, st
        assert np.array_equal(mask, g["ref_mask"])
        os.environ["MN_EXACT_ARENA_WORDS"] = str(int(0.5 * E))
        with pytest.raises(_lib.MergeNetError):
            c_segment.run_segmentation_exact(g["class_pred"], g["adj_pred"].copy(), int(g["num_classes"]), offs, *opts)
    finally:
        if old is None:
            os.environ.pop("MN_EXACT_ARENA_WORDS", None)
        else:
            os.environ["MN_EXACT_ARENA_WORDS"] = old


@pytest.mark.gpu
def test_drop_in_symbol_with_reference_tie_order(monkeypatch):
    """MN_TIE_ORDER=reference: the drop-in symbol itself (what the reference's own Cython binding links) runs the
    tie-exact replay -- raw arrays of the reference on a tie-dependent fixture; without the variable, the hot path."""
    import oracle
    from mergenet_b200 import c_segment
    g, offs, opts = _load([f for f in FIXTURES if f.endswith("quant_24x32_a.npz")][0])
    assert int(g["tie_dependent"]) == 1
    ref_list = [int(v) for v in g["ref_object_class"][0, :int(g["ref_mask"].max())]]
    monkeypatch.setenv("MN_TIE_ORDER", "reference")
    mask, ocls = c_segment.run_segmentation(g["class_pred"], g["adj_pred"].copy(), int(g["num_classes"]), offs, *opts)
    assert np.array_equal(mask, g["ref_mask"]) and [int(c) for c in ocls] == ref_list
    monkeypatch.delenv("MN_TIE_ORDER")
    mask2, ocls2 = c_segment.run_segmentation(g["class_pred"], g["adj_pred"].copy(), int(g["num_classes"]), offs, *opts)
    assert not cases.same_result(oracle, (mask, list(ocls)), (mask2, list(ocls2)))  # the fixed tie rule: another partition
    monkeypatch.setenv("MN_TIE_ORDER", "fixed")  # anything but "reference": the hot path
    mask3, ocls3 = c_segment.run_segmentation(g["class_pred"], g["adj_pred"].copy(), int(g["num_classes"]), offs, *opts)
    assert np.array_equal(mask3, mask2)
