"""CPU: a seeded random sweep over shapes, class counts, offset lists and options -- the unmodified reference (when it
is compiled here), the oracle restatement and the CUDA scheduler's host build must all give the same partition.
Soft random maps: ties do not decide anything (README: known limitation), so equality with the reference is required."""
import ctypes
import os

import numpy as np
import pytest

import cases
from test_emul_scheduler import emul, run_emul  # noqa: F401  (fixture + helper)

from mergenet_b200 import synth


def _random_case(rng):
    h, w = int(rng.integers(1, 40)), int(rng.integers(1, 48))
    C = int(rng.integers(1, 12))
    K = int(rng.integers(1, 9))
    # offsets: distinct, non-zero, no pair (o, -o) (core_config.py:66-73 guarantees one record per pixel pair)
    offs = []
    while len(offs) < K:
        o = (int(rng.integers(-6, 7)), int(rng.integers(-6, 7)))
        if o == (0, 0) or o in offs or (-o[0], -o[1]) in offs:
            continue
        offs.append(o)
    kind = int(rng.integers(0, 3))
    if kind == 0:    # smooth fields
        cp, sp = synth.smooth_random_maps(h, w, C, K, int(rng.integers(0, 1 << 30)))
    elif kind == 1:  # uniform noise
        cp = synth.clip_probs(rng.random((C, h, w)).astype(np.float32))
        sp = synth.clip_probs(rng.random((K, h, w)).astype(np.float32))
    else:            # instances with soft targets
        m, cl = synth.gt_instance_mask(h, w, max(1, h * w // 150), 6, max(2, C), int(rng.integers(0, 1 << 30)))
        C = max(2, C)
        cp, sp = synth.soft_maps(m, cl, C, offs, int(rng.integers(0, 1 << 30)))
    omf = float(rng.choice([1.0, 0.25, 0.1, 2.0]))
    mlb = float(rng.choice([0.0, 0.03, -0.05, 0.2]))
    return cp, sp, C, offs, (0.0, omf, mlb)


def _same(oracle_mod, a, b):
    """cases.same_result, except when EVERY pixel ended as its own instance: the reference wrapper then drops the class
    of its last label (pyx:80-84 cuts the list at the first -1 and, finding none, at its last entry), and which label
    is last follows the reference's unordered_map order -- only the masks are comparable."""
    if int(np.max(a[0])) == a[0].size:
        return np.array_equal(oracle_mod.canonical_result(*a)[0], oracle_mod.canonical_result(*b)[0])
    return cases.same_result(oracle_mod, a, b)


def test_random_sweep_reference_oracle_scheduler(emul, oracle_mod):  # noqa: F811
    rng = np.random.default_rng(20261018)
    have_ref = oracle_mod.have_reference()
    for i in range(600):
        cp, sp, C, offs, opts = _random_case(rng)
        tag = (i, cp.shape, sp.shape, offs, opts)
        m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
        if have_ref:
            ref = oracle_mod.ref_run_segmentation(cp, sp, C, offs, *opts)
            assert _same(oracle_mod, ref, (m0, c0)), ("reference vs oracle", tag)
        rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, opts)
        assert rc == 0, ("scheduler status", rc, tag)
        assert st["merges"] == st0["merges"], ("merge count", tag)
        assert _same(oracle_mod, (m0, c0), (m1, c1)), ("oracle vs scheduler", tag)


def test_soak_generator_short(emul, oracle_mod):  # noqa: F811
    """A short run of the long seeded sweep (tests/manual/soak_sweep.py: wider ranges, block-quantized maps with thousands
    of exact ties, same_different_bias): oracle = scheduler host build everywhere; on non-quantized maps the reference =
    oracle unless the input has a decisive tie (rare, counted: DESIGN.md section 9)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "manual"))
    from soak_sweep import random_case
    have_ref = oracle_mod.have_reference()
    tie_dependent = 0
    for seed, wide, n in [(20261103, False, 250), (20261104, True, 150)]:
        rng = np.random.default_rng(seed)
        for i in range(n):
            cp, sp, C, offs, opts, quant = random_case(rng, 56, 64, wide)
            tag = (seed, i, cp.shape, sp.shape, offs, opts, quant)
            m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
            if opts[0] == 0.0:  # (the scheduler's host build starts after the edge pass, which owns the bias rewrite)
                rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, opts)
                assert rc == 0, ("scheduler status", rc, tag)
                assert st["merges"] == st0["merges"] and _same(oracle_mod, (m0, c0), (m1, c1)), ("oracle vs scheduler", tag)
            if have_ref and not quant:
                ref = oracle_mod.ref_run_segmentation(cp, sp, C, offs, *opts)
                tie_dependent += 0 if _same(oracle_mod, ref, (m0, c0)) else 1
    assert tie_dependent <= 2, tie_dependent


@pytest.mark.gpu
def test_random_sweep_cuda_path_equals_oracle(oracle_mod, lib_mod):
    """The same sweep through the drop-in C ABI on the GPU (every case is a new shape: a new plan each time)."""
    from mergenet_b200 import c_segment
    rng = np.random.default_rng(20261019)
    for i in range(40):
        cp, sp, C, offs, opts = _random_case(rng)
        m0, c0, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
        m1, c1 = c_segment.run_segmentation(cp, sp, C, offs, *opts)
        assert _same(oracle_mod, (m0, c0), (m1, c1)), (i, cp.shape, sp.shape, offs, opts)


@pytest.mark.gpu
def test_random_sweep_wide_cuda_path_equals_oracle(oracle_mod, lib_mod):
    """The generator of the long seeded sweep (tests/manual/soak_sweep.py) through the drop-in C ABI: up to 16 offsets
    reaching 25 pixels, up to 40 classes, quantized maps (thousands of exact ties: oracle and CUDA path share the fixed
    tie rule, so they must agree there too), same_different_bias != 0 (the in-place rewrite of the edge pass), shapes
    whose record count is not a multiple of anything."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "manual"))
    from soak_sweep import random_case
    from mergenet_b200 import c_segment
    for seed, wide, n in [(20261101, True, 50), (20261102, False, 50)]:
        rng = np.random.default_rng(seed)
        for i in range(n):
            cp, sp, C, offs, opts, quant = random_case(rng, 64, 80, wide)
            m0, c0, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
            m1, c1 = c_segment.run_segmentation(cp, sp.copy(), C, offs, *opts)
            assert _same(oracle_mod, (m0, c0), (m1, c1)), (seed, i, cp.shape, sp.shape, offs, opts, quant)
