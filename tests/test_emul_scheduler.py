"""CPU: the CUDA scheduler source (mergenet_b200/csrc/mn_merge.cuh) compiled for the host and run with
one logical thread (tests/emul) must reproduce the oracle: identical masks / classes and the same
number of executed events and merges.  This tests the scheduler's logic; the GPU suite tests the
real kernels."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "emul")
EMUL_SO = os.path.join(EMUL_DIR, "libemul.so")
NAMES = ['rounds', 'events', 'merges', 'restores', 'invalid', 'solo', 'refills', 'flushes', 'splits', 'pairs',
         'cut_conf', 'cut_casc', 'cut_cap', 'qc_bump', 'plc_bump', 'tn_bump']


@pytest.fixture(scope="module")
def emul():
    srcs = [os.path.join(EMUL_DIR, "emul_merge.cpp")] + [
        os.path.join(HERE, "..", "mergenet_b200", "csrc", f) for f in ("mn_merge.cuh", "mn_layout.h", "mn_common.h")]
    if not os.path.exists(EMUL_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMUL_SO) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", EMUL_SO, srcs[0]])
    lib = ctypes.CDLL(EMUL_SO)
    F = ctypes.POINTER(ctypes.c_float); I = ctypes.POINTER(ctypes.c_int); LL = ctypes.POINTER(ctypes.c_longlong)
    lib.emul_run_segmentation.argtypes = [F, ctypes.c_int, F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          I, I, I, ctypes.c_float, ctypes.c_float, ctypes.c_float, LL]
    return lib


def run_emul(emul, oracle, cp, sp, C, offs, opts):
    cpc, apc, off, mask, ocls = oracle._glue(cp, sp, offs)
    st = (ctypes.c_longlong * 16)()
    rc = emul.emul_run_segmentation(oracle._fp(cpc), C, oracle._fp(apc), apc.shape[0], apc.shape[2], apc.shape[1],
                                    C, oracle._ip(off), oracle._ip(mask), oracle._ip(ocls), *opts, st)
    return rc, mask, oracle._trim(ocls), dict(zip(NAMES, list(st)))


@pytest.mark.parametrize("opts", [cases.RECIPE_OPTS, cases.PLAIN_OPTS])
def test_emulated_scheduler_matches_oracle_small(emul, oracle_mod, opts):
    for name, cp, sp, C, offs in cases.small_cases():
        m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
        rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, opts)
        assert rc == 0, (name, rc)
        assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), name
        assert st["merges"] == st0["merges"], name


def test_emulated_scheduler_matches_oracle_medium(emul, oracle_mod):
    for name, cp, sp, C, offs in cases.medium_cases():
        m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
        rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, cases.RECIPE_OPTS)
        assert rc == 0, (name, rc)
        assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), name
        assert st["merges"] == st0["merges"], name


def test_pop_key_is_consistent_with_the_pop_order(emul):
    """Rule (b) of the accept pass is evaluated on 64-bit pop keys (mn_pop_key): a pops before b must
    imply key(a) >= key(b), and a strictly higher priority a strictly larger key."""
    emul.emul_pop_key_violations.restype = ctypes.c_longlong
    emul.emul_pop_key_violations.argtypes = [ctypes.c_longlong, ctypes.c_uint]
    assert emul.emul_pop_key_violations(2_000_000, 7) == 0


STRESS = ["-DMN_LF=2 -DMN_LEAFCAP=64", "-DMN_REFILL_TARGET=16", "-DMN_LEAFCAP=128 -DMN_LF=1",
          "-DMN_LEAFCAP=64 -DMN_REFILL_TARGET=32 -DMN_LF=4"]


@pytest.mark.parametrize("flags", STRESS)
def test_queue_limits_do_not_change_results(oracle_mod, flags):
    """The queue's capacity constants must never change what is computed.  Regression: a refill used to
    stop after MN_LF visited leaves even when none of them held a live entry, took "nothing staged" for
    "the tree is empty" and dropped the entries behind them -- at 1024x2048 the last 1-3 merges of an
    image were lost (341 instead of 338 instances).  Small limits make that situation common."""
    so = os.path.join(EMUL_DIR, "libemul_stress_%s.so" % "".join(ch for ch in flags if ch.isalnum()))
    srcs = [os.path.join(EMUL_DIR, "emul_merge.cpp")] + [
        os.path.join(HERE, "..", "mergenet_b200", "csrc", f) for f in ("mn_merge.cuh", "mn_layout.h", "mn_common.h")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off"] + flags.split() +
                              ["-o", so, srcs[0]])
    lib = ctypes.CDLL(so)
    F = ctypes.POINTER(ctypes.c_float); I = ctypes.POINTER(ctypes.c_int); LL = ctypes.POINTER(ctypes.c_longlong)
    lib.emul_run_segmentation.argtypes = [F, ctypes.c_int, F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          I, I, I, ctypes.c_float, ctypes.c_float, ctypes.c_float, LL]
    for name, cp, sp, C, offs in cases.small_cases()[:5] + cases.medium_cases()[:2]:
        m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
        rc, m1, c1, st = run_emul(lib, oracle_mod, cp, sp, C, offs, cases.RECIPE_OPTS)
        assert rc == 0, (name, rc)
        assert st["merges"] == st0["merges"], name
        assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), name


def test_hash_overflow_area_keeps_results_and_fails_loudly_when_full(emul, oracle_mod, monkeypatch):
    """A tight (lo, hi) -> record table pushes records into the small overflow area (cached in shared
    memory on the GPU): results must not change; past its 128 entries the run must FAIL, not go wrong."""
    name, cp, sp, C, offs = cases.medium_cases()[0]
    m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
    monkeypatch.setenv("EMUL_HASH_PERMILLE", "1300")   # 1.3 slots per record instead of 1.6: tens of overflow records
    rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, cases.RECIPE_OPTS)
    assert rc == 0 and cases.same_result(oracle_mod, (m0, c0), (m1, c1)) and st["merges"] == st0["merges"]
    monkeypatch.setenv("EMUL_HASH_PERMILLE", "1000")
    rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, cases.RECIPE_OPTS)
    assert rc != 0


def test_emulated_scheduler_on_degenerate_inputs(emul, oracle_mod):
    for name, cp, sp, C, offs in cases.pathological_cases():
        for opts in (cases.RECIPE_OPTS, cases.PLAIN_OPTS):
            m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
            rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, opts)
            assert rc == 0 and st["merges"] == st0["merges"], (name, opts, rc)
            assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), (name, opts)


def test_pixel_pool_holds_two_huge_objects_merging(emul, oracle_mod):
    """Regression (found by tests/manual/soak_sweep.py, seed 11 case 1226): an image that collapses to ONE object whose
    last merge joins two objects of just over 2^k pixels each -- pixel-array capacities 2^(k+1) + 2^(k+1), the merged
    array 2^(k+2): four times the pixel count.  With 3 N ints per pool half the run ended with "pixel-list chunk pool
    exhausted" on its very last merge; a half now holds 4 N (mn_api.cu: ws_layout)."""
    rng = np.random.default_rng(11)
    sys_path_mod = os.path.join(HERE, "manual")
    import sys
    sys.path.insert(0, sys_path_mod)
    from soak_sweep import random_case
    for _ in range(1227):
        cp, sp, C, offs, opts, quant = random_case(rng, 64, 80)
    assert cp.shape == (2, 59, 75) and opts == (0.0, 0.25, 0.2)
    m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
    assert st0["final_objects"] == 1 and st0["max_abs_npix"] > 2048 and cp.shape[1] * cp.shape[2] - st0["max_abs_npix"] > 2048
    rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, opts)
    assert rc == 0, rc
    assert st["merges"] == st0["merges"] and cases.same_result(oracle_mod, (m0, c0), (m1, c1))


def test_pixel_pool_two_equal_halves_worst_case(emul, oracle_mod):
    """The constructed worst case: two blobs of 2^k + few pixels each (strong sameness inside, weak across) that merge
    last under a large merge_logprob_bias (cases.layout_regression_cases)."""
    name, cp, sp, C, offs, opts = [c for c in cases.layout_regression_cases() if c[0].startswith("two_halves")][0]
    m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
    assert st0["final_objects"] == 1 and st0["max_abs_npix"] >= 1900, st0
    rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, opts)
    assert rc == 0, rc
    assert st["merges"] == st0["merges"] and cases.same_result(oracle_mod, (m0, c0), (m1, c1))


def test_own_queue_chunks_start_behind_the_last_initial_key(emul, oracle_mod):
    """Regression (found by tests/manual/soak_sweep.py, seed 31 case 3407): the scheduler's own queue chunks share one
    arena with the sorted initial keys and used to start at floor(8 E / 1024) chunks, i.e. ON the last E % 128 keys.
    Those are sentinels of dormant / out-of-image slots on every usual shape; with ONE offset (0, 1), a large
    merge_logprob_bias (every record starts with a priority >= 0) and more than 1024 records they are real entries:
    the first chunk written destroyed them and a later refill decoded garbage record ids (a crash on the host build, an
    illegal address on the GPU).  Chunks now start at ceil(8 E / 1024)."""
    todo = [c for c in cases.layout_regression_cases() if c[0].startswith("keys_tail")]
    assert len(todo) == 3
    for name, cp, sp, C, offs, opts in todo:
        K, (h, w) = len(offs), cp.shape[1:]
        E = h * w * K
        oob = sum((h * w) - (h - abs(dr)) * (w - abs(dc)) for dr, dc in offs)
        assert E > 1024 and E % 128 > oob, (E, E % 128, oob)  # real keys in the last partial kilobyte of the key array
        m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
        assert st0["init_pushes"] == st0["init_records"]  # no dormant record: no sentinel among the valid slots
        rc, m1, c1, st = run_emul(emul, oracle_mod, cp, sp, C, offs, opts)
        assert rc == 0, (name, rc)
        assert st["merges"] == st0["merges"] and cases.same_result(oracle_mod, (m0, c0), (m1, c1)), name


def test_workspace_capacity_invariants(emul):
    """mn_workspace_caps (one definition for the library and this host build): the scheduler's own queue chunks start
    behind the last initial key for EVERY record count, a pixel-pool half holds 4 N ints, and the shapes at the limits
    of mn_plan_create do not overflow the int fields."""
    emul.emul_workspace_caps.restype = None
    emul.emul_workspace_caps.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong)]
    rng = np.random.default_rng(3)
    shapes = [(1, 1), (1643, 1643), (4425, 4425 * 11), (2 ** 21, 10 * 2 ** 21), (2 ** 24 - 1, 2 ** 25), (2 ** 24 - 1, 2 ** 24 - 1)]
    shapes += [(int(n), int(n) * int(k)) for n, k in zip(rng.integers(1, 2 ** 21, 2000), rng.integers(1, 17, 2000))]
    for N, E in shapes:
        if E > 2 ** 25:
            continue
        out = (ctypes.c_longlong * 7)()
        emul.emul_workspace_caps(N, E, out)
        pix_cap, qc_low_n, qc_cap, tn_cap, nb, ovf, qch = list(out)
        chunk_bytes = qch * 16
        assert qc_low_n * chunk_bytes >= 8 * E > (qc_low_n - 1) * chunk_bytes, (N, E)   # ceil(8 E / chunk bytes)
        assert qc_cap > qc_low_n and (qc_cap - qc_low_n) * qch >= 0.45 * E, (N, E)
        assert pix_cap // 2 >= 4 * N + 2048 and pix_cap < 2 ** 31, (N, E)
        assert tn_cap >= 8 * max(4096, E // 512) and nb * 8 >= 1.8 * E and ovf >= 128, (N, E)
