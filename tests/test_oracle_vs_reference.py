"""CPU: the C restatement (oracle/mergenet_oracle.c) against the reference itself and the committed
golden fixtures.  This is what pins the oracle (the reference ships no vectors of its own)."""
import glob
import os

import numpy as np
import pytest

import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("opts", [cases.RECIPE_OPTS, cases.PLAIN_OPTS, cases.QUARTER_OPTS])
def test_oracle_matches_reference_small(oracle_mod, opts):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libsegment_ref.so not built (reference tree absent)")
    for name, cp, sp, C, offs in cases.small_cases():
        ref = oracle_mod.ref_run_segmentation(cp, sp, C, offs, *opts)
        ora = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)[:2]
        assert cases.same_result(oracle_mod, ref, ora), name


def test_oracle_matches_reference_with_bias(oracle_mod):
    if not oracle_mod.have_reference():
        pytest.skip("reference .so absent")
    name, cp, sp, C, offs = cases.small_cases()[0]
    for sdb in (0.5, -0.7):
        ref = oracle_mod.ref_run_segmentation(cp, sp, C, offs, sdb, 1.0, 0.0)
        ora = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, sdb, 1.0, 0.0)[:2]
        assert cases.same_result(oracle_mod, ref, ora), (name, sdb)


def test_oracle_matches_golden_fixtures(oracle_mod):
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
    assert files, "golden fixtures missing"
    for f in files:
        g = np.load(f)
        offs = [tuple(int(v) for v in o) for o in g["offsets"]]
        opts = tuple(float(v) for v in g["opts"])
        ora = oracle_mod.oracle_run_segmentation(g["class_pred"], g["adj_pred"], int(g["num_classes"]), offs, *opts)[:2]
        ref = (g["ref_mask"], [int(v) for v in g["ref_object_class"]])
        assert cases.same_result(oracle_mod, ref, ora), os.path.basename(f)


def test_round_model_equals_sequential(oracle_mod):
    """The plan/commit round rule the CUDA scheduler implements, executed on the host."""
    import ctypes
    lib = oracle_mod.oracle_lib()
    F = ctypes.POINTER(ctypes.c_float); I = ctypes.POINTER(ctypes.c_int); LL = ctypes.POINTER(ctypes.c_longlong)
    lib.mno_run_rounds_model.argtypes = [F, ctypes.c_int, F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         I, I, I, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, LL, LL]
    for name, cp, sp, C, offs in cases.small_cases()[:4]:
        seq = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)[:2]
        cpc, apc, off, mask, ocls = oracle_mod._glue(cp, sp, offs)
        r = ctypes.c_longlong(); e = ctypes.c_longlong()
        rc = lib.mno_run_rounds_model(oracle_mod._fp(cpc), C, oracle_mod._fp(apc), apc.shape[0], apc.shape[2],
                                      apc.shape[1], C, oracle_mod._ip(off), oracle_mod._ip(mask), oracle_mod._ip(ocls),
                                      *cases.RECIPE_OPTS, 32, ctypes.byref(r), ctypes.byref(e))
        assert rc == 0, name
        assert cases.same_result(oracle_mod, seq, (mask, oracle_mod._trim(ocls))), name


def test_from_scratch_logprob_is_partition_function(oracle_mod):
    name, cp, sp, C, offs = cases.small_cases()[0]
    m, oc, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
    a = oracle_mod.total_logprob_from_scratch(m, oc, cp, sp, offs, 1.0)
    cm, ccls = oracle_mod.canonical_result(m, oc)
    b = oracle_mod.total_logprob_from_scratch(cm, ccls, cp, sp, offs, 1.0)
    assert abs(a - b) <= 1e-9 * abs(a)


@pytest.mark.parametrize("name", ["cfg1_256x512_recipe", "cfg1_256x512_plain", "cfg3_256x512_oracle"])
def test_oracle_equals_reference_fixtures_at_named_sizes(oracle_mod, name):
    """The restatement against tests/golden/matrix (results of the unmodified reference at BASELINE.json's
    named sizes; no reference tree needed at test time): mask, classes, from-scratch log-prob."""
    import os
    import matrix_cases
    cp, sp, C, offs, opts = matrix_cases.load(name)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "matrix", name + ".npz"))
    m, c, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
    cm, cc = oracle_mod.canonical_result(m, c)
    assert np.array_equal(cm, g["mask"]) and list(cc) == [int(v) for v in g["cls"]]
    lp = oracle_mod.total_logprob_from_scratch(m, c, cp, sp, offs, opts[1])
    assert abs(lp - float(g["logprob"])) <= 1e-5 * abs(float(g["logprob"]))


def _saturating_cases():
    """Maps on which libm's special values appear: (a) oracle-mode maps with same_different_bias = 2:
    the biased probability of every "same" pair rounds to exactly 1.0f, so log(1 - s) = -inf and the priority
    is +inf (cc:183-195,34); (b) unclipped class maps holding exact 0 and 1 handed to the raw C ABI."""
    name, cp, sp, C, offs = cases.small_cases()[1]  # city_oracle_48x64
    out = [("bias2_oracle_mode", cp, sp, C, offs, (2.0, 1.0, 0.03), True)]
    cp2 = cp.copy()
    cp2[cp2 > 0.5] = 1.0
    cp2[cp2 < 0.5] = 0.0
    out.append(("unclipped_class_0_1", cp2, sp, C, offs, cases.RECIPE_OPTS, False))
    return out


def test_oracle_matches_reference_on_saturated_maps(oracle_mod):
    if not oracle_mod.have_reference():
        pytest.skip("reference .so absent")
    for name, cp, sp, C, offs, opts, clip in _saturating_cases():
        import contextlib
        ctx = contextlib.nullcontext() if clip else oracle_mod.raw_abi()
        with ctx:
            ref = oracle_mod.ref_run_segmentation(cp, sp, C, offs, *opts)
            ora = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)[:2]
        assert cases.same_result(oracle_mod, ref, ora), name


def _quantized_case(h, w, seed, C=3, K=6):
    return cases.quantized(h, w, seed, C, K)


def test_tie_dependent_inputs_known_limitation(oracle_mod):
    """The HOT PATH's tie rule, kept visible: among EQUAL priorities the reference pops in the order its libstdc++
    binary heap happens to hold them (PriorityCompare compares the priority only, segment.h:270-275; the push order
    comes from unordered_map iteration, cc:650-652) -- a function of the whole push/pop history that no parallel
    scheduler can replay.  Oracle and hot path use one fixed total order instead (mn_common.h: mn_tie).  On inputs
    whose partition does not depend on the tie order (every fixture: soft maps, saturated oracle-mode maps,
    constant maps) results are identical to the reference; on block-quantized maps they are not, and this test
    records that instead of hiding it: the oracle is deterministic, produces a partition of comparable total
    log-prob, and differs from the reference's.  (The reference's own result on such inputs is what the tie-exact
    replay returns: tests/test_exact_tie_order.py.)"""
    if not oracle_mod.have_reference():
        pytest.skip("reference .so absent")
    differ = 0
    for h, w, seed in [(17, 23, 0), (24, 32, 1), (24, 32, 2)]:
        cp, sp, C, offs = _quantized_case(h, w, seed)
        ref = oracle_mod.ref_run_segmentation(cp, sp, C, offs, *cases.PLAIN_OPTS)
        ora = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *cases.PLAIN_OPTS)[:2]
        again = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *cases.PLAIN_OPTS)[:2]
        assert cases.same_result(oracle_mod, ora, again)
        lr = oracle_mod.total_logprob_from_scratch(ref[0], ref[1], cp, sp, offs, 1.0)
        lo = oracle_mod.total_logprob_from_scratch(ora[0], ora[1], cp, sp, offs, 1.0)
        assert abs(lr - lo) <= 0.05 * abs(lr), (lr, lo)  # two greedy runs of the same objective
        differ += 0 if cases.same_result(oracle_mod, ref, ora) else 1
    if differ == 0:
        pytest.fail("tie-dependent inputs now equal the reference: update README / DESIGN (limitation lifted)")
