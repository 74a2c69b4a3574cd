"""bench.py attaches a parity check to its timing without touching oracle/: its own first-appearance relabel
must agree with the oracle's canonical_result, and it must notice a wrong mask."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_bench_parity_helper_accepts_any_labelling_of_the_reference_result_and_rejects_a_wrong_one(oracle_mod):
    b = _bench()
    g = np.load(os.path.join(ROOT, "tests", "golden", "full", "bench_image0_seed1000_noise1007.npz"))
    n = len(g["cls"])
    rng = np.random.default_rng(3)
    p = np.concatenate([[0], rng.permutation(n) + 1])       # the reference result under another instance numbering
    mask = p[g["mask"]].astype(np.int32)
    cls = np.zeros(n, np.int64)
    cls[p[1:] - 1] = g["cls"]
    cm, cc = oracle_mod.canonical_result(mask, cls.tolist())
    assert np.array_equal(cm, g["mask"]) and list(cc) == list(g["cls"])
    r = b.parity_with_reference_fixture(mask, cls.tolist(), 1024, 2048)
    assert r["mask_equal"] and r["classes_equal"] and r["instances"] == n
    bad = mask.copy()
    bad[bad == p[5]] = p[6]                                  # two instances merged by mistake
    r = b.parity_with_reference_fixture(bad, cls.tolist(), 1024, 2048)
    assert not r["mask_equal"]
    assert b.parity_with_reference_fixture(mask[:512], cls.tolist(), 512, 2048) is None   # no fixture for other shapes
