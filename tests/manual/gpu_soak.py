"""Time-bounded seeded sweep of the CUDA path against the oracle through the drop-in C ABI (manual, GPU box):
the generator of tests/manual/soak_sweep.py, alternating its normal and wide ranges.  Mismatches and failures are
listed, not fatal.   usage: python tests/manual/gpu_soak.py <seed> <seconds> [max_h max_w]"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
import oracle  # noqa: E402
from soak_sweep import random_case  # noqa: E402
from test_random_sweep import _same  # noqa: E402
from mergenet_b200 import c_segment  # noqa: E402


def main():
    seed, seconds = int(sys.argv[1]), float(sys.argv[2])
    max_h = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    max_w = int(sys.argv[4]) if len(sys.argv) > 4 else 80
    oracle.build()
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = bad = 0
    while time.time() - t0 < seconds:
        cp, sp, C, offs, opts, quant = random_case(rng, max_h, max_w, wide=(n % 2 == 0))
        tag = (seed, n, cp.shape, sp.shape, offs, opts, quant)
        m0, c0, _ = oracle.oracle_run_segmentation(cp, sp, C, offs, *opts)
        try:
            m1, c1 = c_segment.run_segmentation(cp, sp.copy(), C, offs, *opts)
            if not _same(oracle, (m0, c0), (m1, c1)):
                bad += 1
                print("MISMATCH", tag, flush=True)
        except Exception as e:  # a loud failure of the library
            bad += 1
            print("FAILED", repr(e), tag, flush=True)
        n += 1
    print("gpu soak seed %d: %d cases in %.0f s, %d bad" % (seed, n, time.time() - t0, bad), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
