"""Long seeded soak of Mode B (manual, build container): the device source mn_modeb.cuh compiled for the host
(tests/emul/libemul_modeb.so, built by tests/test_modeb.py) against the UNMODIFIED reference class
utils/segmenter.py::ObjectSegmenter imported from /root/reference -- the partition prune() receives and the final
result (or the UnboundLocalError), without relabelling.  Small images (the pure-Python reference is slow), the
generator of tests/manual/soak_sweep.py incl. quantized maps (heapq's own tie order) and same_different_bias.
usage: python tests/manual/soak_modeb.py <seed> <seconds> [max_h max_w]"""
import ctypes
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "golden"))
sys.path.insert(0, HERE)
import make_golden_modeb  # noqa: E402
from soak_sweep import random_case  # noqa: E402
from mergenet_b200 import ObjectSegmenter, SegmenterOptions  # noqa: E402


def load_entry():
    lib = ctypes.CDLL(os.path.join(HERE, "..", "emul", "libemul_modeb.so"))
    V = ctypes.c_void_p
    lib.emul_modeb_segment.restype = ctypes.c_int
    lib.emul_modeb_segment.argtypes = [V, V, V, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_int), ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       V, V, V, V]
    return lambda *a: {0: 0, 1: 3, 2: 5, 3: 9}[lib.emul_modeb_segment(*a)]


def main():
    seed, seconds = int(sys.argv[1]), float(sys.argv[2])
    max_h = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    max_w = int(sys.argv[4]) if len(sys.argv) > 4 else 24
    entry = load_entry()
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = nraise = 0
    while time.time() - t0 < seconds:
        cp, sp, C, offs, opts, quant = random_case(rng, max_h, max_w)
        tag = (seed, n, cp.shape, sp.shape, offs, opts, quant)
        mask, ocls, mask0, ocls0 = make_golden_modeb.reference_modeb(cp, sp, C, offs, opts)
        seg = ObjectSegmenter(cp, sp, C, offs, SegmenterOptions(*opts))
        m0, c0 = seg._run_modeb(entry, -np.inf)
        if not (np.array_equal(m0, mask0) and list(c0) == list(ocls0)):
            print("MISMATCH before prune", tag, flush=True)
            return 1
        seg = ObjectSegmenter(cp, sp, C, offs, SegmenterOptions(*opts))
        if mask is None:
            nraise += 1
            try:
                seg._run_modeb(entry, 200.0)
                print("MISMATCH: the reference raises, the replay does not", tag, flush=True)
                return 1
            except UnboundLocalError:
                pass
        else:
            m1, c1 = seg._run_modeb(entry, 200.0)
            if not (np.array_equal(m1, mask) and list(c1) == list(ocls)):
                print("MISMATCH after prune", tag, flush=True)
                return 1
        n += 1
        if n % 200 == 0:
            print("modeb seed %d: %d cases ok (%d raise), %.0f s" % (seed, n, nraise, time.time() - t0), flush=True)
    print("modeb seed %d: all %d cases ok (%d raise), %.0f s" % (seed, n, nraise, time.time() - t0), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
