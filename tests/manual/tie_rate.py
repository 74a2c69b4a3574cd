"""How often does the fixed tie rule of the hot path change the partition on REALISTIC maps?  (manual, build container)
cfg2-like soft maps (mergenet_b200.synth.cfg_cityscapes) at H x W, the Cityscapes recipe's options (0, 1.0, 0.03):
for every seed the unmodified reference, the oracle (fixed tie rule = the CUDA path's) and the tie-exact replay's
host build.  usage: python tests/manual/tie_rate.py <H> <W> <first_seed> <n> [procs]"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "golden"))
sys.path.insert(0, HERE)


def one(args):
    h, w, seed = args
    import cases
    import oracle
    import make_golden_exact
    from soak_sweep import load_exact, run_exact
    from mergenet_b200 import synth
    cp, sp, offs, _ = synth.cfg_cityscapes(h, w, seed=seed, n_shapes=max(4, int(400 * h * w / (1024 * 2048))),
                                          rmax=max(8, int(120 * h / 1024)), soft=True, noise_seed=seed + 7)
    opts = cases.RECIPE_OPTS
    t = time.time()
    rm, ro, _ = make_golden_exact.reference_raw(cp, sp, 9, offs, opts)
    t_ref = time.time() - t
    m0, c0, _ = oracle.oracle_run_segmentation(cp, sp, 9, offs, *opts)
    rc, xm, xo = run_exact(load_exact(), cp, sp, 9, offs, opts)
    same_fixed = cases.same_result(oracle, (rm, oracle._trim(ro)), (m0, c0))
    same_exact = rc == 0 and np.array_equal(xm, rm) and np.array_equal(xo, ro)
    ndiff = 0
    if not same_fixed:
        a = oracle.canonical_result(rm, oracle._trim(ro))[0]
        b = oracle.canonical_result(m0, c0)[0]
        ndiff = int((a != b).sum())
    return seed, same_fixed, same_exact, ndiff, int(rm.max()), t_ref


def main():
    h, w, s0, n = (int(v) for v in sys.argv[1:5])
    procs = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    with Pool(procs) as pool:
        res = pool.map(one, [(h, w, s0 + i) for i in range(n)], chunksize=1)
    for seed, sf, se, nd, ninst, t_ref in res:
        print("seed %d: %3d instances, fixed tie rule == reference: %s%s, tie-exact replay == reference raw: %s (reference %.1f s)"
              % (seed, ninst, sf, "" if sf else " (%d pixels differ after canonical relabel)" % nd, se, t_ref), flush=True)
    print("%dx%d: fixed tie rule equals the reference on %d of %d images; tie-exact replay on %d of %d"
          % (h, w, sum(r[1] for r in res), n, sum(r[2] for r in res), n))


if __name__ == "__main__":
    main()
