"""checker script (build container): the oracle on the 8 distinct bench images at full size (2 min each, run in
parallel) against gpurun_out/bench8_masks.npz written by bench_images_dump.py on the GPU box."""
import os, sys, time
import multiprocessing as mp
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def work(i):
    import bench, oracle
    cp, sp, offs = bench.make_images(1, 1000 + i)
    m, c, st = oracle.oracle_run_segmentation(cp[0], sp[0], bench.C, offs, *bench.OPTS)
    cm, cc = oracle.canonical_result(m, c)
    return i, cm.astype(np.int32), cc, st["merges"]


if __name__ == "__main__":
    import oracle
    oracle.build()
    g = np.load(os.path.join(ROOT, "gpurun_out", "bench8_masks.npz"))
    with mp.get_context("fork").Pool(8) as pool:
        for i, cm, cc, merges in pool.imap_unordered(work, range(8)):
            n = int(g["ninst"][i])
            gm, gc = oracle.canonical_result(g["masks"][i].astype(np.int32), [int(v) for v in g["classes"][i][:n]])
            print("image", i, "oracle instances", len(cc), "merges", merges, "| GPU instances", n, "| equal:",
                  bool(np.array_equal(cm, gm) and list(cc) == list(gc)), flush=True)
