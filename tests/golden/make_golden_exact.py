"""Generates tests/golden/exact/*.npz: RAW outputs of the UNMODIFIED reference C++ segmenter
(oracle/_ref/libsegment_ref.so, built by oracle/Makefile from /root/reference/utils/csegment/segment.cc with this
image's g++ 13 / libstdc++) on inputs whose result depends on the pop order among exactly equal priorities
(block-quantized maps), plus a few ordinary ones.  "Raw" = the mask and object_class arrays exactly as
c_run_segmentation wrote them, label numbering included -- what mergenet_b200.c_segment.run_segmentation_exact
(the tie-exact replay, mn_exact.cuh) must reproduce byte for byte.  Run from the repo root in the build container:

    python tests/golden/make_golden_exact.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle  # noqa: E402


def reference_raw(cp, sp, C, offs, opts):
    """(mask int32[H,W], object_class int32[1,H*W], adj_pred after the call) from the reference symbol itself,
    called like c_segment.pyx does (clip, int32 offsets, zero-initialised outputs)."""
    cp_, ap_, off, mask, ocls = oracle._glue(cp, sp, offs)
    with oracle._quiet_stdout():
        oracle.ref_lib().c_run_segmentation(oracle._fp(cp_), cp_.shape[0], oracle._fp(ap_), ap_.shape[0], ap_.shape[2],
                                            ap_.shape[1], int(C), oracle._ip(off), oracle._ip(mask), oracle._ip(ocls),
                                            float(opts[0]), float(opts[1]), float(opts[2]))
    return mask, ocls, ap_


def main():
    assert oracle.have_reference(), "oracle/_ref/libsegment_ref.so is missing: python -c 'import oracle; oracle.build()'"
    out_dir = os.path.join(ROOT, "tests", "golden", "exact")
    os.makedirs(out_dir, exist_ok=True)
    for name, cp, sp, C, offs, opts in cases.tie_exact_cases():
        mask, ocls, _ = reference_raw(cp, sp, C, offs, opts)
        again, ocls2, _ = reference_raw(cp, sp, C, offs, opts)
        assert np.array_equal(mask, again) and np.array_equal(ocls, ocls2), name  # the reference is deterministic
        fixed = oracle.oracle_run_segmentation(cp, sp, C, offs, *opts)[:2]
        tie_dependent = not cases.same_result(oracle, (mask, oracle._trim(ocls)), fixed)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), class_pred=cp, adj_pred=sp, num_classes=C,
                            offsets=np.array(offs, np.int32), opts=np.array(opts, np.float64), ref_mask=mask,
                            ref_object_class=ocls, tie_dependent=int(tie_dependent))
        print("%-22s %3d instances, differs from the fixed tie rule: %s" % (name, int(mask.max()), tie_dependent))


if __name__ == "__main__":
    main()
