"""Generates tests/golden/modeb/*.npz by running the UNMODIFIED reference Python segmenter
(/root/reference/utils/segmenter.py::ObjectSegmenter -- "Mode B", the class the COCO recipe calls,
egs/coco/local/segment.py:155-164) on small seeded inputs.  The reference cannot travel, so its outputs are
committed as fixtures.  Run from the repo root in the build container:

    python tests/golden/make_golden_modeb.py

Each fixture holds the inputs, the options, and what `run_segmentation()` returned: the int64 mask (labels
in the reference's own order -- ascending surviving object id, NOT relabelled) and the class list; or
`raises = 1` where the reference raises UnboundLocalError in prune() (no class-0 object); plus the partition
prune() received (the reference's own output_mask() called at the top of its prune()), which at these sizes
is the discriminating part: prune(200) folds almost every small object into the background.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402


def reference_modeb(cp, sp, C, offs, opts):
    sys.path.insert(0, "/root/reference/utils")
    import segmenter as ref  # noqa: E402  (the reference module itself)
    o = ref.SegmenterOptions(same_different_bias=opts[0], object_merge_factor=opts[1], merge_logprob_bias=opts[2])
    pre = {}

    class Recording(ref.ObjectSegmenter):  # the reference's own methods; only records the partition prune() receives
        def prune(self, threshold=200.0):
            pre["mask"], pre["cls"] = self.output_mask()
            return ref.ObjectSegmenter.prune(self, threshold)

    with contextlib.redirect_stdout(io.StringIO()):
        seg = Recording(cp, sp, C, [tuple(int(v) for v in x) for x in offs], o)
        try:
            mask, ocls = seg.run_segmentation()
        except UnboundLocalError:
            mask, ocls = None, None
    return (None if mask is None else np.asarray(mask)), (None if ocls is None else [int(c) for c in ocls]), \
        np.asarray(pre["mask"]), [int(c) for c in pre["cls"]]


def modeb_cases():
    """(name, class_pred, adj_pred, C, offsets, opts).  Sizes the pure-Python reference finishes in seconds."""
    out = []
    out.append(("city_soft_24x32", ) + cases.cityscapes_like(24, 32, 20, True, n_shapes=6, rmax=10) + (cases.RECIPE_OPTS,))
    out.append(("city_soft_32x48_omf", ) + cases.cityscapes_like(32, 48, 21, True, n_shapes=8, rmax=12) + ((0.0, 0.1, 0.0),))
    out.append(("city_oracle_24x32", ) + cases.cityscapes_like(24, 32, 22, False, n_shapes=6, rmax=10) + (cases.RECIPE_OPTS,))
    out.append(("city_soft_sdb_20x28", ) + cases.cityscapes_like(20, 28, 23, True, n_shapes=5, rmax=9) + ((0.5, 1.0, 0.03),))
    out.append(("coco_oracle_20x20", ) + cases.coco_like(20, 20, 24, False) + ((0.0, 1.0 / 16, 0.0),))
    out.append(("coco_soft_18x22", ) + cases.coco_like(18, 22, 25, True) + (cases.RECIPE_OPTS,))
    out.append(("smooth_16x24", ) + cases.smooth(16, 24, 26) + (cases.PLAIN_OPTS,))
    out.append(("tiny_1x7", ) + cases.cityscapes_like(1, 7, 6, True, n_shapes=1, rmax=3) + (cases.RECIPE_OPTS,))
    out.append(("tiny_3x3", ) + cases.cityscapes_like(3, 3, 7, True, n_shapes=1, rmax=3) + (cases.PLAIN_OPTS,))
    out.append(("city_soft_48x64", ) + cases.cityscapes_like(48, 64, 0, True) + (cases.RECIPE_OPTS,))
    out.append(("city_oracle_40x60_omf", ) + cases.cityscapes_like(40, 60, 27, False, n_shapes=10, rmax=14) + ((0.0, 0.1, 0.0),))
    # quantized maps: many exactly equal priorities -> heapq's own tie order decides
    rng = np.random.default_rng(5)
    from mergenet_b200 import synth
    lv = np.array([0.2, 0.5, 0.8], np.float32)
    offs = synth.generate_offsets(40, 6)
    q = lambda n, h, w: np.repeat(np.repeat(lv[rng.integers(0, 3, size=(n, (h + 3) // 4, (w + 3) // 4))], 4, 1), 4, 2)[:, :h, :w].copy()  # noqa: E731
    out.append(("quantized_16x20", q(3, 16, 20), q(6, 16, 20), 3, offs, cases.RECIPE_OPTS))
    out.append(("quantized_12x24", q(3, 12, 24), q(6, 12, 24), 3, offs, cases.PLAIN_OPTS))
    return out


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "modeb")
    os.makedirs(out, exist_ok=True)
    for name, cp, sp, C, offs, opts in modeb_cases():
        mask, ocls, mask0, ocls0 = reference_modeb(cp, sp, C, offs, opts)
        raises = mask is None
        h, w = cp.shape[1:]
        np.savez_compressed(os.path.join(out, name + ".npz"), class_pred=cp, adj_pred=sp, num_classes=np.int32(C),
                            offsets=np.array(offs, np.int32), opts=np.array(opts, np.float64),
                            raises=np.int32(raises),
                            ref_mask=(np.zeros((h, w), np.int64) if raises else mask.astype(np.int64)),
                            ref_object_class=np.array([] if raises else ocls, np.int32),
                            ref_mask_before_prune=mask0.astype(np.int64), ref_object_class_before_prune=np.array(ocls0, np.int32))
        print(name, cp.shape, "raises UnboundLocalError" if raises else "%d instances, mask dtype %s" % (len(ocls), mask.dtype),
              "| before prune:", len(ocls0))


if __name__ == "__main__":
    main()
