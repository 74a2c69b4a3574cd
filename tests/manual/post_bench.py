"""checker script (run by hand, tests/ may use the oracle): time the post-pass (nearest resize 512x1024 -> 1024x2048, COCO RLE of a 1024x2048 mask with
~340 instances) on the GPU, device time from CUDA events, beside the oracle on one host core."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle
from mergenet_b200 import _lib, post, synth
oracle.build()
m, cl = synth.gt_instance_mask(1024, 2048, 400, 120, 9, 2)
m = np.ascontiguousarray(m, np.int32)
n = int(m.max())
for it in range(3):
    t = time.time(); s = post.coco_rle_counts(m, n); wall = time.time() - t
    ms = _lib.lib().mn_post_last_ms()
print("coco_rle 1024x2048 n=%d: device %.3f ms, call wall %.1f ms, bytes %d" % (n, ms, wall * 1e3, sum(len(x) for x in s)))
t = time.time(); s0 = oracle.oracle_coco_rle(m, n); cpu = time.time() - t
print("oracle (1 core) %.1f ms; equal %s; speed-up (device time) %.0fx" % (cpu * 1e3, s == s0, cpu * 1e3 / ms))
small = np.ascontiguousarray(m[::2, ::2])
for it in range(3):
    t = time.time(); big = post.resize_masks_nearest(small, 1024, 2048); wall = time.time() - t
    ms = _lib.lib().mn_post_last_ms()
t = time.time(); b0 = oracle.oracle_resize_nearest(small, 1024, 2048); cpu = time.time() - t
gb = (small.nbytes + big.nbytes) / 1e9
print("resize_nearest 512x1024 -> 1024x2048: device %.4f ms (%.0f GB/s algorithmic), oracle %.1f ms, equal %s" % (ms, gb / (ms * 1e-3), cpu * 1e3, np.array_equal(big, b0)))
