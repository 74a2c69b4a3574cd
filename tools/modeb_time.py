"""dev tool: wall time of Mode B (ObjectSegmenter default mode: one GPU thread replays heapq) on soft maps.
usage: python tools/modeb_time.py [H W]..."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import cases
from mergenet_b200 import ObjectSegmenter, SegmenterOptions
a = [int(x) for x in sys.argv[1:]] or [48, 64, 96, 128, 128, 256]
for h, w in zip(a[0::2], a[1::2]):
    cp, sp, C, offs = cases.cityscapes_like(h, w, 3, True, rmax=30)
    seg = ObjectSegmenter(cp, sp, C, offs, SegmenterOptions(0.0, 1.0, 0.03))
    t = time.time()
    m, c = seg.run_segmentation()
    print("modeb %dx%d: %.2f s  %s  instances %d" % (h, w, time.time() - t, seg.stats, len(c)), flush=True)
