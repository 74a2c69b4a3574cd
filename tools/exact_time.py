"""dev tool: wall time of the tie-exact replay (c_segment.run_segmentation_exact: the edge pass, then one GPU thread
replaying libstdc++'s heap and hash-table orders) on block-quantized maps.
usage: python tools/exact_time.py [H W]..."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import cases
from mergenet_b200 import c_segment
a = [int(x) for x in sys.argv[1:]] or [48, 64, 96, 128, 128, 256]
for h, w in zip(a[0::2], a[1::2]):
    cp, sp, C, offs = cases.quantized(h, w, 11, C=5, K=10)
    st = {}
    t = time.time()
    m, c = c_segment.run_segmentation_exact(cp, sp, C, offs, 0.0, 1.0, 0.03, stats=st)
    print("exact %dx%d: %.2f s  %s  instances %d" % (h, w, time.time() - t, st, len(c)), flush=True)
