"""dev tool: time the edge pass alone.  python tools/edge_bench.py [H W C K B iters]  (env MN_EDGE2_NCONS="ncons,ctas" | "0")"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mergenet_b200 import _lib, synth
a = [int(x) for x in sys.argv[1:]]
H, W, C, K, B, iters = (a + [1024, 2048, 9, 10, 8, 20][len(a):])[:6]
offs = np.ascontiguousarray(np.array(synth.generate_offsets(40, K), np.int32))
L = _lib.lib()
for clip in (0, 1, 2):
    ms = ctypes.c_float(0)
    rc = L.mn_debug_edge_bench(H, W, C, K, offs.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), B, iters, clip, ctypes.byref(ms))
    per_img = ms.value / B
    gb = 4.0 * H * W * ((C + K) + (C + 2 * K)) / 1e9
    print("edge_bench cfg=%s rc=%d clip=%d ms/launch=%.4f ms/image=%.4f GB/s=%.1f frac(6538.9)=%.3f" % (
        os.environ.get("MN_EDGE2_NCONS", "auto"), rc, clip, ms.value, per_img, gb / (per_img * 1e-3), gb / (per_img * 1e-3) / 6538.9))
