"""quick perf probe (dev tool): python tools/probe.py H W B [soft]"""
import sys, time, json
import numpy as np
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from mergenet_b200 import BatchSegmenter, SegmenterOptions, synth
h, w, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
soft = (sys.argv[4] != 'oracle') if len(sys.argv) > 4 else True
cps, sps = [], []
for i in range(min(B, 2)):
    cp, sp, offs, _ = synth.cfg_cityscapes(h, w, seed=1000 + i, n_shapes=max(4, int(400 * h * w / (1024 * 2048))), rmax=max(8, int(120 * h / 1024)), soft=soft, noise_seed=7 + i)
    cps.append(cp); sps.append(sp)
cp = np.ascontiguousarray(np.stack([cps[i % len(cps)] for i in range(B)])); sp = np.ascontiguousarray(np.stack([sps[i % len(sps)] for i in range(B)]))
seg = BatchSegmenter(B, h, w, 9, offs)
opts = SegmenterOptions(0.0, 1.0, 0.03)
for it in range(2):
    t = time.time()
    try:
        m, oc, n = seg.segment_host(cp, sp, opts, clip=False)
    except Exception as e:
        print('ERR', e)
    dt = time.time() - t
    print("iter", it, "wall %.3fs" % dt, json.dumps(seg.timings()), seg.total_logprob(0))
print(json.dumps(seg.stats(0)))
print('ninst', n[:4] if 'n' in dir() else None)
