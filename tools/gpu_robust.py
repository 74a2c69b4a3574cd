"""robustness probe (dev tool): benchmark-family configs other than the headline one, status + time"""
import sys, time, json
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mergenet_b200 import BatchSegmenter, SegmenterOptions, synth
import cases
opts = SegmenterOptions(0.0, 1.0, 0.03)
def run(name, cp, sp, C, offs):
    h, w = cp.shape[1], cp.shape[2]
    seg = BatchSegmenter(1, h, w, C, offs)
    t = time.time()
    try:
        m, oc, n = seg.segment_host(cp[None], sp[None], opts)
        st = seg.stats(0)
        print(name, 'ok %.2fs' % (time.time() - t), 'ninst', int(n[0]), {k: st[k] for k in ('status', 'rounds', 'events', 'merges', 'solo_events', 'splits', 'queue_chunks_used', 'pixel_pool_used', 'tree_nodes_used', 'hash_overflow')}, flush=True)
    except Exception as e:
        print(name, 'FAILED', e, seg.stats(0), flush=True)
    seg.close()
# cfg3: oracle-mode cityscapes, 512x1024 and full-res
for (h, w) in [(512, 1024), (1024, 2048)]:
    cp, sp, offs, _ = synth.cfg_cityscapes(h, w, seed=2, n_shapes=int(400 * h * w / (1024 * 2048)), rmax=int(120 * h / 1024), soft=False, noise_seed=7)
    run('cfg3_oracle_%dx%d' % (h, w), cp, sp, 9, offs)
# cfg4: COCO-shaped 512x512 C=81 K=16, soft and oracle
for soft in (True, False):
    cp, sp, C, offs = cases.coco_like(512, 512, 3, soft)
    run('cfg4_coco_512_%s' % ('soft' if soft else 'oracle'), cp, sp, C, offs)
