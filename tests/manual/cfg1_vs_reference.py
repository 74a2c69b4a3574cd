"""BASELINE config 1 (the reference's own CPU-runnable case): maps from the reference UNet(9,10) at 256x512
(generated in the build container by tests/golden/make_golden.unet_cfg1 together with the outputs of the
compiled reference; the 6.7 MB file is not committed).  Compares the CUDA path with the reference outputs."""
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle, cases
from mergenet_b200 import c_segment, synth, BatchSegmenter, SegmenterOptions
g = np.load('tmp_cfg1_256x512.npz')
cp, sp = np.ascontiguousarray(g['class_pred']), np.ascontiguousarray(g['adj_pred'])
offs = synth.generate_offsets(40, 10)
for tag, opts in (("recipe", cases.RECIPE_OPTS), ("plain", cases.PLAIN_OPTS)):
    t = time.time()
    m, c = c_segment.run_segmentation(cp, sp, 9, offs, *opts)
    dt = time.time() - t
    ref = (g['mask_' + tag], [int(v) for v in g['cls_' + tag]])
    print('cfg1', tag, 'gpu %.2fs' % dt, len(c), 'instances; equals reference:', cases.same_result(oracle, ref, (m, c)), flush=True)
seg = BatchSegmenter(1, 256, 512, 9, offs)
seg.segment_host(cp[None], sp[None], SegmenterOptions(*cases.RECIPE_OPTS))
st = seg.stats(0)
print({k: st[k] for k in ('status', 'rounds', 'events', 'merges', 'solo_events', 'pairs', 'pixel_pool_used', 'queue_chunks_used')}, seg.timings()['merge_ms'])
