"""checker script (GPU box): segment the 8 distinct bench images of rank 0 (bench.make_images(8, 1000)) and save
masks + classes to gpurun_out/bench8_masks.npz; tests/manual/bench_images_vs_oracle.py compares them with the
oracle at full size in the build container."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import bench
from mergenet_b200 import BatchSegmenter, SegmenterOptions
cp, sp, offs = bench.make_images(8, 1000)
seg = BatchSegmenter(8, bench.H, bench.W, bench.C, offs)
m, oc, n = seg.segment_host(cp, sp, SegmenterOptions(*bench.OPTS), clip=False)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "bench8_masks.npz"), masks=m.astype(np.int16), ninst=n,
                    classes=np.stack([oc[i][:512] for i in range(8)]))
print("instances", n.tolist(), "merges", [seg.stats(i)["merges"] for i in range(8)])
