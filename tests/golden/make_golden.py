"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libsegment_ref.so,
built from /root/reference/utils/csegment/segment.cc by oracle/Makefile) on small seeded inputs.
Run from the repo root in the build container (the reference tree cannot travel):

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle  # noqa: E402


def main():
    assert oracle.have_reference() or os.path.exists("/root/reference"), "reference needed"
    oracle.build()
    out = os.path.dirname(os.path.abspath(__file__))
    sel = {"city_soft_48x64": cases.RECIPE_OPTS, "city_oracle_48x64": cases.RECIPE_OPTS,
           "smooth_40x56": cases.PLAIN_OPTS, "coco_oracle_48x48": cases.RECIPE_OPTS,
           "city_soft_odd_45x67": cases.QUARTER_OPTS}
    for name, cp, sp, C, offs in cases.small_cases():
        if name not in sel:
            continue
        opts = sel[name]
        mask, ocls = oracle.ref_run_segmentation(cp, sp, C, offs, *opts)
        np.savez_compressed(os.path.join(out, name + ".npz"), class_pred=cp,
                            adj_pred=sp, num_classes=np.int32(C), offsets=np.array(offs, np.int32),
                            opts=np.array(opts, np.float32), ref_mask=mask.astype(np.int16),
                            ref_object_class=np.array(ocls, np.int32))
        print(name, mask.shape, len(ocls), "instances")


def unet_cfg1(h=64, w=128):
    """BASELINE config 1 at a fixture-friendly size: the REFERENCE's own UNet(9, 10)
    (models/Unet.py:118), torch.manual_seed(0), eval mode, input torch.rand(1, 3, h, w); class maps =
    sigmoid channels 0..8, sameness maps = channels 9..18.  models/__init__.py is bypassed (it imports
    caffe / pspnet code that this container lacks)."""
    import types
    import torch
    pkg = types.ModuleType("models")
    pkg.__path__ = ["/root/reference/models"]
    sys.modules["models"] = pkg
    from models.Unet import UNet
    torch.manual_seed(0)
    net = UNet(9, 10).eval()
    with torch.no_grad():
        out = torch.sigmoid(net(torch.rand(1, 3, h, w)))[0].numpy().astype(np.float32)
    return np.ascontiguousarray(out[:9]), np.ascontiguousarray(out[9:19])


def main_unet():
    from mergenet_b200 import synth
    out = os.path.dirname(os.path.abspath(__file__))
    cp, sp = unet_cfg1()
    offs = synth.generate_offsets(40, 10)
    for tag, opts in (("recipe", cases.RECIPE_OPTS), ("plain", cases.PLAIN_OPTS)):
        mask, ocls = oracle.ref_run_segmentation(cp, sp, 9, offs, *opts)
        np.savez_compressed(os.path.join(out, "unet_cfg1_64x128_%s.npz" % tag), class_pred=cp, adj_pred=sp,
                            num_classes=np.int32(9), offsets=np.array(offs, np.int32),
                            opts=np.array(opts, np.float32), ref_mask=mask.astype(np.int16),
                            ref_object_class=np.array(ocls, np.int32))
        print("unet_cfg1", tag, mask.shape, len(ocls), "instances")


def main_full(images=(0, 1)):
    """BASELINE config 2 at its full size (1024x2048, the bench generator with seeds 1000 + i): canonical
    mask and classes of the UNMODIFIED reference -> tests/golden/full/ (75 KB each; the inputs are
    regenerated from the seeds).  ~5 min and 7 GB of RAM per image:  python tests/golden/make_golden.py full"""
    from mergenet_b200 import synth
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "full")
    os.makedirs(out, exist_ok=True)
    for i in images:
        cp, sp, offs, _ = synth.cfg_cityscapes(1024, 2048, seed=1000 + i, n_shapes=400, rmax=120, soft=True, noise_seed=7 + i)
        mask, ocls = oracle.ref_run_segmentation(cp, sp, 9, offs, *cases.RECIPE_OPTS)
        cm, cc = oracle.canonical_result(mask, ocls)
        np.savez_compressed(os.path.join(out, "cfg2_1024x2048_seed%d.npz" % (1000 + i)), mask=cm.astype(np.int32),
                            cls=np.array(cc, np.int32), opts=np.array(cases.RECIPE_OPTS, np.float32))
        print("full", i, len(cc), "instances")


def _save_result(path, cp, sp, C, offs, opts, store_inputs=False, **extra):
    """Run the unmodified reference and store its canonical mask / classes and the float64 from-scratch
    log-prob of ITS partition (segment.cc:314-350 as oracle.total_logprob_from_scratch evaluates it)."""
    import time
    t0 = time.time()
    mask, ocls = oracle.ref_run_segmentation(cp, sp, C, offs, *opts)
    secs = time.time() - t0
    cm, cc = oracle.canonical_result(mask, ocls)
    lp = oracle.total_logprob_from_scratch(mask, ocls, cp, sp, offs, opts[1])
    d = dict(mask=cm.astype(np.int32), cls=np.array(cc, np.int32), opts=np.array(opts, np.float32),
             logprob=np.float64(lp), ref_seconds=np.float64(secs), **extra)
    if store_inputs:
        d.update(class_pred=cp, adj_pred=sp, num_classes=np.int32(C), offsets=np.array(offs, np.int32))
    np.savez_compressed(path, **d)
    print(os.path.basename(path), cm.shape, len(cc), "instances, logprob %.6f, reference %.1f s" % (lp, secs), flush=True)


def main_matrix(which=None):
    """The BASELINE.json configurations at their named sizes -> tests/golden/matrix/:
    cfg1 256x512 (the reference's UNet(9,10); inputs stored once, they cannot be regenerated without the
    reference tree), cfg3 oracle-mode 256x512 and 1024x2048, cfg4 512x512 C=81 K=16 oracle-mode and soft.
    Inputs other than cfg1's are regenerated from the seeds by tests/matrix_cases.py."""
    import matrix_cases
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "matrix")
    os.makedirs(out, exist_ok=True)
    for name in matrix_cases.NAMES:
        if which and name not in which:
            continue
        if name.startswith("cfg1"):
            inp = os.path.join(out, "cfg1_256x512_inputs.npz")
            if not os.path.exists(inp):
                cp, sp = unet_cfg1(256, 512)
                np.savez_compressed(inp, class_pred=cp, adj_pred=sp)
        cp, sp, C, offs, opts = matrix_cases.load(name)
        _save_result(os.path.join(out, name + ".npz"), cp, sp, C, offs, opts)


def main_full_logprob():
    """Adds the reference partition's float64 from-scratch log-prob to the tests/golden/full fixtures (the
    partition is the reference's: recomputed from the stored canonical mask, no reference run needed)."""
    from mergenet_b200 import synth
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "full")
    for fn, seed, noise in (("cfg2_1024x2048_seed1000.npz", 1000, 7), ("cfg2_1024x2048_seed1001.npz", 1001, 8),
                            ("bench_image0_seed1000_noise1007.npz", 1000, 1007)):
        z = dict(np.load(os.path.join(out, fn)))
        cp, sp, offs, _ = synth.cfg_cityscapes(1024, 2048, seed=seed, n_shapes=400, rmax=120, soft=True, noise_seed=noise)
        z["logprob"] = np.float64(oracle.total_logprob_from_scratch(z["mask"], list(z["cls"]), cp, sp, offs, float(z["opts"][1])))
        np.savez_compressed(os.path.join(out, fn), **z)
        print(fn, "logprob %.6f" % float(z["logprob"]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        main_full()
    elif len(sys.argv) > 1 and sys.argv[1] == "matrix":
        main_matrix(sys.argv[2:] or None)
    elif len(sys.argv) > 1 and sys.argv[1] == "full_logprob":
        main_full_logprob()
    else:
        main()
        main_unet()
