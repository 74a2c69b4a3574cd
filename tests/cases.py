"""Seeded test inputs shared by the CPU and GPU suites (sizes the oracle finishes in seconds)."""
import numpy as np

from mergenet_b200 import synth

RECIPE_OPTS = (0.0, 1.0, 0.03)   # egs/cityscape/local/segment.py:134-136
PLAIN_OPTS = (0.0, 1.0, 0.0)
QUARTER_OPTS = (0.0, 0.25, 0.0)  # ObjectSegmenterOption default omf (segment.h:250-254)


def cityscapes_like(h, w, seed, soft, n_shapes=None, rmax=20):
    offs = synth.generate_offsets(40, 10)
    n_shapes = n_shapes if n_shapes is not None else max(3, h * w // 2000)
    m, cl = synth.gt_instance_mask(h, w, n_shapes, rmax, 9, seed)
    if soft:
        cp, sp = synth.soft_maps(m, cl, 9, offs, seed + 5)
    else:
        cp, sp = synth.oracle_mode_maps(m, cl, 9, offs)
    return cp, sp, 9, offs


def coco_like(h, w, seed, soft):
    offs = synth.generate_offsets(40, 16)
    m, cl = synth.gt_instance_mask(h, w, max(4, h * w // 300), 8, 81, seed)
    if soft:
        cp, sp = synth.soft_maps(m, cl, 81, offs, seed + 11)
    else:
        cp, sp = synth.oracle_mode_maps(m, cl, 81, offs)
    return cp, sp, 81, offs


def smooth(h, w, seed):
    offs = synth.generate_offsets(40, 10)
    cp, sp = synth.smooth_random_maps(h, w, 9, 10, seed)
    return cp, sp, 9, offs


def quantized(h, w, seed, C=3, K=6):
    """Block-quantized maps (levels 0.2 / 0.5 / 0.8 on 4x4 blocks): thousands of EXACTLY equal priorities whose pop
    order decides which of several equally good merges happens first -- the inputs on which the reference's result
    depends on libstdc++'s heap / hash-table orders."""
    rng = np.random.default_rng(seed)
    lv = np.array([0.2, 0.5, 0.8], np.float32)
    up = lambda a: np.kron(a, np.ones((4, 4), np.float32))[:h, :w]  # noqa: E731
    cp = np.stack([up(lv[rng.integers(0, 3, ((h + 3) // 4, (w + 3) // 4))]) for _ in range(C)])
    sp = np.stack([up(lv[rng.integers(0, 3, ((h + 3) // 4, (w + 3) // 4))]) for _ in range(K)])
    return synth.clip_probs(cp), synth.clip_probs(sp), C, synth.generate_offsets(40, K)


def tie_exact_cases():
    """(name, class_pred, adj_pred, C, offsets, opts) -- the fixtures of tests/golden/exact (raw reference outputs)."""
    out = []
    out.append(("quant_17x23", ) + quantized(17, 23, 0) + (PLAIN_OPTS,))
    out.append(("quant_24x32_a", ) + quantized(24, 32, 1) + (PLAIN_OPTS,))
    out.append(("quant_24x32_b", ) + quantized(24, 32, 2) + (RECIPE_OPTS,))
    out.append(("quant_32x32_C4_K16", ) + quantized(32, 32, 3, C=4, K=16) + (PLAIN_OPTS,))
    out.append(("quant_sdb_20x28", ) + quantized(20, 28, 4) + ((0.5, 1.0, 0.03),))
    out.append(("quant_omf_40x56", ) + quantized(40, 56, 5, C=5, K=10) + (QUARTER_OPTS,))
    out.append(("quant_96x128", ) + quantized(96, 128, 6, C=5, K=10) + (PLAIN_OPTS,))
    out.append(("city_oracle_48x64", ) + cityscapes_like(48, 64, 1, False) + (RECIPE_OPTS,))
    out.append(("smooth_40x56", ) + smooth(40, 56, 3) + (QUARTER_OPTS,))
    out.append(("coco_soft_44x52", ) + coco_like(44, 52, 5, True) + (RECIPE_OPTS,))
    out.append(("tiny_1x7", ) + cityscapes_like(1, 7, 6, True, n_shapes=1, rmax=3) + (RECIPE_OPTS,))
    out.append(("one_pixel", ) + quantized(1, 1, 7) + (PLAIN_OPTS,))
    return out


def layout_regression_cases():
    """(name, class_pred, adj_pred, C, offsets, opts): inputs that broke the workspace layout until the end of round 2
    (found by tests/manual/soak_sweep.py; see tests/test_emul_scheduler.py for what each one hit)."""
    out = []
    # (1) own queue chunks on top of the last E % 128 initial keys: one small offset, every priority >= 0, E > 1024
    for (h, w, offs, seed) in [(31, 53, [(0, 1)], 1), (40, 47, [(0, 1)], 2), (30, 34, [(1, 0), (0, 1)], 3)]:
        rng = np.random.default_rng(seed)
        cp = rng.random((3, h, w)).astype(np.float32) * 0.2 + 0.1
        cp[1] += 0.5                                                      # one dominant class: class deltas are 0
        sp = rng.random((len(offs), h, w)).astype(np.float32) * 0.45 + 0.52  # "same" everywhere: every priority > 0
        out.append(("keys_tail_%dx%d" % (h, w), synth.clip_probs(cp), synth.clip_probs(sp), 3, offs, (0.0, 1.0, 0.2)))
    # (2) pixel pool: two blobs of 2^k + few pixels each that merge last (capacities 2^(k+1) + 2^(k+1) -> 2^(k+2))
    h, w = 2, 2 * 1030
    offs = [(0, 1), (1, 0), (1, 1)]
    cp = np.full((2, h, w), 0.5, np.float32)
    cp[1] = 0.6
    sp = np.full((3, h, w), 0.9, np.float32)
    rng = np.random.default_rng(5)
    sp += (rng.random(sp.shape).astype(np.float32) - 0.5) * 0.05   # no ties
    for k, (dr, dc) in enumerate(offs):                            # pairs across the middle are "different"
        sp[k, :, 1030 - dc:1030] = 0.2
    out.append(("two_halves_2x2060", synth.clip_probs(cp), synth.clip_probs(sp), 2, offs, (0.0, 1.0, 0.5)))
    return out


def small_cases():
    """(name, class_pred, adj_pred, C, offsets) -- the parity matrix."""
    out = []
    out.append(("city_soft_48x64", ) + cityscapes_like(48, 64, 0, True))
    out.append(("city_oracle_48x64", ) + cityscapes_like(48, 64, 1, False))
    out.append(("city_soft_odd_45x67", ) + cityscapes_like(45, 67, 2, True))
    out.append(("smooth_40x56", ) + smooth(40, 56, 3))
    out.append(("coco_oracle_48x48", ) + coco_like(48, 48, 4, False))
    out.append(("coco_soft_44x52", ) + coco_like(44, 52, 5, True))
    out.append(("tiny_1x7", ) + cityscapes_like(1, 7, 6, True, n_shapes=1, rmax=3))
    out.append(("tiny_3x3", ) + cityscapes_like(3, 3, 7, True, n_shapes=1, rmax=3))
    return out


def medium_cases():
    out = []
    out.append(("city_soft_128x192", ) + cityscapes_like(128, 192, 10, True, rmax=30))
    out.append(("city_oracle_96x160", ) + cityscapes_like(96, 160, 11, False, rmax=30))
    out.append(("smooth_96x128", ) + smooth(96, 128, 12))
    return out


def pathological_cases():
    """Degenerate shapes and maps: 1x1 and one-row images, one class, one or sixteen offsets, constant 0.5
    maps (every priority equal), all-same and all-different sameness, uniform random maps."""
    rng = np.random.default_rng(0)
    out = []
    for (h, w, C, K) in [(1, 1, 2, 1), (1, 2, 2, 1), (2, 1, 3, 2), (5, 5, 1, 1), (8, 8, 2, 16), (16, 16, 9, 10),
                         (32, 32, 9, 10), (3, 50, 4, 3)]:
        offs = [(1, 0)] if K == 1 else synth.generate_offsets(40, K)
        tag = "%dx%d_C%d_K%d" % (h, w, C, K)
        f = lambda a: synth.clip_probs(np.asarray(a, np.float32))  # noqa: E731
        out.append(("const_" + tag, f(np.full((C, h, w), 0.5)), f(np.full((K, h, w), 0.5)), C, offs))
        out.append(("allsame_" + tag, f(np.full((C, h, w), 0.9)), f(np.full((K, h, w), 1.0)), C, offs))
        out.append(("alldiff_" + tag, f(rng.random((C, h, w))), f(np.full((K, h, w), 0.0)), C, offs))
        out.append(("random_" + tag, f(rng.random((C, h, w))), f(rng.random((K, h, w))), C, offs))
    return out


def same_result(oracle, a, b):
    ca = oracle.canonical_result(*a)
    cb = oracle.canonical_result(*b)
    return np.array_equal(ca[0], cb[0]) and list(ca[1]) == list(cb[1])
