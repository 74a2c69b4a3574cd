"""Synthetic inputs of the shapes BASELINE.json names (there is no dataset / checkpoint here).

Everything is numpy on the host and deterministic in its seed.  Recipes follow SURVEY.md 8(d):

* ``generate_offsets``      -- the reference's log-spiral offset list
                               (/root/reference/utils/train_utils.py:317-328), restated.
* ``gt_instance_mask``      -- random occluding ellipses / rectangles, earlier shapes on top
                               (mirrors the occlusion rule of utils/dataset.py:501).
* ``oracle_mode_maps``      -- near-binary maps from a GT mask, sameness target per
                               utils/dataset.py:259-277 (roll rule, out-of-image rows/cols = 1).
* ``soft_maps``             -- p = sigmoid(3*(2t-1) + N(0,1)) on the same targets.
* ``smooth_random_maps``    -- low-contrast maps around 0.5 (stand-in for a random-init CNN, cfg1).
"""
import math

import numpy as np

EPS = float(np.finfo(np.float32).eps)


def generate_offsets(max_offset=20, num_offsets=10):
    """Log-spiral offsets, (row delta, col delta) tuples (train_utils.py:317-328)."""
    out = []
    angle = math.pi * 5 / 9
    tri = max(abs(math.cos((num_offsets - 1) * angle)), abs(math.sin((num_offsets - 1) * angle)))
    base = abs(max_offset / tri)
    ratio = math.pow(base, 1 / float(num_offsets - 1))
    for n in range(num_offsets):
        x = int(round(math.cos(n * angle) * math.pow(ratio, n)))
        y = int(round(math.sin(n * angle) * math.pow(ratio, n)))
        out.append((x, y))
    return out


def clip_probs(p):
    """The wrapper's clip (c_segment.pyx:53-55): [eps, 1-eps], eps = 2**-23, stays float32."""
    return np.ascontiguousarray(np.clip(p, EPS, 1.0 - EPS).astype(np.float32))


def gt_instance_mask(h, w, n_shapes, rmax, num_classes, seed):
    """Instance-id mask (0 = background) and per-instance class (index 0 unused -> class 0)."""
    rng = np.random.default_rng(seed)
    mask = np.zeros((h, w), dtype=np.int32)
    classes = [0]
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(n_shapes):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        ry, rx = rng.uniform(3, rmax), rng.uniform(3, rmax)
        y0, y1 = max(0, int(cy - ry)), min(h, int(cy + ry) + 1)
        x0, x1 = max(0, int(cx - rx)), min(w, int(cx + rx) + 1)
        if y0 >= y1 or x0 >= x1:
            classes.append(int(rng.integers(1, num_classes)))
            continue
        sub_y, sub_x = yy[y0:y1, x0:x1], xx[y0:y1, x0:x1]
        if rng.random() < 0.5:
            inside = ((sub_y - cy) / ry) ** 2 + ((sub_x - cx) / rx) ** 2 <= 1.0
        else:
            inside = np.ones_like(sub_y, dtype=bool)
        region = mask[y0:y1, x0:x1]
        region[inside & (region == 0)] = i + 1  # only where empty: earlier shapes occlude later
        classes.append(int(rng.integers(1, num_classes)))
    return mask, np.asarray(classes, dtype=np.int32)


def sameness_targets(mask, offsets):
    """Binary same/different targets (K,H,W) from an instance mask (dataset.py:259-277)."""
    h, w = mask.shape
    t = np.zeros((len(offsets), h, w), dtype=bool)
    for n, (i, j) in enumerate(offsets):
        rolled = np.roll(np.roll(mask, -i, axis=0), -j, axis=1)
        t[n] = rolled == mask
        if i < 0:
            t[n, :-i, :] = 1
        elif i > 0:
            t[n, -i:, :] = 1
        if j < 0:
            t[n, :, :-j] = 1
        elif j > 0:
            t[n, :, -j:] = 1
    return t


def class_targets(mask, classes, num_classes):
    cls_map = classes[mask]
    t = np.zeros((num_classes,) + mask.shape, dtype=bool)
    for c in range(num_classes):
        t[c] = cls_map == c
    return t


def oracle_mode_maps(mask, classes, num_classes, offsets):
    """Near-binary (clipped one-hot) class map (C,H,W) and sameness map (K,H,W), float32."""
    ct = class_targets(mask, classes, num_classes).astype(np.float32)
    st = sameness_targets(mask, offsets).astype(np.float32)
    return clip_probs(ct), clip_probs(st)


def soft_maps(mask, classes, num_classes, offsets, seed, sharp=3.0, noise=1.0):
    """p = sigmoid(sharp*(2t-1) + noise*N(0,1)), clipped, float32."""
    rng = np.random.default_rng(seed)
    ct = class_targets(mask, classes, num_classes).astype(np.float32)
    st = sameness_targets(mask, offsets).astype(np.float32)

    def soften(t):
        z = sharp * (2.0 * t - 1.0) + noise * rng.standard_normal(t.shape, dtype=np.float32)
        return (1.0 / (1.0 + np.exp(-z))).astype(np.float32)

    return clip_probs(soften(ct)), clip_probs(soften(st))


def smooth_random_maps(h, w, num_classes, num_offsets, seed, spread=0.03):
    """Low-contrast maps in ~[0.47, 0.53] with spatially smooth structure (cfg1 stand-in: a
    random-init UNet emits exactly this kind of map; the reference UNet itself cannot travel)."""
    rng = np.random.default_rng(seed)

    def field(n):
        coarse = rng.standard_normal((n, (h + 7) // 8 + 1, (w + 7) // 8 + 1)).astype(np.float32)
        up = np.repeat(np.repeat(coarse, 8, axis=1), 8, axis=2)[:, :h, :w]
        fine = rng.standard_normal((n, h, w)).astype(np.float32)
        return 0.5 + spread * (0.6 * up + 0.4 * fine) / 1.5

    return clip_probs(field(num_classes)), clip_probs(field(num_offsets))


# ---- the named configurations of BASELINE.json ---------------------------------
def cfg_cityscapes(h=1024, w=2048, seed=2, n_shapes=400, rmax=120, soft=True, noise_seed=7):
    """cfg2 (soft) / cfg3 (oracle mode): C=9, K=10, offsets generate_offsets(40,10)."""
    offsets = generate_offsets(40, 10)
    mask, classes = gt_instance_mask(h, w, n_shapes, rmax, 9, seed)
    if soft:
        cp, sp = soft_maps(mask, classes, 9, offsets, noise_seed)
    else:
        cp, sp = oracle_mode_maps(mask, classes, 9, offsets)
    return cp, sp, offsets, mask


def cfg_coco(h=512, w=512, seed=3, n_shapes=900, rmax=14, soft=False, noise_seed=11):
    """cfg4: C=81, K=16, offsets generate_offsets(40,16), many small instances."""
    offsets = generate_offsets(40, 16)
    mask, classes = gt_instance_mask(h, w, n_shapes, rmax, 81, seed)
    if soft:
        cp, sp = soft_maps(mask, classes, 81, offsets, noise_seed)
    else:
        cp, sp = oracle_mode_maps(mask, classes, 81, offsets)
    return cp, sp, offsets, mask
