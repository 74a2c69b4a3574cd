"""The step after the path: nearest-neighbour mask resize and COCO RLE.
CPU: the oracle restatements against cv2 (present in this image) and against themselves (encode ->
decode round trip, hand-checked strings).  GPU: mergenet_b200.post against the oracle, bit for bit."""
import numpy as np
import pytest


def _label_mask(h, w, n, seed):
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), np.int32)
    for i in range(1, n + 1):
        r, c = int(rng.integers(0, h)), int(rng.integers(0, w))
        rh, rw = int(rng.integers(1, max(2, h // 4))), int(rng.integers(1, max(2, w // 4)))
        sub = m[r:r + rh, c:c + rw]
        sub[sub == 0] = i
    return m


SIZES = [(7, 9, 20, 31), (33, 57, 100, 99), (64, 128, 1024, 2048), (10, 10, 3, 4), (50, 70, 1023, 2047), (5, 1, 1, 7)]


def test_oracle_resize_nearest_equals_cv2(oracle_mod):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (h, w, oh, ow) in SIZES:
        a = rng.integers(0, 1000, (h, w)).astype(np.int32)
        want = cv2.resize(a, (ow, oh), interpolation=cv2.INTER_NEAREST)
        assert np.array_equal(oracle_mod.oracle_resize_nearest(a, oh, ow), want.reshape(oh, ow)), (h, w, oh, ow)


def test_oracle_coco_rle_hand_checked_strings(oracle_mod):
    m = np.zeros((4, 5), np.int32)
    m[1:3, 1:3] = 1          # column-major runs 5,2,2,2,9 -> '5','2','2', 2-2 -> '0', 9-2 -> '7'
    m[3, 4] = 2              # runs 19,1 -> 19 = 0b10011: sign bit set, needs a second char: 'c','0'; then '1'
    assert oracle_mod.oracle_coco_rle(m, 2) == [b"52207", b"c01"]
    z = np.zeros((3, 3), np.int32)
    assert oracle_mod.oracle_coco_rle(z, 1) == [b"9"]          # an instance without pixels: the single count 9
    f = np.ones((3, 3), np.int32)
    assert oracle_mod.oracle_coco_rle(f, 1) == [b"09"]         # starts with a zero run of length 0


def test_oracle_coco_rle_round_trip(oracle_mod):
    for (h, w, n, seed) in [(17, 23, 5, 1), (64, 48, 40, 2), (128, 200, 150, 3)]:
        m = _label_mask(h, w, n, seed)
        s = oracle_mod.oracle_coco_rle(m, n)
        assert np.array_equal(oracle_mod.oracle_coco_rle_decode(s, h, w), m)


@pytest.mark.gpu
def test_gpu_resize_nearest_equals_oracle(oracle_mod, lib_mod):
    from mergenet_b200 import post
    rng = np.random.default_rng(1)
    for (h, w, oh, ow) in SIZES + [(256, 512, 1024, 2048)]:
        a = rng.integers(0, 1000, (2, h, w)).astype(np.int32)
        got = post.resize_masks_nearest(a, oh, ow)
        for b in range(2):
            assert np.array_equal(got[b], oracle_mod.oracle_resize_nearest(a[b], oh, ow)), (h, w, oh, ow)


@pytest.mark.gpu
def test_gpu_coco_rle_equals_oracle(oracle_mod, lib_mod):
    from mergenet_b200 import post
    cases = [(4, 5, 2, 0), (17, 23, 5, 1), (64, 48, 40, 2), (128, 200, 150, 3), (1, 9, 3, 4), (9, 1, 3, 5), (300, 257, 700, 6)]
    for (h, w, n, seed) in cases:
        m = _label_mask(h, w, n, seed)
        assert post.coco_rle_counts(m, n) == oracle_mod.oracle_coco_rle(m, n), (h, w, n)
    z = np.zeros((6, 7), np.int32)
    assert post.coco_rle_counts(z, 3) == oracle_mod.oracle_coco_rle(z, 3) == [b"Z1"] * 3   # 42 = the single count
    f = np.full((6, 7), 2, np.int32)   # instance 1 has no pixels, instance 2 all of them
    assert post.coco_rle_counts(f, 2) == oracle_mod.oracle_coco_rle(f, 2)
    assert post.coco_rle_counts(f, 0) == []                       # no instances asked for
    g = _label_mask(20, 30, 9, 8)
    assert post.coco_rle_counts(g, 4) == oracle_mod.oracle_coco_rle(g, 4)   # labels above n belong to no instance
    assert post.convert_to_coco_result(np.zeros((5, 5), np.int32), [], 1, [0]) == []


@pytest.mark.gpu
def test_gpu_post_pass_on_a_segmentation_full_size_round_trip(oracle_mod, lib_mod):
    """Segment a 128x192 image on the GPU, resize the mask to 1024x2048 like the Cityscapes recipe and
    encode it: the strings must decode (oracle rleFrString/rleDecode) to the resized mask, and the
    reference-shaped result dicts must carry them."""
    import cases
    from mergenet_b200 import c_segment, post
    name, cp, sp, C, offs = cases.medium_cases()[0]
    mask, ocls = c_segment.run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
    big = post.resize_masks_nearest(mask, 1024, 2048)
    assert np.array_equal(big, oracle_mod.oracle_resize_nearest(mask, 1024, 2048))
    n = int(big.max())
    res = post.convert_to_coco_result(big, ocls, 7, list(range(100, 100 + C)))
    assert len(res) == n and all(r["image_id"] == 7 and r["segmentation"]["size"] == [1024, 2048] for r in res)
    strings = [r["segmentation"]["counts"] for r in res]
    assert np.array_equal(oracle_mod.oracle_coco_rle_decode(strings, 1024, 2048), big)
    assert [r["category_id"] for r in res] == [100 + c for c in ocls[:n]]


BILINEAR_SIZES = [(33, 47, 100, 64), (31, 47, 93, 47), (64, 64, 17, 23), (5, 7, 50, 70), (50, 70, 5, 7), (1, 9, 4, 30),
                  (9, 1, 30, 4), (40, 80, 80, 160)]


def test_oracle_resize_bilinear_equals_cv2(oracle_mod):
    """egs/cityscape/local/segment.py:116-123: the restatement against this image's cv2 itself, BIT FOR BIT, for the
    channel counts the class / offset maps have (OpenCV's generic float path: 2 or >= 5 channels; its 1-, 3- and
    4-channel paths round differently and are not claimed).  Tolerance: none (float bits equal)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for cn in (2, 5, 9, 10, 16, 81):
        for (h, w, oh, ow) in BILINEAR_SIZES:
            src = rng.random((cn, h, w)).astype(np.float32)
            want = np.moveaxis(cv2.resize(np.moveaxis(src, 0, -1), (ow, oh)).reshape(oh, ow, cn), -1, 0)
            got = oracle_mod.oracle_resize_bilinear(src, oh, ow)
            assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want).view(np.uint32)), (cn, h, w, oh, ow)


@pytest.mark.gpu
def test_gpu_resize_bilinear_equals_oracle_and_cv2(oracle_mod, lib_mod):
    from mergenet_b200 import post
    rng = np.random.default_rng(2)
    for cn, (h, w, oh, ow) in [(9, s) for s in BILINEAR_SIZES] + [(10, (128, 256, 256, 512)), (81, (40, 40, 64, 64)),
                                                                    (16, (64, 64, 40, 40))]:
        src = rng.random((cn, h, w)).astype(np.float32)
        got = post.resize_maps_bilinear(src, oh, ow)
        assert np.array_equal(got.view(np.uint32), oracle_mod.oracle_resize_bilinear(src, oh, ow).view(np.uint32)), (cn, h, w, oh, ow)
    try:
        import cv2
    except ImportError:
        return
    src = rng.random((10, 512, 1024)).astype(np.float32)  # the Cityscapes recipe: maps at half size -> 1024 x 2048
    want = np.moveaxis(cv2.resize(np.moveaxis(src, 0, -1), (2048, 1024)), -1, 0)
    got = post.resize_maps_bilinear(src, 1024, 2048)
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))


@pytest.mark.gpu
def test_driver_loop_body_end_to_end(oracle_mod, lib_mod):
    """The body of egs/cityscape/local/segment.py:112-186 with every step on the GPU -- maps resized to the
    segmentation size (cv2.resize, :116-123), c_segment.run_segmentation (:138-143), mask back at the image size
    (INTER_NEAREST, :147-149), COCO results (:165-186) -- against the same chain built from the checkers
    (cv2 / the oracle restatements): identical masks after canonical relabel, identical classes, identical RLE."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    import cases
    from mergenet_b200 import c_segment, post, synth
    cp, sp, C, offs = cases.cityscapes_like(48, 96, 31, True, rmax=20)   # network output at half size
    seg_h, seg_w, img_h, img_w = 96, 192, 150, 300
    cls_r = synth.clip_probs(post.resize_maps_bilinear(cp, seg_h, seg_w))
    adj_r = synth.clip_probs(post.resize_maps_bilinear(sp, seg_h, seg_w))
    mask, ocls = c_segment.run_segmentation(cls_r, adj_r, C, offs, *cases.RECIPE_OPTS)
    big = post.resize_masks_nearest(mask, img_h, img_w)
    res = post.convert_to_coco_result(big, ocls, 7, list(range(100, 100 + C)))
    # the checkers' chain
    cls_o = synth.clip_probs(oracle_mod.oracle_resize_bilinear(cp, seg_h, seg_w))
    adj_o = synth.clip_probs(oracle_mod.oracle_resize_bilinear(sp, seg_h, seg_w))
    assert np.array_equal(cls_r.view(np.uint32), cls_o.view(np.uint32)) and np.array_equal(adj_r.view(np.uint32), adj_o.view(np.uint32))
    m0, c0, _ = oracle_mod.oracle_run_segmentation(cls_o, adj_o, C, offs, *cases.RECIPE_OPTS)
    assert cases.same_result(oracle_mod, (m0, c0), (mask, ocls))
    assert np.array_equal(big, oracle_mod.oracle_resize_nearest(mask, img_h, img_w))
    want = oracle_mod.oracle_coco_rle(big, len(ocls))
    assert len(res) == len(ocls) and len(ocls) > 0
    for i, r in enumerate(res):
        assert r["segmentation"]["counts"] == want[i] and r["segmentation"]["size"] == [img_h, img_w]
        assert r["category_id"] == 100 + ocls[i] and r["image_id"] == 7
