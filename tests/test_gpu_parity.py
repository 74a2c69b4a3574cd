"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs, against the committed golden fixtures (outputs of the reference itself), and through
size-independent invariants.  Bit-exact: masks after canonical relabel, per-instance classes."""
import ctypes
import glob
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _edge_dump(lib_mod, cp, sp, C, offs, opts):
    L = lib_mod.lib()
    F = ctypes.POINTER(ctypes.c_float); I = ctypes.POINTER(ctypes.c_int)
    K, H, W = sp.shape
    N = H * W
    off = np.ascontiguousarray(np.array(offs, np.int32))
    sp = sp.copy()
    d = dict(clp=np.zeros((N, C), np.float32), cls=np.zeros(N, np.int32), same=np.zeros(N * K, np.float32),
             diff=np.zeros(N * K, np.float32), oml=np.zeros(N * K, np.float32), mp=np.zeros(N * K, np.float32),
             lo=np.zeros(N * K, np.int32), hi=np.zeros(N * K, np.int32))
    st = L.mn_debug_edge_dump(H, W, C, K, off.ctypes.data_as(I), cp.ctypes.data_as(F), sp.ctypes.data_as(F),
                              *[float(o) for o in opts], d["clp"].ctypes.data_as(F), d["cls"].ctypes.data_as(I),
                              d["same"].ctypes.data_as(F), d["diff"].ctypes.data_as(F), d["oml"].ctypes.data_as(F),
                              d["mp"].ctypes.data_as(F), d["lo"].ctypes.data_as(I), d["hi"].ctypes.data_as(I))
    assert st == 0
    d["adj"] = sp
    return d


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("opts", [cases.RECIPE_OPTS, cases.QUARTER_OPTS, (0.5, 1.0, 0.03)])
def test_edge_pass_bitwise_equals_reference_constructor(oracle_mod, lib_mod, opts):
    """segment.cc:153-232: clp, cls, same, diff, oml, initial mp per record -- every bit."""
    for name, cp, sp, C, offs in cases.small_cases() + cases.medium_cases()[:1]:
        ref = oracle_mod.oracle_init_dump(cp, sp, C, offs, *opts)
        got = _edge_dump(lib_mod, cp, sp, C, offs, opts)
        valid = ref["valid"].astype(bool)
        assert np.array_equal(valid, got["lo"] >= 0), name
        assert np.array_equal(_bits(ref["clp"]), _bits(got["clp"])), name
        assert np.array_equal(ref["cls"], got["cls"]), name
        for k in ("same", "diff", "oml", "mp"):
            assert np.array_equal(_bits(ref[k])[valid], _bits(got[k])[valid]), (name, k)
        if opts[0] != 0:  # in-place rewrite of the caller's sameness buffer (segment.cc:187-191)
            assert np.array_equal(_bits(ref["adj_pred"]), _bits(got["adj"])), name


def test_edge_pass_bitwise_at_full_resolution(oracle_mod, lib_mod):
    """The warp-pipeline edge kernel over a whole 1024x2048 cfg2 image (9363 tiles of 224 pixels, ragged last
    tile) against the oracle's constructor: every clp / same / diff / oml / initial-priority bit."""
    from mergenet_b200 import synth
    cp, sp, offs, _ = synth.cfg_cityscapes(1024, 2048, seed=1003, n_shapes=400, rmax=120, soft=True, noise_seed=11)
    ref = oracle_mod.oracle_init_dump(cp, sp, 9, offs, *cases.RECIPE_OPTS)
    got = _edge_dump(lib_mod, cp, sp, 9, offs, cases.RECIPE_OPTS)
    valid = ref["valid"].astype(bool)
    assert np.array_equal(valid, got["lo"] >= 0)
    assert np.array_equal(_bits(ref["clp"]), _bits(got["clp"]))
    assert np.array_equal(ref["cls"], got["cls"])
    for k in ("same", "diff", "oml", "mp"):
        assert np.array_equal(_bits(ref[k])[valid], _bits(got[k])[valid]), k


def test_edge_kernels_agree_bitwise_incl_unclipped_input(lib_mod, monkeypatch):
    """The warp-pipeline edge kernel (fast path) and the tile kernel (general path) must produce the
    same bits, also when the caller did not clip (values below 2^-23 / equal to 1 - 2^-24 leave the
    (k, i) table and take the generic recipe) and on a ragged last tile."""
    rng = np.random.default_rng(5)
    for (h, w) in ((64, 100), (37, 64)):
        name, cp, sp, C, offs = ("x", ) + cases.cityscapes_like(h, w, 21, True)
        cp = cp.copy(); sp = sp.copy()
        for arr in (cp, sp):
            flat = arr.reshape(-1)
            idx = rng.choice(flat.size, flat.size // 50, replace=False)
            flat[idx[0::3]] = np.float32(1e-12)
            flat[idx[1::3]] = np.float32(3.0e-8)
            flat[idx[2::3]] = np.nextafter(np.float32(1.0), np.float32(0.0))
        monkeypatch.setenv("MN_EDGE2_NCONS", "0")
        a = _edge_dump(lib_mod, cp, sp, C, offs, cases.PLAIN_OPTS)
        monkeypatch.delenv("MN_EDGE2_NCONS")
        b = _edge_dump(lib_mod, cp, sp, C, offs, cases.PLAIN_OPTS)
        for k in ("clp", "same", "diff", "oml", "mp"):
            assert np.array_equal(_bits(a[k]), _bits(b[k])), (h, w, k)
        assert np.array_equal(a["cls"], b["cls"])


@pytest.mark.parametrize("opts", [cases.RECIPE_OPTS, cases.PLAIN_OPTS, cases.QUARTER_OPTS])
def test_drop_in_c_abi_matches_oracle_small(oracle_mod, lib_mod, opts):
    from mergenet_b200 import c_segment
    for name, cp, sp, C, offs in cases.small_cases():
        m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
        m1, c1 = c_segment.run_segmentation(cp, sp, C, offs, *opts)
        assert m1.dtype == np.int32 and m1.shape == m0.shape
        assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), name
        lp0 = oracle_mod.total_logprob_from_scratch(m0, c0, cp, sp, offs, opts[1])
        lp1 = oracle_mod.total_logprob_from_scratch(m1, c1, cp, sp, offs, opts[1])
        assert abs(lp0 - lp1) <= 1e-5 * abs(lp0), name  # north-star tolerance (identical partitions)


def test_drop_in_matches_oracle_medium(oracle_mod, lib_mod):
    from mergenet_b200 import c_segment
    for name, cp, sp, C, offs in cases.medium_cases():
        m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
        m1, c1 = c_segment.run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
        assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), name


def test_degenerate_inputs(oracle_mod, lib_mod):
    """1x1 / one-row images, one class, one or sixteen offsets, constant maps (all priorities tie),
    all-same / all-different sameness: the drop-in entry against the oracle."""
    from mergenet_b200 import c_segment
    for name, cp, sp, C, offs in cases.pathological_cases():
        for opts in (cases.RECIPE_OPTS, cases.PLAIN_OPTS):
            m0, c0, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
            m1, c1 = c_segment.run_segmentation(cp, sp, C, offs, *opts)
            assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), (name, opts)


def test_same_different_bias_path(oracle_mod, lib_mod):
    from mergenet_b200 import c_segment
    name, cp, sp, C, offs = cases.small_cases()[0]
    for sdb in (0.5, -0.7):
        m0, c0, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, sdb, 1.0, 0.0)
        m1, c1 = c_segment.run_segmentation(cp, sp, C, offs, sdb, 1.0, 0.0)
        assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), (name, sdb)


def test_golden_fixtures_from_the_reference(oracle_mod, lib_mod):
    from mergenet_b200 import c_segment
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
    assert files
    for f in files:
        g = np.load(f)
        offs = [tuple(int(v) for v in o) for o in g["offsets"]]
        opts = tuple(float(v) for v in g["opts"])
        m1, c1 = c_segment.run_segmentation(np.ascontiguousarray(g["class_pred"]), np.ascontiguousarray(g["adj_pred"]),
                                            int(g["num_classes"]), offs, *opts)
        ref = (g["ref_mask"].astype(np.int32), [int(v) for v in g["ref_object_class"]])
        assert cases.same_result(oracle_mod, ref, (m1, c1)), os.path.basename(f)


def test_python_class_facade(oracle_mod, lib_mod):
    from mergenet_b200 import ObjectSegmenter, SegmenterOptions
    name, cp, sp, C, offs = cases.small_cases()[1]
    opts = SegmenterOptions(0.0, 1.0, 0.03)
    m1, c1 = ObjectSegmenter(cp, sp, C, offs, opts, mode="csegment").run_segmentation()
    m0, c0, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
    assert cases.same_result(oracle_mod, (m0, c0), (m1, c1))


def test_batch_api_host_and_device_agree_and_are_deterministic(oracle_mod, lib_mod):
    import torch
    from mergenet_b200 import BatchSegmenter, SegmenterOptions
    H, W = 48, 64
    items = [cases.cityscapes_like(H, W, s, s % 2 == 0) for s in range(5)]
    C, offs = items[0][2], items[0][3]
    cp = np.ascontiguousarray(np.stack([it[0] for it in items]))
    sp = np.ascontiguousarray(np.stack([it[1] for it in items]))
    opts = SegmenterOptions(*cases.RECIPE_OPTS)
    seg = BatchSegmenter(8, H, W, C, offs)
    m_h, oc_h, n_h = seg.segment_host(cp, sp, opts)
    stats = [seg.stats(b) for b in range(5)]
    m_d, oc_d, n_d = seg.segment_device(torch.from_numpy(cp).cuda(), torch.from_numpy(sp).cuda(), opts)
    m_d2, oc_d2, n_d2 = seg.segment_device(torch.from_numpy(cp).cuda(), torch.from_numpy(sp).cuda(), opts)
    assert np.array_equal(m_h, m_d.cpu().numpy()) and np.array_equal(oc_h, oc_d.cpu().numpy())
    assert torch.equal(m_d, m_d2) and torch.equal(oc_d, oc_d2) and torch.equal(n_d, n_d2)
    for b, it in enumerate(items):
        m0, c0, st0 = oracle_mod.oracle_run_segmentation(it[0], it[1], C, offs, *cases.RECIPE_OPTS)
        got = (m_h[b], [int(v) for v in oc_h[b][:n_h[b]]])
        assert cases.same_result(oracle_mod, (m0, c0), got), b
        assert stats[b]["merges"] == st0["merges"], b
        assert stats[b]["status"] == 0
    seg.close()


def test_host_entry_chunked_upload_equals_device_entry(oracle_mod, lib_mod):
    """mn_segment_batch_host uploads in chunks of 8 images on a second stream while the front end of the
    previous chunk runs: 21 images = chunks of 8, 8, 5; results must equal the device entry's."""
    import torch
    from mergenet_b200 import BatchSegmenter, SegmenterOptions
    H, W = 32, 48
    items = [cases.cityscapes_like(H, W, 40 + s, s % 3 != 0) for s in range(7)]
    C, offs = items[0][2], items[0][3]
    cp = np.ascontiguousarray(np.stack([items[i % 7][0] for i in range(21)]))
    sp = np.ascontiguousarray(np.stack([items[i % 7][1] for i in range(21)]))
    opts = SegmenterOptions(*cases.PLAIN_OPTS)
    seg = BatchSegmenter(21, H, W, C, offs)
    for _ in range(2):  # (the second call reuses the staging buffers and the copy stream)
        m_h, oc_h, n_h = seg.segment_host(cp, sp, opts)
        m_d, oc_d, n_d = seg.segment_device(torch.from_numpy(cp).cuda(), torch.from_numpy(sp).cuda(), opts)
        assert np.array_equal(m_h, m_d.cpu().numpy()) and np.array_equal(oc_h, oc_d.cpu().numpy())
        assert np.array_equal(n_h, n_d.cpu().numpy())
    for b in range(7):
        assert np.array_equal(m_h[b], m_h[b + 7]) and np.array_equal(m_h[b], m_h[b + 14])
    seg.close()


def test_logits_input_equals_torch_sigmoid_then_segment(oracle_mod, lib_mod):
    """MN_INPUT_LOGITS: the edge pass applies F.sigmoid + the wrapper's clip while it reads the maps
    (utils/inference_utils.py:43-44,95-96; c_segment.pyx:53-55).  Same masks as sigmoid on the device
    with torch, then the plain path -- on a shape served by the warp-pipeline kernel and on one served
    by the tile kernel (odd pixel count), including saturating logits."""
    import torch
    from mergenet_b200 import BatchSegmenter, SegmenterOptions
    opts = SegmenterOptions(*cases.RECIPE_OPTS)
    for (h, w) in ((64, 96), (45, 67)):
        name, cp, sp, C, offs = ("x", ) + cases.cityscapes_like(h, w, 31, True)
        lc = np.log(cp.astype(np.float64) / (1.0 - cp.astype(np.float64))).astype(np.float32)
        ls = np.log(sp.astype(np.float64) / (1.0 - sp.astype(np.float64))).astype(np.float32)
        lc[0, :2, :5] = 40.0; ls[1, :3, :3] = -120.0; ls[2, 5:7, :] = 95.0   # sigmoid saturates: the clip matters
        dlc = torch.from_numpy(lc[None]).cuda(); dls = torch.from_numpy(ls[None]).cuda()
        seg = BatchSegmenter(1, h, w, C, offs)
        m1, c1, n1 = seg.segment_device(dlc, dls, opts, logits=True)
        pc = torch.sigmoid(dlc).contiguous(); ps = torch.sigmoid(dls).contiguous()
        m2, c2, n2 = seg.segment_device(pc, ps, opts, clip=True)
        assert torch.equal(m1, m2) and torch.equal(c1, c2) and torch.equal(n1, n2), (h, w)
        # and against the oracle on the host copy of torch's probabilities
        from mergenet_b200 import synth
        m0, c0, _ = oracle_mod.oracle_run_segmentation(synth.clip_probs(pc[0].cpu().numpy()), synth.clip_probs(ps[0].cpu().numpy()),
                                                       C, offs, *cases.RECIPE_OPTS)
        k = int(n1[0])
        assert cases.same_result(oracle_mod, (m0, c0), (m1[0].cpu().numpy(), list(c1[0, :k].cpu().numpy()))), (h, w)
        seg.close()


def test_total_logprob_partition_pass_matches_oracle(oracle_mod, lib_mod):
    """segment.cc:314-350: the GPU partition-statistics pass (float64, from the maps and the final label mask)
    against the oracle's float64 from-scratch evaluation, term by term (north star: 1e-5 relative; these agree to
    summation order)."""
    from mergenet_b200 import BatchSegmenter, SegmenterOptions
    for name, cp, sp, C, offs in cases.small_cases()[:6] + cases.medium_cases()[:1]:
        H, W = cp.shape[1], cp.shape[2]
        seg = BatchSegmenter(1, H, W, C, offs)
        opts = SegmenterOptions(*cases.RECIPE_OPTS)
        m, oc, n = seg.segment_host(cp[None], sp[None], opts)
        got = seg.total_logprob(0)
        k = int(n[0])
        ref = oracle_mod.total_logprob_terms_from_scratch(m[0], [int(v) for v in oc[0][:k]], cp, sp, offs)
        ref = ref + (ref[0] + cases.RECIPE_OPTS[1] * (ref[1] + ref[2]),)
        for g, r in zip(got, ref):
            assert abs(g - r) <= 1e-9 * max(1.0, abs(r)), (name, got, ref)
        seg.close()


@pytest.mark.gpu
def test_invariants_at_larger_size(oracle_mod, lib_mod):
    """Size-independent properties (no oracle needed): labels are 1..n, every labelled object is one
    4... offset-connected set of pixels, idempotent determinism, and merges = N - surviving objects."""
    from mergenet_b200 import BatchSegmenter, SegmenterOptions
    H, W = 256, 384
    cp, sp, C, offs = cases.cityscapes_like(H, W, 21, True, rmax=40)
    seg = BatchSegmenter(1, H, W, C, offs)
    opts = SegmenterOptions(*cases.RECIPE_OPTS)
    m, oc, n = seg.segment_host(cp[None], sp[None], opts)
    st = seg.stats(0)
    m2, oc2, n2 = seg.segment_host(cp[None], sp[None], opts)
    assert np.array_equal(m, m2) and np.array_equal(oc, oc2)
    n = int(n[0])
    labs = np.unique(m[0])
    assert labs.min() >= 0 and labs.max() == n and len(labs[labs > 0]) == n
    assert np.all(oc[0][:n] > 0) and np.all(oc[0][n:] == -1)
    assert st["status"] == 0 and st["events"] == st["merges"] + st["restores"]
    assert st["merges"] <= H * W - 1
    seg.close()


@pytest.mark.gpu
def test_coco_shaped_medium_matches_oracle(oracle_mod, lib_mod):
    """cfg4 family (81 classes, 16 offsets up to (23,40)): soft and oracle-mode maps at 96x96."""
    from mergenet_b200 import c_segment
    for soft in (True, False):
        cp, sp, C, offs = cases.coco_like(96, 96, 31, soft)
        m0, c0, st0 = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
        m1, c1 = c_segment.run_segmentation(cp, sp, C, offs, *cases.RECIPE_OPTS)
        assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), soft


@pytest.mark.gpu
def test_full_resolution_equals_the_reference(oracle_mod, lib_mod):
    """BASELINE config 2 at its full size: two 1024x2048 cfg2 images (the bench generator, seeds 1000 and
    1001) against tests/golden/full/*.npz = canonical masks and classes computed by the UNMODIFIED compiled
    reference in the build container (tests/golden/make_golden.py full; ~5 min and 7 GB per image there).
    Regression for the refill leaf-limit bug, which only showed at this size."""
    from mergenet_b200 import BatchSegmenter, SegmenterOptions, synth
    h, w = 1024, 2048
    ims = [synth.cfg_cityscapes(h, w, seed=1000 + i, n_shapes=400, rmax=120, soft=True, noise_seed=7 + i) for i in range(2)]
    offs = ims[0][2]
    cp = np.ascontiguousarray(np.stack([im[0] for im in ims])); sp = np.ascontiguousarray(np.stack([im[1] for im in ims]))
    seg = BatchSegmenter(2, h, w, 9, offs)
    m, oc, n = seg.segment_host(cp, sp, SegmenterOptions(*cases.RECIPE_OPTS), clip=False)
    for i in range(2):
        g = np.load(os.path.join(GOLDEN, "full", "cfg2_1024x2048_seed%d.npz" % (1000 + i)))
        cm, cc = oracle_mod.canonical_result(m[i], [int(v) for v in oc[i][:n[i]]])
        assert len(cc) == len(g["cls"]), (i, len(cc), len(g["cls"]))
        assert np.array_equal(cm, g["mask"]) and list(cc) == list(g["cls"]), i
        assert seg.stats(i)["status"] == 0
    seg.close()


def test_full_resolution_properties(oracle_mod, lib_mod):
    """BASELINE.json's full size (1024x2048, C=9, K=10), where the oracle is too slow for the suite:
    size-independent properties only.  Two images in one batch (soft and oracle-mode maps):
    status 0, labels exactly 1..n, every event accounted for, merges = N - surviving objects,
    run twice -> identical bytes (determinism), and the maintained log-prob total is finite and equals
    class + omf * (differentness + sameness)."""
    from mergenet_b200 import BatchSegmenter, SegmenterOptions, synth
    H, W, C = 1024, 2048, 9
    cps, sps = [], []
    for soft in (True, False):
        cp, sp, offs, _ = synth.cfg_cityscapes(H, W, seed=5, n_shapes=400, rmax=120, soft=soft, noise_seed=9)
        cps.append(cp); sps.append(sp)
    cp = np.ascontiguousarray(np.stack(cps)); sp = np.ascontiguousarray(np.stack(sps))
    seg = BatchSegmenter(2, H, W, C, offs)
    opts = SegmenterOptions(*cases.RECIPE_OPTS)
    m, oc, n = seg.segment_host(cp, sp, opts)
    stats = [seg.stats(b) for b in range(2)]
    tot = [seg.total_logprob(b) for b in range(2)]
    m2, oc2, n2 = seg.segment_host(cp, sp, opts)
    assert np.array_equal(m, m2) and np.array_equal(oc, oc2) and np.array_equal(n, n2)
    for b in range(2):
        st = stats[b]
        assert st["status"] == 0
        k = int(n[b])
        labs = np.unique(m[b])
        assert labs.min() >= 0 and labs.max() == k and len(labs[labs > 0]) == k
        assert np.all(oc[b][:k] > 0) and np.all(oc[b][k:] == -1)
        assert st["events"] == st["merges"] + st["restores"]
        # every merge removes one object: surviving objects = N - merges >= instances (+ class-0 objects)
        assert H * W - st["merges"] >= k
        t = tot[b]
        assert np.isfinite(t[3]) and abs(t[3] - (t[0] + opts.object_merge_factor * (t[2] + t[1]))) <= 1e-9 * abs(t[3])
    seg.close()


def _logprob_close(a, b):
    return abs(a - b) <= 1e-5 * abs(b)


@pytest.mark.parametrize("name", ["cfg1_256x512_recipe", "cfg1_256x512_plain", "cfg3_256x512_oracle",
                                  "cfg4_512x512_oracle", "cfg4_512x512_soft", "cfg3_1024x2048_oracle"])
def test_baseline_config_matrix_equals_the_reference(oracle_mod, lib_mod, name):
    """BASELINE.json's configurations at their NAMED sizes against results of the unmodified compiled reference
    (tests/golden/matrix/, made by `python tests/golden/make_golden.py matrix` in the build container):
    cfg1 256x512 from the reference's own UNet (recipe and plain options), cfg3 oracle-mode maps (tie-heavy)
    at 256x512 and 1024x2048, cfg4 512x512 with 81 classes and 16 offsets (oracle-mode and soft).  Three legs:
    identical canonical mask, identical per-instance classes, and the float64 from-scratch log-prob of the
    partition (segment.cc:314-350) within 1e-5 relative of the reference partition's."""
    import matrix_cases
    from mergenet_b200 import c_segment
    cp, sp, C, offs, opts = matrix_cases.load(name)
    g = np.load(os.path.join(GOLDEN, "matrix", name + ".npz"))
    assert tuple(float(v) for v in g["opts"]) == tuple(np.float32(o) for o in opts)
    m1, c1 = c_segment.run_segmentation(cp, sp, C, offs, *opts)
    cm, cc = oracle_mod.canonical_result(m1, c1)
    assert len(cc) == len(g["cls"]), (len(cc), len(g["cls"]))
    assert np.array_equal(cm, g["mask"]) and list(cc) == [int(v) for v in g["cls"]]
    lp = oracle_mod.total_logprob_from_scratch(m1, c1, cp, sp, offs, opts[1])
    assert _logprob_close(lp, float(g["logprob"])), (lp, float(g["logprob"]))


def test_full_resolution_logprob_within_tolerance(oracle_mod, lib_mod):
    """North star, third leg at BASELINE's full size: the float64 from-scratch log-prob of the CUDA partition of
    a 1024x2048 cfg2 image within 1e-5 relative of the reference partition's (stored in tests/golden/full), and
    the scheduler's own maintained total (fp32 sums in merge order, segment.cc:272-287) within 1e-5 of it too."""
    from mergenet_b200 import BatchSegmenter, SegmenterOptions, synth
    h, w = 1024, 2048
    cp, sp, offs, _ = synth.cfg_cityscapes(h, w, seed=1001, n_shapes=400, rmax=120, soft=True, noise_seed=8)
    seg = BatchSegmenter(1, h, w, 9, offs)
    m, oc, n = seg.segment_host(cp[None], sp[None], SegmenterOptions(*cases.RECIPE_OPTS), clip=False)
    g = np.load(os.path.join(GOLDEN, "full", "cfg2_1024x2048_seed1001.npz"))
    lp = oracle_mod.total_logprob_from_scratch(m[0], [int(v) for v in oc[0][:n[0]]], cp, sp, offs, 1.0)
    assert _logprob_close(lp, float(g["logprob"])), (lp, float(g["logprob"]))
    assert _logprob_close(seg.total_logprob(0)[3], float(g["logprob"]))
    seg.close()


def test_saturated_maps_take_libm_special_values(oracle_mod, lib_mod):
    """(a) same_different_bias = 2 on oracle-mode maps rounds every "same" probability to exactly 1.0f
    (log(1 - s) = -inf, priority +inf: cc:183-195,34); (b) a caller of the raw C symbol hands in unclipped class
    maps holding exact 0 and 1 (logf(0) = -inf).  The drop-in entry checks the domain of its fast log recipes and
    lets the general edge kernel, which honours libm's special values, redo such a batch."""
    import test_oracle_vs_reference as t
    from mergenet_b200 import c_segment
    L = lib_mod.lib()
    F = ctypes.POINTER(ctypes.c_float); I = ctypes.POINTER(ctypes.c_int)
    for name, cp, sp, C, offs, opts, clip in t._saturating_cases():
        if clip:
            m0, c0, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
            m1, c1 = c_segment.run_segmentation(cp, sp.copy(), C, offs, *opts)
        else:
            with oracle_mod.raw_abi():
                m0, c0, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
            K, H, W = sp.shape
            cpc = np.ascontiguousarray(cp, np.float32).copy(); spc = np.ascontiguousarray(sp, np.float32).copy()
            off = np.ascontiguousarray(np.array(offs, np.int32))
            m1 = np.zeros((H, W), np.int32); oc = np.zeros((1, H * W), np.int32)
            L.c_run_segmentation(cpc.ctypes.data_as(F), C, spc.ctypes.data_as(F), K, W, H, C, off.ctypes.data_as(I),
                                 m1.ctypes.data_as(I), oc.ctypes.data_as(I), *[ctypes.c_float(o) for o in opts])
            assert L.mn_last_error() == 0
            c1 = oracle_mod._trim(oc)
        assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), name


def test_workspace_layout_regressions(oracle_mod, lib_mod):
    """Inputs that broke the workspace layout until the end of round 2 (cases.layout_regression_cases): the
    scheduler's own queue chunks overlaid the last E % 128 initial keys; a pixel-pool half of 3 N ints was too small
    for two objects of just over 2^k pixels merging.  Through the drop-in C ABI on the real kernels."""
    from mergenet_b200 import c_segment
    for name, cp, sp, C, offs, opts in cases.layout_regression_cases():
        m0, c0, _ = oracle_mod.oracle_run_segmentation(cp, sp, C, offs, *opts)
        m1, c1 = c_segment.run_segmentation(cp, sp, C, offs, *opts)
        assert cases.same_result(oracle_mod, (m0, c0), (m1, c1)), name
