"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the merge-segmenter path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product (``mergenet_b200``) never does.

Two checkers live here:

* ``ref_run_segmentation``    -- the UNMODIFIED reference ``segment.cc`` compiled into
  ``oracle/_ref/libsegment_ref.so`` (``make -C oracle ref``), driven through ctypes with the same
  glue as ``/root/reference/utils/csegment/c_segment.pyx:53-84``.
* ``oracle_run_segmentation`` -- the plain-C restatement ``mergenet_oracle.c`` (``liboracle.so``),
  deterministic tie-break (mp desc, then the (u, D) scatter rule of mn_common.h: mn_tie), with counters and an optional merge log.

Parity pin: the reference has no golden vectors for this path; the restatement is pinned against
the compiled reference itself (tests/test_oracle_vs_reference.py) and against fixtures generated
from it (tests/golden/, script tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libsegment_ref.so")
EPS = float(np.finfo(np.float32).eps)


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_longlong) for n in (
        "pops", "valid_pops", "merges", "repushes", "pushes", "adj_visits", "folds",
        "init_records", "init_pushes", "sum_abs_npix", "max_abs_npix", "merges_abs_gt32",
        "merges_abs_gt1024")] + [("final_objects", ctypes.c_int),
                                           ("final_instances", ctypes.c_int)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    """Compile liboracle.so (always possible: gcc) and, when /root/reference is present,
    oracle/_ref/libsegment_ref.so.  Building the checker is not using it."""
    src = os.path.join(_HERE, "mergenet_oracle.c")
    stale = (not os.path.exists(_ORACLE_SO)) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so", "-B"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/utils/csegment/segment.cc") and (
            force or not os.path.exists(_REF_SO)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_oracle = None
_ref = None
_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


def oracle_lib():
    global _oracle
    if _oracle is None:
        build()
        lib = ctypes.CDLL(_ORACLE_SO)
        lib.mno_run_segmentation.restype = ctypes.c_int
        lib.mno_run_segmentation.argtypes = [
            _F, ctypes.c_int, _F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _I, _I, _I,
            ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.POINTER(Stats), _I,
            ctypes.c_longlong]
        lib.mno_run_segmentation_totals.restype = ctypes.c_int
        lib.mno_run_segmentation_totals.argtypes = [
            _F, ctypes.c_int, _F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _I, _I, _I,
            ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.POINTER(ctypes.c_double)]
        lib.mno_init_dump.restype = ctypes.c_int
        lib.mno_init_dump.argtypes = [
            _F, ctypes.c_int, _F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _I,
            ctypes.c_float, ctypes.c_float, ctypes.c_float, _F, _I, _F, _F, _F, _F,
            ctypes.POINTER(ctypes.c_ubyte)]
        lib.mno_logf_recipe_mismatches.restype = ctypes.c_longlong
        lib.mno_logf_recipe_mismatches.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        for name in ("mno_host_logf_table", "mno_host_log1m_table"):
            fn = getattr(lib, name)
            fn.restype = None
            fn.argtypes = [ctypes.c_uint32, ctypes.c_uint32, _F]
        lib.mno_host_bias_table.restype = None
        lib.mno_host_bias_table.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, _F]
        _oracle = lib
    return _oracle


def have_reference():
    return os.path.exists(_REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            build()
        lib = ctypes.CDLL(_REF_SO)
        lib.c_run_segmentation.restype = None
        lib.c_run_segmentation.argtypes = [
            _F, ctypes.c_int, _F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _I, _I, _I,
            ctypes.c_float, ctypes.c_float, ctypes.c_float]
        _ref = lib
    return _ref


def _fp(a):
    return a.ctypes.data_as(_F)


def _ip(a):
    return a.ctypes.data_as(_I)


_CLIP = True  # tests of the raw C ABI (unclipped maps, what a caller that skips the .pyx glue hands in) switch it off


class raw_abi:
    """with oracle.raw_abi(): ... -- call the checkers like the C symbol itself, without the wrapper's clip."""

    def __enter__(self):
        global _CLIP
        self._saved, _CLIP = _CLIP, False

    def __exit__(self, *a):
        global _CLIP
        _CLIP = self._saved


def _glue(class_pred, adj_pred, offset_list):
    """c_segment.pyx:53-67 -- clip, offsets -> int32, allocate outputs."""
    class_pred = np.asarray(class_pred, dtype=np.float32)
    adj_pred = np.asarray(adj_pred, dtype=np.float32)
    if _CLIP:
        class_pred = class_pred.clip(EPS, 1.0 - EPS)
        adj_pred = adj_pred.clip(EPS, 1.0 - EPS)
    class_pred = np.ascontiguousarray(class_pred).copy()
    adj_pred = np.ascontiguousarray(adj_pred).copy()
    off = np.ascontiguousarray(np.array(offset_list).astype(np.int32))
    k, h, w = adj_pred.shape
    mask = np.zeros((h, w), dtype=np.int32)
    ocls = np.zeros((1, h * w), dtype=np.int32)
    return class_pred, adj_pred, off, mask, ocls


def _trim(ocls):
    """c_segment.pyx:80-84"""
    out = []
    for i in range(ocls.shape[1] - 1):
        if ocls[0, i] == -1:
            break
        out.append(int(ocls[0, i]))
    return out


class _quiet_stdout:
    """The reference prints progress with std::cout (segment.cc:540-569); silence fd 1."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        os.dup2(self._saved, 1)
        os.close(self._saved)
        os.close(self._null)


def ref_run_segmentation(class_pred, adj_pred, num_classes, offset_list, same_different_bias,
                         object_merge_factor, merge_logprob_bias, quiet=True):
    """The reference itself: returns (mask int32[H,W], object_class list)."""
    cp, ap, off, mask, ocls = _glue(class_pred, adj_pred, offset_list)
    lib = ref_lib()
    args = (_fp(cp), cp.shape[0], _fp(ap), ap.shape[0], ap.shape[2], ap.shape[1], int(num_classes),
            _ip(off), _ip(mask), _ip(ocls), float(same_different_bias),
            float(object_merge_factor), float(merge_logprob_bias))
    if quiet:
        with _quiet_stdout():
            lib.c_run_segmentation(*args)
    else:
        lib.c_run_segmentation(*args)
    return mask, _trim(ocls)


def oracle_run_segmentation(class_pred, adj_pred, num_classes, offset_list, same_different_bias,
                            object_merge_factor, merge_logprob_bias, want_merge_log=False):
    """The C restatement: returns (mask, object_class list, stats dict[, merge_log (M,2)])."""
    cp, ap, off, mask, ocls = _glue(class_pred, adj_pred, offset_list)
    st = Stats()
    n = mask.size
    log = np.zeros((n, 2), dtype=np.int32) if want_merge_log else None
    oracle_lib().mno_run_segmentation(
        _fp(cp), cp.shape[0], _fp(ap), ap.shape[0], ap.shape[2], ap.shape[1], int(num_classes),
        _ip(off), _ip(mask), _ip(ocls), float(same_different_bias), float(object_merge_factor),
        float(merge_logprob_bias), ctypes.byref(st), _ip(log) if log is not None else None, n)
    out = (mask, _trim(ocls), st.as_dict())
    if want_merge_log:
        out = out + (log[:st.merges].copy(),)
    return out


def oracle_total_logprob(class_pred, adj_pred, num_classes, offset_list, same_different_bias,
                         object_merge_factor, merge_logprob_bias):
    """segment.cc:272-287 from the oracle's own accumulators: (class term, object sameness term, record
    differentness term, total)."""
    cp, ap, off, mask, ocls = _glue(class_pred, adj_pred, offset_list)
    tot = (ctypes.c_double * 4)()
    oracle_lib().mno_run_segmentation_totals(
        _fp(cp), cp.shape[0], _fp(ap), ap.shape[0], ap.shape[2], ap.shape[1], int(num_classes),
        _ip(off), _ip(mask), _ip(ocls), ctypes.c_float(same_different_bias), ctypes.c_float(object_merge_factor),
        ctypes.c_float(merge_logprob_bias), tot)
    return tuple(float(v) for v in tot)


def oracle_init_dump(class_pred, adj_pred, num_classes, offset_list, same_different_bias,
                     object_merge_factor, merge_logprob_bias):
    """Constructor results (segment.cc:153-232): dict of clp[N,C], cls[N], and per record slot
    (pixel*K+k) same/diff/oml/mp/valid."""
    cp, ap, off, mask, _ = _glue(class_pred, adj_pred, offset_list)
    k, h, w = ap.shape
    n = h * w
    c = cp.shape[0]
    d = dict(clp=np.zeros((n, c), np.float32), cls=np.zeros(n, np.int32),
             same=np.zeros(n * k, np.float32), diff=np.zeros(n * k, np.float32),
             oml=np.zeros(n * k, np.float32), mp=np.zeros(n * k, np.float32),
             valid=np.zeros(n * k, np.uint8))
    oracle_lib().mno_init_dump(
        _fp(cp), c, _fp(ap), k, w, h, int(num_classes), _ip(off), float(same_different_bias),
        float(object_merge_factor), float(merge_logprob_bias), _fp(d["clp"]), _ip(d["cls"]),
        _fp(d["same"]), _fp(d["diff"]), _fp(d["oml"]), _fp(d["mp"]),
        d["valid"].ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    d["adj_pred"] = ap  # biased in place when same_different_bias != 0
    return d


# ---- parity helpers (SURVEY 8(d) "parity check attached to every timing") ---------------------
def canonical_relabel(mask):
    """Relabel instance ids by first appearance in raster order; 0 (background) stays 0.
    Returns (canonical mask int32, perm) where perm[k] = canonical label of original label k."""
    flat = np.asarray(mask).ravel()
    labels, first = np.unique(flat, return_index=True)
    order = np.argsort(first, kind="stable")
    perm = np.zeros(int(labels.max()) + 1 if labels.size else 1, dtype=np.int64)
    nxt = 1
    for idx in order:
        lab = int(labels[idx])
        if lab == 0:
            continue
        perm[lab] = nxt
        nxt += 1
    return perm[flat].reshape(np.asarray(mask).shape).astype(np.int32), perm


def canonical_result(mask, object_class):
    """(canonical mask, classes listed in canonical label order)."""
    cm, perm = canonical_relabel(mask)
    ncls = [0] * int(cm.max())
    for k, c in enumerate(object_class, start=1):
        if k < len(perm) and perm[k] > 0:
            ncls[perm[k] - 1] = int(c)
    return cm, ncls


def total_logprob_terms_from_scratch(mask, object_class, class_pred, adj_pred, offset_list):
    """float64 evaluation of segment.cc:314-350 restricted to what the output mask determines: (class term over
    all pixels -- the class of the pixel's instance, class 0 for background --, same-term over the in-image
    (pixel, offset) pairs inside one label, different-term over the pairs across two); all class-0 (label 0)
    pixels are one region."""
    cp = np.asarray(class_pred, dtype=np.float32).clip(EPS, 1.0 - EPS).astype(np.float64)
    ap = np.asarray(adj_pred, dtype=np.float32).clip(EPS, 1.0 - EPS).astype(np.float64)
    m = np.asarray(mask)
    h, w = m.shape
    lab_cls = np.zeros(int(m.max()) + 1, dtype=np.int64)
    for k, c in enumerate(object_class, start=1):
        if k < lab_cls.size:
            lab_cls[k] = c
    cls_map = lab_cls[m]
    tot_class = np.take_along_axis(np.log(cp), cls_map[None], axis=0).sum()
    tot_same = 0.0
    tot_diff = 0.0
    for k, (di, dj) in enumerate(offset_list):
        r0, r1 = max(0, -di), min(h, h - di)
        c0, c1 = max(0, -dj), min(w, w - dj)
        if r0 >= r1 or c0 >= c1:
            continue
        a = m[r0:r1, c0:c1]
        b = m[r0 + di:r1 + di, c0 + dj:c1 + dj]
        s = ap[k, r0:r1, c0:c1]
        same = a == b
        tot_same += np.log(s[same]).sum()
        tot_diff += np.log(1.0 - s[~same]).sum()
    return float(tot_class), float(tot_same), float(tot_diff)


def total_logprob_from_scratch(mask, object_class, class_pred, adj_pred, offset_list,
                               object_merge_factor):
    """class term + omf * (same-term + different-term) of total_logprob_terms_from_scratch."""
    tc, ts, td = total_logprob_terms_from_scratch(mask, object_class, class_pred, adj_pred, offset_list)
    return float(tc + (td + ts) * object_merge_factor)


# ---- the step after the path: checkers for mergenet_b200/csrc/mn_post.cuh ------------
def oracle_resize_bilinear(maps, out_h, out_w):
    """cv2.resize(maps, (out_w, out_h)) (INTER_LINEAR, float32, generic many-channel path) restated in numpy for
    planar maps [..., H, W]: OpenCV resize.cpp computes per destination index f = (float)((d + 0.5) * scale - 0.5),
    s = floor(f), f -= s in float; columns clamp (f, s) to (0, 0) / (0, W - 1) outside the image, rows keep the
    weight and clip the two source rows; horizontal then vertical pass, every product and sum rounded to float."""
    import math
    m = np.ascontiguousarray(maps, dtype=np.float32)
    H, W = m.shape[-2:]

    def coef(n_src, n_dst, clamp):
        scale = n_src / n_dst
        i0 = np.zeros(n_dst, np.int64); i1 = np.zeros(n_dst, np.int64); w1 = np.zeros(n_dst, np.float32)
        for d in range(n_dst):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = math.floor(f)
            f = np.float32(f - np.float32(s))
            if clamp:
                if s < 0:
                    f = np.float32(0); s = 0
                if s >= n_src - 1:
                    f = np.float32(0); s = n_src - 1
            i0[d] = min(max(s, 0), n_src - 1); i1[d] = min(max(s + 1, 0), n_src - 1); w1[d] = f
        return i0, i1, w1

    y0, y1, yb = coef(H, int(out_h), False)
    x0, x1, xa = coef(W, int(out_w), True)
    one = np.float32(1)
    rows = m[..., :, x0] * (one - xa) + m[..., :, x1] * xa
    out = rows[..., y0, :] * (one - yb)[:, None] + rows[..., y1, :] * yb[:, None]
    return out.astype(np.float32)


def oracle_resize_nearest(mask, out_h, out_w):
    """cv2.resize(mask, (out_w, out_h), interpolation=cv2.INTER_NEAREST), restated (resizeNN)."""
    L = oracle_lib()
    m = np.ascontiguousarray(mask, np.int32)
    out = np.empty((out_h, out_w), np.int32)
    L.mno_resize_nearest(m.ctypes.data_as(_I), m.shape[0], m.shape[1], out.ctypes.data_as(_I), out_h, out_w)
    return out


def oracle_coco_rle(mask, n):
    """[counts bytes of maskUtils.encode(asfortranarray(mask == i)) for i in 1..n] (restated maskApi.c)."""
    L = oracle_lib()
    L.mno_coco_rle.restype = ctypes.c_longlong
    L.mno_coco_rle.argtypes = [_I, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong,
                               ctypes.c_void_p]
    m = np.ascontiguousarray(mask, np.int32)
    offs = np.zeros(n + 1, np.int64)
    need = L.mno_coco_rle(m.ctypes.data_as(_I), m.shape[0], m.shape[1], n, None, 0, offs.ctypes.data)
    buf = np.zeros(max(1, need), np.uint8)
    L.mno_coco_rle(m.ctypes.data_as(_I), m.shape[0], m.shape[1], n, buf.ctypes.data, need, offs.ctypes.data)
    return [bytes(buf[offs[i]:offs[i + 1]]) for i in range(n)]


def oracle_coco_rle_decode(strings, h, w):
    """Label mask painted from the per-instance counts strings (rleFrString + rleDecode)."""
    L = oracle_lib()
    L.mno_coco_rle_decode.restype = ctypes.c_int
    L.mno_coco_rle_decode.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, _I]
    out = np.zeros((h, w), np.int32)
    for i, s in enumerate(strings):
        rc = L.mno_coco_rle_decode(s, len(s), h, w, i + 1, out.ctypes.data_as(_I))
        assert rc == 0, (i, rc)
    return out
