"""Drop-in for the reference's Cython module ``utils/csegment/c_segment.pyx``.

``run_segmentation`` has the reference's signature, argument meaning, return types and error
behaviour (c_segment.pyx:30-86); the work is done by the CUDA library through its C ABI
(``c_run_segmentation``, the symbol the .pyx binds at :16-25).  A caller switches with::

    # import utils.csegment.c_segment as cseg      (egs/cityscape/local/segment.py:19)
    import mergenet_b200.c_segment as cseg
    mask, object_class = cseg.run_segmentation(class_mask, bound_mask, num_classes, offset_list,
                                               same_different_bias, object_merge_factor,
                                               merge_logprob_bias)

``run_segmentation_exact`` (additive) has the same signature and returns the reference's result INCLUDING its
order among exactly equal priorities and its label numbering -- the raw arrays of the reference, at sequential
speed (mn_exact.cuh); see DESIGN.md section 2.
"""
import ctypes

import numpy as np

from . import _lib


def _check_map(a, name):
    # the Cython buffer declaration `np.ndarray[float, ndim=3, mode="c"] x not None` raises
    # TypeError for None / non-arrays and ValueError for wrong dtype / ndim / non-contiguity
    if a is None:
        raise TypeError("Argument '%s' must not be None" % name)
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)"
                        % (name, type(a).__name__))
    if a.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float' but got '%s'" % a.dtype.name)
    if a.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % a.ndim)
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("ndarray is not C-contiguous")


def _glue(class_pred, adj_pred, offset_list):
    _check_map(class_pred, "class_pred")
    _check_map(adj_pred, "adj_pred")
    if offset_list is None:
        raise TypeError("Argument 'offset_list' must not be None")
    if not isinstance(offset_list, list):
        raise TypeError("Argument 'offset_list' has incorrect type (expected list, got %s)"
                        % type(offset_list).__name__)
    _lib.require_device()
    # pyx:53-55 clip; pyx:57 offsets -> int32
    epsilon = np.finfo(np.float32).eps
    class_pred = np.ascontiguousarray(class_pred.clip(epsilon, 1.0 - epsilon))
    adj_pred = np.ascontiguousarray(adj_pred.clip(epsilon, 1.0 - epsilon))
    offset_array = np.ascontiguousarray(np.array(offset_list).astype(np.int32))
    img_height, img_width = adj_pred.shape[1], adj_pred.shape[2]
    # pyx:66-67 outputs
    mask_pred = np.zeros((img_height, img_width)).astype(np.int32)
    object_class_pred = np.zeros((1, img_height * img_width)).astype(np.int32)
    return class_pred, adj_pred, offset_array, mask_pred, object_class_pred


def _object_class_list(object_class_pred):
    # pyx:80-84
    object_class = []
    for i in range(object_class_pred.shape[1] - 1):
        if object_class_pred[0, i] == -1:
            break
        object_class.append(object_class_pred[0, i])
    return object_class


def run_segmentation(class_pred, adj_pred, num_classes, offset_list, same_different_bias,
                     object_merge_factor, merge_logprob_bias):
    """(mask int32[H,W], object_class list[int]) -- c_segment.pyx:30-86."""
    num_classes = int(num_classes)
    class_pred, adj_pred, offset_array, mask_pred, object_class_pred = _glue(class_pred, adj_pred, offset_list)
    class_dim = class_pred.shape[0]
    offset_dim = adj_pred.shape[0]
    img_height, img_width = adj_pred.shape[1], adj_pred.shape[2]
    L = _lib.lib()
    F = ctypes.POINTER(ctypes.c_float)
    I = ctypes.POINTER(ctypes.c_int)
    L.c_run_segmentation(class_pred.ctypes.data_as(F), class_dim, adj_pred.ctypes.data_as(F), offset_dim,
                         img_width, img_height, num_classes, offset_array.ctypes.data_as(I),
                         mask_pred.ctypes.data_as(I), object_class_pred.ctypes.data_as(I),
                         float(same_different_bias), float(object_merge_factor), float(merge_logprob_bias))
    st = L.mn_last_error()
    if st != 0:
        # the reference would exit(1) the process on internal errors (segment.cc:42,667,672)
        raise _lib.MergeNetError(st)
    return mask_pred, _object_class_list(object_class_pred)


def run_segmentation_exact(class_pred, adj_pred, num_classes, offset_list, same_different_bias,
                           object_merge_factor, merge_logprob_bias, stats=None):
    """Same call, the reference's TIE ORDER as well: (mask, object_class) equal the RAW outputs of the reference's
    c_segment.run_segmentation -- also on inputs whose partition depends on the pop order among exactly equal
    priorities (an artefact of libstdc++'s heap and hash-table orders, which ``run_segmentation`` replaces by a fixed
    rule), and including the label numbering.  One GPU thread replays segment.cc:539-727 after the same edge pass
    (mn_exact.cuh): seconds at 256x512; for validation and small / medium images.  ``stats`` (a dict) receives the
    pop / merge / push counts."""
    num_classes = int(num_classes)
    class_pred, adj_pred, offset_array, mask_pred, object_class_pred = _glue(class_pred, adj_pred, offset_list)
    if class_pred.shape[0] != num_classes:
        raise _lib.MergeNetError(1, "class_pred.shape[0] != num_classes")
    L = _lib.lib()
    F = ctypes.POINTER(ctypes.c_float)
    I = ctypes.POINTER(ctypes.c_int)
    n = ctypes.c_int(0)
    st4 = (ctypes.c_longlong * 4)()
    object_class_pred[:] = -1
    rc = L.mn_exact_segment_host(class_pred.ctypes.data_as(F), num_classes, adj_pred.ctypes.data_as(F), adj_pred.shape[0],
                                 adj_pred.shape[1], adj_pred.shape[2], offset_array.ctypes.data_as(I), 0,
                                 float(same_different_bias), float(object_merge_factor), float(merge_logprob_bias),
                                 mask_pred.ctypes.data_as(I), object_class_pred.ctypes.data_as(I), ctypes.byref(n), st4)
    if rc != 0:
        raise _lib.MergeNetError(rc, "mn_exact_segment_host")
    if stats is not None:
        stats.update(zip(("pops", "merges", "pushes", "bucket_arena_collections"), list(st4)))
    return mask_pred, _object_class_list(object_class_pred)
