"""ctypes binding of libmergenet_b200.so (the C ABI in include/mergenet_b200.h).

The library is CUDA-only.  Loading it works without a GPU (the CPU suite checks the exported
symbols); every computing entry point fails loudly with MN_STATUS_CUDA when no device is usable.
There is deliberately no fallback of any kind here.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# MN_LIB_PATH selects another BUILD of this same library (e.g. the -DMN_PHASE_CYCLES profiling build)
LIB_PATH = os.environ.get("MN_LIB_PATH") or os.path.join(_HERE, "libmergenet_b200.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["mn_api.cu", "mn_edge.cuh", "mn_merge.cuh", "mn_modeb.cuh", "mn_exact.cuh", "mn_stl_order.h", "mn_stl_primes.h", "mn_post.cuh", "mn_layout.h", "mn_common.h", "mn_log1m_tab.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-fmad=false"]

EXPORTS = ["c_run_segmentation", "mn_shutdown", "mn_last_error", "mn_status_string", "mn_device_count",
           "mn_workspace_bytes_per_image", "mn_plan_create", "mn_plan_destroy",
           "mn_segment_batch_device", "mn_segment_batch_host", "mn_modeb_segment_host", "mn_exact_segment_host", "mn_plan_image_stats",
           "mn_plan_timings", "mn_plan_image_logprob", "mn_debug_edge_dump", "mn_debug_libm", "mn_debug_edge_bench",
           "mn_resize_masks_nearest_device", "mn_resize_masks_nearest_host", "mn_resize_maps_bilinear_device", "mn_resize_maps_bilinear_host", "mn_mask_to_coco_rle_host", "mn_post_last_ms"]


class MergeNetError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        msg = "mergenet_b200: status %d" % status
        try:
            msg += " (%s)" % lib().mn_status_string(status).decode()
        except Exception:
            pass
        super().__init__(msg + (": " + what if what else ""))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    hdr = os.path.join(_HERE, "..", "include", "mergenet_b200.h")
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES) or os.path.getmtime(hdr) > t


def build(force=False, verbose=False):
    """nvcc cross-compiles for sm_100a without a GPU; the .so stays in-tree."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("MN_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB_PATH, os.path.join(CSRC, "mn_api.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


class ImageStats(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int), ("fail_line", ctypes.c_int), ("n_instances", ctypes.c_int),
                ("n_init_entries", ctypes.c_int)] + [(n, ctypes.c_longlong) for n in (
                    "rounds", "events", "merges", "restores", "invalid_pops", "solo_events", "refills",
                    "flushes", "splits", "pairs", "cuts_conflict", "cuts_cascade", "cuts_capacity",
                    "queue_chunks_used", "pixel_pool_used", "tree_nodes_used", "requeues", "hash_overflow", "pixel_pool_collections", "cycles_total")] + [
        ("cycles", ctypes.c_longlong * 16)]

    CYCLE_NAMES = ("conflict", "plan", "accept", "commit", "hot", "flush", "refill", "split", "solo", "pairlist",
                   "rf_leaves", "rf_init", "rf_sort", "sel_stage", "sel_class", "sel_pix")

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "cycles"}
        d["cycles"] = dict(zip(self.CYCLE_NAMES, list(self.cycles)))
        return d


class Timings(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("h2d_ms", "edge_ms", "record_init_sort_ms", "merge_ms",
                                               "label_ms", "d2h_ms", "total_ms", "aggregate_ms")] + [
        ("edge_launches", ctypes.c_longlong), ("other_launches", ctypes.c_longlong)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None
_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)
_V = ctypes.c_void_p


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libmergenet_b200.so is not built: run `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.c_run_segmentation.restype = None
    L.c_run_segmentation.argtypes = [_F, ctypes.c_int, _F, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, _I, _I, _I, ctypes.c_float, ctypes.c_float, ctypes.c_float]
    L.mn_shutdown.restype = None
    L.mn_shutdown.argtypes = []
    L.mn_last_error.restype = ctypes.c_int
    L.mn_status_string.restype = ctypes.c_char_p
    L.mn_status_string.argtypes = [ctypes.c_int]
    L.mn_device_count.restype = ctypes.c_int
    L.mn_workspace_bytes_per_image.restype = ctypes.c_size_t
    L.mn_workspace_bytes_per_image.argtypes = [ctypes.c_int] * 4
    L.mn_plan_create.restype = ctypes.c_int
    L.mn_plan_create.argtypes = [ctypes.POINTER(_V), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, _I, ctypes.c_int]
    L.mn_plan_destroy.restype = None
    L.mn_plan_destroy.argtypes = [_V]
    L.mn_segment_batch_device.restype = ctypes.c_int
    L.mn_segment_batch_device.argtypes = [_V, ctypes.c_int, _V, _V, _V, _V, _V, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_float, ctypes.c_float, _V]
    L.mn_segment_batch_host.restype = ctypes.c_int
    L.mn_segment_batch_host.argtypes = [_V, ctypes.c_int, _V, _V, _V, _V, _V, ctypes.c_int, ctypes.c_float,
                                        ctypes.c_float, ctypes.c_float]
    L.mn_modeb_segment_host.restype = ctypes.c_int
    L.mn_modeb_segment_host.argtypes = [_V, _V, _V, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _I,
                                        ctypes.c_double, ctypes.c_double, ctypes.c_double, _V, _V, _V, _V]
    L.mn_exact_segment_host.restype = ctypes.c_int
    L.mn_exact_segment_host.argtypes = [_F, ctypes.c_int, _F, ctypes.c_int, ctypes.c_int, ctypes.c_int, _I, ctypes.c_int,
                                        ctypes.c_float, ctypes.c_float, ctypes.c_float, _I, _I, _I, _V]
    L.mn_plan_image_stats.restype = ctypes.c_int
    L.mn_plan_image_stats.argtypes = [_V, ctypes.c_int, ctypes.POINTER(ImageStats)]
    L.mn_plan_image_logprob.restype = ctypes.c_int
    L.mn_plan_image_logprob.argtypes = [_V, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    L.mn_plan_timings.restype = ctypes.c_int
    L.mn_plan_timings.argtypes = [_V, ctypes.POINTER(Timings)]
    L.mn_debug_edge_dump.restype = ctypes.c_int
    L.mn_debug_edge_dump.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _I, _F, _F,
                                     ctypes.c_float, ctypes.c_float, ctypes.c_float, _F, _I, _F, _F, _F, _F, _I, _I]
    L.mn_debug_edge_bench.restype = ctypes.c_int
    L.mn_debug_edge_bench.argtypes = [ctypes.c_int] * 4 + [_I] + [ctypes.c_int] * 3 + [_F]
    L.mn_resize_masks_nearest_device.restype = ctypes.c_int
    L.mn_resize_masks_nearest_device.argtypes = [_V, ctypes.c_int, ctypes.c_int, ctypes.c_int, _V, ctypes.c_int, ctypes.c_int, _V]
    L.mn_resize_masks_nearest_host.restype = ctypes.c_int
    L.mn_resize_masks_nearest_host.argtypes = [_V, ctypes.c_int, ctypes.c_int, ctypes.c_int, _V, ctypes.c_int, ctypes.c_int]
    L.mn_resize_maps_bilinear_device.restype = ctypes.c_int
    L.mn_resize_maps_bilinear_device.argtypes = [_V, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, _V, ctypes.c_int, ctypes.c_int, _V]
    L.mn_resize_maps_bilinear_host.restype = ctypes.c_int
    L.mn_resize_maps_bilinear_host.argtypes = [_V, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, _V, ctypes.c_int, ctypes.c_int]
    L.mn_mask_to_coco_rle_host.restype = ctypes.c_int
    L.mn_mask_to_coco_rle_host.argtypes = [_V, ctypes.c_int, ctypes.c_int, ctypes.c_int, _V, ctypes.c_longlong, _V]
    L.mn_post_last_ms.restype = ctypes.c_float
    L.mn_debug_libm.restype = ctypes.c_int
    L.mn_debug_libm.argtypes = [ctypes.c_int, ctypes.c_uint, ctypes.c_uint, ctypes.c_float, _F]
    _lib = L
    return L


def require_device():
    n = lib().mn_device_count()
    if n <= 0:
        raise MergeNetError(7, "no usable CUDA device; this library has no CPU path")
    return n
