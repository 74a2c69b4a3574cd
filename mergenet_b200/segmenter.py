"""Host-side mirror of the reference's Python segmenter API (``utils/segmenter.py``).

``SegmenterOptions`` (segmenter.py:21-24) and ``ObjectSegmenter(...).run_segmentation()``
(segmenter.py:225-260, 432-483) keep their names, arguments and shape checks.  Two semantics:

* ``mode="segmenter"`` (the default, "Mode B"): what the reference CLASS itself computes -- priority
  ``(oml*omf + cdl + mlb) / (n1*n2)`` (py:189-193), merge on ``>=`` (py:470), float64 class accumulators
  (py:51), heapq's own order among equal priorities, ``prune(200)`` (py:351-375, incl. the
  UnboundLocalError when there is no class-0 object), int64 mask with labels in ascending surviving id
  (py:377-389).  A drop-in for ``egs/coco/local/segment.py:155-164``.  Strictly sequential on the GPU
  (mn_modeb.cuh): the small-image mode the pure-Python reference is, not the hot path.
* ``mode="csegment"`` ("Mode A"): the semantics of the reference's C++ port, which the Cityscapes recipe
  calls (egs/cityscape/local/segment.py:138-143) -- the hot path of this library.
* ``mode="csegment-exact"``: Mode A including the reference's order among exactly equal priorities and its label
  numbering (``c_segment.run_segmentation_exact``; sequential, for validation and small / medium images).

``BatchSegmenter`` is the additive batched interface (device tensors in, device tensors out).
"""
import ctypes
from collections import namedtuple

import numpy as np

from . import _lib, c_segment

SegmenterOptions = namedtuple('SegmenterOptions',
                              ['same_different_bias', 'object_merge_factor', 'merge_logprob_bias'])


class ObjectSegmenter:
    def __init__(self, nnet_class_probs, nnet_sameness_probs, num_classes, offsets, opts=None,
                 mode="segmenter"):
        if mode not in ("segmenter", "csegment", "csegment-exact"):
            raise ValueError("mode must be 'segmenter' (utils/segmenter.py semantics), 'csegment' or 'csegment-exact'")
        self.mode = mode
        self.opts = opts
        if self.opts is None:
            self.opts = self.default_options()
        self.num_classes = num_classes
        self.offsets = offsets  # should be a list of tuples
        if mode == "segmenter":
            # segmenter.py:230-241, with the reference's own NumPy arithmetic (float32 maps stay float32 under
            # NumPy 2; the device consumes the logarithms, so they carry exactly the reference's bits)
            epsilon = np.finfo(np.float32).eps
            self.class_probs = np.asarray(nnet_class_probs, dtype=np.float32).clip(epsilon, 1.0 - epsilon)
            self.sameness_probs = np.asarray(nnet_sameness_probs, dtype=np.float32).clip(epsilon, 1.0 - epsilon)
            if self.opts.same_different_bias != 0.0:
                sameness_probs_biased_logit = (np.log(self.sameness_probs) - np.log(1.0 - self.sameness_probs) +
                                               self.opts.same_different_bias)
                self.sameness_probs = 1.0 / (1.0 + np.exp(-sameness_probs_biased_logit))
        else:
            self.class_probs = np.ascontiguousarray(nnet_class_probs, dtype=np.float32)
            self.sameness_probs = np.ascontiguousarray(nnet_sameness_probs, dtype=np.float32)
        class_dim, self.img_height, self.img_width = self.class_probs.shape
        offset_dim, img_height, img_width = self.sameness_probs.shape
        # segmenter.py:245-250
        assert class_dim == self.num_classes
        assert offset_dim == len(self.offsets)
        assert self.img_height == img_height
        assert self.img_width == img_width
        self.stats = None

    def default_options(self):
        # segmenter.py:257-260
        return SegmenterOptions(same_different_bias=0.0, object_merge_factor=1.0, merge_logprob_bias=0.0)

    def run_segmentation(self, prune_threshold=200.0):
        """(mask [H,W], object_class list) -- segmenter.py:432-483.  mode "segmenter": int64 mask, labels in
        ascending surviving object id, after prune(prune_threshold) (the reference always prunes at 200.0);
        raises UnboundLocalError where the reference does.  mode "csegment": the C++ port's result (int32)."""
        if self.mode != "segmenter":
            run = c_segment.run_segmentation if self.mode == "csegment" else c_segment.run_segmentation_exact
            return run(self.class_probs, self.sameness_probs, self.num_classes,
                       [tuple(o) for o in self.offsets],
                       self.opts.same_different_bias, self.opts.object_merge_factor,
                       self.opts.merge_logprob_bias)
        _lib.require_device()
        return self._run_modeb(_lib.lib().mn_modeb_segment_host, prune_threshold)

    def _run_modeb(self, entry, prune_threshold):
        h, w = self.img_height, self.img_width
        log_class = np.ascontiguousarray(np.log(self.class_probs), dtype=np.float32)           # py:305
        log_same = np.ascontiguousarray(np.log(self.sameness_probs), dtype=np.float32)         # py:135
        log_diff = np.ascontiguousarray(np.log(1.0 - self.sameness_probs), dtype=np.float32)   # py:136
        off = np.ascontiguousarray(np.array([tuple(o) for o in self.offsets], dtype=np.int32))
        mask = np.zeros((h, w), dtype=np.int64)
        ocls = np.full(h * w, -1, dtype=np.int32)
        n = ctypes.c_int(0)
        st = (ctypes.c_longlong * 4)()
        rc = entry(log_class.ctypes.data, log_same.ctypes.data, log_diff.ctypes.data, int(self.num_classes),
                   len(self.offsets), h, w, off.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                   float(self.opts.object_merge_factor), float(self.opts.merge_logprob_bias), float(prune_threshold),
                   mask.ctypes.data, ocls.ctypes.data, ctypes.byref(n), st)
        self.stats = dict(zip(("pops", "merges", "pushes", "pruned"), list(st)))
        if rc == 9:  # MN_STATUS_NO_BACKGROUND
            raise UnboundLocalError("cannot access local variable 'background_obj' where it is not associated with a value")
        if rc != 0:
            raise _lib.MergeNetError(rc, "mn_modeb_segment_host")
        return mask, [int(c) for c in ocls[:n.value]]


class BatchSegmenter:
    """A plan for up to ``max_batch`` images of one shape on one GPU.

    ``segment_device`` takes torch CUDA tensors class_probs [B,C,H,W] and same_probs [B,K,H,W]
    (float32, contiguous) and returns (masks int32 [B,H,W], object_class int32 [B,H*W],
    num_instances int32 [B]) on the same device.  ``segment_host`` takes / returns numpy arrays and
    includes the host<->device copies.
    """

    def __init__(self, max_batch, height, width, num_classes, offsets, device=0):
        _lib.require_device()
        self.max_batch, self.H, self.W, self.C = int(max_batch), int(height), int(width), int(num_classes)
        self.offsets = [tuple(int(v) for v in o) for o in offsets]
        self.K = len(self.offsets)
        self.device = int(device)
        off = np.ascontiguousarray(np.array(self.offsets, dtype=np.int32))
        self._plan = ctypes.c_void_p()
        st = _lib.lib().mn_plan_create(ctypes.byref(self._plan), self.max_batch, self.H, self.W, self.C,
                                       self.K, off.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), self.device)
        if st != 0:
            self._plan = None
            raise _lib.MergeNetError(st, "mn_plan_create")

    @staticmethod
    def workspace_bytes_per_image(height, width, num_classes, num_offsets):
        return int(_lib.lib().mn_workspace_bytes_per_image(height, width, num_classes, num_offsets))

    def close(self):
        if getattr(self, "_plan", None):
            _lib.lib().mn_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def segment_device(self, class_probs, same_probs, opts, clip=True, out=None, logits=False):
        """logits=True: the maps are the network's raw outputs; the edge pass applies F.sigmoid
        (utils/inference_utils.py:43-44,95-96) and the clip while it reads them."""
        import torch
        B = class_probs.shape[0]
        assert class_probs.is_cuda and same_probs.is_cuda and class_probs.dtype == torch.float32
        assert class_probs.is_contiguous() and same_probs.is_contiguous()
        assert tuple(class_probs.shape) == (B, self.C, self.H, self.W)
        assert tuple(same_probs.shape) == (B, self.K, self.H, self.W)
        dev = class_probs.device
        if out is None:
            masks = torch.empty((B, self.H, self.W), dtype=torch.int32, device=dev)
            ocls = torch.empty((B, self.H * self.W), dtype=torch.int32, device=dev)
            ninst = torch.empty((B,), dtype=torch.int32, device=dev)
        else:
            masks, ocls, ninst = out
        stream = torch.cuda.current_stream(dev).cuda_stream
        st = _lib.lib().mn_segment_batch_device(
            self._plan, B, class_probs.data_ptr(), same_probs.data_ptr(), masks.data_ptr(), ocls.data_ptr(),
            ninst.data_ptr(), (2 if logits else 0) | (1 if clip else 0), float(opts.same_different_bias),
            float(opts.object_merge_factor), float(opts.merge_logprob_bias), ctypes.c_void_p(stream))
        if st != 0:
            raise _lib.MergeNetError(st, "mn_segment_batch_device: " + str(self.failed_images(B)))
        return masks, ocls, ninst

    def segment_host(self, class_probs, same_probs, opts, clip=True, out=None, logits=False):
        B = class_probs.shape[0]
        assert class_probs.dtype == np.float32 and same_probs.dtype == np.float32
        assert class_probs.flags["C_CONTIGUOUS"] and same_probs.flags["C_CONTIGUOUS"]
        assert tuple(class_probs.shape) == (B, self.C, self.H, self.W)
        assert tuple(same_probs.shape) == (B, self.K, self.H, self.W)
        if out is None:
            masks = np.empty((B, self.H, self.W), dtype=np.int32)
            ocls = np.empty((B, self.H * self.W), dtype=np.int32)
            ninst = np.empty((B,), dtype=np.int32)
        else:
            masks, ocls, ninst = out
        st = _lib.lib().mn_segment_batch_host(
            self._plan, B, class_probs.ctypes.data, same_probs.ctypes.data, masks.ctypes.data,
            ocls.ctypes.data, ninst.ctypes.data, (2 if logits else 0) | (1 if clip else 0), float(opts.same_different_bias),
            float(opts.object_merge_factor), float(opts.merge_logprob_bias))
        if st != 0:
            raise _lib.MergeNetError(st, "mn_segment_batch_host: " + str(self.failed_images(B)))
        return masks, ocls, ninst

    def stats(self, image):
        s = _lib.ImageStats()
        _lib.lib().mn_plan_image_stats(self._plan, int(image), ctypes.byref(s))
        return s.as_dict()

    def total_logprob(self, image):
        """(class term, sameness term, differentness term, total) of the last run's segmentation of
        ``image`` -- the number the reference only prints (segment.cc:314-350, ComputeTotalLogprobFromScratch),
        evaluated on the GPU in float64 from the maps and the final label mask."""
        out = (ctypes.c_double * 4)()
        st = _lib.lib().mn_plan_image_logprob(self._plan, int(image), out)
        if st != 0:
            raise _lib.MergeNetError(st, "mn_plan_image_logprob")
        return tuple(float(v) for v in out)

    def failed_images(self, B):
        out = []
        for b in range(B):
            s = self.stats(b)
            if s["status"] != 0:
                out.append((b, s["status"], s["fail_line"]))
        return out

    def timings(self):
        t = _lib.Timings()
        _lib.lib().mn_plan_timings(self._plan, ctypes.byref(t))
        return t.as_dict()
