"""Host-side mirror of the reference's Python segmenter API (``utils/segmenter.py``).

``SegmenterOptions`` (segmenter.py:21-24) and ``ObjectSegmenter(...).run_segmentation()``
(segmenter.py:225-260, 432-483) keep their names, arguments and shape checks.  The merge itself
runs on the GPU with the csegment ("Mode A") semantics -- the variant the Cityscapes recipe uses
(egs/cityscape/local/segment.py:138-143).  The reference's pure-Python variant differs from its own
C++ port in the priority formula, accept rule and pruning; that Mode B is a
"next" row of the scope table and is not implemented: asking for it raises NotImplementedError
rather than silently returning Mode-A results.

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
                 mode="csegment"):
        if mode != "csegment":
            raise NotImplementedError("only the csegment (Mode A) semantics are implemented")
        self.opts = opts
        if self.opts is None:
            self.opts = self.default_options()
        self.class_probs = np.ascontiguousarray(nnet_class_probs, dtype=np.float32)
        self.sameness_probs = np.ascontiguousarray(nnet_sameness_probs, dtype=np.float32)
        self.num_classes = num_classes
        self.offsets = offsets  # should be a list of tuples
        class_dim, self.img_height, self.img_width = self.class_probs.shape
        offset_dim, img_height, img_width = self.sameness_probs.shape
        # segmenter.py:245-250
        assert class_dim == self.num_classes
        assert offset_dim == len(self.offsets)
        assert self.img_height == img_height
        assert self.img_width == img_width

    def default_options(self):
        # segmenter.py:257-260
        return SegmenterOptions(same_different_bias=0.0, object_merge_factor=1.0, merge_logprob_bias=0.0)

    def run_segmentation(self):
        """(mask int[H,W], object_class list) -- segmenter.py:432-483 (csegment semantics)."""
        return c_segment.run_segmentation(self.class_probs, self.sameness_probs, self.num_classes,
                                          [tuple(o) for o in self.offsets],
                                          self.opts.same_different_bias, self.opts.object_merge_factor,
                                          self.opts.merge_logprob_bias)


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
        """(class term, object sameness term, record differentness term, total) of the last run's
        segmentation of ``image`` -- the reference's printed-only ComputeTotalLogprob
        (segment.cc:272-287), evaluated on the GPU from the maintained sufficient statistics."""
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
