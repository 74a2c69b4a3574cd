"""The step after the merge path, on the GPU: instance masks back at the image size and
their COCO run-length encoding -- what `egs/cityscape/local/segment.py:147-149,165-186` and
`egs/coco/local/segment.py:190-204` do with cv2 and pycocotools on the host.

    resize_maps_bilinear(maps, out_h, out_w)           cv2.resize(maps, (out_w, out_h)) of the class / offset maps
    resize_masks_nearest(masks, out_h, out_w)          cv2.resize(mask, (out_w, out_h), INTER_NEAREST)
    coco_rle_counts(mask, n_instances)                 [maskUtils.encode(asfortranarray(mask == i))["counts"]]
    convert_to_coco_result(mask, object_class, image_id, catIds)   the reference function, same dicts

CUDA-only like the rest of the package (no fallback).
"""
import ctypes

import numpy as np

from . import _lib


def _check(rc, what):
    if rc != 0:
        raise _lib.MergeNetError(rc, what)


def resize_masks_nearest(masks, out_h, out_w):
    """masks: int32 [H, W] or [B, H, W] (host).  Returns the resized int32 array of the same rank."""
    m = np.ascontiguousarray(masks, dtype=np.int32)
    single = m.ndim == 2
    if single:
        m = m[None]
    if m.ndim != 3:
        raise ValueError("masks must be [H, W] or [B, H, W]")
    _lib.require_device()
    B, H, W = m.shape
    out = np.empty((B, int(out_h), int(out_w)), np.int32)
    _check(_lib.lib().mn_resize_masks_nearest_host(m.ctypes.data, B, H, W, out.ctypes.data, int(out_h), int(out_w)),
           "mn_resize_masks_nearest_host")
    return out[0] if single else out


def resize_maps_bilinear(maps, out_h, out_w):
    """maps: float32 [C, H, W] (or [B, C, H, W]) on the host.  Returns cv2.resize(channel-last maps, (out_w, out_h))
    moved back to channel-first, i.e. what egs/cityscape/local/segment.py:116-123 feeds the segmenter (bit-identical
    to OpenCV 4.13's generic float path, which maps of 2 or >= 5 channels take)."""
    m = np.ascontiguousarray(maps, dtype=np.float32)
    if m.ndim not in (3, 4):
        raise ValueError("maps must be [C, H, W] or [B, C, H, W]")
    _lib.require_device()
    H, W = m.shape[-2:]
    planes = int(np.prod(m.shape[:-2]))
    out = np.empty(m.shape[:-2] + (int(out_h), int(out_w)), np.float32)
    _check(_lib.lib().mn_resize_maps_bilinear_host(m.ctypes.data, planes, H, W, out.ctypes.data, int(out_h), int(out_w)),
           "mn_resize_maps_bilinear_host")
    return out


def coco_rle_counts(mask, n_instances):
    """mask: int32 [H, W] with labels 0..n_instances.  Returns the list of `counts` byte strings."""
    m = np.ascontiguousarray(mask, dtype=np.int32)
    if m.ndim != 2:
        raise ValueError("mask must be [H, W]")
    _lib.require_device()
    n = int(n_instances)
    H, W = m.shape
    offs = np.zeros(n + 1, np.int64)
    cap = max(64, 16 * n + H * W // 8)
    L = _lib.lib()
    for _ in range(2):
        buf = np.empty(cap, np.uint8)
        rc = L.mn_mask_to_coco_rle_host(m.ctypes.data, H, W, n, buf.ctypes.data, cap, offs.ctypes.data)
        if rc == 0:
            return [bytes(buf[offs[i]:offs[i + 1]]) for i in range(n)]
        if offs[n] > cap:  # the library reports the size it needs
            cap = int(offs[n])
            continue
        break
    _check(rc, "mn_mask_to_coco_rle_host")


def convert_to_coco_result(mask, object_class, image_id, catIds):
    """egs/cityscape/local/segment.py:165-186 with the encoding done on the GPU: one dict per instance,
    "segmentation" = {"size": [h, w], "counts": bytes} as pycocotools' maskUtils.encode returns it."""
    mask = np.asarray(mask)
    num_objects = int(mask.max()) if mask.size else 0
    counts = coco_rle_counts(mask, num_objects)
    h, w = mask.shape
    results = []
    for i in range(1, num_objects + 1):
        class_id = object_class[i - 1]
        results.append({"image_id": image_id, "score": 1, "category_id": catIds[class_id],
                        "segmentation": {"size": [int(h), int(w)], "counts": counts[i - 1]}})
    return results
