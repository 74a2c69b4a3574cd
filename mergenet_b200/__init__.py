"""mergenet_b200 -- B200-native MergeNet merge segmenter (post-network instance merging path).

Public surface (mirrors the reference's segmenter API):
    mergenet_b200.c_segment.run_segmentation      drop-in for utils/csegment/c_segment.pyx
    mergenet_b200.segmenter.ObjectSegmenter       drop-in for utils/segmenter.py (csegment semantics)
    mergenet_b200.segmenter.SegmenterOptions
    mergenet_b200.segmenter.BatchSegmenter        additive batched device/host interface
    mergenet_b200.synth                           synthetic maps of the benchmark shapes
The compute path is CUDA only (sm_100a) behind libmergenet_b200.so; nothing here falls back to CPU.
"""
from .segmenter import BatchSegmenter, ObjectSegmenter, SegmenterOptions  # noqa: F401
from . import c_segment, synth  # noqa: F401
