"""The BASELINE.json configurations at their named sizes, shared by tests/golden/make_golden.py
(which runs the unmodified reference on them) and the GPU suite (which compares against those results).
Everything but cfg1's maps is regenerated from seeds; cfg1's maps come from the reference's own UNet and
are stored next to the fixtures."""
import os

import numpy as np

from mergenet_b200 import synth

RECIPE = (0.0, 1.0, 0.03)
PLAIN = (0.0, 1.0, 0.0)
MATRIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matrix")
NAMES = ["cfg1_256x512_recipe", "cfg1_256x512_plain", "cfg3_256x512_oracle", "cfg4_512x512_oracle",
         "cfg4_512x512_soft", "cfg3_1024x2048_oracle"]


def load(name):
    """(class_pred, adj_pred, C, offsets, opts)"""
    if name.startswith("cfg1"):
        z = np.load(os.path.join(MATRIX, "cfg1_256x512_inputs.npz"))
        return z["class_pred"], z["adj_pred"], 9, synth.generate_offsets(40, 10), RECIPE if name.endswith("recipe") else PLAIN
    if name == "cfg3_256x512_oracle":
        cp, sp, offs, _ = synth.cfg_cityscapes(256, 512, seed=1, n_shapes=60, rmax=40, soft=False)
        return cp, sp, 9, offs, RECIPE
    if name == "cfg3_1024x2048_oracle":
        cp, sp, offs, _ = synth.cfg_cityscapes(1024, 2048, seed=2, n_shapes=400, rmax=120, soft=False)
        return cp, sp, 9, offs, RECIPE
    if name == "cfg4_512x512_oracle":
        cp, sp, offs, _ = synth.cfg_coco(512, 512, seed=3, n_shapes=900, rmax=14, soft=False)
        return cp, sp, 81, offs, RECIPE
    if name == "cfg4_512x512_soft":
        cp, sp, offs, _ = synth.cfg_coco(512, 512, seed=3, n_shapes=900, rmax=14, soft=True, noise_seed=11)
        return cp, sp, 81, offs, RECIPE
    raise KeyError(name)
