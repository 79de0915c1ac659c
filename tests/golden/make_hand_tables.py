"""Generates tests/golden/shadow_tables.npz: the kinematic tables of the Shadow hand (right and left, scales 1.5 and
2.5) produced by dexdeform_b200.mujoco_parser from the MJCF files of the DexDeform checkout (read-only, not copied).
GPU tests build HandSimulator from these tables because the asset files do not exist on the GPU box."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from dexdeform_b200.mujoco_parser import hand_tables, load_hand  # noqa: E402

out = {}
for tag, sides, scale in (("rh15", ["right_hand"], 1.5), ("rh25", ["right_hand"], 2.5), ("dual15", ["left_hand", "right_hand"], 1.5)):
    t = hand_tables([load_hand(s, scale) for s in sides])
    for k, v in t.__dict__.items():
        out[f"{tag}.{k}"] = np.asarray(v)
np.savez_compressed(os.path.join(HERE, "shadow_tables.npz"), **out)
print("written", os.path.getsize(os.path.join(HERE, "shadow_tables.npz")), "bytes")
