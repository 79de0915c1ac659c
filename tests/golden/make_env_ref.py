"""Generates tests/golden/env_ref.npz: what the REFERENCE's environment construction (mpm/hand.py:436-476, HandEnv.__init__)
derives from each of the seven shipped env YAMLs before a simulator exists -- particle positions of the SHAPES section
(mpm/shapes.py, imported with an empty stand-in for open3d), the final particle count, the tool entries of the hand primitives
(HandEnv.parse_sim_cfg -> CudaEnv.parse_tools arguments), and the initial wrist frames / joint positions
(HandEnv.parse_manip_cfgs with a minimal stand-in for yacs' CfgNode).  The reference code runs from /root/reference under the
stubs of make_hand_ref.py; particle clouds are stored as a SHA-256 of their float32 bytes plus their first 8 rows.
Re-run: ``python tests/golden/make_env_ref.py``."""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_hand_ref import REF, load_reference_hand, stub  # noqa: E402

ENVS = ["folding", "rope", "bun", "dumpling", "wrap", "flip", "lift_box"]


class CN(dict):                                   # the part of yacs.config.CfgNode that parse_manip_cfgs touches (hand.py:506-512)
    def __init__(self, *a, new_allowed=False, **k):
        super().__init__(*a, **k)

    __getattr__ = dict.__getitem__

    @classmethod
    def _load_cfg_from_yaml_str(cls, s):
        return cls(yaml.safe_load(s))


def main():
    cwd = os.getcwd()
    os.chdir(REF)
    hand = load_reference_hand()
    y = stub("yacs")
    y.config = stub("yacs.config", CfgNode=CN)
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    spec = importlib.util.spec_from_file_location("ref_shapes", os.path.join(REF, "mpm", "shapes.py"))
    shapes = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shapes)
    out = {}
    for name in ENVS:
        cfg = yaml.safe_load(open(os.path.join(REF, "mpm", "assets", "env_cfgs", f"{name}.yml")))
        sim = cfg["SIMULATOR"]
        objects, colors, _, mly = shapes.Shapes(cfg["SHAPES"]).get()
        n = max(int(sim["n_particles"]), len(objects))
        env = hand.HandEnv.__new__(hand.HandEnv)
        params = hand.HandEnv.parse_sim_cfg(env, types.SimpleNamespace(mode=sim["mode"], scale=sim["scale"], hand_friction=sim["hand_friction"]))
        prims = params["primitives"]
        roots, qpos = hand.HandEnv.parse_manip_cfgs(env, cfg["MANIPULATORS"])
        obj = np.ascontiguousarray(objects, np.float32)
        out.update({f"{name}.n_particles": n, f"{name}.n_objects": len(objects), f"{name}.objects_sha256": hashlib.sha256(obj.tobytes()).hexdigest(),
                    f"{name}.objects_head": obj[:8], f"{name}.has_mly": mly is not None,
                    f"{name}.prim_type": np.array([0 if p["shape"] == "Box" else 1 for p in prims], np.int32),
                    f"{name}.prim_args": np.array([[*p["size"], 0] if p["shape"] == "Box" else [*p["size"], 0, 0] for p in prims], np.float32),
                    f"{name}.prim_friction": np.array([p["friction"] for p in prims], np.float32),
                    f"{name}.root_matrix": np.float64(roots), f"{name}.joint_pos": np.float64(qpos),
                    f"{name}.n_hands": params["hand_cfg"]["n_hands"]})
    os.chdir(cwd)
    path = os.path.join(HERE, "env_ref.npz")
    np.savez_compressed(path, **out)
    print("written", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
