"""Generates tests/golden/shadow_ref.npz from the REFERENCE's own code, run in this container:

* ``mpm/mujoco_parser.py`` (``prepare``) and ``mpm/robots/interface.py`` need only the standard library and numpy; they are
  imported from /root/reference through a stub ``mpm`` package whose ``__path__`` points at the checkout, so the package's
  ``__init__`` (yacs) never runs;
* ``mpm/hand.py`` (``HandSimulator.__init__`` chain tables, ``hand_forward_kinematics``, ``JointVel_Fk``,
  ``rigid_body_motion_hand``, ``HandEnv.parse_sim_cfg``) is imported as it stands with stand-ins for what is missing here:
  ``mpm.cuda_env`` / ``mpm.simulator`` / ``tools`` (bases that only keep the constructor arguments the hand code reads),
  ``tqdm``, and -- because ``pytorch3d==0.7.2`` and ``transforms3d==0.4.1`` (environment.yml:154,233) are not installed --
  ``pytorch3d.transforms.rotation_conversions`` and ``transforms3d`` modules that expose dexdeform_b200.rotations'
  restatements of the four published conversions the hand code calls (axis_angle_to_matrix, matrix_to_quaternion,
  quaternion_to_matrix, axangle2mat / euler2mat).  Everything kinematic (tree walk, chain order, coupled joints, action
  scaling and clamping, capsule frame fix-up, primitive order and sizes) is therefore the reference's code, not ours.

Stored per configuration (right hand scale 1.5 / 2.5 fixed base, dual hands 1.5): the reference's tables, its primitive
list (type, args as cuda_env.parse_tools builds them) and the poses of one 40-substep env step from seeded inputs.
The asset files and the reference sources are read in place and not copied.  Re-run: ``python tests/golden/make_hand_ref.py``."""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from dexdeform_b200 import rotations  # noqa: E402


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_hand():
    pkg = stub("mpm")
    pkg.__path__ = [os.path.join(REF, "mpm")]            # sub-modules resolve to the checkout, mpm/__init__.py never runs

    class Configurable:
        def __init__(self, *a, **k):
            pass

    class CudaEnvBase(Configurable):                      # only what HandEnv.parse_sim_cfg needs: the default tool entry (cuda_env.py:34-48)
        def default_tool_config(self):
            return {"shape": "", "init_pos": (0.3, 0.3, 0.3), "init_rot": (1.0, 0.0, 0.0, 0.0), "friction": 0.9, "action": {"dim": 0, "scale": ()}}

    class MPMSimulatorBase:                               # HandSimulator.__init__ reads max_steps / substeps only
        def __init__(self, n_bodies, dt=None, dx=None, substeps=None, **k):
            self.n_bodies, self.dt, self.dx, self.substeps, self.max_steps = n_bodies, dt, dx, substeps, 2 * substeps + 1

    stub("tools", Configurable=Configurable)
    stub("mpm.cuda_env", CudaEnv=CudaEnvBase)
    stub("mpm.simulator", MPMSimulator=MPMSimulatorBase)
    stub("mpm.shapes", Shapes=object)
    stub("tqdm", tqdm=lambda *a, **k: None)
    stub("pytorch3d")
    stub("pytorch3d.transforms")
    stub("pytorch3d.transforms.rotation_conversions", axis_angle_to_matrix=rotations.axis_angle_to_matrix,
         matrix_to_quaternion=rotations.matrix_to_quaternion, quaternion_to_matrix=rotations.quaternion_to_matrix)
    t3 = stub("transforms3d")
    t3.axangles = stub("transforms3d.axangles", axangle2mat=rotations.axangle2mat)
    t3.euler = stub("transforms3d.euler", euler2mat=rotations.euler2mat)
    return importlib.import_module("mpm.hand")


def main():
    cwd = os.getcwd()
    os.chdir(REF)
    hand = load_reference_hand()
    out = {}
    for tag, mode, scale, fixed_base in (("rh15", "rh", 1.5, False), ("rh25", "rh", 2.5, True), ("dual15", "dual", 1.5, False)):
        env = hand.HandEnv.__new__(hand.HandEnv)
        cfg = types.SimpleNamespace(mode=mode, scale=scale, hand_friction=0.9)
        params = hand.HandEnv.parse_sim_cfg(env, cfg)
        prims, hcfg = params["primitives"], params["hand_cfg"]
        # mpm/cuda_env.py:78-85 on the reference's tool entries
        ptype = np.array([0 if p["shape"] == "Box" else 1 for p in prims], np.int32)
        pargs = np.array([[*p["size"], 0] if p["shape"] == "Box" else [*p["size"], 0, 0] for p in prims], np.float32)
        sim_cfg = types.SimpleNamespace(ctrl_type="vel")
        sim = hand.HandSimulator(len(prims), hcfg, cfg=sim_cfg, quality=1, device="cpu", fixed_base=fixed_base)
        nh = hcfg["n_hands"]
        g = torch.Generator().manual_seed(11)
        base = torch.tensor(np.asarray(hcfg["root_frame"]), dtype=torch.float32).reshape(nh, 4, 4).clone()
        base[:, :3, :3] = rotations.axis_angle_to_matrix(torch.randn(nh, 3, generator=g) * 0.6) @ base[:, :3, :3]
        q0 = torch.rand(nh, 24, generator=g) * 0.3
        act = torch.rand(nh, 26, generator=g) * 3 - 1.5     # beyond [-1, 1]: exercises the clamp
        pos, rot, (nb_, nq_) = sim.JointVel_Fk(0, act, pos_rot=(base, q0))
        pos0, rot0 = sim.hand_forward_kinematics(sim.base_pose[0][None], sim.joint_rot[0][None])
        out.update({f"{tag}.n_hands": nh, f"{tag}.prim_type": ptype, f"{tag}.prim_args": pargs,
                    f"{tag}.root_frame": np.float32(np.asarray(hcfg["root_frame"]).reshape(nh, 4, 4)),
                    f"{tag}.joint_pos": sim.joint_pos.numpy(), f"{tag}.joint_axis": sim.joint_axis.numpy(),
                    f"{tag}.geometries": sim.geometries.numpy(), f"{tag}.geom_index": sim.geom_index.numpy(),
                    f"{tag}.q_lower": sim.q_lower.numpy().reshape(-1), f"{tag}.q_upper": sim.q_upper.numpy().reshape(-1),
                    f"{tag}.action_map": sim.action_map.numpy(), f"{tag}.action_scale": sim.torch_action_scale.numpy(),
                    f"{tag}.default_qpos": sim.joint_rot[0].numpy(), f"{tag}.substeps": sim.substeps,
                    f"{tag}.in_base": base.numpy(), f"{tag}.in_q": q0.numpy(), f"{tag}.in_action": act.numpy(),
                    f"{tag}.out_pos": pos.numpy(), f"{tag}.out_rot": rot.numpy(), f"{tag}.out_base": nb_.numpy(), f"{tag}.out_q": nq_.numpy(),
                    f"{tag}.rest_pos": pos0.numpy(), f"{tag}.rest_rot": rot0.numpy()})
    os.chdir(cwd)
    path = os.path.join(HERE, "shadow_ref.npz")
    np.savez_compressed(path, **out)
    print("written", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
