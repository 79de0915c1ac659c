"""Host-side mirror of the reference's ``mpm/hand.py``: ``HandSimulator`` (joint-velocity control + Shadow-hand forward
kinematics producing the per-substep poses of the 19/38 collision primitives) and ``HandEnv`` / ``make`` (scene
construction from DexDeform's YAML + MJCF assets, without yacs / open3d / pytorch3d / transforms3d).

State format is the reference's (hand.py:183-210): ``(x, v, F (N,3,3), C (N,3,3), nb x (pos|quat), base_pose (nh,4,4),
joint_rot (nh,24))``; actions are ``(n_hands, 26)`` = 20 actuators + 3 translation + 3 axis-angle (hand.py:383-428).

Two FK paths produce identical poses: ``hand_forward_kinematics`` in torch (differentiable, the reference's algorithm,
hand.py:347-381) and the hand-written CUDA kernel behind ``dd_hand_fk`` (device-resident poses written straight into the
engine's pose table, no host round trip per substep); tests compare them."""
import os

import numpy as np
import torch

from .cuda_env import CudaEnv
from .mujoco_parser import default_assets_dir, hand_tables, load_hand
from .robots import ACTUATORS, DEFAULT_INITIAL_QPOS, JOINTS, N_ACTUATORS, N_JOINTS, actuator_of_joint, joint_limits
from .rotations import axis_angle_to_matrix, euler2mat, matrix_to_quaternion, quaternion_to_matrix
from .scenes import _qmul  # noqa: F401  (numpy quaternion product, re-exported for tests)
from .shapes import Shapes
from .simulator import MPMSimulator

SUPPORTED_ENVS = ["folding", "rope", "bun", "dumpling", "wrap", "flip", "lift_box"]
# primitive index ranges of the hands (hand.py:14-17): 19 collision primitives per hand, left hand first
LH_PRIM_RANGE = list(range(0, 19))
RH_PRIM_RANGE = list(range(19, 38))
SINGLE_PRIM_RANGE = list(range(0, 19))
DUAL_PRIM_RANGE = list(range(0, 38))


def rigid_body_motion_hand(state, actions, T):
    """hand.py:20-65: wrist pose (..., nh, 4, 4) ramped linearly by `actions` (..., nh, 6) over T substeps -> (T, ..., nh, 4, 4).
    Leading batch axes (environments) are carried along."""
    state = state[None].expand(T, *state.shape).clone()
    ramp = ((torch.arange(T, device=actions.device) + 1) / T).reshape((T,) + (1,) * actions.dim())
    actions = actions[None].expand(T, *actions.shape) * ramp
    trans = state[..., :3, 3] + actions[..., :3]
    q = matrix_to_quaternion(state[..., :3, :3])
    rot = actions[..., 3:]
    w = torch.sqrt((rot * rot).sum(-1, keepdim=True) + 1e-16)
    dq = torch.cat((torch.cos(w / 2), (rot / torch.clamp(w, 1e-7, 1e9)) * torch.sin(w / 2)), -1)
    t = q[..., None] * dq[..., None, :]
    out = torch.stack([t[..., 0, 0] - t[..., 1, 1] - t[..., 2, 2] - t[..., 3, 3], t[..., 0, 1] + t[..., 1, 0] - t[..., 2, 3] + t[..., 3, 2],
                       t[..., 0, 2] + t[..., 1, 3] + t[..., 2, 0] - t[..., 3, 1], t[..., 0, 3] - t[..., 1, 2] + t[..., 2, 1] + t[..., 3, 0]], -1)
    out = out / torch.linalg.norm(out, dim=-1, keepdim=True)
    new = torch.zeros_like(state)
    new[..., :3, :3] = quaternion_to_matrix(out)
    new[..., :3, 3] = trans
    new[..., 3, 3] = 1.0
    return new


class HandKinematics:
    """Tables + torch FK, independent of the simulator (so it can be tested on CPU)."""

    def __init__(self, tables, device="cpu"):
        self.t, self.device = tables, device
        g = lambda a, dt=torch.float32: torch.tensor(np.asarray(a), dtype=dt, device=device)
        self.n_hands = tables.n_hands
        self.joint_pos, self.joint_axis = g(tables.joint_pos), g(tables.joint_axis)
        self.mats, self.geom_local = g(tables.mats), g(tables.geom_local)
        self.geom_index = g(tables.geom_joint, torch.long)
        self.ops = list(zip(tables.op_kind.tolist(), tables.op_index.tolist(), tables.op_reset.tolist()))
        lim = joint_limits()
        self.q_lower, self.q_upper = g(lim[:, 0])[None, None, :], g(lim[:, 1])[None, None, :]
        self.action_map = g(actuator_of_joint(), torch.long)

    def forward(self, base_pose, q):
        """hand.py:347-381.  base_pose (S, [E,] nh, 4, 4), q (S, [E,] nh, 24) -> pos (S, [E,] nb, 3), quat (S, [E,] nb, 4 wxyz)."""
        T = torch.zeros(q.shape + (4, 4), device=q.device, dtype=q.dtype)
        T[..., :3, :3] = axis_angle_to_matrix(self.joint_axis * q[..., None])
        T[..., :3, 3] = self.joint_pos
        T[..., 3, 3] = 1
        joint_pose = [None] * N_JOINTS
        base = None
        for kind, idx, reset in self.ops:
            if reset:
                base = base_pose
            if kind == 0:
                base = base @ self.mats[:, idx]
            else:
                base = base @ T[..., idx, :, :]
                joint_pose[idx] = base
        jp = torch.stack([joint_pose[j] for j in self.geom_index.tolist()], -3)  # (S, [E,] nh, n_geoms, 4, 4)
        geom = (jp @ self.geom_local).flatten(-4, -3)                            # hands side by side: primitive = hand * n_geoms + geom
        return geom[..., :3, 3], matrix_to_quaternion(geom[..., :3, :3])


class DeviceFK:
    """Handle on the device-resident kinematic tables (``dd_hand_*`` in include/dexdeform_mpm.h)."""

    def __init__(self, tables, action_scale, library=None):
        import ctypes
        from .types import lib as default_lib
        self.lib = library if library is not None else default_lib
        t = tables
        lim = joint_limits()
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        self._keep = [i32(t.op_kind), i32(t.op_index), i32(t.op_reset), f32(t.mats), f32(t.joint_pos), f32(t.joint_axis), i32(t.geom_joint),
                      f32(t.geom_local), f32(lim[:, 0]), f32(lim[:, 1]), i32(actuator_of_joint()), f32(action_scale)]
        k = self._keep
        p = lambda a: a.ctypes.data
        h = ctypes.c_void_p()
        rc = self.lib.dd_hand_create(t.n_hands, len(k[0]), p(k[0]), p(k[1]), p(k[2]), t.mats.shape[1], p(k[3]), p(k[4]), p(k[5]), len(k[6]), p(k[6]),
                                     p(k[7]), p(k[8]), p(k[9]), p(k[10]), p(k[11]), ctypes.byref(h))
        if rc:
            raise RuntimeError(self.lib.dd_last_error().decode())
        self._h, self.n_hands, self.n_geoms = h, t.n_hands, len(k[6])

    def __del__(self):
        try:
            self.lib.dd_hand_destroy(self._h)
        except Exception:
            pass

    def run(self, engine, f, S, base_pose, joint_rot, action, has_base_action=True):
        """base_pose (E, nh, 4, 4), joint_rot (E, nh, 24), action (E, nh, 26): CUDA tensors.  Writes the poses of states
        f+1..f+S into `engine` and returns the end-of-step (base_pose, joint_rot) as new CUDA tensors."""
        base_pose, joint_rot, action = (a.detach().contiguous().float() for a in (base_pose, joint_rot, action))
        nb_, nq_ = torch.empty_like(base_pose), torch.empty_like(joint_rot)
        rc = self.lib.dd_hand_fk(self._h, engine._h, f, S, base_pose.data_ptr(), joint_rot.data_ptr(), action.data_ptr(), nb_.data_ptr(), nq_.data_ptr(),
                                 int(has_base_action), engine.stream)
        if rc:
            raise RuntimeError(self.lib.dd_last_error().decode())
        return nb_, nq_

    def run_grad(self, engine, f, S, base_pose, joint_rot, action, g_next_base=None, g_next_q=None, has_base_action=True):
        """Reverse mode of ``run`` for the same inputs (after the engine's backward pass over f .. f+S): returns
        (dL/daction (E, nh, 26), dL/dbase_pose (E, nh, 4, 4), dL/djoint_rot (E, nh, 24)) as CUDA tensors."""
        base_pose, joint_rot, action = (a.detach().contiguous().float() for a in (base_pose, joint_rot, action))
        ga, gb, gq = torch.zeros_like(action), torch.zeros_like(base_pose), torch.zeros_like(joint_rot)
        ptr = lambda a: None if a is None else a.detach().contiguous().float().data_ptr()
        keep = [None if a is None else a.detach().contiguous().float() for a in (g_next_base, g_next_q)]
        rc = self.lib.dd_hand_fk_grad(self._h, engine._h, f, S, base_pose.data_ptr(), joint_rot.data_ptr(), action.data_ptr(),
                                      None if keep[0] is None else keep[0].data_ptr(), None if keep[1] is None else keep[1].data_ptr(),
                                      gb.data_ptr(), gq.data_ptr(), ga.data_ptr(), int(has_base_action), engine.stream)
        if rc:
            raise RuntimeError(self.lib.dd_last_error().decode())
        return ga, gb, gq


def read_poses(engine, f0, count):
    """Poses of states f0..f0+count-1 as CUDA tensors (count, E, nb, 3) / (count, E, nb, 4), copied device-to-device."""
    import ctypes
    pos, rot = ctypes.c_void_p(), ctypes.c_void_p()
    slots, E, nb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    engine._check(engine.lib.dd_sim_pose_table(engine._h, ctypes.byref(pos), ctypes.byref(rot), ctypes.byref(slots), ctypes.byref(E), ctypes.byref(nb)))
    n = count * E.value * nb.value
    out_p = torch.empty((count, E.value, nb.value, 4), dtype=torch.float32, device="cuda")
    out_r = torch.empty((count, E.value, nb.value, 4), dtype=torch.float32, device="cuda")
    off = f0 * E.value * nb.value * 16
    engine.lib.cuda_copy_async(out_p.data_ptr(), pos.value + off, n * 16, engine.stream)
    engine.lib.cuda_copy_async(out_r.data_ptr(), rot.value + off, n * 16, engine.stream)
    engine.sync()
    return out_p[..., :3].contiguous(), out_r


class HandSimulator(MPMSimulator):
    def __init__(self, n_bodies, hand_cfg, cfg=None, quality=None, action_scale=None, device=None, mode=None, scale=None, ctrl_type=None,
                 hand_friction=None, fixed_base=False, **engine_kwargs):
        cfg = dict(cfg or {})
        if device is None:  # the reference hard-codes cuda:0 (hand.py:74); one process per GPU uses its current device
            device = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cpu"
        quality = cfg.get("quality", quality if quality is not None else 1)
        fixed_base = cfg.get("fixed_base", fixed_base)
        dt = 0.5e-4 / quality
        dx = 1.0 / (64 * quality)
        substeps = int(np.ceil(2e-3 / dt))
        super().__init__(n_bodies, cfg=cfg, dt=dt, dx=dx, substeps=substeps, **engine_kwargs)
        self.device = device
        self.ctrl_type = cfg.get("ctrl_type", ctrl_type or "vel")
        if self.ctrl_type != "vel":
            raise ValueError(f"ctrl type {self.ctrl_type} is not supported!")
        self.tables = hand_cfg["tables"]
        self.n_hands = self.tables.n_hands
        self.n_joints_per_hand, self.n_actuators_per_hand = N_JOINTS, N_ACTUATORS
        self.kin = HandKinematics(self.tables, device)
        self.q_lower, self.q_upper, self.action_map = self.kin.q_lower, self.kin.q_upper, self.kin.action_map
        if action_scale is None:
            action_scale = [0.33 * 0.002] * N_ACTUATORS + ([0.01] * 3 + [0.015] * 3 if not fixed_base else [0.0] * 6)
        self.torch_action_scale = self.togpu(np.array(action_scale))
        self.base_pose = [None] * (self.max_steps + 1)
        self.joint_rot = [None] * (self.max_steps + 1)
        self.base_pose[0] = self.togpu(self.tables.root_frame)
        self.joint_rot[0] = self.togpu([DEFAULT_INITIAL_QPOS[j] for j in JOINTS])[None, :].expand(self.n_hands, -1)
        self.compute_forward_kinematics = self.JointVel_Fk
        self.fixed_base = bool(fixed_base)
        self.device_fk = DeviceFK(self.tables, np.array(action_scale)) if str(device).startswith("cuda") else None

    def togpu(self, x, dtype=torch.float32):
        return torch.tensor(np.array(x), dtype=dtype, device=self.device)

    def hand_forward_kinematics(self, base_pose, q):
        return self.kin.forward(base_pose, q)

    def download_pos_rot(self, cur, device):
        assert cur == 0
        for i in range(1, len(self.base_pose)):
            self.base_pose[i] = self.joint_rot[i] = None
        return self.base_pose[cur].to(device), self.joint_rot[cur].to(device)

    def get_state(self, index):
        assert index % self.substeps == 0
        return super().get_state(index) + (self.base_pose[index].detach().cpu().numpy(), self.joint_rot[index].detach().cpu().numpy())

    def set_state(self, index, state):
        assert index % self.substeps == 0
        self.base_pose[index] = torch.tensor(np.asarray(state[-2]), device=self.device, dtype=torch.float32)
        self.joint_rot[index] = torch.tensor(np.asarray(state[-1]), device=self.device, dtype=torch.float32)
        pos, rot = self.hand_forward_kinematics(self.base_pose[index][None, :], self.joint_rot[index][None, :])
        pose = np.concatenate([pos[0].detach().cpu().numpy(), rot[0].detach().cpu().numpy()], 1)
        super().set_state(index, tuple(state[:4]) + tuple(pose))

    # ---- renderer-facing state (hand.py:212-234); only positions and the kinematic state travel
    def get_state_render_only(self, f=0, device="numpy"):
        return self.get_x(f), self.base_pose[f].cpu().numpy(), self.joint_rot[f].cpu().numpy()

    def set_state_render_only(self, p, base_pose, joint_rot, f=0):
        base_pose = torch.tensor(np.asarray(base_pose), device=self.device, dtype=torch.float32)
        joint_rot = torch.tensor(np.asarray(joint_rot), device=self.device, dtype=torch.float32)
        pos, rot = self.hand_forward_kinematics(base_pose[None, :], joint_rot[None, :])
        self.states[f].x.upload(p)
        self.states[f].body_pos.upload(pos[0].detach().cpu().numpy())
        self.states[f].body_rot.upload(rot[0].detach().cpu().numpy())

    # ---- signed distances of arbitrary points to the hand primitives (hand.py:236-341; used by policy/preprocess)
    def lh_sdf_given_p(self, p):
        return self.primitive_sdf_given_p(p, LH_PRIM_RANGE)

    def rh_sdf_given_p(self, p):
        return self.primitive_sdf_given_p(p, RH_PRIM_RANGE)

    def _dists_of_points(self, pts):
        """(n_particles, nb) distances with state 0's particle positions replaced by `pts` (hand.py:248-252)."""
        self.states[0].x.upload(pts)
        return self.get_dists(f=0, device="numpy")

    def primitive_sdf_given_p(self, p, hand_inds):
        init_state = self.get_state(0)
        num_points = len(p)
        assert num_points <= self.n_particles
        tmp_p = np.ones((self.n_particles, 3))
        tmp_p[:num_points] = p[:num_points]
        sdf_vals = self._dists_of_points(tmp_p)[:num_points, hand_inds].min(-1)
        self.set_state(0, init_state)
        return sdf_vals

    def sample_pts_helper(self, num_points, hand_inds, center=None, hot_start=None):
        """Rejection sampling of points inside the primitives `hand_inds` from the unit cube around `center` (hand.py:254-290)."""
        points = np.ones((num_points, 3)) * 5
        remain_cnt, started = num_points, False
        while remain_cnt > 0:
            p_samples = (np.random.random((self.n_particles, 3)) - 0.5) + center
            if not started and hot_start is not None:
                tmp_len = min(self.n_particles, len(hot_start))
                p_samples[:tmp_len] = hot_start[:tmp_len]
                started = True
            accept_map = self._dists_of_points(p_samples)[:, hand_inds].min(-1) <= 0
            accept_cnt = int(accept_map.sum())
            start = num_points - remain_cnt
            points[start:start + accept_cnt] = p_samples[accept_map][:min(accept_cnt, remain_cnt)]
            remain_cnt -= accept_cnt
        assert np.all(points != 5)
        return points

    def sample_pts_inside_primitives(self, n_pts, mode=None, hot_start=None):
        """hand.py:292-341 -> (points, their distances to every primitive, hand labels)."""
        init_state = self.get_state(0)
        body_pos = self.states[0].body_pos.download()
        if mode == "dual":
            half = n_pts // 2
            p1 = self.sample_pts_helper(half, LH_PRIM_RANGE, body_pos[LH_PRIM_RANGE, :].mean(0), hot_start)
            p2 = self.sample_pts_helper(half, RH_PRIM_RANGE, body_pos[RH_PRIM_RANGE, :].mean(0), hot_start)
            p = np.concatenate((p1, p2), axis=0)
            labels = np.concatenate((np.zeros((len(p1), 1)), np.ones((len(p2), 1))), axis=0)
        elif mode in ("lh", "rh"):
            hand_inds = (LH_PRIM_RANGE if mode == "lh" else RH_PRIM_RANGE) if self.n_hands == 2 else SINGLE_PRIM_RANGE
            p = self.sample_pts_helper(n_pts, hand_inds, body_pos[hand_inds, :].mean(0), hot_start)
            labels = np.zeros((len(p), 0 if mode == "lh" else 1))  # (sic: hand.py:322 / 331)
        else:
            raise ValueError("Mode is not supported!")
        num_valid = len(p)
        p_samples = np.zeros(shape=(self.n_particles, 3))
        p_samples[:num_valid] = p
        p_sdf_vals = self._dists_of_points(p_samples)[:num_valid]
        self.set_state(0, init_state)
        return p, p_sdf_vals, labels

    def mat2pos_rot(self, mat):
        return mat[..., :3, 3], matrix_to_quaternion(mat[..., :3, :3])

    def JointVel_Fk(self, f, actions, pos_rot=None):
        """hand.py:383-428: action ([E,] nh, 26) -> poses of the S substeps + the end-of-step kinematic state."""
        curr_base, curr_q = (self.base_pose[f], self.joint_rot[f]) if pos_rot is None else pos_rot
        if not isinstance(actions, torch.Tensor):
            actions = torch.tensor(np.asarray(actions), device=self.device, dtype=torch.float32)
        actions = actions.to(self.device)
        S, na = self.substeps, N_ACTUATORS
        assert actions.shape[-2] == self.n_hands
        if actions.dim() == 3 and curr_base.dim() == 3:  # one kinematic state shared by all environments so far
            curr_base, curr_q = curr_base[None].expand(actions.shape[0], -1, -1, -1), curr_q[None].expand(actions.shape[0], -1, -1)
        if actions.shape[-1] == na + 6:
            next_base = rigid_body_motion_hand(curr_base, actions[..., -6:] * self.torch_action_scale[-6:], S)
        else:
            next_base = curr_base[None].expand(S, *curr_base.shape)
        a = (actions[..., :na].clamp(-1.0, 1.0) * self.torch_action_scale[:na])[..., self.action_map]
        ramp = (torch.arange(S, device=self.device) + 1).reshape((S,) + (1,) * a.dim())
        next_q = (curr_q[None] + a[None] * ramp).clamp(self.q_lower, self.q_upper)
        geom_pos, geom_rot = self.hand_forward_kinematics(next_base, next_q)
        nb_, nq_ = next_base[-1], next_q[-1]
        if f + S < len(self.base_pose):
            self.base_pose[f + S], self.joint_rot[f + S] = nb_, nq_
        return geom_pos, geom_rot, (nb_, nq_)

    def step(self, action, q_state=None):
        """hand.py:430-433 + mpm/simulator.py:626-634.  Kinematics, the S substeps and the rolling window (state S -> state 0,
        re-sorted) all stay on the device."""
        S = self.substeps
        if self.device_fk is not None and self.n_envs == 1:
            if not isinstance(action, torch.Tensor):
                action = torch.tensor(np.asarray(action), dtype=torch.float32)
            base, q = (self.base_pose[0], self.joint_rot[0]) if q_state is None else q_state
            nb_, nq_ = self.device_fk.run(self.engine, 0, S, base[None].to(self.device), q[None].to(self.device), action[None].to(self.device),
                                          has_base_action=action.shape[1] == N_ACTUATORS + 6)
            nb_, nq_ = nb_[0], nq_[0]
        else:
            pos, rot, (nb_, nq_) = self.JointVel_Fk(self.cur if self.cur % self.substeps == 0 else 0, action, q_state)
            self.set_poses(1, pos, rot)
        self.engine.forward(0, S)
        self.engine.roll(S)
        self.base_pose[0], self.joint_rot[0] = nb_.detach(), nq_.detach()
        self.cur = 0


def _parse_tuple(v):
    return eval(v, {"np": np}) if isinstance(v, str) else v


class HandEnv(CudaEnv):
    """hand.py:436-649 without the renderer: cfg (dict from the env YAML) -> particles, MJCF primitives, HandSimulator."""

    def __init__(self, cfg, MANIPULATORS=None, env_name=None, assets_dir=None, device=None, tables=None, **engine_kwargs):
        """``tables``: pre-built kinematic tables (tests on machines without the MJCF assets); otherwise parsed from the assets."""
        self.cfg = cfg
        self.assets_dir = assets_dir
        d = self.describe(cfg, MANIPULATORS, tables)
        self.fixed_base = d["fixed_base"]
        self.particle_colors = np.zeros(d["n_particles"]) + 0xdf73ff
        n_bodies, kwargs = self.parse_tools(d["primitives"])
        kwargs.pop("pos"), kwargs.pop("rot")  # (the reference uploads the tool defaults first; the kinematics below overwrite them)
        self.simulator = HandSimulator(n_bodies, d["hand_cfg"], cfg=d["sim_cfg"], device=device, **engine_kwargs)
        self.simulator.init_bodies(**kwargs)
        n = self.simulator.n_particles
        # State buffers start as zeros (mpm/types.py:303-306); SHAPES fill the first len(objects) positions (hand.py:465-467)
        x = np.zeros((n, 3), np.float32)
        if d["objects"] is not None:
            x[: len(d["objects"])] = d["objects"]
        zeros = np.zeros((1, n, 9), np.float32)
        self.simulator.engine.set_state(0, x[None], zeros[..., :3].copy(), zeros, zeros)   # every buffer of a fresh State is zero
        self.renderer = None
        self.initialize(d["root_matrix"], d["joint_pos"])
        self.init_state = self.simulator.get_state(0)

    def describe(self, cfg, MANIPULATORS=None, tables=None):
        """Everything hand.py:436-476 derives from the configuration before a simulator exists (no GPU needed): particle
        positions of the SHAPES section, the simulator section with the final particle count, the tool entries of the hand's
        collision primitives, the kinematic description, and the initial wrist frames / joint positions."""
        sim_cfg = dict(cfg["SIMULATOR"])
        objects = None
        if cfg.get("SHAPES"):
            objects, colors, _, mly = Shapes(cfg["SHAPES"]).get()
            sim_cfg["n_particles"] = max(int(sim_cfg.get("n_particles", 0)), len(objects))
        env_params = self.parse_sim_cfg(sim_cfg, tables)
        root_matrix, joint_pos = self.parse_manip_cfgs(MANIPULATORS if MANIPULATORS is not None else cfg["MANIPULATORS"])
        return {"sim_cfg": sim_cfg, "objects": objects, "n_particles": int(sim_cfg["n_particles"]), "fixed_base": sim_cfg.get("fixed_base", False),
                "primitives": env_params["primitives"], "hand_cfg": env_params["hand_cfg"], "root_matrix": root_matrix, "joint_pos": joint_pos}

    def initialize(self, root_frame=None, joint_pos=None):
        """hand.py:478-489: F = I, the hands at their configured root frames / joint positions."""
        state = list(self.simulator.get_state(0))
        state[2] = np.tile(np.eye(3, dtype=np.float32)[None], (self.simulator.n_particles, 1, 1))
        if root_frame is not None:
            state[-2] = np.float32(np.broadcast_to(root_frame, state[-2].shape))
        if joint_pos is not None:
            state[-1] = np.float32(np.broadcast_to(joint_pos, state[-1].shape))
        self.simulator.set_color(self.particle_colors)
        self.simulator.set_state(0, tuple(state))

    def set_particle_color(self, col):
        self.particle_colors[:] = col
        self.simulator.set_color(self.particle_colors)

    def parse_sim_cfg(self, cfg, tables=None):
        """hand.py:540-624: MJCF -> tool configs of the collision primitives (``primitives``) + kinematic description (``hand_cfg``)."""
        mode, scale, hand_friction = cfg.get("mode", "rh"), float(cfg.get("scale", 1.0)), float(cfg.get("hand_friction", 0.9))
        if mode in ("lh", "rh"):
            sides = ["left_hand" if mode == "lh" else "right_hand"]
        elif mode in ("dual", "lh+rh"):
            sides = ["left_hand", "right_hand"]
        else:
            raise ValueError(f"incorrect hand mode: {mode}")
        if tables is None:
            tables = hand_tables([load_hand(s, scale, getattr(self, "assets_dir", None)) for s in sides])
        assert tables.n_hands == len(sides)
        primitives = []
        for t, size in zip(tables.prim_type, tables.prim_size):
            tool = self.default_tool_config()
            tool["shape"] = "Capsule" if int(t) == 1 else "Box"
            tool["size"] = tuple(float(v) for v in (size[:2] if int(t) == 1 else size[:3]))
            tool["round"] = 0
            tool["friction"] = hand_friction
            primitives.append(tool)
        hand_cfg = {"n_hands": len(sides), "tables": tables, "root_frame": tables.root_frame}
        return {"primitives": primitives, "hand_cfg": hand_cfg}

    @staticmethod
    def get_root_matrix(pos, rot):
        m = np.eye(4)
        m[:3, :3] = euler2mat(*rot)
        m[:3, 3] = pos
        return m

    def parse_manip_cfgs(self, cfgs):
        roots, qposes, idx = [], [], []
        for c in cfgs:
            rot, pos = _parse_tuple(c["init_rot"]), _parse_tuple(c["init_pos"])
            q = c["init_qpos"]
            if q == "default":
                q = list(DEFAULT_INITIAL_QPOS.values())
            elif q == "zero":
                q = [0.0] * N_JOINTS
            else:
                q = list(_parse_tuple(q))
                assert len(q) == N_JOINTS
            idx.append(c["hand_idx"])
            roots.append(self.get_root_matrix(pos, rot))
            qposes.append(np.array(q))
        if len(idx) == 2 and idx[1] < idx[0]:
            raise ValueError("hand config is out of order!")
        return np.stack(roots), np.stack(qposes)

    def set_single_hand_pose(self, hand_idx=0, pos=None, rot=None, joint_pos=None):
        state = list(self.simulator.get_state(0))
        if pos is not None and rot is not None:
            state[-2][hand_idx] = self.get_root_matrix(pos, rot)
        if joint_pos is not None:
            assert self.simulator.n_joints_per_hand == len(joint_pos)
            state[-1][hand_idx] = joint_pos
        self.simulator.set_state(0, tuple(state))

    def set_dual_hand_pose(self, pos=None, rot=None, joint_pos=None):
        """hand.py:640-649."""
        state = list(self.simulator.get_state(0))
        if pos is not None and rot is not None:
            assert self.simulator.n_hands == len(pos) == len(rot) == 2
            state[-2][:] = np.stack([self.get_root_matrix(p, r) for p, r in zip(pos, rot)], axis=0)
        if joint_pos is not None:
            assert self.simulator.n_hands == len(joint_pos) and self.simulator.n_joints_per_hand == len(joint_pos[0])
            state[-1][:] = np.stack([x for x in joint_pos], axis=0)
        self.simulator.set_state(0, tuple(state))


def load_env_cfg(env_name, sim_cfg=None, assets_dir=None):
    import yaml
    path = os.path.join(assets_dir or default_assets_dir(), "env_cfgs", f"{env_name}.yml")
    with open(path) as f:
        cfg = yaml.safe_load(f)
    cfg["env_name"] = env_name
    if sim_cfg:
        cfg["SIMULATOR"].update(sim_cfg)
    return cfg


def make(env_name, sim_cfg=None, assets_dir=None, **kw):
    """mpm/__init__.py:25-34."""
    if env_name not in SUPPORTED_ENVS:
        raise ValueError(f"input environment name *{env_name}* is not supported!")
    return HandEnv(load_env_cfg(env_name, sim_cfg, assets_dir), assets_dir=assets_dir, **kw)
