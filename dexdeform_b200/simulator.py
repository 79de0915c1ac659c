"""Host-side mirror of the reference's ``mpm/simulator.py`` on top of the fused engine.

Same constructor arguments, attributes and methods as ``MPMSimulator`` (mpm/simulator.py:151-634): ``states[f]`` with
``x, v, F, C, body_pos, body_rot`` and their ``*_grad`` buffers, ``set_state / get_state``, ``get_x / get_v``,
``get_dists``, ``init_particles``, ``init_bodies``, ``set_pose``, ``substep``, ``substep_grad``, ``step``, ``sync``.
What is different underneath:

* ``states[f]`` are views into the engine's per-substep checkpoint ring (SoA, cell-sorted); the ``*_grad`` members of a
  state are views into two ping-pong gradient slots -- ``states[f].x_grad`` is readable/addable exactly while the
  backward sweep stands at state ``f`` (which is the only time the reference's GradModel touches it,
  mpm/torch_wrapper.py:79-141);
* ``substep(f)`` / ``substep_grad(f)`` are single engine calls; ``forward_range`` / ``backward_range`` run whole
  env steps as one CUDA-graph launch with no host work per substep (the reference does ~10-20 ctypes calls and a
  host<->device pose round trip per substep, mpm/simulator.py:553-585);
* an extra leading environment axis: ``n_envs`` independent copies of the scene share one engine (default 1, in which
  case every array has the reference's shape).
"""
import numpy as np

from .engine import FusedSim


def rigid_body_motion(states, actions):
    """mpm/simulator.py:19-43 -- poses of free-moving tools after applying (translation, axis-angle) actions."""
    import torch
    T, B = actions.shape[:2]
    pos, q = states
    pos = pos[None, :].expand(T, -1, -1).reshape(-1, 3)
    q = q[None, :].expand(T, -1, -1).reshape(-1, 4)
    actions = actions.reshape(-1, 6)
    rot = actions[:, 3:]
    w = torch.sqrt((rot * rot).sum(axis=-1, keepdims=True) + 1e-16)
    quat = torch.cat((torch.cos(w / 2), (rot / torch.clamp(w, 1e-7, 1e9)) * torch.sin(w / 2)), 1)
    next_pos = pos + actions[:, :3]
    t = q[:, :, None] * quat[:, None, :]
    out = torch.stack([t[:, 0, 0] - t[:, 1, 1] - t[:, 2, 2] - t[:, 3, 3], t[:, 0, 1] + t[:, 1, 0] - t[:, 2, 3] + t[:, 3, 2],
                       t[:, 0, 2] + t[:, 1, 3] + t[:, 2, 0] - t[:, 3, 1], t[:, 0, 3] - t[:, 1, 2] + t[:, 2, 1] + t[:, 3, 0]], 1)
    next_rot = out / torch.linalg.norm(out, dim=-1, keepdims=True)
    return next_pos.reshape(T, B, 3), next_rot.reshape(T, B, 4)


class _Field:
    """One member of a state (``states[f].x`` ...): upload / download / cuda_add / zero in the reference's vocabulary
    (mpm/types.py:294-400), forwarded to the engine."""

    def __init__(self, sim, f, name, grad):
        self.sim, self.f, self.name, self.grad = sim, f, name, grad

    def _squeeze(self, a):
        return a[0] if self.sim.n_envs == 1 else a

    def download(self, n=None, device="numpy", stream=None):
        s, eng = self.sim, self.sim.engine
        if self.name in ("body_pos", "body_rot"):
            if self.grad:
                gp, gr = eng.get_pose_grads(self.f, 1)
                out = (gp if self.name == "body_pos" else gr)[0]
            else:
                out = (s._pos if self.name == "body_pos" else s._rot)[self.f]
        else:
            out = (eng.get_state_grad if self.grad else eng.get_state)(self.f, (self.name,))[self.name]
        out = self._squeeze(out)
        if n is not None and self.name not in ("body_pos", "body_rot"):
            out = out[..., :n, :]
        if device != "numpy":
            import torch
            return torch.tensor(out, device=device)
        return out

    def upload(self, arr, strict=False):
        s = self.sim
        arr = np.ascontiguousarray(arr, np.float32)
        if self.name in ("body_pos", "body_rot") and not self.grad:
            buf = s._pos if self.name == "body_pos" else s._rot
            buf[self.f] = arr.reshape(buf[self.f].shape)
            s.engine.set_poses(self.f, s._pos[self.f:self.f + 1], s._rot[self.f:self.f + 1])
            return
        raise NotImplementedError(f"upload of states[{self.f}].{self.name}{'_grad' if self.grad else ''}: use set_state / cuda_add")

    def cuda_add(self, values, stream=None):
        s, eng = self.sim, self.sim.engine
        v = np.ascontiguousarray(values, np.float32)
        assert self.grad, "cuda_add is only meaningful on gradient buffers"
        if self.name in ("body_pos", "body_rot"):
            v = v.reshape(s.n_envs, s.n_bodies, -1)
            eng.add_pose_grads(self.f, gpos=v if self.name == "body_pos" else None, grot=v if self.name == "body_rot" else None)
        else:
            v = v.reshape(s.n_envs, s.n_particles, -1)
            eng.add_state_grad(self.f, **{"g" + self.name: v})

    def zero(self, stream=None):
        if self.grad:
            self.sim._zero_grad_pending.add(self.f)

    zero_async = zero


class State:
    """View of slot ``f`` of the checkpoint ring with the member names of mpm/simulator.py:88-145."""

    def __init__(self, sim, f):
        self.sim, self.f = sim, f
        for name in ("x", "v", "F", "C", "body_pos", "body_rot"):
            setattr(self, name, _Field(sim, f, name, False))
            setattr(self, name + "_grad", _Field(sim, f, name, True))

    def clear_grad(self, stream=None):
        self.sim.engine.zero_grad(self.f)


class MPMSimulator:
    def __init__(self, n_bodies, cfg=None, ground_friction=0.0, gravity=(0.0, -1.0, 0.0), n_particles=20000, dx=1.0 / 64, dt=0.0001,
                 grid_size=(1.0, 1.0, 1.0), max_steps=30, substeps=20, yield_stress=30.0, vol=(1.0 / 64 / 2) ** 2,
                 mass=(1.0 / 64 / 2) ** 2, E=5000.0, nu=0.2, n_envs=1, **engine_kwargs):
        cfg = dict(cfg or {})  # YAML ``SIMULATOR:`` keys override the constructor defaults (tools/config/configurable.py:249-250)
        ground_friction = cfg.get("ground_friction", ground_friction)
        gravity = cfg.get("gravity", gravity)
        if isinstance(gravity, str):
            gravity = eval(gravity)
        n_particles = int(cfg.get("n_particles", n_particles))
        max_steps = int(cfg.get("max_steps", max_steps))
        yield_stress, E, nu = float(cfg.get("yield_stress", yield_stress)), float(cfg.get("E", E)), float(cfg.get("nu", nu))
        self.dx, self.dt, self.inv_dx = float(dx), float(dt), 1.0 / float(dx)
        self.n_particles = self.max_particles = n_particles
        self.n_bodies, self.substeps, self.max_steps, self.n_envs = int(n_bodies), int(substeps), max_steps, int(n_envs)
        gd = np.ceil(np.array(grid_size) / self.dx / 4).astype(int) * 4  # mpm/simulator.py:181
        self.grid_dim = tuple(int(g) for g in gd)
        self.n_grid = self.grid_dim[0]
        self._ground_friction = float(ground_friction)
        self.gravity = np.float32(np.array(gravity) * 30)  # mpm/simulator.py:385
        self.engine = FusedSim(self.n_envs, n_particles, self.n_bodies, self.grid_dim, self.dx, self.dt, max_steps, self._ground_friction,
                               self.get_ground_height(), self.gravity, **engine_kwargs)
        self.states = [State(self, f) for f in range(max_steps + 1)]
        self._pos = np.zeros((max_steps + 1, self.n_envs, max(self.n_bodies, 1), 3), np.float32)
        self._rot = np.zeros((max_steps + 1, self.n_envs, max(self.n_bodies, 1), 4), np.float32)
        self._rot[..., 0] = 1.0
        self._zero_grad_pending = set()
        mu = E / (2 * (1 + nu))
        lam = E * nu / ((1 + nu) * (1 - 2 * nu))
        self.init_particles(np.zeros(n_particles) + vol, np.zeros(n_particles) + mass, np.zeros((n_particles, 3)) + np.array([mu, lam, yield_stress]))
        self.object_id, self.torch_scale, self.cur, self.action_scales = None, None, 0, None

    # ---- configuration
    def get_ground_friction(self):
        return self._ground_friction

    def get_ground_height(self):
        return 3.0

    def init_particles(self, vol, mass, mu_lam_yield):
        """mpm/simulator.py:375-386; arrays of shape (N,), (N,), (N, 3) -- or with a leading environment axis."""
        n, E = self.n_particles, self.n_envs
        vol, mass, mly = np.float32(vol), np.float32(mass), np.float32(mu_lam_yield)
        assert vol.shape[-1] == n and mass.shape[-1] == n and mly.shape[-2:] == (n, 3)
        full = lambda a, shp: np.ascontiguousarray(np.broadcast_to(a, shp), np.float32)
        self.engine.set_material(full(mass, (E, n)), full(vol, (E, n)), full(mly, (E, n, 3)))

    def set_object_id(self, object_id):
        self.object_id = np.ascontiguousarray(object_id, np.int32)
        assert self.object_id.shape[-1] == self.n_particles

    def compute_grid_mass(self, f, id=-1, device="numpy", backward_grad=None):
        """mpm/simulator.py:323-354 for an integer state index."""
        ids = None if id == -1 else np.ascontiguousarray(np.broadcast_to(self.object_id, (self.n_envs, self.n_particles)), np.int32)
        if backward_grad is not None:
            g = backward_grad.detach().cpu().numpy() if hasattr(backward_grad, "detach") else np.asarray(backward_grad)
            self.engine.compute_grid_mass_grad(f, np.float32(g).reshape((self.n_envs,) + self.grid_dim), ids, id)
            return None
        out = self.engine.compute_grid_mass(f, ids, id)
        out = out[0] if self.n_envs == 1 else out
        if device == "numpy":
            return out
        import torch
        return torch.tensor(out, device=device)

    def init_bodies(self, types, softness, mu, round, args, action_scales, pos=None, rot=None):
        assert len(mu) == len(args) == self.n_bodies
        tfsr = np.stack((np.float32(types), np.float32(mu), np.float32(softness), np.float32(round)), 1).reshape(-1, 4)  # simulator.py:396-397
        self._tfsr = tfsr
        self._args = np.float32(args).reshape(-1, 4)
        self.engine.set_bodies(self._tfsr, self._args)
        self.action_scales = action_scales
        if pos is not None and rot is not None:
            self._pos[0] = np.float32(pos).reshape(1, self.n_bodies, 3)
            self._rot[0] = np.float32(rot).reshape(1, self.n_bodies, 4)
            self.engine.set_poses(0, self._pos[0:1], self._rot[0:1])

    def set_softness(self, softness):
        self._tfsr[:, 2] = softness
        self.engine.set_bodies(self._tfsr, self._args)

    def get_softness(self):
        return self._tfsr[:, 2]

    # ---- state in the reference's tuple format (mpm/simulator.py:232-251)
    def get_state(self, index):
        st = self.engine.get_state(index)
        E, n = self.n_envs, self.n_particles
        if E == 1:
            x, v, F, C = st["x"][0], st["v"][0], st["F"][0].reshape(n, 3, 3), st["C"][0].reshape(n, 3, 3)
        else:
            x, v, F, C = st["x"], st["v"], st["F"].reshape(E, n, 3, 3), st["C"].reshape(E, n, 3, 3)
        tools = [np.r_[a, b] for a, b in zip(self._pos[index][0], self._rot[index][0])] if self.n_bodies else []
        return (x, v, F, C) + tuple(tools)

    def set_state(self, index, state):
        E, n = self.n_envs, self.n_particles
        x, v, F, C = [np.ascontiguousarray(np.broadcast_to(np.float32(a).reshape((-1, n, d)), (E, n, d)), np.float32)
                      for a, d in zip(state[:4], (3, 3, 9, 9))]
        self.engine.set_state(index, x, v, F, C)
        if self.n_bodies and len(state) > 4:
            pose = np.float32(state[4:4 + self.n_bodies])
            self._pos[index] = pose[None, :, :3]
            self._rot[index] = pose[None, :, 3:7]
            self.engine.set_poses(index, self._pos[index:index + 1], self._rot[index:index + 1])
        self.cur = index

    def get_x(self, index, device="numpy"):
        return self.states[index].x.download(self.n_particles, device=device)

    def get_v(self, index, device="numpy"):
        return self.states[index].v.download(self.n_particles, device=device)

    def get_tool_state(self, index=0, device=None):
        state = [np.r_[a, b] for a, b in zip(self._pos[index][0], self._rot[index][0])]
        if device == "numpy" or device is None:
            return state
        import torch
        return [torch.tensor(i).to(device) for i in state]

    def set_tool_state(self, index, pose):
        pose = np.float32(pose)
        assert pose.shape == (self.n_bodies, 7), f"{pose.shape}, {self.n_bodies}"
        self._pos[index] = pose[None, :, :3]
        self._rot[index] = pose[None, :, 3:7]
        self.engine.set_poses(index, self._pos[index:index + 1], self._rot[index:index + 1])

    def get_dists(self, f, grad=None, device="cuda:0"):
        """mpm/simulator.py:294-321: (N, n_bodies) signed distances; with ``grad`` given, their adjoint is accumulated into
        states[f].x_grad and the body pose gradients instead."""
        if grad is None:
            d = self.engine.compute_dist(f)
            d = d[0] if self.n_envs == 1 else d
            if device == "numpy":
                return d
            import torch
            return torch.tensor(d, device=device)
        g = grad.detach().cpu().numpy() if hasattr(grad, "detach") else np.asarray(grad)
        self.engine.compute_dist_grad(f, np.float32(g).reshape(self.n_envs, self.n_particles, self.n_bodies))

    # ---- stepping
    def set_pose(self, state, pos, rot, stream=None):
        """mpm/simulator.py:553-559.  ``state`` is a State (or its index); pos (nb,3) / rot (nb,4) numpy or torch."""
        f = state.f if isinstance(state, State) else int(state)
        if hasattr(pos, "detach"):
            pos, rot = pos.detach().cpu().numpy(), rot.detach().cpu().numpy()
        self._pos[f] = np.float32(pos).reshape(self._pos[f].shape)
        self._rot[f] = np.float32(rot).reshape(self._rot[f].shape)
        self.engine.set_poses(f, self._pos[f:f + 1], self._rot[f:f + 1])

    def set_poses(self, f0, pos, rot):
        """All poses of an env step at once: pos (S, nb, 3) / rot (S, nb, 4) for states f0 .. f0+S-1 (device tensors stay on device)."""
        if hasattr(pos, "detach"):
            p = pos.detach().reshape(pos.shape[0], self.n_envs, self.n_bodies, 3).contiguous()
            r = rot.detach().reshape(rot.shape[0], self.n_envs, self.n_bodies, 4).contiguous()
            self.engine.set_poses(f0, p, r)
            self._pos[f0:f0 + len(p)] = p.cpu().numpy()
            self._rot[f0:f0 + len(r)] = r.cpu().numpy()
        else:
            self._pos[f0:f0 + len(pos)] = np.float32(pos).reshape((-1,) + self._pos.shape[1:])
            self._rot[f0:f0 + len(rot)] = np.float32(rot).reshape((-1,) + self._rot.shape[1:])
            self.engine.set_poses(f0, self._pos[f0:f0 + len(pos)], self._rot[f0:f0 + len(rot)])

    def substep(self, f, clear_grad=False):
        self.engine.forward(f, 1)

    def substep_grad(self, f):
        self.engine.backward(f, 1)

    def forward_range(self, f0, n):
        self.engine.forward(f0, n)

    def backward_range(self, f0, n):
        self.engine.backward(f0, n)

    def sync(self):
        self.engine.sync()

    def download_pos_rot(self, cur, device):
        import torch
        return (torch.tensor(self._pos[cur][0], device=device, dtype=torch.float32), torch.tensor(self._rot[cur][0], device=device, dtype=torch.float32))

    def compute_forward_kinematics(self, f, action, pos_rot=None):
        """mpm/simulator.py:597-624: free tools driven by (translation, axis-angle) velocity actions."""
        import torch
        device = "cpu" if not isinstance(action, torch.Tensor) else action.device
        pos, rot = self.download_pos_rot(f, device) if pos_rot is None else pos_rot
        if not isinstance(action, torch.Tensor):
            action = torch.tensor(np.array(action, np.float32), device=device)
        if self.torch_scale is None:
            self.torch_scale = torch.tensor(np.array(self.action_scales), device=device, dtype=torch.float32)
        action = action.reshape(-1, 6).clamp(-1.0, 1.0) * self.torch_scale.to(device)
        S = self.substeps
        pos, rot = rigid_body_motion((pos, rot), action[None, :].expand(S, -1, -1) * (torch.arange(S, device=device)[:, None, None] + 1) / S)
        return pos, rot, (pos[-1], rot[-1])

    def step(self, action, pos_rot=None):
        """mpm/simulator.py:626-634: one env step from states[0]; the result becomes the new states[0]."""
        pos, rot, _ = self.compute_forward_kinematics(self.cur, action, pos_rot=pos_rot)
        S = self.substeps
        self.set_poses(1, pos, rot)
        self.engine.forward(0, S)
        st = self.engine.get_state(S)
        self.engine.set_state(0, st["x"], st["v"], st["F"], st["C"])  # rolling window + re-sort at the env-step boundary
        self._pos[0], self._rot[0] = self._pos[S], self._rot[S]
        self.engine.set_poses(0, self._pos[0:1], self._rot[0:1])
        self.sync()
