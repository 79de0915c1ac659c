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
    (mpm/types.py:294-400), forwarded to the engine.  ``device`` other than "numpy" returns torch CUDA tensors without a
    host round trip."""

    def __init__(self, sim, f, name, grad):
        self.sim, self.f, self.name, self.grad = sim, f, name, grad

    def _squeeze(self, a):
        return a[0] if self.sim.n_envs == 1 else a

    def download(self, n=None, device="numpy", stream=None):
        s, eng = self.sim, self.sim.engine
        dev = device != "numpy" and str(device).startswith("cuda")
        if self.name in ("body_pos", "body_rot"):
            gp, gr = (eng.get_pose_grads if self.grad else eng.get_poses)(self.f, 1, device=dev)
            out = (gp if self.name == "body_pos" else gr)[0]
        else:
            out = (eng.get_state_grad if self.grad else eng.get_state)(self.f, (self.name,), device=dev)[self.name]
        out = self._squeeze(out)
        if n is not None and self.name not in ("body_pos", "body_rot"):
            out = out[..., :n, :]
        if device != "numpy" and not dev:
            import torch
            return torch.tensor(out, device=device)
        return out

    def upload(self, arr, strict=False):
        s = self.sim
        if self.grad:
            raise NotImplementedError(f"upload of states[{self.f}].{self.name}_grad: gradients are accumulated, use cuda_add (after State.clear_grad)")
        if self.name in ("body_pos", "body_rot"):
            pos, rot = s.engine.get_poses(self.f, 1)
            (pos if self.name == "body_pos" else rot)[0] = np.float32(arr).reshape((pos if self.name == "body_pos" else rot)[0].shape)
            s.engine.set_poses(self.f, pos, rot)
            return
        # one particle field of state f: read the others back, replace this one, store (re-sorts, like any set_state)
        st = s.engine.get_state(self.f)
        st[self.name] = np.ascontiguousarray(np.broadcast_to(np.float32(arr).reshape((-1, s.n_particles, st[self.name].shape[-1])), st[self.name].shape))
        s.engine.set_state(self.f, st["x"], st["v"], st["F"], st["C"])

    def cuda_add(self, values, stream=None):
        s, eng = self.sim, self.sim.engine
        assert self.grad, "cuda_add is only meaningful on gradient buffers"
        v = values if hasattr(values, "is_cuda") and values.is_cuda else np.ascontiguousarray(values, np.float32)
        if self.name in ("body_pos", "body_rot"):
            v = v.reshape(s.n_envs, s.n_bodies, -1)
            eng.add_pose_grads(self.f, gpos=v if self.name == "body_pos" else None, grot=v if self.name == "body_rot" else None)
        else:
            v = v.reshape(s.n_envs, s.n_particles, -1)
            eng.add_state_grad(self.f, **{"g" + self.name: v})

    def zero(self, stream=None):
        """Gradient members only.  Particle gradients of a state live in one ping-pong slot and are cleared together
        (the reference clears them together too, mpm/simulator.py:136-145)."""
        if not self.grad:
            raise NotImplementedError("zero() of a state member: use set_state")
        if self.name in ("body_pos", "body_rot"):
            self.sim.engine.zero_pose_grads(self.f, 1)
        else:
            self.sim.engine.zero_grad(self.f)

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
        # particles are re-sorted on the device at every env-step boundary of a longer rollout (GradModel chains env steps
        # without set_state, mpm/torch_wrapper.py:110-118)
        engine_kwargs.setdefault("resort_interval", self.substeps if max_steps > self.substeps else 0)
        self.engine = FusedSim(self.n_envs, n_particles, self.n_bodies, self.grid_dim, self.dx, self.dt, max_steps, self._ground_friction,
                               self.get_ground_height(), self.gravity, **engine_kwargs)
        self.states = [State(self, f) for f in range(max_steps + 1)]
        if self.n_bodies:  # identity quaternions until poses are set
            rot = np.zeros((max_steps + 1, self.n_envs, self.n_bodies, 4), np.float32)
            rot[..., 0] = 1.0
            self.engine.set_poses(0, np.zeros((max_steps + 1, self.n_envs, self.n_bodies, 3), np.float32), rot)
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

    def set_color(self, inp):
        """mpm/simulator.py:263-266: per-particle colours for the renderer (kept on the host; rendering is out of scope)."""
        self.particle_color = np.zeros(self.n_particles, np.float64) if getattr(self, "particle_color", None) is None else self.particle_color
        self.particle_color[:] = inp

    def get_object_id(self, device="numpy"):
        assert device == "numpy"
        return self.object_id

    def set_object_id(self, object_id):
        self.object_id = np.ascontiguousarray(object_id, np.int32)
        assert self.object_id.shape[-1] == self.n_particles

    def compute_grid_mass(self, f, id=-1, device="numpy", backward_grad=None):
        """mpm/simulator.py:323-354 for an integer state index."""
        ids = None if id == -1 else np.ascontiguousarray(np.broadcast_to(self.object_id, (self.n_envs, self.n_particles)), np.int32)
        if backward_grad is not None:
            g = backward_grad
            if hasattr(g, "is_cuda") and g.is_cuda:
                g = g.detach().float().reshape((self.n_envs,) + self.grid_dim).contiguous()
            else:
                g = np.float32(g.detach().cpu().numpy() if hasattr(g, "detach") else g).reshape((self.n_envs,) + self.grid_dim)
            self.engine.compute_grid_mass_grad(f, g, ids, id)
            return None
        dev = str(device).startswith("cuda")
        out = self.engine.compute_grid_mass(f, ids, id, device=dev)
        out = out[0] if self.n_envs == 1 else out
        if device == "numpy" or dev:
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
            self._upload_pose(0, np.float32(pos), np.float32(rot))

    def _upload_pose(self, f, pos, rot):
        """pos (nb, 3) / rot (nb, 4) shared by all environments, or with a leading environment axis."""
        E, nb = self.n_envs, self.n_bodies
        full = lambda a, d: np.ascontiguousarray(np.broadcast_to(np.float32(a).reshape((-1, nb, d)), (E, nb, d)))[None]
        self.engine.set_poses(f, full(pos, 3), full(rot, 4))

    def _pose7(self, f):
        """(E, nb, 7) numpy: position | quaternion of every primitive at state f."""
        pos, rot = self.engine.get_poses(f, 1)
        return np.concatenate((pos[0], rot[0]), -1)

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
        tools = list(self._pose7(index)[0]) if self.n_bodies else []
        return (x, v, F, C) + tuple(tools)

    def set_state(self, index, state):
        E, n = self.n_envs, self.n_particles
        x, v, F, C = [np.ascontiguousarray(np.broadcast_to(np.float32(a).reshape((-1, n, d)), (E, n, d)), np.float32)
                      for a, d in zip(state[:4], (3, 3, 9, 9))]
        self.engine.set_state(index, x, v, F, C)
        if self.n_bodies and len(state) > 4:
            pose = np.float32(state[4:4 + self.n_bodies])
            self._upload_pose(index, pose[:, :3], pose[:, 3:7])
        self.cur = index

    def get_x(self, index, device="numpy"):
        return self.states[index].x.download(self.n_particles, device=device)

    def get_v(self, index, device="numpy"):
        return self.states[index].v.download(self.n_particles, device=device)

    def get_tool_state(self, index=0, device=None):
        state = list(self._pose7(index)[0])
        if device == "numpy" or device is None:
            return state
        import torch
        return [torch.tensor(i).to(device) for i in state]

    def set_tool_state(self, index, pose):
        pose = np.float32(pose)
        assert pose.shape == (self.n_bodies, 7), f"{pose.shape}, {self.n_bodies}"
        self._upload_pose(index, pose[:, :3], pose[:, 3:7])

    def get_dists(self, f, grad=None, device="cuda:0"):
        """mpm/simulator.py:294-321: (N, n_bodies) signed distances; with ``grad`` given, their adjoint is accumulated into
        states[f].x_grad and the body pose gradients instead.  CUDA tensors in and out stay on the device."""
        if grad is None:
            dev = str(device).startswith("cuda")
            d = self.engine.compute_dist(f, device=dev)
            d = d[0] if self.n_envs == 1 else d
            if device == "numpy" or dev:
                return d
            import torch
            return torch.tensor(d, device=device)
        if hasattr(grad, "is_cuda") and grad.is_cuda:
            g = grad.detach().float().reshape(self.n_envs, self.n_particles, self.n_bodies).contiguous()
        else:
            g = np.float32(grad.detach().cpu().numpy() if hasattr(grad, "detach") else grad).reshape(self.n_envs, self.n_particles, self.n_bodies)
        self.engine.compute_dist_grad(f, g)

    # ---- stepping
    def set_pose(self, state, pos, rot, stream=None):
        """mpm/simulator.py:553-559.  ``state`` is a State (or its index); pos (nb,3) / rot (nb,4) numpy or torch."""
        f = state.f if isinstance(state, State) else int(state)
        if hasattr(pos, "is_cuda") and pos.is_cuda:
            E, nb = self.n_envs, self.n_bodies
            p = pos.detach().float().reshape(-1, nb, 3).expand(E, nb, 3).contiguous()[None]
            r = rot.detach().float().reshape(-1, nb, 4).expand(E, nb, 4).contiguous()[None]
            self.engine.set_poses(f, p, r)
            return
        if hasattr(pos, "detach"):
            pos, rot = pos.detach().cpu().numpy(), rot.detach().cpu().numpy()
        self._upload_pose(f, pos, rot)

    def set_poses(self, f0, pos, rot):
        """All poses of an env step at once: pos (S, [E,] nb, 3) / rot (S, [E,] nb, 4) for states f0 .. f0+S-1 (device tensors
        stay on the device)."""
        E, nb = self.n_envs, self.n_bodies
        if hasattr(pos, "detach"):
            S = pos.shape[0]
            p = pos.detach().float().reshape(S, -1, nb, 3).expand(S, E, nb, 3).contiguous()
            r = rot.detach().float().reshape(S, -1, nb, 4).expand(S, E, nb, 4).contiguous()
        else:
            S = len(pos)
            p = np.ascontiguousarray(np.broadcast_to(np.float32(pos).reshape(S, -1, nb, 3), (S, E, nb, 3)))
            r = np.ascontiguousarray(np.broadcast_to(np.float32(rot).reshape(S, -1, nb, 4), (S, E, nb, 4)))
        self.engine.set_poses(f0, p, r)

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
        dev = str(device).startswith("cuda")
        pos, rot = self.engine.get_poses(cur, 1, device=dev)
        pos, rot = pos[0], rot[0]
        if self.n_envs == 1:
            pos, rot = pos[0], rot[0]
        if dev:
            return pos, rot
        return torch.tensor(pos, device=device, dtype=torch.float32), torch.tensor(rot, device=device, dtype=torch.float32)

    def compute_forward_kinematics(self, f, action, pos_rot=None):
        """mpm/simulator.py:597-624: free tools driven by (translation, axis-angle) velocity actions.  action ([E,] nb, 6)."""
        import torch
        device = "cpu" if not isinstance(action, torch.Tensor) else action.device
        pos, rot = self.download_pos_rot(f, device) if pos_rot is None else pos_rot
        if not isinstance(action, torch.Tensor):
            action = torch.tensor(np.array(action, np.float32), device=device)
        if self.torch_scale is None:
            self.torch_scale = torch.tensor(np.array(self.action_scales), device=device, dtype=torch.float32)
        nb, S = self.n_bodies, self.substeps
        batched = action.dim() == 3 or pos.dim() == 3
        a = action.reshape(-1, nb, 6).clamp(-1.0, 1.0) * self.torch_scale.to(device)
        B = max(a.shape[0], pos.reshape(-1, nb, 3).shape[0])
        a = a.expand(B, nb, 6).reshape(B * nb, 6)
        pos, rot = pos.reshape(-1, nb, 3).expand(B, nb, 3).reshape(B * nb, 3), rot.reshape(-1, nb, 4).expand(B, nb, 4).reshape(B * nb, 4)
        ramp = (torch.arange(S, device=device)[:, None, None] + 1) / S
        pos, rot = rigid_body_motion((pos, rot), a[None].expand(S, -1, -1) * ramp)
        if batched:
            pos, rot = pos.reshape(S, B, nb, 3), rot.reshape(S, B, nb, 4)
        return pos, rot, (pos[-1], rot[-1])

    def step(self, action, pos_rot=None):
        """mpm/simulator.py:626-634: one env step from states[0]; the result becomes the new states[0].  The rolling window
        (state and poses of states[S] -> states[0], re-sorted) never leaves the device."""
        pos, rot, _ = self.compute_forward_kinematics(self.cur, action, pos_rot=pos_rot)
        S = self.substeps
        self.set_poses(1, pos, rot)
        self.engine.forward(0, S)
        self.engine.roll(S)
        self.cur = 0
