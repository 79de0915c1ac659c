"""``GradModel`` -- the reference's ``torch.autograd.Function`` bridge (mpm/torch_wrapper.py:7-184) on the fused engine.

Contract kept: ``get_obs(s, device) -> (cat[x, v, dist] (N, 6+nb), tool (nb, 7)[, grids...])``; ``forward(s, action,
*past_obs)`` runs env step ``s`` (``substeps`` substeps) and returns the observations of step ``s+1``;
``diff_forward(s, pos (S,nb,3), rot (S,nb,4), *past_obs)`` is the autograd Function whose backward returns
``(None, pos_grad (S,nb,3), rot_grad (S,nb,4), zeros_like(past_obs)...)``; state-to-state gradients never pass through
torch -- they live in the simulator and observation gradients are *added* to them at step boundaries.

What changed:
* the S poses of a step go to the device in one call and the S substeps (and their reverse) are one CUDA-graph launch each;
  the reference performs a device->host->device pose round trip, ~10 library calls and one synchronous gradient download
  per substep (mpm/torch_wrapper.py:115-134);
* nothing crosses PCIe: observations, observation gradients and pose gradients are torch CUDA tensors handed to / filled by
  the engine through raw device pointers (the reference downloads to numpy and re-uploads, mpm/torch_wrapper.py:54-105);
* particles are re-sorted on the device at every env-step boundary, so a rollout may be as long as ``max_steps`` allows;
* ``n_envs > 1``: every tensor gains a leading environment axis -- obs ``(E, N, 6+nb)``, tool ``(E, nb, 7)``, poses
  ``(S, E, nb, 3|4)``, actions ``(E, n_hands, 26)``.
"""
import torch
from torch.autograd import Function


class GradModel:
    def __init__(self, env, return_grid=(-1,), return_svd=False, windowed=None):
        """``windowed`` (default: automatically, when the simulator was created with ``max_steps == substeps``): two-level
        checkpointing.  The simulator then holds the per-substep checkpoints (states, constitutive factors, grids) of ONE env
        step; the state at every env-step boundary of the rollout is kept here as device tensors (96 B per particle), and the
        backward pass re-runs the forward substeps of an env step from its boundary state before differentiating it.  A rollout
        of any length then needs the memory of one env step plus 96 B x particles per step, at the price of one extra forward
        pass (the reference keeps 192 B x particles x substeps, mpm/simulator.py:206)."""
        self.env = env
        self.sim = env.simulator if hasattr(env, "simulator") else env
        self.windowed = (self.sim.max_steps == self.sim.substeps) if windowed is None else bool(windowed)
        if self.windowed and self.sim.max_steps < self.sim.substeps:
            raise ValueError("windowed GradModel: the simulator needs max_steps >= substeps")
        self._boundary, self._carry, self._resident = {}, None, None  # windowed mode: boundary states, gradient of the boundary above, window held
        self.dim = 3
        assert len(return_grid) == 0 or return_grid[0] == -1
        self.primitives = list(range(self.sim.n_bodies))
        self._forward_func = None
        self._set_pose_func = None
        self.return_grid = tuple(return_grid)
        self.return_svd = return_svd
        self.device = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cpu"
        self._frontier = None  # state index whose gradient slot currently holds a valid gradient
        self.pos_rot = None
        self._forward_fk_func = None
        # hand kinematics and their adjoint on the device (dd_hand_fk / dd_hand_fk_grad) instead of the ~100-op torch graph per env
        # step; DD_TORCH_FK=1 keeps the torch mirror of the reference's FK in the loop
        import os
        self.use_device_fk = getattr(self.sim, "device_fk", None) is not None and os.environ.get("DD_TORCH_FK", "0") != "1"

    @property
    def substeps(self):
        return self.sim.substeps

    def zero_grad(self, return_grid=None, return_svd=None, **kwargs):
        """mpm/torch_wrapper.py:24-30: a new optimisation iteration starts -- states[0]'s gradients are cleared (the forward
        pass clears those of every later state, mpm/simulator.py:570-571)."""
        self._frontier = None
        self._boundary, self._carry, self._resident = {}, None, None
        self.sim.engine.zero_pose_grads(0, 1)
        if return_grid is not None:
            self.return_grid = tuple(return_grid)
        if return_svd is not None:
            self.return_svd = return_svd or self.return_svd

    def wrap_obs(self, obs):
        output = {"pos": obs[0][..., :3], "vel": obs[0][..., 3:6], "tool": obs[1], "dist": obs[0][..., 6:]}
        obs = obs[2:]
        if len(self.return_grid) > 0:
            ng = len(self.return_grid)
            grid, obs = obs[:ng], obs[ng:]
            output["grid"] = {k: v for k, v in zip(self.return_grid, grid)}
        assert len(obs) == 0
        return output

    def _squeeze(self, t):
        return t[0] if self.sim.n_envs == 1 else t

    def _slot(self, s):
        """Checkpoint slot that holds the state of env-step boundary ``s``."""
        if not self.windowed:
            return s * self.sim.substeps
        if self._resident is not None and s == self._resident + 1:
            return self.sim.substeps
        if self._resident == s or (self._resident is None and s == 0):
            return 0
        raise RuntimeError(f"windowed GradModel: the state of env step {s} is not in the simulator (window {self._resident} is)")

    # ---- two-level checkpointing (no-ops unless windowed)
    def _window_forward(self, s):
        """Before the substeps of env step ``s``: returns the slot its first state sits in."""
        if not self.windowed:
            return s * self.sim.substeps
        eng, S = self.sim.engine, self.sim.substeps
        if self._resident is not None and s == self._resident + 1:
            self._save_boundary(self._resident)           # (saved only now: the last window of a rollout never needs a copy)
            eng.roll(S)                                   # state and poses of slot S become slot 0, re-sorted on the device
        elif not (self._resident is None and s == 0) and self._resident != s:
            raise RuntimeError("windowed GradModel: env steps must be run in order")
        self._resident = s
        return 0

    def _save_boundary(self, s):
        """Keep the first state of window ``s`` (slot 0) before the window is replaced: the backward pass restarts from it."""
        eng = self.sim.engine
        st = eng.get_state(0, device=True)
        pos, rot = eng.get_poses(0, 1, device=True)
        self._boundary[s] = (st["x"], st["v"], st["F"], st["C"], pos, rot)

    def _window_backward(self, s, replay):
        """Before the adjoint of env step ``s``: its checkpoints are in the simulator (re-running the forward substeps from the
        boundary state if another window is resident) and the gradient of the boundary above is seeded."""
        if not self.windowed:
            return s * self.sim.substeps
        eng, S = self.sim.engine, self.sim.substeps
        if self._resident != s:
            if self._resident is not None and self._resident not in self._boundary:
                self._save_boundary(self._resident)       # (a second backward pass over the same rollout may come back to it)
            x, v, F, C, pos, rot = self._boundary[s]
            eng.set_state(0, x, v, F, C)
            eng.set_poses(0, pos, rot)
            replay()
            eng.forward(0, S)
            self._resident = s
        eng.zero_grad(S)
        eng.zero_pose_grads(0, 1)   # slot 0 is shared by all windows (the forward pass only clears slots 1 .. S)
        self._frontier = S
        if self._carry is not None and self._carry[0] == s + 1:
            _, g, gp, gr = self._carry
            eng.add_state_grad(S, gx=g["x"], gv=g["v"], gF=g["F"], gC=g["C"])
            eng.add_pose_grads(S, gpos=gp[0], grot=gr[0])
        return 0

    def _window_backward_done(self, s):
        if not self.windowed:
            return
        if s == 0:   # no window below: the gradient of the rollout's first state stays in the simulator (get_state_grad(0))
            self._carry = None
            return
        eng = self.sim.engine
        gp, gr = eng.get_pose_grads(0, 1, device=True)
        self._carry = (s, eng.get_state_grad(0, device=True), gp, gr)

    def get_obs(self, s, device):
        sim, eng = self.sim, self.sim.engine
        f = self._slot(s)
        if sim.n_bodies:
            pos, rot = eng.get_poses(f, 1, device=True)
            tool = torch.cat((pos[0], rot[0]), -1)
        else:
            tool = torch.zeros((sim.n_envs, 0, 7), device="cuda")
        outputs = [self._squeeze(eng.get_obs(f)).to(device), self._squeeze(tool).to(device)]   # [x | v | dist] in one kernel
        for i in self.return_grid:
            outputs.append(sim.compute_grid_mass(f, i, device=device))
        return tuple(outputs)

    def _ensure_frontier(self, f):
        """The first observation gradient of a backward pass arrives at the last state: open a zeroed gradient slot there."""
        if self._frontier != f:
            self.sim.engine.zero_grad(f)
            self._frontier = f

    def set_obs_grad(self, s, particle_grad, tool_grad, *args):
        sim, eng = self.sim, self.sim.engine
        f = self._slot(s)
        self._ensure_frontier(f)
        for idx, i in enumerate(self.return_grid):
            if idx < len(args) and args[idx] is not None:
                sim.compute_grid_mass(f, i, backward_grad=args[idx])
        E, n, nb = sim.n_envs, sim.n_particles, sim.n_bodies
        eng.add_obs_grad(f, particle_grad.detach().to("cuda", torch.float32).reshape(E, n, 6 + nb).contiguous())
        if nb:
            c = tool_grad.detach().to("cuda", torch.float32).reshape(E, nb, 7)
            eng.add_pose_grads(f, gpos=c[..., :3].contiguous(), grot=c[..., 3:].contiguous())

    @property
    def diff_forward(self):
        if self._forward_func is None:
            model = self

            class forward(Function):
                @staticmethod
                def forward(ctx, s, pos, rot, *past_obs):
                    ctx.s = s
                    ctx.shapes = (pos.shape, rot.shape)
                    ctx.zero = [torch.zeros_like(i) for i in past_obs]
                    S, f = model.substeps, model._window_forward(s)
                    ctx.poses = (pos.detach(), rot.detach()) if model.windowed else None
                    model.sim.set_poses(f + 1, pos, rot)   # poses of states f+1 .. f+S in one call
                    model.sim.forward_range(f, S)          # (re-sorts at f when it continues a rollout)
                    return model.get_obs(s + 1, pos.device)

                @staticmethod
                def backward(ctx, *obs_grad):
                    s = ctx.s
                    S = model.substeps
                    f = model._window_backward(s, lambda: model.sim.set_poses(1, *ctx.poses))
                    model.set_obs_grad(s + 1, *obs_grad)
                    model.sim.backward_range(f, S)
                    model._frontier = f
                    model._window_backward_done(s)
                    gp, gr = model.sim.engine.get_pose_grads(f + 1, S, device=True)  # (S, E, nb, 3|4), on the device
                    return (None, gp.reshape(ctx.shapes[0]), gr.reshape(ctx.shapes[1])) + tuple(ctx.zero)

            self._forward_func = forward
        return self._forward_func.apply

    @property
    def diff_forward_fk(self):
        """One env step with the hand kinematics on the device: ``(s, action, base_pose, joint_rot, *past_obs) -> obs(s+1) +
        (next_base_pose, next_joint_rot)``.  Forward: dd_hand_fk writes the S poses straight into the engine, then the S substeps
        run; backward: observation gradients in, the substeps' adjoint, then dd_hand_fk_grad turns the S pose gradients (and the
        gradient of the end-of-step kinematic state) into gradients of the action and of the kinematic state the step began with."""
        if self._forward_fk_func is None:
            model = self

            class forward_fk(Function):
                @staticmethod
                def forward(ctx, s, action, base, q, *past_obs):
                    sim = model.sim
                    S, f, E, nh = model.substeps, model._window_forward(s), sim.n_envs, sim.n_hands
                    ctx.s, ctx.shapes = s, (action.shape, base.shape, q.shape)
                    ctx.zero = [torch.zeros_like(i) for i in past_obs]
                    a = action.detach().float().reshape(-1, nh, action.shape[-1]).expand(E, -1, -1).contiguous()
                    b = base.detach().float().reshape(-1, nh, 4, 4).expand(E, -1, -1, -1).contiguous()
                    j = q.detach().float().reshape(-1, nh, q.shape[-1]).expand(E, -1, -1).contiguous()
                    if a.shape[-1] < 26:  # fixed base: 20 actuator commands only
                        a = torch.cat((a, torch.zeros((E, nh, 26 - a.shape[-1]), device=a.device)), -1)
                    ctx.has_base = action.shape[-1] == 26
                    ctx.inputs = (a, b, j)
                    nb_, nq_ = sim.device_fk.run(sim.engine, f, S, b, j, a, has_base_action=ctx.has_base)
                    sim.forward_range(f, S)
                    if not model.windowed and f + S < len(sim.base_pose):
                        sim.base_pose[f + S], sim.joint_rot[f + S] = (nb_[0], nq_[0]) if E == 1 else (nb_, nq_)
                    kin = (nb_[0], nq_[0]) if E == 1 else (nb_, nq_)
                    return model.get_obs(s + 1, action.device) + kin

                @staticmethod
                def backward(ctx, *grads):
                    sim, s = model.sim, ctx.s
                    S, E, nh = model.substeps, sim.n_envs, sim.n_hands
                    a, b, j = ctx.inputs
                    f = model._window_backward(s, lambda: sim.device_fk.run(sim.engine, 0, S, b, j, a, has_base_action=ctx.has_base))
                    obs_grad, g_nb, g_nq = grads[:-2], grads[-2], grads[-1]
                    model.set_obs_grad(s + 1, *obs_grad)
                    sim.backward_range(f, S)
                    model._frontier = f
                    model._window_backward_done(s)
                    ga, gb, gq = sim.device_fk.run_grad(sim.engine, f, S, b, j, a, g_nb.reshape(E, nh, 4, 4), g_nq.reshape(E, nh, -1), has_base_action=ctx.has_base)

                    def fit(g, shape):  # an input shared by all environments receives the sum of their gradients
                        g = g[..., :shape[-1]] if g.shape[-1] != shape[-1] else g
                        return g.reshape(shape) if g.numel() == int(torch.Size(shape).numel()) else g.sum(0).reshape(shape)

                    return (None, fit(ga, ctx.shapes[0]), fit(gb, ctx.shapes[1]), fit(gq, ctx.shapes[2])) + tuple(ctx.zero)

            self._forward_fk_func = forward_fk
        return self._forward_fk_func.apply

    @property
    def diff_set_pose(self):
        if self._set_pose_func is None:
            model = self

            class SetPose(Function):
                @staticmethod
                def forward(ctx, s, pos, rot):
                    ctx.s = s
                    ctx.shapes = (pos.shape, rot.shape)
                    model.sim.set_pose(model._slot(s), pos, rot)
                    return model.get_obs(s, pos.device)

                @staticmethod
                def backward(ctx, *obs_grad):
                    s = ctx.s
                    f = model._slot(s)
                    model.set_obs_grad(s, *obs_grad)
                    gp, gr = model.sim.engine.get_pose_grads(f, 1, device=True)
                    return (None, gp.reshape(ctx.shapes[0]), gr.reshape(ctx.shapes[1]))

            self._set_pose_func = SetPose.apply
        return self._set_pose_func

    def forward(self, s, action, *past_obs, pos_rot=None):
        if s == 0 and pos_rot is None:
            self.pos_rot = self.sim.download_pos_rot(0, action.device)  # gradients flow along the kinematic state
        if pos_rot is not None:
            assert isinstance(pos_rot[0], torch.Tensor)
            past_obs = self.diff_set_pose(s, *pos_rot)
        if self.use_device_fk and pos_rot is None:
            out = self.diff_forward_fk(s, action, *self.pos_rot, *past_obs)
            self.pos_rot = out[-2:]
            return out[:-2]
        pos, rot, q_state = self.sim.compute_forward_kinematics(0 if self.windowed else s * self.substeps, action, pos_rot=self.pos_rot)
        self.pos_rot = q_state
        return self.diff_forward(s, pos, rot, *past_obs)
