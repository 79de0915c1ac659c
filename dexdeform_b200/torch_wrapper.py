"""``GradModel`` -- the reference's ``torch.autograd.Function`` bridge (mpm/torch_wrapper.py:7-184) on the fused engine.

Contract kept: ``get_obs(s, device) -> (cat[x, v, dist] (N, 6+nb), tool (nb, 7)[, grids...])``; ``forward(s, action,
*past_obs)`` runs env step ``s`` (``substeps`` substeps) and returns the observations of step ``s+1``;
``diff_forward(s, pos (S,nb,3), rot (S,nb,4), *past_obs)`` is the autograd Function whose backward returns
``(None, pos_grad (S,nb,3), rot_grad (S,nb,4), zeros_like(past_obs)...)``; state-to-state gradients never pass through
torch -- they live in the simulator and observation gradients are *added* to them at step boundaries.

What changed: the S poses of a step go to the device in one call and the S substeps (and their reverse) are one
CUDA-graph launch each; the reference performs a device->host->device pose round trip, ~10 library calls and one
synchronous gradient download per substep (mpm/torch_wrapper.py:115-134).
"""
import numpy as np
import torch
from torch.autograd import Function


class GradModel:
    def __init__(self, env, return_grid=(-1,), return_svd=False):
        self.env = env
        self.sim = env.simulator if hasattr(env, "simulator") else env
        self.dim = 3
        assert len(return_grid) == 0 or return_grid[0] == -1
        self.primitives = list(range(self.sim.n_bodies))
        self._forward_func = None
        self._set_pose_func = None
        self.return_grid = tuple(return_grid)
        self.return_svd = return_svd
        self.device = "cuda:0"
        self._frontier = None  # state index whose gradient slot currently holds a valid gradient
        self.pos_rot = None

    @property
    def substeps(self):
        return self.sim.substeps

    def zero_grad(self, return_grid=None, return_svd=None, **kwargs):
        self._frontier = None
        if return_grid is not None:
            self.return_grid = tuple(return_grid)
        if return_svd is not None:
            self.return_svd = return_svd or self.return_svd

    def wrap_obs(self, obs):
        output = {"pos": obs[0][:, :3], "vel": obs[0][:, 3:6], "tool": obs[1], "dist": obs[0][:, 6:]}
        obs = obs[2:]
        if len(self.return_grid) > 0:
            ng = len(self.return_grid)
            grid, obs = obs[:ng], obs[ng:]
            output["grid"] = {k: v for k, v in zip(self.return_grid, grid)}
        assert len(obs) == 0
        return output

    def get_obs(self, s, device):
        f = s * self.sim.substeps
        st = self.sim.engine.get_state(f, ("x", "v"))
        x = torch.tensor(st["x"][0], device=device)
        v = torch.tensor(st["v"][0], device=device)
        c = torch.tensor(np.concatenate((self.sim._pos[f][0], self.sim._rot[f][0]), 1), device=device)
        dists = self.sim.get_dists(f, device=device) if self.sim.n_bodies else torch.zeros((self.sim.n_particles, 0), device=device)
        outputs = [torch.cat((x, v, dists), 1), c]
        for i in self.return_grid:
            outputs.append(self.sim.compute_grid_mass(f, i, device=device))
        return tuple(outputs)

    def _ensure_frontier(self, f):
        """The first observation gradient of a backward pass arrives at the last state: open a zeroed gradient slot there."""
        if self._frontier != f:
            self.sim.engine.zero_grad(f)
            self._frontier = f

    def set_obs_grad(self, s, particle_grad, tool_grad, *args):
        f = s * self.sim.substeps
        self._ensure_frontier(f)
        for idx, i in enumerate(self.return_grid):
            if args[idx] is not None:
                self.sim.compute_grid_mass(f, i, backward_grad=args[idx])
        nb = self.sim.n_bodies
        if nb:
            self.sim.get_dists(f, particle_grad[:, 6:])
        pg = particle_grad[:, :6].detach().cpu().numpy().astype(np.float32)
        E, n = self.sim.n_envs, self.sim.n_particles
        self.sim.engine.add_state_grad(f, gx=np.ascontiguousarray(pg[:, :3]).reshape(E, n, 3), gv=np.ascontiguousarray(pg[:, 3:]).reshape(E, n, 3))
        if nb:
            c = tool_grad.reshape(nb, 7).detach().cpu().numpy().astype(np.float32)
            self.sim.engine.add_pose_grads(f, gpos=np.ascontiguousarray(c[:, :3]).reshape(E, nb, 3), grot=np.ascontiguousarray(c[:, 3:]).reshape(E, nb, 4))

    @property
    def diff_forward(self):
        if self._forward_func is None:
            model = self

            class forward(Function):
                @staticmethod
                def forward(ctx, s, pos, rot, *past_obs):
                    ctx.s = s
                    ctx.zero = [torch.zeros_like(i) for i in past_obs]
                    S, f = model.substeps, s * model.substeps
                    model.sim.set_poses(f + 1, pos, rot)   # poses of states f+1 .. f+S in one call
                    model.sim.forward_range(f, S)
                    model.sim.sync()
                    return model.get_obs(s + 1, pos.device)

                @staticmethod
                def backward(ctx, *obs_grad):
                    s = ctx.s
                    S, f = model.substeps, s * model.substeps
                    model.set_obs_grad(s + 1, *obs_grad)
                    model.sim.backward_range(f, S)
                    model._frontier = f
                    gp, gr = model.sim.engine.get_pose_grads(f + 1, S)
                    pos_grad = torch.tensor(gp[:, 0], device=model.device)
                    rot_grad = torch.tensor(gr[:, 0], device=model.device)
                    return (None, pos_grad, rot_grad) + tuple(ctx.zero)

            self._forward_func = forward
        return self._forward_func.apply

    @property
    def diff_set_pose(self):
        if self._set_pose_func is None:
            model = self

            class SetPose(Function):
                @staticmethod
                def forward(ctx, s, pos, rot):
                    ctx.s = s
                    model.sim.set_pose(s * model.substeps, pos, rot)
                    return model.get_obs(s, pos.device)

                @staticmethod
                def backward(ctx, *obs_grad):
                    s = ctx.s
                    f = s * model.substeps
                    model.set_obs_grad(s, *obs_grad)
                    gp, gr = model.sim.engine.get_pose_grads(f, 1)
                    return (None, torch.tensor(gp[0, 0], device=model.device), torch.tensor(gr[0, 0], device=model.device))

            self._set_pose_func = SetPose.apply
        return self._set_pose_func

    def forward(self, s, action, *past_obs, pos_rot=None):
        if s == 0 and pos_rot is None:
            self.pos_rot = self.sim.download_pos_rot(0, action.device)  # gradients flow along the kinematic state
        if pos_rot is not None:
            assert isinstance(pos_rot[0], torch.Tensor)
            past_obs = self.diff_set_pose(s, *pos_rot)
        pos, rot, q_state = self.sim.compute_forward_kinematics(s * self.substeps, action, pos_rot=self.pos_rot)
        self.pos_rot = q_state
        return self.diff_forward(s, pos, rot, *past_obs)
