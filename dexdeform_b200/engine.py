"""``FusedSim`` -- thin Python handle on the fused engine (ABI-2).  No arithmetic happens here: every method is one
C-ABI call.  Arrays may be numpy arrays (host) or torch CUDA tensors (device); both are passed as raw pointers.  Getters
return numpy by default and torch CUDA tensors with ``device=True`` -- in that case nothing waits for the GPU (the engine
enqueues on ``stream``; give it the stream torch is using, or leave both on the legacy default stream)."""
import ctypes

import numpy as np

from .engine_abi import dd_sim_config
from .types import lib as _default_lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], "float32 C-contiguous arrays only"
        return a.ctypes.data
    # torch tensor
    assert a.dtype.is_floating_point and a.element_size() == 4 and a.is_contiguous()
    return a.data_ptr()


def _is_host(*arrays):
    return any(isinstance(a, np.ndarray) or (a is not None and hasattr(a, "is_cuda") and not a.is_cuda) for a in arrays)


class EngineError(RuntimeError):
    pass


class FusedSim:
    def __init__(self, n_envs, n_particles, n_bodies, grid_dim, dx, dt, max_steps, ground_friction=0.0, ground_height=3.0,
                 gravity=(0.0, -30.0, 0.0), svd_mode=1, use_graphs=True, sort_particles=True, tile_mode=True, grid_ckpt=True, chunk_max=0,
                 resort_interval=0, library=None, stream=None):
        self.lib = library if library is not None else _default_lib
        self.E, self.N, self.nb = int(n_envs), int(n_particles), int(n_bodies)
        self.grid_dim = tuple(int(g) for g in grid_dim)
        self.max_steps = int(max_steps)
        self.dx, self.dt = float(dx), float(dt)
        self.stream = stream  # raw cudaStream_t (int) or None for the legacy default stream
        cfg = dd_sim_config(self.E, self.N, self.nb, *self.grid_dim, self.max_steps, self.dx, self.dt, float(ground_friction),
                            float(ground_height), (ctypes.c_float * 3)(*[float(g) for g in gravity]), int(svd_mode), int(bool(use_graphs)),
                            int(bool(sort_particles)), int(bool(tile_mode)), int(grid_ckpt), int(chunk_max), int(resort_interval))
        handle = ctypes.c_void_p()
        self._h = None
        self._check(self.lib.dd_sim_create(ctypes.byref(cfg), ctypes.byref(handle)))
        self._h = handle

    @classmethod
    def from_scene(cls, scene, n_envs=1, max_steps=None, **kw):
        sim = cls(n_envs, scene["n"], scene["nb"], scene["grid_dim"], scene["dx"], scene["dt"],
                  max_steps if max_steps is not None else scene["steps"], scene["ground_friction"], scene["ground_height"],
                  scene["gravity"].reshape(3), **kw)
        E = n_envs
        tile = lambda a: np.ascontiguousarray(np.broadcast_to(a[None], (E,) + a.shape), dtype=np.float32)
        sim.set_material(tile(scene["mass"]), tile(scene["vol"]), tile(scene["mu_lam_yield"]))
        if scene["nb"]:
            sim.set_bodies(scene["tfsr"], scene["args"])
            cnt = min(len(scene["pos"]), sim.max_steps + 1)
            pos = np.ascontiguousarray(np.broadcast_to(scene["pos"][:cnt, None], (cnt, E) + scene["pos"].shape[1:]), dtype=np.float32)
            rot = np.ascontiguousarray(np.broadcast_to(scene["rot"][:cnt, None], (cnt, E) + scene["rot"].shape[1:]), dtype=np.float32)
            sim.set_poses(0, pos, rot)
        sim.set_state(0, tile(scene["x"]), tile(scene["v"]), tile(scene["F"]), tile(scene["C"]))
        return sim

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.lib.dd_last_error().decode())

    def close(self):
        if self._h is not None:
            self.lib.dd_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _empty(self, shape, device):
        if device:
            import torch
            return torch.empty(shape, dtype=torch.float32, device="cuda")
        return np.empty(shape, np.float32)

    # ---- setup
    def set_material(self, mass, vol, mu_lam_yield):
        self._check(self.lib.dd_sim_set_material(self._h, _ptr(mass), _ptr(vol), _ptr(mu_lam_yield), self.stream))
        self.sync()

    def set_bodies(self, tfsr, args):
        tfsr, args = np.ascontiguousarray(tfsr, np.float32), np.ascontiguousarray(args, np.float32)
        self._check(self.lib.dd_sim_set_bodies(self._h, _ptr(tfsr), _ptr(args)))

    def set_state(self, f, x, v, F, C, non_blocking=False):
        """(E, N, 3|9) arrays in the caller's particle order.  Starts a new particle order (cell sort) for f's segment.
        ``non_blocking`` (pinned host tensors only, torch's ``copy_(non_blocking=True)`` contract): do not wait for the uploads --
        the caller keeps the buffers alive and unchanged until the stream has passed them, and the host goes on enqueueing work
        (kinematics, the substeps) while the copies run."""
        self._check(self.lib.dd_sim_set_state(self._h, f, _ptr(x), _ptr(v), _ptr(F), _ptr(C), self.stream))
        if _is_host(x, v, F, C):
            if non_blocking and all(hasattr(a, "is_pinned") and a.is_pinned() for a in (x, v, F, C)):
                return
            self.sync()  # host arrays may be released by the caller

    def roll(self, f_src):
        """State f_src (particles and poses) becomes state 0, re-sorted, without leaving the device (mpm/simulator.py:634)."""
        self._check(self.lib.dd_sim_roll(self._h, int(f_src), self.stream))

    def set_poses(self, f0, pos, rot):
        """pos (count, E, nb, 3), rot (count, E, nb, 4 wxyz) for slots f0 .. f0+count-1."""
        count = pos.shape[0]
        chunk = max(1, (self.E * self.N * 24) // (7 * self.E * max(self.nb, 1)))
        for c0 in range(0, count, chunk):
            c1 = min(count, c0 + chunk)
            p, r = pos[c0:c1], rot[c0:c1]
            if isinstance(p, np.ndarray):
                p, r = np.ascontiguousarray(p, np.float32), np.ascontiguousarray(r, np.float32)
            self._check(self.lib.dd_sim_set_poses(self._h, f0 + c0, c1 - c0, _ptr(p), _ptr(r), self.stream))
            if _is_host(p, r):
                self.sync()

    def get_poses(self, f0, count, device=False):
        pos, rot = self._empty((count, self.E, self.nb, 3), device), self._empty((count, self.E, self.nb, 4), device)
        if self.nb:
            self._check(self.lib.dd_sim_get_poses(self._h, f0, count, _ptr(pos), _ptr(rot), self.stream))
        return pos, rot

    # ---- simulation
    def forward(self, f0, n):
        self._check(self.lib.dd_sim_forward(self._h, f0, n, self.stream))

    def backward(self, f0, n):
        self._check(self.lib.dd_sim_backward(self._h, f0, n, self.stream))

    def zero_grad(self, f):
        self._check(self.lib.dd_sim_zero_grad(self._h, f, self.stream))

    def zero_pose_grads(self, f0, count=1):
        self._check(self.lib.dd_sim_zero_pose_grads(self._h, f0, count, self.stream))

    def add_state_grad(self, f, gx=None, gv=None, gF=None, gC=None):
        self._check(self.lib.dd_sim_add_state_grad(self._h, f, _ptr(gx), _ptr(gv), _ptr(gF), _ptr(gC), self.stream))
        if _is_host(gx, gv, gF, gC):
            self.sync()

    def sync(self):
        self._check(self.lib.dd_sim_sync(self._h, self.stream))

    def launch_count(self):
        return int(self.lib.dd_sim_launch_count(self._h))

    def segment_info(self, f=0):
        """dict(segment, chunks, active_bricks, occupied_bricks, epoch, linked, n_segments, interval) of state f's segment."""
        out = (ctypes.c_int * 8)()
        self._check(self.lib.dd_sim_segment_info(self._h, int(f), out, self.stream))
        keys = ("segment", "chunks", "active_bricks", "occupied_bricks", "epoch", "linked", "n_segments", "interval")
        return dict(zip(keys, [int(v) for v in out]))

    # ---- readback (original particle order)
    def get_state(self, f, names=("x", "v", "F", "C"), device=False):
        shp = dict(x=3, v=3, F=9, C=9)
        out = {k: self._empty((self.E, self.N, shp[k]), device) for k in names}
        self._check(self.lib.dd_sim_get_state(self._h, f, _ptr(out.get("x")), _ptr(out.get("v")), _ptr(out.get("F")), _ptr(out.get("C")), self.stream))
        return out

    def get_state_grad(self, f, names=("x", "v", "F", "C"), device=False):
        shp = dict(x=3, v=3, F=9, C=9)
        out = {k: self._empty((self.E, self.N, shp[k]), device) for k in names}
        self._check(self.lib.dd_sim_get_state_grad(self._h, f, _ptr(out.get("x")), _ptr(out.get("v")), _ptr(out.get("F")), _ptr(out.get("C")), self.stream))
        return out

    def get_pose_grads(self, f0, count, device=False):
        gp, gr = self._empty((count, self.E, self.nb, 3), device), self._empty((count, self.E, self.nb, 4), device)
        if self.nb:
            self._check(self.lib.dd_sim_get_pose_grads(self._h, f0, count, _ptr(gp), _ptr(gr), self.stream))
        return gp, gr

    def add_pose_grads(self, f, gpos=None, grot=None):
        self._check(self.lib.dd_sim_add_pose_grads(self._h, f, _ptr(gpos), _ptr(grot), self.stream))
        if _is_host(gpos, grot):
            self.sync()

    def compute_dist(self, f, device=False):
        d = self._empty((self.E, self.N, self.nb), device)
        if self.nb:
            self._check(self.lib.dd_sim_compute_dist(self._h, f, _ptr(d), self.stream))
        return d

    def compute_dist_grad(self, f, dist_grad):
        if isinstance(dist_grad, np.ndarray):
            dist_grad = np.ascontiguousarray(dist_grad, np.float32)
        self._check(self.lib.dd_sim_compute_dist_grad(self._h, f, _ptr(dist_grad), self.stream))
        if _is_host(dist_grad):
            self.sync()

    def get_obs(self, f):
        """Particle observation of state f as one CUDA tensor (E, N, 6 + nb) = [x | v | dist] (mpm/torch_wrapper.py:46-66)."""
        obs = self._empty((self.E, self.N, 6 + self.nb), True)
        self._check(self.lib.dd_sim_get_obs(self._h, f, _ptr(obs), self.stream))
        return obs

    def add_obs_grad(self, f, gobs):
        """Adjoint of get_obs: gobs is a contiguous CUDA tensor (E, N, 6 + nb) (mpm/torch_wrapper.py:68-105)."""
        self._check(self.lib.dd_sim_add_obs_grad(self._h, f, _ptr(gobs), self.stream))

    def profile_substep(self, f, reps=5):
        """[(kernel label, ms)] for one forward + backward substep (needs a gradient seeded for state f+1)."""
        ms = (ctypes.c_float * 32)()
        names = ctypes.create_string_buffer(2048)
        n = ctypes.c_int(0)
        self._check(self.lib.dd_sim_profile_substep(self._h, f, reps, ms, names, 2048, ctypes.byref(n), self.stream))
        labels = names.value.decode().strip().split("\n")
        return [(labels[i], float(ms[i])) for i in range(n.value)]

    def compute_grid_mass(self, f, ids=None, id=-1, device=False):
        out = self._empty((self.E,) + self.grid_dim, device)
        ids_a = None if ids is None else np.ascontiguousarray(ids, np.int32)
        self._check(self.lib.dd_sim_compute_grid_mass(self._h, f, None if ids_a is None else ids_a.ctypes.data, int(id), _ptr(out), self.stream))
        return out

    def compute_grid_mass_grad(self, f, grid_m_grad, ids=None, id=-1):
        g = np.ascontiguousarray(grid_m_grad, np.float32) if isinstance(grid_m_grad, np.ndarray) else grid_m_grad
        ids_a = None if ids is None else np.ascontiguousarray(ids, np.int32)
        self._check(self.lib.dd_sim_compute_grid_mass_grad(self._h, f, None if ids_a is None else ids_a.ctypes.data, int(id), _ptr(g), self.stream))
        self.sync()
