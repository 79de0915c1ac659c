"""Particle initialisation of the DexDeform scenes (mpm/shapes.py:50-170): box, sphere and cylinder samplers with the
reference's fixed seed 0, so the shipped environments start from the same particles.  (The open3d mesh sampler of
shapes.py:173-338 is not used by any shipped YAML and is out of scope.)"""
import numpy as np


def _eval(v):
    return eval(v, {"np": np}) if isinstance(v, str) else v


class Shapes:
    COLORS = [(127 << 16) + 127, (127 << 8), 127, 127 << 16]

    def __init__(self, cfg, dim=3):
        self.objects, self.colors, self.object_id, self.mu_lam_yield, self.dim = [], [], [], [], dim
        state = np.random.get_state()
        np.random.seed(0)  # shapes.py:66
        for item in cfg:
            kw = {k: _eval(v) for k, v in item.items() if k != "shape"}
            getattr(self, "add_" + item["shape"])(**kw)
        np.random.set_state(state)

    @staticmethod
    def get_n_particles(volume):
        return max(int(volume / 0.2 ** 3) * 10000, 1)

    def add_object(self, particles, color=None, init_rot=None, **extras):
        if init_rot is not None:
            from .rotations import quat2mat
            origin = particles.mean(axis=0)
            particles = (particles[:, :self.dim] - origin) @ quat2mat(init_rot).T + origin
        self.objects.append(particles[:, :self.dim])
        if color is None or isinstance(color, int):
            c = np.zeros(len(particles), np.int32)
            c[:] = self.COLORS[len(self.objects) - 1] if color is None else color
            color = c
        self.object_id.append([len(self.object_id)] * len(particles))
        self.colors.append(color)
        if any(k in extras for k in ("yield_stress", "E", "nu")):
            ys, E, nu = extras.get("yield_stress", 30.0), extras.get("E", 5000.0), extras.get("nu", 0.2)
            self.mu_lam_yield.append(np.zeros((len(particles), 3)) + np.array([E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu)), ys]))

    def add_box(self, init_pos, width, n_particles=10000, color=None, init_rot=None, **extras):
        width = np.array([width] * self.dim) if isinstance(width, float) else np.array(width)
        if n_particles is None:
            n_particles = self.get_n_particles(np.prod(width))
        p = (np.random.random((n_particles, self.dim)) * 2 - 1) * (0.5 * width) + np.array(init_pos)
        self.add_object(p, color, init_rot=init_rot, **extras)

    def add_sphere(self, init_pos, radius, n_particles=10000, color=None, init_rot=None, **extras):
        if n_particles is None:
            n_particles = self.get_n_particles((radius ** 3) * 4 * np.pi / 3)
        p = np.random.normal(size=(n_particles, self.dim))
        p /= np.linalg.norm(p, axis=-1, keepdims=True)
        u = np.random.random(size=(n_particles, 1)) ** (1.0 / self.dim)
        self.add_object(p * u * radius + np.array(init_pos)[:self.dim], color, init_rot=init_rot, **extras)

    def add_cylinder(self, init_pos, h, r, n_particles=10000, color=None, init_rot=None, **extras):
        """Rejection sampling of a y-axis cylinder (half height h, radius r) from its bounding box; consumes the random stream
        exactly like shapes.py:10-47,150-159 (each round draws only the particles still missing)."""
        half, rh = np.array([r, h, r]), np.array([[r, h]])
        length = lambda a: np.sqrt(np.einsum("ij,ij->i", a, a) + 1e-14)
        p = np.ones((n_particles, 3)) * 5
        remain = n_particles
        while remain > 0:
            cand = (np.random.random((remain, 3)) * 2 - 1) * half
            d = np.abs(np.stack([length(cand[:, [0, 2]]), cand[:, 1]], axis=1)) - rh
            sdf = np.minimum(np.maximum(d[:, 0], d[:, 1]), 0.0) + length(np.maximum(d, 0.0))
            ok = sdf <= 0
            cnt = int(ok.sum())
            start = n_particles - remain
            p[start:start + cnt] = cand[ok]
            remain -= cnt
        self.add_object(p + np.array(init_pos), color, init_rot=init_rot, **extras)

    def get(self):
        assert len(self.objects) > 0, "please add at least one shape into the scene"
        mly = None if not self.mu_lam_yield else np.concatenate(self.mu_lam_yield, axis=0)
        return np.concatenate(self.objects), np.concatenate(self.colors), np.concatenate(self.object_id), mly
