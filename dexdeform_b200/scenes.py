"""Synthetic scenes of the BASELINE.json shapes.  Pure numpy, deterministic.

A scene is a plain dict: particle state in the reference's order and layouts (mpm/simulator.py:88-145), material
arrays (mpm/simulator.py:381-383), primitive tables (mpm/simulator.py:388-404, mpm/cuda_env.py:76-90) and a pose
trajectory ``pos (S+1, nb, 3)``, ``rot (S+1, nb, 4 wxyz)`` -- entry f is the pose stored in ``states[f]``.

The primitives are a hand-like cluster (16 capsules + 3 boxes, the Shadow hand's primitive mix, hand.py / robot.xml)
placed so that some of them press into the material while it also touches the floor; this exercises every branch
of the grid update (contact, friction, soft influence band, floor Coulomb friction, walls).
"""
import numpy as np


def _unit_quat(rng, n, spread=1.0):
    q = np.concatenate([np.ones((n, 1)), spread * rng.normal(size=(n, 3))], 1)
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


def _qmul(a, b):
    w1, x1, y1, z1 = a.T
    w2, x2, y2, z2 = b.T
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], 1)


def hand_like_bodies(rng, center, extent, nb=19, scale=1.5):
    """tfsr (type, friction, softness=666, round) and args per primitive + initial poses around ``center``."""
    types = np.array([1] * nb, np.float32)
    types[[i for i in (1, 2, 12) if i < nb]] = 0.0  # palm boxes + LF metacarpal box (robot.xml:18-19,99)
    tfsr = np.stack([types, np.full(nb, 0.9, np.float32), np.full(nb, 666.0, np.float32), np.zeros(nb, np.float32)], 1)
    args = np.zeros((nb, 4), np.float32)
    for b in range(nb):
        if types[b] == 0:
            args[b, :3] = np.array([0.011, 0.016, 0.006]) * scale * rng.uniform(0.8, 1.2, 3)
        else:
            args[b, 0] = 0.006 * scale * rng.uniform(0.8, 1.3)   # radius
            args[b, 1] = 0.012 * scale * rng.uniform(0.6, 1.4)   # (half-)length as passed by cuda_env.py:85
    pos = center[None] + rng.uniform(-0.5, 0.5, (nb, 3)) * extent[None]
    pos[:, 1] = center[1] + extent[1] * rng.uniform(0.30, 0.55, nb)  # around the top surface
    return tfsr.astype(np.float32), args, pos.astype(np.float32), _unit_quat(rng, nb, 0.4)


def pose_trajectory(rng, pos0, rot0, steps, dt, speed=0.4, spin=3.0):
    """Poses for states 0..steps: each primitive moves with a constant velocity (mostly downwards) and spin."""
    nb = len(pos0)
    vel = rng.normal(size=(nb, 3)) * 0.3 * speed
    vel[:, 1] = -np.abs(rng.normal(size=nb)) * speed
    omega = rng.normal(size=(nb, 3)) * spin
    pos = np.zeros((steps + 1, nb, 3), np.float32)
    rot = np.zeros((steps + 1, nb, 4), np.float32)
    for f in range(steps + 1):
        t = f * dt
        pos[f] = pos0 + vel * t
        ang = omega * t
        w = np.sqrt((ang * ang).sum(1, keepdims=True) + 1e-16)
        dq = np.concatenate([np.cos(w / 2), ang / w * np.sin(w / 2)], 1)
        q = _qmul(rot0.astype(np.float64), dq)
        rot[f] = q / np.linalg.norm(q, axis=1, keepdims=True)
    return pos, rot


def make_scene(n_particles=10000, grid=64, quality=None, box_center=(0.49, 0.22, 0.45), box_width=(0.09, 0.09, 0.09),
               E=5e3, nu=0.2, yield_stress=50.0, gravity=(0.0, -2.0, 0.0), ground_friction=0.3, nb=19, steps=50, seed=0,
               perturb=0.0, vel_scale=0.0, on_floor=False, hand_scale=1.5):
    """Box of particles + hand-like primitives.  Defaults follow lift_box.yml (BASELINE config A).

    ``perturb`` adds N(0, perturb) noise to F (and perturb*10 to C) so that both the elastic and the plastic branch of
    the return mapping are exercised in single-substep parity tests; ``vel_scale`` gives particles random velocities.
    """
    rng = np.random.default_rng(seed)
    quality = quality if quality is not None else grid / 64.0
    dx = 1.0 / grid
    dt = 0.5e-4 / quality
    center = np.array(box_center, np.float64)
    width = np.array(box_width, np.float64)
    if on_floor:
        center[1] = 3 * dx + 0.5 * width[1] + 0.25 * dx
    x = ((rng.random((n_particles, 3)) * 2 - 1) * (0.5 * width) + center).astype(np.float32)
    v = (rng.normal(size=(n_particles, 3)) * vel_scale).astype(np.float32)
    F = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (n_particles, 1))
    C = np.zeros((n_particles, 9), np.float32)
    if perturb > 0:
        F = (F + rng.normal(size=F.shape) * perturb).astype(np.float32)
        C = (rng.normal(size=C.shape) * perturb * 10).astype(np.float32)
    mu = E / (2 * (1 + nu))
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    p_mass = (1.0 / 64 / 2) ** 2  # mpm/simulator.py:166-167 (independent of dx)
    scene = dict(
        n=n_particles, grid_dim=np.array([grid, grid, grid], np.int32), dx=dx, inv_dx=float(grid), dt=dt, steps=steps,
        ground_friction=float(ground_friction), ground_height=3.0,
        gravity=(np.array(gravity, np.float32) * 30).reshape(1, 3),  # mpm/simulator.py:385
        x=x, v=v, F=F, C=C,
        mass=np.full(n_particles, p_mass, np.float32), vol=np.full(n_particles, p_mass, np.float32),
        mu_lam_yield=np.tile(np.array([[mu, lam, yield_stress]], np.float32), (n_particles, 1)),
        nb=nb,
    )
    if nb > 0:
        tfsr, args, pos0, rot0 = hand_like_bodies(rng, center.astype(np.float32), width.astype(np.float32), nb, hand_scale)
        pos, rot = pose_trajectory(rng, pos0, rot0, steps, dt)
        scene.update(tfsr=tfsr, args=args, pos=pos, rot=rot)
    return scene


# the BASELINE.json configurations
def scene_tutorial(steps=50, **kw):          # config A: lift_box.yml
    return make_scene(10000, 64, steps=steps, **kw)


def scene_flip(n=50000, steps=40, **kw):      # config B: flip.yml scaled to ~50k particles (thin disc approximated by a slab)
    return make_scene(n, 64, box_center=(0.5, 0.38, 0.6), box_width=(0.18, 0.012, 0.18), E=4e3, yield_stress=130.0,
                      gravity=(0.0, -3.0, 0.0), ground_friction=500.0, steps=steps, hand_scale=2.5, **kw)


def scene_highres(n=1000000, steps=80, **kw):  # config D: 1M particles, 128^3, quality 2
    return make_scene(n, 128, box_center=(0.5, 0.3, 0.5), box_width=(0.4, 0.4, 0.4), steps=steps, hand_scale=6.0, **kw)
