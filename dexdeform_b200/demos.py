"""Demonstration files and batched replay (BASELINE config C).

* ``load_gzip_file`` / ``save_gzip_file`` -- the reference's demonstration container (policy/utils/io.py:32-52): a gzip'ed
  pickle (protocol 4) of ``{"states": [state, ...], "actions": [...], ...}``; torch storages inside are mapped to the CPU.
* A *state* is the tuple ``MPMSimulator.get_state`` returns (mpm/simulator.py:232-251), extended by ``HandSimulator`` with the
  hand's base pose and joint angles (mpm/hand.py:183-189): ``(x (N,3), v (N,3), F (N,3,3), C (N,3,3), body_pos (nb,3),
  body_rot (nb,4 wxyz)[, base_pose (n_hands,6|7), joint_rot (n_hands,24)])``.  ``split_state`` / ``join_state`` convert between
  that tuple and named float32 arrays in the engine's layouts.
* ``replay_batch`` -- forward replay of E independent demonstrations as ONE batched launch sequence of the fused engine
  (the reference replays them one environment at a time through ``set_pose`` + ``substep``, mpm/simulator.py:626-634).
* ``chamfer_scores`` -- a dependency-free default score of final against goal particle sets on the device; the reference
  leaves the metric (Sinkhorn / EMD) to ``geomloss``, which callers can plug in through ``score_fn``.
"""
import gzip
import io
import pickle

import numpy as np
import torch

STATE_FIELDS = ("x", "v", "F", "C", "body_pos", "body_rot", "base_pose", "joint_rot")


class _CpuUnpickler(pickle.Unpickler):  # policy/utils/io.py:32-40
    def find_class(self, module, name):
        if module == "torch.storage" and name == "_load_from_bytes":
            return lambda b: torch.load(io.BytesIO(b), map_location="cpu")
        return super().find_class(module, name)


def load_gzip_file(file_name):
    with gzip.open(file_name, "rb") as f:
        return _CpuUnpickler(f).load()


def save_gzip_file(data, file_name):
    if not str(file_name).endswith("pkl"):
        raise ValueError("demonstration files end in .pkl (policy/utils/io.py:50)")
    with gzip.open(file_name, "wb") as f:
        pickle.dump(data, f, protocol=4)


def split_state(state):
    """State tuple -> dict of float32 arrays (F and C flattened to (N, 9) row-major, the engine's layout)."""
    if not 6 <= len(state) <= 8:
        raise ValueError(f"a state has 6 fields (+ base_pose, joint_rot for hands), got {len(state)}")
    out = {}
    for name, a in zip(STATE_FIELDS, state):
        a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
        out[name] = np.ascontiguousarray(a, np.float32)
    n = out["x"].shape[0]
    out["F"], out["C"] = out["F"].reshape(n, 9), out["C"].reshape(n, 9)
    return out


def join_state(fields):
    """Inverse of ``split_state``: the tuple the reference's ``set_state`` accepts (F, C as (N,3,3))."""
    n = fields["x"].shape[0]
    t = [fields["x"], fields["v"], fields["F"].reshape(n, 3, 3), fields["C"].reshape(n, 3, 3), fields["body_pos"], fields["body_rot"]]
    if "base_pose" in fields:
        t += [fields["base_pose"], fields["joint_rot"]]
    return tuple(np.asarray(a, np.float32) for a in t)


def replay_batch(engine, states, poses_pos, poses_rot, n_substeps, f0=0):
    """Replay E demonstrations forward in one batch.

    engine      FusedSim with n_envs = E and max_steps >= n_substeps
    states      E state tuples (or dicts from ``split_state``) with the same particle count: the start of every demonstration
    poses_pos   (n_substeps + 1, E, nb, 3) body positions of every substep boundary (slot f0 ...), poses_rot likewise (.., 4)
    returns     final particle positions (E, N, 3) as a numpy array, original particle order
    """
    E = engine.E
    if len(states) != E:
        raise ValueError(f"{len(states)} demonstrations for an engine with {E} environments")
    fs = [s if isinstance(s, dict) else split_state(s) for s in states]
    stack = lambda k: np.ascontiguousarray(np.stack([f[k] for f in fs]), np.float32)
    engine.set_poses(f0, np.ascontiguousarray(poses_pos, np.float32), np.ascontiguousarray(poses_rot, np.float32))
    engine.set_state(f0, stack("x"), stack("v"), stack("F"), stack("C"))
    engine.forward(f0, n_substeps)
    out = engine.get_state(f0 + n_substeps, names=("x",))["x"]
    engine.sync()
    return out


def chamfer_scores(final_x, goal_x, device="cuda", chunk=4096):
    """Symmetric Chamfer distance per environment: final_x (E, N, 3), goal_x (E, M, 3) or (M, 3) -> (E,) torch tensor."""
    a = torch.as_tensor(final_x, dtype=torch.float32, device=device)
    b = torch.as_tensor(goal_x, dtype=torch.float32, device=device)
    if b.dim() == 2:
        b = b[None].expand(a.shape[0], -1, -1)

    def one_way(p, q):  # mean over p of the distance to the nearest q, chunked so that (chunk, M) fits easily
        acc = torch.zeros(p.shape[0], device=p.device)
        for i in range(0, p.shape[1], chunk):
            acc += torch.cdist(p[:, i:i + chunk], q, compute_mode="donot_use_mm_for_euclid_dist").min(dim=2).values.sum(dim=1)
        return acc / p.shape[1]

    return one_way(a, b) + one_way(b, a)


def score_demos(engine, states, poses_pos, poses_rot, n_substeps, goal_x, n_envs_total=None, score_fn=None, group=None):
    """Config C: replay this rank's demonstrations (its contiguous share of ``n_envs_total``, see ``batch.partition_envs``),
    score them, and gather the scores of all ranks into one (n_envs_total,) tensor."""
    from .batch import gather_scores
    final_x = replay_batch(engine, states, poses_pos, poses_rot, n_substeps)
    local = (score_fn or chamfer_scores)(final_x, goal_x)
    return gather_scores(local, int(n_envs_total) if n_envs_total is not None else local.numel(), group)
