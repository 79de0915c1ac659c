"""Demonstration container round trip (CPU) and batched replay / scoring (GPU) -- SURVEY.md 8f row f2, BASELINE config C."""
import numpy as np
import pytest
import torch

from dexdeform_b200 import demos
from dexdeform_b200.scenes import make_scene


def _state(rng, n=50, nb=3, hand=True):
    x, v = rng.random((n, 3), np.float32), rng.standard_normal((n, 3)).astype(np.float32)
    F = np.tile(np.eye(3, dtype=np.float32), (n, 1, 1)) + 0.01 * rng.standard_normal((n, 3, 3)).astype(np.float32)
    C = rng.standard_normal((n, 3, 3)).astype(np.float32)
    t = [x, v, F, C, rng.random((nb, 3), np.float32), rng.random((nb, 4), np.float32)]
    if hand:
        t += [rng.random((1, 6), np.float32), rng.random((1, 24), np.float32)]
    return tuple(t)


def test_gzip_container_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    demo = {"states": [_state(rng), _state(rng)], "actions": [rng.random((1, 26), np.float32)], "tensor": torch.arange(4.0)}
    path = tmp_path / "demo_0.pkl"
    demos.save_gzip_file(demo, str(path))
    back = demos.load_gzip_file(str(path))
    assert back.keys() == demo.keys()
    for a, b in zip(demo["states"][1], back["states"][1]):
        np.testing.assert_array_equal(a, b)
    assert torch.equal(back["tensor"], demo["tensor"])
    with pytest.raises(ValueError):
        demos.save_gzip_file(demo, str(tmp_path / "demo_0.bin"))


@pytest.mark.parametrize("hand", [False, True])
def test_state_tuple_roundtrip(hand):
    st = _state(np.random.default_rng(1), hand=hand)
    f = demos.split_state(st)
    assert f["F"].shape == (50, 9) and f["C"].dtype == np.float32 and ("base_pose" in f) == hand
    back = demos.join_state(f)
    assert len(back) == len(st)
    for a, b in zip(st, back):
        np.testing.assert_array_equal(np.asarray(a, np.float32), b)
    with pytest.raises(ValueError):
        demos.split_state(st[:3])


def test_chamfer_scores_cpu():
    a = np.zeros((2, 4, 3), np.float32)
    b = a.copy()
    b[1, :, 0] = 0.5
    s = demos.chamfer_scores(a, b, device="cpu")
    assert s.shape == (2,) and float(s[0]) == 0.0 and abs(float(s[1]) - 1.0) < 1e-6


@pytest.mark.gpu
def test_batched_replay_equals_single_env_replays_and_scores():
    from dexdeform_b200.engine import FusedSim
    S, E = 6, 3
    base = make_scene(3000, 32, box_width=(0.12, 0.12, 0.12), steps=S, perturb=0.02, vel_scale=0.3, on_floor=True, seed=10)
    rng = np.random.default_rng(3)
    scenes = []
    for e in range(E):  # same material and body shapes (they are per engine), different particle states and hand trajectories
        sc = dict(base)
        sc["x"] = (base["x"] + np.float32(0.004 * e) * np.array([1, 0, 1], np.float32)).astype(np.float32)
        sc["v"] = (base["v"] + 0.1 * rng.standard_normal(base["v"].shape)).astype(np.float32)
        sc["pos"] = (base["pos"] + np.float32(0.003 * e)).astype(np.float32)
        scenes.append(sc)
    pos = np.stack([sc["pos"][:S + 1] for sc in scenes], axis=1)
    rot = np.stack([sc["rot"][:S + 1] for sc in scenes], axis=1)
    as_state = lambda sc: (sc["x"], sc["v"], sc["F"].reshape(-1, 3, 3), sc["C"].reshape(-1, 3, 3), sc["pos"][0], sc["rot"][0])
    batched = FusedSim.from_scene(scenes[0], n_envs=E, max_steps=S)
    final = demos.replay_batch(batched, [as_state(sc) for sc in scenes], pos, rot, S)
    for e, sc in enumerate(scenes):
        single = FusedSim.from_scene(sc, n_envs=1, max_steps=S)
        single.forward(0, S)
        ref = single.get_state(S, names=("x",))["x"][0]
        single.close()
        np.testing.assert_allclose(final[e], ref, atol=2e-6)  # same kernels; only the reduction order on the grid differs
    scores = demos.score_demos(batched, [as_state(sc) for sc in scenes], pos, rot, S, goal_x=final[0])
    assert scores.shape == (E,) and float(scores[0]) < 1e-6 and float(scores[1]) > 1e-4
    batched.close()
