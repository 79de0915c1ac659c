"""Hand layer: MJCF tables, rotation helpers, torch FK (CPU) and the device FK kernel (GPU) against the torch FK."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from dexdeform_b200.hand import HandKinematics, rigid_body_motion_hand
from dexdeform_b200.mujoco_parser import HandTables, default_assets_dir
from dexdeform_b200.robots import JOINTS, N_ACTUATORS, actuator_of_joint, joint_limits
from dexdeform_b200.rotations import axis_angle_to_matrix, euler2mat, matrix_to_quaternion, quaternion_to_matrix

FIX = os.path.join(ROOT, "tests", "golden", "shadow_tables.npz")
HAVE_ASSETS = os.path.isdir(os.path.join(default_assets_dir(), "robots", "shadow"))


def tables(tag="rh15"):
    z = np.load(FIX)
    return HandTables(**{k.split(".", 1)[1]: (int(z[k]) if k.endswith("n_hands") else z[k]) for k in z.files if k.startswith(tag + ".")})


def test_joint_tables():
    assert len(JOINTS) == 24 and N_ACTUATORS == 20
    m = actuator_of_joint()
    assert m[4] == m[5] == 4 and m[23] == 19 and sorted(set(m)) == list(range(20))
    lim = joint_limits()
    assert np.allclose(lim[4], [0.0, 3.1416 / 2]) and np.allclose(lim[0], [-0.4887, 0.1396])  # coupled joints get half the range


def test_rotation_helpers_roundtrip():
    g = torch.Generator().manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(200, 4, generator=g), dim=-1)
    R = quaternion_to_matrix(q)
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3).expand(200, 3, 3), atol=1e-5)
    q2 = matrix_to_quaternion(R)
    assert torch.allclose(torch.minimum((q - q2).norm(dim=-1), (q + q2).norm(dim=-1)), torch.zeros(200), atol=2e-6)
    aa = torch.randn(50, 3, generator=g)
    Ra = axis_angle_to_matrix(aa)
    ang = aa.norm(dim=-1)
    assert torch.allclose((Ra.diagonal(dim1=-2, dim2=-1).sum(-1) - 1) / 2, torch.cos(ang), atol=1e-5)   # trace = 1 + 2 cos
    assert torch.allclose(torch.einsum("nij,nj->ni", Ra, aa), aa, atol=1e-5)                              # the axis is invariant
    assert np.allclose(euler2mat(0.3, 0.0, 0.0), [[1, 0, 0], [0, np.cos(0.3), -np.sin(0.3)], [0, np.sin(0.3), np.cos(0.3)]])
    assert np.allclose(euler2mat(0.0, 0.0, np.pi) @ [1, 0, 0], [-1, 0, 0], atol=1e-12)


@pytest.mark.skipif(not HAVE_ASSETS, reason="DexDeform asset files not present")
def test_parser_reproduces_fixture_tables():
    from dexdeform_b200.mujoco_parser import hand_tables, load_hand
    t = hand_tables([load_hand("right_hand", 1.5)])
    ref = tables("rh15")
    for k, v in ref.__dict__.items():
        assert np.array_equal(np.asarray(getattr(t, k)), np.asarray(v)), k
    m = load_hand("right_hand", 1.5)
    kinds = [p.kind for p in m.primitives]
    assert len(kinds) == 19 and kinds.count("box") == 3 and kinds.count("capsule") == 16
    assert [len(c) for c in m.chains] == [12, 12, 12, 14, 14]


def test_torch_fk_geometry_and_gradients():
    t = tables()
    kin = HandKinematics(t)
    base = torch.tensor(t.root_frame, dtype=torch.float32)[None]
    pos0, rot0 = kin.forward(base, torch.zeros(1, 1, 24))
    assert pos0.shape == (1, 19, 3) and torch.allclose(rot0.norm(dim=-1), torch.ones(1, 19), atol=1e-5)
    # wrist capsule sits at the wrist frame origin; bending FFJ3..FFJ0 moves only the first finger's primitives (3, 4, 5)
    assert torch.allclose(pos0[0, 0], base[0, 0, :3, 3], atol=1e-6)
    q = torch.zeros(1, 1, 24)
    q[..., 3] = 0.5
    pos1, _ = kin.forward(base, q)
    moved = ((pos1 - pos0).norm(dim=-1) > 1e-6)[0]
    assert moved.nonzero().flatten().tolist() == [3, 4, 5]
    # a rigid motion of the wrist moves every primitive rigidly
    T = torch.eye(4)
    T[:3, :3] = axis_angle_to_matrix(torch.tensor([0.2, -0.4, 0.3]))
    T[:3, 3] = torch.tensor([0.1, 0.2, -0.3])
    pos2, _ = kin.forward((T @ base[0])[None], torch.zeros(1, 1, 24))
    assert torch.allclose(pos2[0], (T[:3, :3] @ pos0[0].T).T + T[:3, 3], atol=1e-5)
    # differentiable w.r.t. joints and base
    q = torch.full((1, 1, 24), 0.1, requires_grad=True)
    p, r = kin.forward(base, q)
    (p.sum() + r.sum()).backward()
    assert torch.isfinite(q.grad).all() and q.grad.abs().sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("tag,fixed_base,E", [("rh15", False, 1), ("rh25", True, 3), ("dual15", False, 2)])
def test_device_fk_matches_torch_fk(tag, fixed_base, E):
    from dexdeform_b200.engine import FusedSim
    from dexdeform_b200.hand import DeviceFK, read_poses
    t = tables(tag)
    nh, ng = t.n_hands, len(t.geom_joint)
    S = 40
    scale = np.array([0.33 * 0.002] * 20 + ([0.0] * 6 if fixed_base else [0.01] * 3 + [0.015] * 3))
    eng = FusedSim(E, 64, nh * ng, (32, 32, 32), 1 / 32, 1e-4, S)
    fk = DeviceFK(t, scale)
    g = torch.Generator().manual_seed(3)
    kin = HandKinematics(t, "cuda")
    base = torch.tensor(t.root_frame, dtype=torch.float32, device="cuda")[None].repeat(E, 1, 1, 1)
    base[..., :3, :3] = axis_angle_to_matrix(torch.randn(E, nh, 3, generator=g).cuda() * 0.7) @ base[..., :3, :3]
    q0 = (torch.rand(E, nh, 24, generator=g).cuda() * 0.3)
    act = (torch.rand(E, nh, 26, generator=g).cuda() * 3 - 1.5)       # beyond [-1, 1]: exercises the clamp
    nb_, nq_ = fk.run(eng, 0, S, base, q0, act, has_base_action=not fixed_base)
    pos_d, rot_d = read_poses(eng, 1, S)
    for e in range(E):
        sc = torch.tensor(scale, dtype=torch.float32, device="cuda")
        nbase = rigid_body_motion_hand(base[e], act[e, :, -6:] * sc[None, -6:], S) if not fixed_base else base[e][None].expand(S, -1, -1, -1)
        a = (act[e, :, :20].clamp(-1, 1) * sc[None, :20])[:, kin.action_map]
        nq = (q0[e][None] + a[None] * (torch.arange(S, device="cuda")[:, None, None] + 1)).clamp(kin.q_lower, kin.q_upper)
        pos_t, rot_t = kin.forward(nbase, nq)
        assert torch.allclose(pos_d[:, e], pos_t, atol=2e-6), (pos_d[:, e] - pos_t).abs().max()
        sign = torch.sign((rot_d[:, e] * rot_t).sum(-1, keepdim=True))
        assert torch.allclose(rot_d[:, e] * sign, rot_t, atol=5e-6), (rot_d[:, e] * sign - rot_t).abs().max()
        assert torch.allclose(nb_[e], nbase[-1], atol=2e-6) and torch.allclose(nq_[e], nq[-1], atol=1e-7)
    eng.close()


@pytest.mark.skipif(not os.path.isfile("/root/reference/mpm/shapes.py"), reason="reference checkout not present")
def test_shapes_reproduce_reference_particles_bit_for_bit():
    """mpm/shapes.py imports open3d at module level only for its (unused) mesh sampler; with an empty stand-in module the
    reference's own box / cylinder / sphere samplers run here and must produce the very same particles (seed 0)."""
    import importlib.util
    import sys
    import types
    from dexdeform_b200.shapes import Shapes
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    spec = importlib.util.spec_from_file_location("ref_shapes", "/root/reference/mpm/shapes.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cfgs = [
        [dict(shape="box", width="(0.09, 0.09, 0.09)", init_pos="(0.49, 0.22, 0.45)", n_particles=10000)],               # lift_box.yml
        [dict(shape="cylinder", h=0.003, r=0.1, init_pos="(0.5, 0.38, 0.6)")],                                             # flip.yml
        [dict(shape="sphere", radius=0.05, init_pos="(0.5, 0.2, 0.5)", n_particles=3000, E=4000.0, yield_stress=80.0),
         dict(shape="box", width=0.05, init_pos="(0.3, 0.1, 0.3)", n_particles=500)],
    ]
    for cfg in cfgs:
        a, b = ref.Shapes(cfg).get(), Shapes(cfg).get()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert (a[3] is None and b[3] is None) or np.array_equal(a[3], b[3])


@pytest.mark.gpu
def test_hand_simulator_step_and_gradmodel():
    """HandSimulator built from the fixture tables (the MJCF files do not exist on the GPU box): lift_box-like scene, two env
    steps forward-only with device FK, then a differentiable step through GradModel with gradients w.r.t. the action."""
    from dexdeform_b200.hand import HandSimulator
    from dexdeform_b200.torch_wrapper import GradModel
    from dexdeform_b200.hand import HandEnv
    t = tables("rh15")
    nb = len(t.prim_type)
    n = 4000
    cfg = dict(n_particles=n, E=5e3, nu=0.2, yield_stress=50.0, ground_friction=0.3, quality=1, max_steps=90, gravity=(0.0, -2.0, 0.0), fixed_base=False)
    sim = HandSimulator(nb, {"tables": t}, cfg=cfg)
    assert sim.substeps == 40 and abs(sim.dt - 5e-5) < 1e-12 and sim.grid_dim == (64, 64, 64)
    sim.init_bodies(t.prim_type.astype(np.float32), np.full(nb, 666.0, np.float32), np.full(nb, 0.9, np.float32), np.zeros(nb, np.float32), t.prim_size,
                    action_scales=[()] * nb)
    rng = np.random.default_rng(0)
    x = ((rng.random((n, 3)) * 2 - 1) * 0.045 + np.array([0.49, 0.22, 0.45])).astype(np.float32)
    root = HandEnv.get_root_matrix((0.5, 0.2, 0.3), (0.0, 0.0, np.pi))[None]
    state = (x, np.zeros((n, 3), np.float32), np.tile(np.eye(3, dtype=np.float32)[None], (n, 1, 1)), np.zeros((n, 3, 3), np.float32), np.float32(root),
             np.zeros((1, 24), np.float32))
    sim.set_state(0, state)
    st = sim.get_state(0)
    assert len(st) == 4 + nb + 2 and st[-2].shape == (1, 4, 4) and st[-1].shape == (1, 24)
    act = np.zeros((1, 26), np.float32)
    act[0, :20] = 0.5
    act[0, 21] = -0.5
    sim.step(act)
    sim.step(act)
    x2 = sim.get_x(0)
    assert np.isfinite(x2).all() and x2[:, 1].mean() < x[:, 1].mean()
    assert float(sim.joint_rot[0][0, 3]) > 0.0                     # FFJ2 closed by the positive actuator command
    assert float(sim.base_pose[0][0, 1, 3]) < 0.2                  # wrist translated down
    # differentiable step
    model = GradModel(sim, return_grid=())
    model.zero_grad()
    a = torch.tensor(act, device="cuda:0", requires_grad=True)
    obs = model.get_obs(0, "cuda:0")
    obs = model.forward(0, a, *obs)
    loss = obs[0][:, 1].mean() + 0.1 * obs[0][:, 6:].mean()
    loss.backward()
    assert torch.isfinite(a.grad).all() and a.grad.abs().sum() > 0


# ---- pinned against the reference's own parser and kinematics code (tests/golden/make_hand_ref.py ran mpm/mujoco_parser.py and
# ---- mpm/hand.py from the checkout; only the pytorch3d / transforms3d conversions were stand-ins)
REF_FIX = os.path.join(ROOT, "tests", "golden", "shadow_ref.npz")


def ref_case(tag):
    z = np.load(REF_FIX)
    return {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(tag + ".")}


@pytest.mark.parametrize("tag", ["rh15", "rh25", "dual15"])
def test_tables_equal_reference_parser(tag):
    t, r = tables(tag), ref_case(tag)
    nh = int(r["n_hands"])
    assert t.n_hands == nh
    assert np.array_equal(t.prim_type, r["prim_type"])                                   # order and kind of the 19 / 38 primitives
    assert np.allclose(t.prim_size, r["prim_args"], rtol=0, atol=1e-7)                   # args as cuda_env.parse_tools passes them
    assert np.allclose(t.root_frame, r["root_frame"], atol=1e-6)
    assert np.allclose(t.joint_pos, r["joint_pos"], atol=1e-7) and np.allclose(t.joint_axis, r["joint_axis"], atol=1e-7)
    assert np.array_equal(np.asarray(t.geom_joint), r["geom_index"])
    assert np.allclose(np.asarray(t.geom_local).reshape(r["geometries"].shape), r["geometries"], atol=1e-6)   # incl. the capsule Rx(90 deg) fix-up
    lim = joint_limits()
    assert np.allclose(lim[:, 0], r["q_lower"], atol=1e-7) and np.allclose(lim[:, 1], r["q_upper"], atol=1e-7)
    assert np.array_equal(actuator_of_joint(), r["action_map"])


@pytest.mark.parametrize("tag,fixed_base", [("rh15", False), ("rh25", True), ("dual15", False)])
def test_torch_fk_equals_reference_fk(tag, fixed_base):
    """One env step of joint-velocity control (hand.py:383-428) and the rest pose: our tables + FK against poses computed by the
    reference's HandSimulator code."""
    from dexdeform_b200.robots import DEFAULT_INITIAL_QPOS
    t, r = tables(tag), ref_case(tag)
    kin = HandKinematics(t)
    S = int(r["substeps"])
    sc = torch.tensor(r["action_scale"])
    base, q0, act = torch.tensor(r["in_base"]), torch.tensor(r["in_q"]), torch.tensor(r["in_action"])
    assert np.allclose(sc.numpy(), [0.33 * 0.002] * 20 + ([0.0] * 6 if fixed_base else [0.01] * 3 + [0.015] * 3))
    nbase = rigid_body_motion_hand(base, act[:, -6:] * sc[None, -6:], S)
    a = (act[:, :20].clamp(-1, 1) * sc[None, :20])[:, kin.action_map]
    nq = (q0[None] + a[None] * (torch.arange(S)[:, None, None] + 1)).clamp(kin.q_lower, kin.q_upper)
    pos, rot = kin.forward(nbase, nq)
    assert np.allclose(pos.numpy(), r["out_pos"], atol=2e-6), np.abs(pos.numpy() - r["out_pos"]).max()
    sign = np.sign((rot.numpy() * r["out_rot"]).sum(-1, keepdims=True))
    assert np.allclose(rot.numpy() * sign, r["out_rot"], atol=5e-6)
    assert np.allclose(nbase[-1].numpy(), r["out_base"], atol=2e-6) and np.allclose(nq[-1].numpy(), r["out_q"], atol=1e-7)
    q_def = torch.tensor([DEFAULT_INITIAL_QPOS[j] for j in JOINTS], dtype=torch.float32)[None].expand(t.n_hands, -1)
    assert np.allclose(q_def.numpy(), r["default_qpos"])
    p0, r0 = kin.forward(torch.tensor(t.root_frame, dtype=torch.float32)[None], q_def[None])
    assert np.allclose(p0.numpy(), r["rest_pos"], atol=2e-6)


def test_cuda_env_parse_tools_matches_reference_semantics():
    """mpm/cuda_env.py:51-103: softness is 666 whatever the entry says, Box args = (hx, hy, hz, 0), Capsule args = (r, half length, 0, 0)."""
    from dexdeform_b200.cuda_env import CudaEnv
    env = CudaEnv.__new__(CudaEnv)
    n, kw = env.parse_tools([dict(shape="Box", size=(0.1, 0.2, 0.3), round=0.01, friction=0.5, softness=3.0, init_pos=(0.5, 0.5, 0.5)),
                             dict(shape="Capsule", size=(0.02, 0.06), round=0, action=dict(dim=6, scale=(0.01,) * 6))])
    assert n == 2 and kw["types"] == [0, 1] and kw["softness"] == [666.0, 666.0] and kw["mu"] == [0.5, 0.9] and kw["round"] == [0.01, 0]
    assert kw["args"] == [[0.1, 0.2, 0.3, 0], [0.02, 0.06, 0, 0]] and kw["action_scales"] == [(), (0.01,) * 6]
    assert kw["pos"] == [(0.5, 0.5, 0.5), (0.3, 0.3, 0.3)] and kw["rot"] == [(1.0, 0.0, 0.0, 0.0)] * 2
    d = env.default_tool_config()
    assert d.friction == 0.9 and d.action.dim == 0 and d["action"]["scale"] == () and d.mass == 1.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/mpm"), reason="reference checkout not present")
def test_reference_fixture_is_current():
    """Re-runs the generator (the reference's parser + kinematics under stubs) and compares with the committed fixture."""
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ)
        src = open(os.path.join(ROOT, "tests", "golden", "make_hand_ref.py")).read().replace('path = os.path.join(HERE, "shadow_ref.npz")', f'path = os.path.join({d!r}, "shadow_ref.npz")')
        script = os.path.join(d, "gen.py")
        open(script, "w").write(src.replace("HERE = os.path.dirname(os.path.abspath(__file__))", f"HERE = {os.path.join(ROOT, 'tests', 'golden')!r}"))
        subprocess.run([sys.executable, script], check=True, capture_output=True, env=env)
        a, b = np.load(os.path.join(d, "shadow_ref.npz")), np.load(REF_FIX)
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert np.allclose(a[k], b[k], atol=1e-6), k


ENV_FIX = os.path.join(ROOT, "tests", "golden", "env_ref.npz")


@pytest.mark.skipif(not HAVE_ASSETS, reason="DexDeform asset files not present")
@pytest.mark.parametrize("name", ["folding", "rope", "bun", "dumpling", "wrap", "flip", "lift_box"])
def test_env_description_equals_reference_for_every_shipped_yaml(name):
    """HandEnv.describe (everything mpm/hand.py:436-476 derives from an env YAML before a simulator exists) against the same
    quantities computed by the reference's own Shapes / parse_sim_cfg / parse_manip_cfgs (tests/golden/make_env_ref.py)."""
    import hashlib
    from dexdeform_b200.hand import HandEnv, load_env_cfg
    z = np.load(ENV_FIX)
    r = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}
    env = HandEnv.__new__(HandEnv)
    env.assets_dir = None
    d = env.describe(load_env_cfg(name))
    assert d["n_particles"] == int(r["n_particles"]) and len(d["objects"]) == int(r["n_objects"])
    obj = np.ascontiguousarray(d["objects"], np.float32)
    assert np.array_equal(obj[:8], r["objects_head"]) and hashlib.sha256(obj.tobytes()).hexdigest() == str(r["objects_sha256"])   # bit for bit
    n, kw = env.parse_tools(d["primitives"])
    assert n == len(r["prim_type"]) and kw["types"] == r["prim_type"].tolist() and kw["softness"] == [666.0] * n
    assert np.allclose(np.float32(kw["args"]), r["prim_args"], rtol=0, atol=1e-7) and np.allclose(kw["mu"], r["prim_friction"])
    assert d["hand_cfg"]["n_hands"] == int(r["n_hands"])
    assert np.allclose(d["root_matrix"], r["root_matrix"], atol=1e-12) and np.allclose(d["joint_pos"], r["joint_pos"], atol=0)


def _env_cfg(mode="rh", n=3000, max_steps=45):
    """lift_box.yml's sections with a smaller block (the YAML files themselves do not exist on the GPU box)."""
    manip = [dict(hand_idx=0, init_pos="(0.5, 0.2, 0.3)", init_rot="(0., 0., np.pi)", init_qpos="zero")]
    if mode == "dual":
        manip = [dict(hand_idx=0, init_pos="(0.3, 0.25, 0.3)", init_rot="(0., 0., np.pi)", init_qpos="default"),
                 dict(hand_idx=1, init_pos="(0.7, 0.25, 0.3)", init_rot="(0., 0., np.pi)", init_qpos="zero")]
    return {"SIMULATOR": dict(n_particles=n, E=5e3, nu=0.2, yield_stress=50.0, ground_friction=0.3, quality=1, max_steps=max_steps, gravity="(0., -2., 0.)",
                              mode=mode, ctrl_type="vel", scale=1.5, hand_friction=0.9),
            "SHAPES": [dict(shape="box", width="(0.09, 0.09, 0.09)", init_pos="(0.49, 0.22, 0.45)", n_particles=n)],
            "MANIPULATORS": manip}


@pytest.mark.gpu
def test_hand_env_construction_and_sdf_helpers():
    """HandEnv (hand.py:436-649) from a config + fixture tables: initialize, set_single_hand_pose, the signed-distance helpers
    policy/preprocess uses (hand.py:236-341) and the render-only state."""
    from dexdeform_b200.hand import HandEnv, SINGLE_PRIM_RANGE
    env = HandEnv(_env_cfg(), tables=tables("rh15"))
    sim = env.simulator
    st = env.init_state
    assert len(st) == 4 + 19 + 2 and np.array_equal(st[2], np.tile(np.eye(3, dtype=np.float32)[None], (sim.n_particles, 1, 1)))
    assert np.allclose(st[-2][0], HandEnv.get_root_matrix((0.5, 0.2, 0.3), (0.0, 0.0, np.pi)), atol=1e-6) and np.all(st[-1] == 0)
    assert np.abs(st[0] - np.array([0.49, 0.22, 0.45])).max() <= 0.045 + 1e-6       # the block of the SHAPES section
    # body poses of state 0 are the forward kinematics of the configured wrist frame
    kin = HandKinematics(tables("rh15"))
    p0, _ = kin.forward(torch.tensor(st[-2])[None], torch.tensor(st[-1])[None])
    assert np.allclose(np.stack(st[4:4 + 19])[:, :3], p0[0].numpy(), atol=2e-6)
    # signed distances of given points: primitive centres are inside, far points outside; the state is restored afterwards
    centres = np.stack(st[4:4 + 19])[:, :3]
    far = centres + np.array([0.0, 0.3, 0.0])
    d_in, d_out = sim.primitive_sdf_given_p(centres, SINGLE_PRIM_RANGE), sim.primitive_sdf_given_p(far, SINGLE_PRIM_RANGE)
    assert d_in.shape == (19,) and (d_in < 0).all() and (d_out > 0.05).all()
    assert np.array_equal(sim.get_state(0)[0], st[0])
    np.random.seed(0)
    pts, sdf, labels = sim.sample_pts_inside_primitives(200, mode="rh")
    assert pts.shape == (200, 3) and sdf.shape == (200, 19) and labels.shape == (200, 1) and (sdf.min(-1) <= 0).all()
    # render-only state round trip and a new wrist pose
    p, base, q = sim.get_state_render_only(0)
    sim.set_state_render_only(p + 0.01, base, q)
    assert np.allclose(sim.get_x(0), p + 0.01, atol=1e-7)
    env.set_single_hand_pose(0, pos=(0.5, 0.3, 0.3), rot=(0.0, 0.0, np.pi), joint_pos=[0.1] * 24)
    st2 = sim.get_state(0)
    assert np.allclose(st2[-2][0][:3, 3], (0.5, 0.3, 0.3)) and np.allclose(st2[-1][0], 0.1)
    assert not np.allclose(np.stack(st2[4:4 + 19])[:, :3], centres)
    env.set_particle_color(0xff0000)
    # one env step still runs from this state
    sim.step(np.zeros((1, 26), np.float32))
    assert np.isfinite(sim.get_x(0)).all()


@pytest.mark.gpu
def test_dual_hand_env():
    from dexdeform_b200.hand import HandEnv, LH_PRIM_RANGE, RH_PRIM_RANGE
    env = HandEnv(_env_cfg("dual"), tables=tables("dual15"))
    sim = env.simulator
    assert sim.n_hands == 2 and sim.n_bodies == 38
    st = sim.get_state(0)
    assert st[-2].shape == (2, 4, 4) and st[-1].shape == (2, 24) and np.all(st[-1][1] == 0) and np.abs(st[-1][0]).sum() > 0
    env.set_dual_hand_pose(pos=[(0.3, 0.3, 0.3), (0.7, 0.3, 0.3)], rot=[(0.0, 0.0, np.pi)] * 2, joint_pos=[[0.05] * 24, [0.1] * 24])
    st = sim.get_state(0)
    poses = np.stack(st[4:4 + 38])
    assert np.allclose(st[-2][:, :3, 3], [(0.3, 0.3, 0.3), (0.7, 0.3, 0.3)]) and np.allclose(st[-1][1], 0.1)
    assert poses[LH_PRIM_RANGE, 0].mean() < 0.5 < poses[RH_PRIM_RANGE, 0].mean()
    lh = sim.lh_sdf_given_p(poses[LH_PRIM_RANGE, :3])
    rh = sim.rh_sdf_given_p(poses[LH_PRIM_RANGE, :3])
    assert (lh < 0).all() and (rh > 0).all()
    act = np.zeros((2, 26), np.float32)
    act[:, :20] = 0.3
    sim.step(act)
    assert np.isfinite(sim.get_x(0)).all() and float(sim.joint_rot[0][1, 3]) > 0.1


@pytest.mark.gpu
@pytest.mark.parametrize("tag,fixed_base,E", [("rh15", False, 1), ("rh15", True, 2), ("dual15", False, 2)])
def test_device_fk_adjoint_matches_torch_autograd(tag, fixed_base, E, monkeypatch):
    """Action gradients of a two-step rollout through GradModel: kinematics + adjoint on the device (dd_hand_fk / dd_hand_fk_grad)
    against the torch mirror of the reference's FK differentiated by autograd (hand.py:347-428).  The MPM part is the same engine
    in both runs, so any difference is the kinematics adjoint."""
    from dexdeform_b200.hand import HandEnv, HandSimulator
    from dexdeform_b200.torch_wrapper import GradModel
    t = tables(tag)
    nb, nh, n = len(t.prim_type), t.n_hands, 2000
    cfg = dict(n_particles=n, E=5e3, nu=0.2, yield_stress=50.0, ground_friction=0.3, quality=1, max_steps=85, gravity=(0.0, -2.0, 0.0), fixed_base=fixed_base)
    rng = np.random.default_rng(1)
    x = ((rng.random((n, 3)) * 2 - 1) * 0.04 + np.array([0.5, 0.2, 0.45])).astype(np.float32)
    roots = np.stack([HandEnv.get_root_matrix((0.5 + 0.12 * (2 * h - nh + 1), 0.2, 0.3), (0.0, 0.0, np.pi)) for h in range(nh)])
    state = (x, np.zeros((n, 3), np.float32), np.tile(np.eye(3, dtype=np.float32)[None], (n, 1, 1)), np.zeros((n, 3, 3), np.float32), np.float32(roots),
             np.float32(rng.random((nh, 24)) * 0.2))
    act0 = np.float32(rng.uniform(-1.3, 1.3, (2, E, nh, 26)))     # some commands beyond the clamp
    act0[..., 21] = -0.6
    grads, losses = {}, {}
    for mode in ("device", "torch"):
        monkeypatch.setenv("DD_TORCH_FK", "1" if mode == "torch" else "0")
        sim = HandSimulator(nb, {"tables": t}, cfg=cfg, n_envs=E)
        sim.init_bodies(t.prim_type.astype(np.float32), np.full(nb, 666.0, np.float32), np.full(nb, 0.9, np.float32), np.zeros(nb, np.float32), t.prim_size,
                        action_scales=[()] * nb)
        sim.set_state(0, state)
        model = GradModel(sim, return_grid=())
        assert model.use_device_fk == (mode == "device")
        model.zero_grad()
        a = torch.tensor(act0 if E > 1 else act0[:, 0], device="cuda", requires_grad=True)
        obs = model.get_obs(0, "cuda")
        for s in range(2):
            obs = model.forward(s, a[s], *obs)
        loss = obs[0][..., 1].mean() + 0.3 * obs[0][..., 6:].mean() + 0.1 * obs[1][..., :3].sum() + 0.05 * obs[1][..., 3:].square().sum()
        loss.backward()
        grads[mode], losses[mode] = a.grad.detach().cpu().numpy().copy(), float(loss)
        sim.engine.close()
    assert abs(losses["device"] - losses["torch"]) < 1e-5 * max(1.0, abs(losses["torch"]))
    gd, gt = grads["device"].ravel(), grads["torch"].ravel()
    assert np.abs(gt).max() > 0
    rel = np.linalg.norm(gd - gt) / np.linalg.norm(gt)
    cos = float(gd @ gt / (np.linalg.norm(gd) * np.linalg.norm(gt)))
    assert rel < 2e-3 and cos > 0.99999, (rel, cos)
    if fixed_base:
        assert np.all(grads["device"][..., 20:] == 0)
