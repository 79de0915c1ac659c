"""GPU tests of the host-side mirrors (MPMSimulator / GradModel): the autograd contract of mpm/torch_wrapper.py, checked
against the same chain driven through the reference's CUDA library with the reference's own per-substep call sequence."""
import numpy as np
import pytest
import torch

from abi1_driver import Abi1Sim
from conftest import cosine, rel_err, rel_l2
from dexdeform_b200.scenes import make_scene
from dexdeform_b200.simulator import MPMSimulator, rigid_body_motion
from dexdeform_b200.torch_wrapper import GradModel
from dexdeform_b200.types import array, float32

pytestmark = pytest.mark.gpu

S, STEPS, N, NB = 6, 2, 3000, 3


def build_scene():
    sc = make_scene(N, 32, box_width=(0.14, 0.1, 0.14), steps=S * STEPS, perturb=0.0, vel_scale=0.0, on_floor=True, seed=31, nb=NB)
    sc["pos"][0, :, 1] -= 0.004  # press the tools into the block so that every step has contact
    return sc


def make_sim(sc, **kw):
    sim = MPMSimulator(NB, ground_friction=sc["ground_friction"], gravity=tuple(sc["gravity"].reshape(3) / 30), n_particles=N, dx=sc["dx"],
                       dt=sc["dt"], max_steps=S * STEPS, substeps=S, yield_stress=50.0, **kw)
    sim.init_bodies(sc["tfsr"][:, 0], sc["tfsr"][:, 2], sc["tfsr"][:, 1], sc["tfsr"][:, 3], sc["args"], action_scales=[[0.01] * 3 + [0.02] * 3] * NB,
                    pos=sc["pos"][0], rot=sc["rot"][0])
    sim.set_state(0, (sc["x"], sc["v"], sc["F"].reshape(-1, 3, 3), sc["C"].reshape(-1, 3, 3)) + tuple(np.r_[p, r] for p, r in zip(sc["pos"][0], sc["rot"][0])))
    return sim


def reference_chain(ref_gpu, sc, action, dist_weight):
    """The reference's autograd chain re-enacted with its own library: torch FK -> poses -> substeps -> loss -> substep_grads."""
    sim = Abi1Sim(ref_gpu, sc, S * STEPS)
    scale = torch.tensor([[0.01] * 3 + [0.02] * 3] * NB, dtype=torch.float32)
    pos_rot = (torch.tensor(sc["pos"][0]), torch.tensor(sc["rot"][0]))
    poses = []
    for s in range(STEPS):
        a = action[s].reshape(-1, 6).clamp(-1.0, 1.0) * scale
        pos, rot = rigid_body_motion(pos_rot, a[None].expand(S, -1, -1) * (torch.arange(S)[:, None, None] + 1) / S)
        pos_rot = (pos[-1], rot[-1])
        poses.append((pos, rot))
        for i in range(S):
            f = s * S + i
            sim.states[f + 1]["body_pos"].upload(pos[i].detach().numpy())
            sim.states[f + 1]["body_rot"].upload(rot[i].detach().numpy())
            sim.substep(f)
    fe = S * STEPS
    x = sim.get(fe, "x")["x"]
    dist = array(dtype=float32, length=N * NB, library=ref_gpu)
    sim.compute_dist(sim.states[fe], dist, dist, 0)
    sim.sync()
    d = dist.download().reshape(N, NB)
    loss = -x[:, 1].mean() + dist_weight * d.mean()
    gx = np.zeros((N, 3), np.float32)
    gx[:, 1] = -1.0 / N
    sim.states[fe]["x_grad"].upload(gx)
    dg = array(dtype=float32, length=N * NB, library=ref_gpu)
    dg.upload(np.full(N * NB, dist_weight / (N * NB), np.float32))
    sim.compute_dist(sim.states[fe], dist, dg, 1)
    for f in range(fe - 1, -1, -1):
        sim.substep_grad(f)
    sim.sync()
    total = 0.0
    for s in range(STEPS):
        gp = np.stack([sim.get(s * S + i + 1, "body_pos_grad")["body_pos_grad"] for i in range(S)])
        gr = np.stack([sim.get(s * S + i + 1, "body_rot_grad")["body_rot_grad"] for i in range(S)])
        total = total + (poses[s][0] * torch.tensor(gp)).sum() + (poses[s][1] * torch.tensor(gr)).sum()
    total.backward()
    return float(loss), action.grad.clone()


def test_gradmodel_trajectory_gradient_matches_reference_chain(ref_gpu):
    sc = build_scene()
    rng = np.random.default_rng(0)
    a0 = np.float32(rng.uniform(-0.8, 0.8, (STEPS, NB, 6)))
    a0[:, :, 1] = -0.9  # push down
    w = 0.3
    act_ref = torch.tensor(a0, requires_grad=True)
    loss_ref, grad_ref = reference_chain(ref_gpu, sc, act_ref, w)

    sim = make_sim(sc, svd_mode=1)
    model = GradModel(sim, return_grid=())
    model.zero_grad()
    action = torch.tensor(a0, device="cuda:0", requires_grad=True)
    obs = model.get_obs(0, "cuda:0")
    assert obs[0].shape == (N, 6 + NB) and obs[1].shape == (NB, 7)
    for j in range(STEPS):
        obs = model.forward(j, action[j], *obs)
    loss = -obs[0][:, 1].mean() + w * obs[0][:, 6:].mean()
    loss.backward()
    assert abs(float(loss) - loss_ref) < 1e-5 * max(1.0, abs(loss_ref))
    g = action.grad.cpu()
    assert torch.isfinite(g).all() and g.abs().max() > 0
    assert rel_l2(g.numpy(), grad_ref.numpy()) < 1e-2, rel_l2(g.numpy(), grad_ref.numpy())   # action gradients: rel-L2 <= 1e-2
    assert cosine(g.numpy(), grad_ref.numpy()) > 0.999


def test_gradmodel_grid_mass_observation(oracle_lib):
    sc = build_scene()
    sim = make_sim(sc)
    model = GradModel(sim)  # default return_grid=(-1,) like the tutorial (mpm/torch_wrapper.py:8)
    obs = model.get_obs(0, "cuda:0")
    assert len(obs) == 3 and obs[2].shape == (32, 32, 32)
    assert abs(float(obs[2].sum()) - float(sc["mass"].sum())) < 1e-5 * float(sc["mass"].sum())
    a = Abi1Sim(oracle_lib, sc, 1)
    gm = array(dtype=float32, length=32 ** 3, library=oracle_lib)
    ids = array(dtype=int, length=N, library=oracle_lib)
    gd = a.grid_dim
    oracle_lib.particle2mass(a.states[0]["x"].data_ptr, a.mass.data_ptr, a.grid_lower.data_ptr, gd, a.dx, a.inv_dx, gm.data_ptr, gm.data_ptr,
                             a.states[0]["x_grad"].data_ptr, ids.data_ptr, -1, 0, N, None)
    assert rel_err(obs[2].cpu().numpy().reshape(-1), gm.download()) < 1e-5
    # adjoint: d(sum(w * grid_m))/dx against the oracle
    rng = np.random.default_rng(1)
    wgt = np.float32(rng.normal(size=32 ** 3))
    sim.engine.zero_grad(0)
    sim.compute_grid_mass(0, -1, backward_grad=torch.tensor(wgt.reshape(32, 32, 32)))
    gw = array(dtype=float32, length=32 ** 3, library=oracle_lib)
    gw.upload(wgt)
    oracle_lib.particle2mass(a.states[0]["x"].data_ptr, a.mass.data_ptr, a.grid_lower.data_ptr, gd, a.dx, a.inv_dx, gm.data_ptr, gw.data_ptr,
                             a.states[0]["x_grad"].data_ptr, ids.data_ptr, -1, 1, N, None)
    assert rel_err(sim.engine.get_state_grad(0, ("x",))["x"][0], a.get(0, "x_grad")["x_grad"]) < 1e-4


def test_simulator_state_roundtrip_and_step():
    sc = build_scene()
    sim = make_sim(sc)
    st = sim.get_state(0)
    assert len(st) == 4 + NB and st[2].shape == (N, 3, 3) and st[4].shape == (7,)
    for a, k in zip(st[:4], ("x", "v", "F", "C")):
        assert np.array_equal(a.reshape(N, -1), sc[k])   # original particle order, bit-exact round trip through the sorted SoA layout
    y0 = sim.get_x(0)[:, 1].mean()
    sim.step(np.zeros((NB, 6), np.float32))
    x1 = sim.get_x(0)
    assert x1.shape == (N, 3) and np.isfinite(x1).all() and x1[:, 1].mean() < y0 + 1e-6   # gravity pulls the block down
    d = sim.get_dists(0, device="numpy")
    assert d.shape == (N, NB)


# ---------------------------------------------------------------------------------------------------- long horizon
def lift_scene(n=3000, steps=50, S=40):
    """lift_box-like: a plasticine block resting on a two-piece platform that is raised by the actions -- the block travels
    more than ten cells over the rollout (the tutorial's horizon, tutorials/1_trajectory_optimization.ipynb:184-199)."""
    sc = make_scene(n, 64, box_center=(0.5, 0.25, 0.5), box_width=(0.08, 0.06, 0.08), steps=steps * S, seed=11, nb=2)
    bottom = 0.25 - 0.03
    sc["tfsr"] = np.float32([[0, 0.9, 666.0, 0], [0, 0.9, 666.0, 0]])          # two boxes, friction 0.9, softness 666
    sc["args"] = np.float32([[0.035, 0.008, 0.07, 0], [0.035, 0.008, 0.07, 0]])
    pos0 = np.float32([[0.5 - 0.035, bottom - 0.008 - 0.5 / 64, 0.5], [0.5 + 0.035, bottom - 0.008 - 0.5 / 64, 0.5]])
    rot0 = np.float32([[1, 0, 0, 0], [1, 0, 0, 0]])
    sc["pos"] = np.repeat(pos0[None], steps * S + 1, 0)
    sc["rot"] = np.repeat(rot0[None], steps * S + 1, 0)
    return sc


def reference_rollout(ref_gpu, sc, action, S, steps, scale):
    """mpm/torch_wrapper.py:110-141 re-enacted on the reference's library: per-substep pose upload + substep, then substep_grad
    and per-substep pose-gradient downloads, torch autograd through the tool kinematics."""
    n, nb = sc["n"], sc["nb"]
    sim = Abi1Sim(ref_gpu, sc, S * steps)
    pos_rot = (torch.tensor(sc["pos"][0]), torch.tensor(sc["rot"][0]))
    poses = []
    for s_ in range(steps):
        a = action[s_].reshape(-1, 6).clamp(-1.0, 1.0) * scale
        pos, rot = rigid_body_motion(pos_rot, a[None].expand(S, -1, -1) * (torch.arange(S)[:, None, None] + 1) / S)
        pos_rot = (pos[-1], rot[-1])
        poses.append((pos, rot))
        for i in range(S):
            f = s_ * S + i
            sim.states[f + 1]["body_pos"].upload(pos[i].detach().numpy())
            sim.states[f + 1]["body_rot"].upload(rot[i].detach().numpy())
            sim.substep(f)
    fe = S * steps
    x = sim.get(fe, "x")["x"]
    loss = -x[:, 1].mean()
    gx = np.zeros((n, 3), np.float32)
    gx[:, 1] = -1.0 / n
    sim.states[fe]["x_grad"].upload(gx)
    for f in range(fe - 1, -1, -1):
        sim.substep_grad(f)
    sim.sync()
    total = 0.0
    for s_ in range(steps):
        gp = np.stack([sim.get(s_ * S + i + 1, "body_pos_grad")["body_pos_grad"] for i in range(S)])
        gr = np.stack([sim.get(s_ * S + i + 1, "body_rot_grad")["body_rot_grad"] for i in range(S)])
        total = total + (poses[s_][0] * torch.tensor(gp)).sum() + (poses[s_][1] * torch.tensor(gr)).sum()
    total.backward()
    return float(loss), action.grad.clone(), x


def test_gradmodel_tutorial_horizon_50_steps_of_40_substeps(ref_gpu):
    """The reference's tutorial horizon through GradModel: 50 env steps x 40 substeps with no set_state in between (the engine
    re-sorts on the device at every env-step boundary), block lifted by more than 10 cells; loss and action gradients against
    the reference chain.  Tolerance: rel-L2 <= 1e-2, cosine >= 0.999, or 5x the reference's own spread under a
    permutation of the particle order if that is larger (2000 substeps of contact dynamics amplify its atomics' noise)."""
    S_, STEPS_, n = 40, 50, 3000
    sc = lift_scene(n, STEPS_, S_)
    nb = sc["nb"]
    scale = torch.tensor([[0.01] * 3 + [0.015] * 3] * nb, dtype=torch.float32)
    rng = np.random.default_rng(3)
    a0 = np.float32(rng.uniform(-0.05, 0.05, (STEPS_, nb, 6)))
    a0[:, :, 1] = 0.45   # raise the platform: 0.0045 per env step = 0.29 cells, 14 cells over the rollout
    a0[:, :, 3:] *= 0.2
    act_ref = torch.tensor(a0, requires_grad=True)
    loss_ref, grad_ref, x_ref = reference_rollout(ref_gpu, sc, act_ref, S_, STEPS_, scale)
    assert (x_ref[:, 1].mean() - sc["x"][:, 1].mean()) * 64 > 10, "the block must travel more than ten cells"
    # the reference against itself with the particles listed in another order
    perm = np.random.default_rng(0).permutation(n)
    sc2 = dict(sc)
    for k in ("x", "v", "F", "C", "mass", "vol", "mu_lam_yield"):
        sc2[k] = np.ascontiguousarray(sc[k][perm])
    act2 = torch.tensor(a0, requires_grad=True)
    _, grad_ref2, _ = reference_rollout(ref_gpu, sc2, act2, S_, STEPS_, scale)
    spread = rel_l2(grad_ref2.numpy(), grad_ref.numpy())

    sim = MPMSimulator(nb, ground_friction=sc["ground_friction"], gravity=tuple(sc["gravity"].reshape(3) / 30), n_particles=n, dx=sc["dx"],
                       dt=sc["dt"], max_steps=S_ * STEPS_, substeps=S_, yield_stress=50.0)
    sim.init_bodies(sc["tfsr"][:, 0], sc["tfsr"][:, 2], sc["tfsr"][:, 1], sc["tfsr"][:, 3], sc["args"], action_scales=scale.tolist(),
                    pos=sc["pos"][0], rot=sc["rot"][0])
    sim.set_state(0, (sc["x"], sc["v"], sc["F"].reshape(-1, 3, 3), sc["C"].reshape(-1, 3, 3)) + tuple(np.r_[p, r] for p, r in zip(sc["pos"][0], sc["rot"][0])))
    info = sim.engine.segment_info(0)
    assert info["n_segments"] == STEPS_ and info["interval"] == S_
    model = GradModel(sim, return_grid=())
    grads = []
    for it in range(2):   # two optimisation iterations on one engine: cached graphs and segment tables are reused
        model.zero_grad()
        if it:
            sim.set_state(0, (sc["x"], sc["v"], sc["F"].reshape(-1, 3, 3), sc["C"].reshape(-1, 3, 3)) + tuple(np.r_[p, r] for p, r in zip(sc["pos"][0], sc["rot"][0])))
        action = torch.tensor(a0, device="cuda:0", requires_grad=True)
        obs = model.get_obs(0, "cuda:0")
        for j in range(STEPS_):
            obs = model.forward(j, action[j], *obs)
        loss = -obs[0][:, 1].mean()
        loss.backward()
        grads.append(action.grad.cpu().numpy())
    sim.sync()
    x_end = obs[0][:, :3].detach().cpu().numpy()
    assert abs(float(loss) - loss_ref) < 2e-4 * max(1.0, abs(loss_ref)), (float(loss), loss_ref)
    print('final position difference', np.abs(x_end - x_ref).max())
    assert np.abs(x_end - x_ref).max() < 2e-3, np.abs(x_end - x_ref).max()   # after 2000 substeps (0.13 cells)
    g = grads[0]
    assert np.isfinite(g).all() and np.abs(g).max() > 0
    assert rel_l2(grads[1], g) < 1e-3, rel_l2(grads[1], g)   # the second iteration reproduces the first
    e = rel_l2(g, grad_ref.numpy())
    print(f"tutorial horizon: action-gradient rel-L2 {e:.3e} cos {cosine(g, grad_ref.numpy()):.6f}; reference self-spread {spread:.3e}")
    assert e < max(1e-2, 5 * spread), (e, spread)
    assert cosine(g, grad_ref.numpy()) > min(0.999, 1 - 10 * spread ** 2), cosine(g, grad_ref.numpy())


def test_gradmodel_batched_environments():
    """n_envs = 3: GradModel carries a leading environment axis; every environment's gradient equals the single-environment run."""
    sc = build_scene()
    rng = np.random.default_rng(2)
    E = 3
    a0 = np.float32(rng.uniform(-0.8, 0.8, (E, STEPS, NB, 6)))
    a0[..., 1] = -0.9
    singles = []
    for e in range(E):
        sim = make_sim(sc)
        model = GradModel(sim, return_grid=())
        model.zero_grad()
        action = torch.tensor(a0[e], device="cuda:0", requires_grad=True)
        obs = model.get_obs(0, "cuda:0")
        for j in range(STEPS):
            obs = model.forward(j, action[j], *obs)
        loss = -obs[0][:, 1].mean() + 0.3 * obs[0][:, 6:].mean()
        loss.backward()
        singles.append((float(loss), action.grad.cpu().numpy()))
        sim.engine.close()
    sim = make_sim(sc, n_envs=E)
    model = GradModel(sim, return_grid=())
    model.zero_grad()
    action = torch.tensor(a0, device="cuda:0", requires_grad=True)   # (E, STEPS, NB, 6)
    obs = model.get_obs(0, "cuda:0")
    assert obs[0].shape == (E, N, 6 + NB) and obs[1].shape == (E, NB, 7)
    for j in range(STEPS):
        obs = model.forward(j, action[:, j], *obs)
    losses = -obs[0][..., 1].mean(1) + 0.3 * obs[0][..., 6:].mean((1, 2))
    losses.sum().backward()
    g = action.grad.cpu().numpy()
    for e in range(E):
        assert abs(float(losses[e]) - singles[e][0]) < 1e-5 * max(1.0, abs(singles[e][0]))
        assert rel_l2(g[e], singles[e][1]) < 2e-3, (e, rel_l2(g[e], singles[e][1]))


@pytest.mark.parametrize("E", [1, 2])
def test_two_level_checkpointing_matches_full_checkpointing(E):
    """Windowed GradModel (simulator with max_steps == substeps: per-substep checkpoints of ONE env step, boundary states kept as
    device tensors, forward substeps re-run in the backward pass) against the GradModel that checkpoints every substep of the
    rollout: same loss, same action gradients, over 6 env steps with observation gradients entering at every boundary."""
    S_, STEPS_, n = 20, 6, 2500
    sc = lift_scene(n, STEPS_, S_)
    nb = sc["nb"]
    scale = [[0.01] * 3 + [0.015] * 3] * nb
    rng = np.random.default_rng(5)
    a0 = np.float32(rng.uniform(-0.1, 0.1, (STEPS_, E, nb, 6)))
    a0[..., 1] = 0.4
    out = {}
    for mode, max_steps in (("full", S_ * STEPS_), ("windowed", S_)):
        sim = MPMSimulator(nb, ground_friction=sc["ground_friction"], gravity=tuple(sc["gravity"].reshape(3) / 30), n_particles=n, dx=sc["dx"],
                           dt=sc["dt"], max_steps=max_steps, substeps=S_, yield_stress=50.0, n_envs=E)
        sim.init_bodies(sc["tfsr"][:, 0], sc["tfsr"][:, 2], sc["tfsr"][:, 1], sc["tfsr"][:, 3], sc["args"], action_scales=scale, pos=sc["pos"][0], rot=sc["rot"][0])
        sim.set_state(0, (sc["x"], sc["v"], sc["F"].reshape(-1, 3, 3), sc["C"].reshape(-1, 3, 3)) + tuple(np.r_[p, r] for p, r in zip(sc["pos"][0], sc["rot"][0])))
        model = GradModel(sim, return_grid=())
        assert model.windowed == (mode == "windowed")
        model.zero_grad()
        action = torch.tensor(a0 if E > 1 else a0[:, 0], device="cuda", requires_grad=True)
        obs = model.get_obs(0, "cuda")
        loss = 0.0
        for j in range(STEPS_):
            obs = model.forward(j, action[j], *obs)
            loss = loss - obs[0][..., 1].mean() + 0.05 * obs[0][..., 3:6].square().mean() + 0.02 * obs[0][..., 6:].mean() + 0.1 * obs[1][..., :3].sum()
        loss.backward()
        out[mode] = (float(loss), action.grad.cpu().numpy().copy(), obs[0][..., :3].detach().cpu().numpy())
        sim.engine.close()
    assert abs(out["full"][0] - out["windowed"][0]) < 1e-5 * max(1.0, abs(out["full"][0]))
    assert np.abs(out["full"][2] - out["windowed"][2]).max() < 1e-5
    e, c = rel_l2(out["windowed"][1], out["full"][1]), cosine(out["windowed"][1], out["full"][1])
    print(f"two-level checkpointing (E={E}): action-gradient rel-L2 {e:.3e} cos {c:.7f}")
    assert e < 2e-3 and c > 0.99999, (e, c)
