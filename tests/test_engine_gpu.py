"""GPU parity tests of the fused engine (ABI-2) against the reference's CUDA library, the per-stage ABI-1 path and the
C oracle, all on identical inputs.  svd_mode=0 uses the reference-order fp64 SVD (tight tolerances); svd_mode=1 is the
production fp32 in-register SVD (tolerances of SURVEY.md 8c)."""
import numpy as np
import pytest

from abi1_driver import Abi1Sim, loss_seed
from conftest import assert_close_rows, cosine, rel_err, rel_l2
from dexdeform_b200.engine import EngineError, FusedSim
from dexdeform_b200.scenes import make_scene, scene_tutorial

pytestmark = pytest.mark.gpu


def run_abi1(lib, sc, S, seedg):
    sim = Abi1Sim(lib, sc, S)
    for f in range(S):
        sim.substep(f)
    for k, v in seedg.items():
        sim.states[S][k].upload(v)
    for f in range(S - 1, -1, -1):
        sim.substep_grad(f)
    out = dict(state=sim.get(S), grad=sim.get(0, "x_grad", "v_grad", "F_grad", "C_grad"))
    if sc["nb"]:
        out["gpos"] = np.stack([sim.get(f, "body_pos_grad")["body_pos_grad"] for f in range(S + 1)])
        out["grot"] = np.stack([sim.get(f, "body_rot_grad")["body_rot_grad"] for f in range(S + 1)])
    return out


def permuted(sc, seedg, seed=0):
    """The same scene with particles listed in another order (mathematically identical problem)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(sc["n"])
    sc2 = dict(sc)
    for k in ("x", "v", "F", "C", "mass", "vol", "mu_lam_yield"):
        sc2[k] = np.ascontiguousarray(sc[k][perm])
    return sc2, {k: np.ascontiguousarray(v[perm]) for k, v in seedg.items()}, np.argsort(perm)


def reference_spread(lib, sc, S, seedg, ref):
    """The reference's own sensitivity to summation order: re-run it on the permuted scene and compare.  Float atomics
    make the reference irreproducible at this level, so no implementation can be asked to match it tighter."""
    sc2, seed2, inv = permuted(sc, seedg)
    other = run_abi1(lib, sc2, S, seed2)
    spread = {}
    for grp in ("state", "grad"):
        for k, v in ref[grp].items():
            spread[k] = rel_err(other[grp][k][inv], v)
    for k in ("gpos", "grot"):
        if k in ref:
            spread[k] = rel_err(other[k], ref[k])
    return spread


def run_engine(sc, S, seedg, E=1, **kw):
    sim = FusedSim.from_scene(sc, n_envs=E, max_steps=S, **kw)
    l0 = sim.launch_count()  # (the cell sort of set_state is counted too; the assertions below are about the substeps)
    sim.forward(0, S)
    sim.zero_grad(S)
    t = lambda a: np.ascontiguousarray(np.broadcast_to(a[None], (E,) + a.shape))
    sim.add_state_grad(S, t(seedg["x_grad"]), t(seedg["v_grad"]), t(seedg["F_grad"]), t(seedg["C_grad"]))
    sim.backward(0, S)
    out = dict(state=sim.get_state(S), grad=sim.get_state_grad(0))
    out["gpos"], out["grot"] = sim.get_pose_grads(0, S + 1)
    out["launches"] = sim.launch_count() - l0
    out["info"] = sim.segment_info(0) if kw.get("tile_mode", True) else None
    sim.close()
    return out


@pytest.mark.parametrize("svd_mode,graphs,tile", [(0, False, False), (0, False, True), (1, True, True), (1, True, False)])
def test_engine_matches_reference_cuda_short(ref_gpu, svd_mode, graphs, tile):
    S = 4
    sc = scene_tutorial(steps=S, perturb=0.02, vel_scale=0.3, on_floor=True, seed=2)
    seedg = loss_seed(sc["n"], 3)
    ref = run_abi1(ref_gpu, sc, S, seedg)
    spread = reference_spread(ref_gpu, sc, S, seedg, ref)
    eng = run_engine(sc, S, seedg, svd_mode=svd_mode, use_graphs=graphs, tile_mode=tile)
    tight = svd_mode == 0
    tol_of = lambda key, base: max(base * (1 if tight else 5), 5 * spread[key])
    assert np.abs(eng["state"]["x"][0] - ref["state"]["x"]).max() < (2e-7 if tight else 1e-6)
    for k, tol in dict(v=2e-5, F=5e-6, C=1e-4).items():
        e = rel_err(eng["state"][k][0], ref["state"][k])
        assert e < tol_of(k, tol), (k, e, spread[k])
    for k, tol in dict(x=2e-4, v=2e-4, C=1e-2, F=1e-2).items():
        assert_close_rows(eng["grad"][k][0], ref["grad"][k + "_grad"], tol_of(k + "_grad", tol), k + "_grad")
        assert cosine(eng["grad"][k][0], ref["grad"][k + "_grad"]) > 0.999
    assert rel_err(eng["gpos"][:, 0], ref["gpos"]) < tol_of("gpos", 2e-4)
    assert rel_err(eng["grot"][:, 0], ref["grot"]) < tol_of("grot", 2e-4)
    assert eng["launches"] == (3 * S + 1 + 3 * S if tile else 3 * S + 5 * S)


def test_engine_rollout_50_substeps_vs_reference(ref_gpu):
    """BASELINE config A through the production path (fp32 SVD, CUDA graphs): tolerances of SURVEY.md 8c."""
    S = 50
    sc = scene_tutorial(steps=S, seed=0, on_floor=True)
    n = sc["n"]
    seedg = dict(x_grad=np.zeros((n, 3), np.float32), v_grad=np.zeros((n, 3), np.float32), F_grad=np.zeros((n, 9), np.float32),
                 C_grad=np.zeros((n, 9), np.float32))
    seedg["x_grad"][:, 1] = -1.0 / n
    ref = run_abi1(ref_gpu, sc, S, seedg)
    eng = run_engine(sc, S, seedg)
    assert np.abs(eng["state"]["x"][0] - ref["state"]["x"]).max() < 1e-4
    for k in ("v", "F", "C"):
        assert rel_err(eng["state"][k][0], ref["state"][k]) < 1e-3, (k, rel_err(eng["state"][k][0], ref["state"][k]))
    assert np.abs(ref["gpos"]).max() > 0
    for a, b in ((eng["gpos"][:, 0], ref["gpos"]), (eng["grot"][:, 0], ref["grot"])):
        assert rel_l2(a, b) < 1e-2, rel_l2(a, b)
        assert cosine(a, b) > 0.999
    for k in ("x", "v"):
        assert rel_l2(eng["grad"][k][0], ref["grad"][k + "_grad"]) < 1e-2
        assert cosine(eng["grad"][k][0], ref["grad"][k + "_grad"]) > 0.999


def test_engine_vs_oracle_and_env_batching(oracle_lib):
    """E=3 environments with different states and poses must each equal a single-environment oracle run."""
    S, E = 3, 3
    scs = [make_scene(1200, 32, box_width=(0.12, 0.1, 0.12), steps=S, perturb=0.03, vel_scale=0.4, on_floor=True, seed=40 + e, nb=5)
           for e in range(E)]
    for sc in scs[1:]:  # shapes are shared by all environments of one engine
        sc["tfsr"], sc["args"] = scs[0]["tfsr"], scs[0]["args"]
    seedg = loss_seed(1200, 8)
    refs = [run_abi1(oracle_lib, sc, S, seedg) for sc in scs]
    sc0 = scs[0]
    sim = FusedSim(E, 1200, 5, sc0["grid_dim"], sc0["dx"], sc0["dt"], S, sc0["ground_friction"], sc0["ground_height"], sc0["gravity"].reshape(3),
                   svd_mode=0)
    st = lambda k: np.ascontiguousarray(np.stack([sc[k] for sc in scs]))
    sim.set_material(st("mass"), st("vol"), st("mu_lam_yield"))
    sim.set_bodies(sc0["tfsr"], sc0["args"])
    sim.set_poses(0, np.ascontiguousarray(np.stack([sc["pos"] for sc in scs], 1)), np.ascontiguousarray(np.stack([sc["rot"] for sc in scs], 1)))
    sim.set_state(0, st("x"), st("v"), st("F"), st("C"))
    sim.forward(0, S)
    sim.zero_grad(S)
    t = lambda a: np.ascontiguousarray(np.broadcast_to(a[None], (E,) + a.shape))
    sim.add_state_grad(S, t(seedg["x_grad"]), t(seedg["v_grad"]), t(seedg["F_grad"]), t(seedg["C_grad"]))
    sim.backward(0, S)
    state, grad = sim.get_state(S), sim.get_state_grad(0)
    gpos, grot = sim.get_pose_grads(0, S + 1)
    for e in range(E):
        for k, tol in dict(x=2e-6, v=5e-5, F=5e-6, C=2e-4).items():
            assert rel_err(state[k][e], refs[e]["state"][k]) < tol, (e, k)
        for k in ("x", "v"):
            assert rel_err(grad[k][e], refs[e]["grad"][k + "_grad"]) < 5e-4, (e, k)
        # an environment may have (almost) no contact at all: tolerance relative to max(|ref|, 1)
        for a, b in ((gpos[:, e], refs[e]["gpos"]), (grot[:, e], refs[e]["grot"])):
            assert np.abs(a - b).max() < 5e-4 * max(np.abs(b).max(), 1.0), (e, np.abs(a - b).max(), np.abs(b).max())
    # observations: signed distances and their adjoint against the oracle
    d = sim.compute_dist(S)
    a1 = Abi1Sim(oracle_lib, scs[1], S)
    for f in range(S):
        a1.substep(f)
    from dexdeform_b200.types import array, float32
    dist = array(dtype=float32, length=1200 * 5, library=oracle_lib)
    a1.compute_dist(a1.states[S], dist, dist, 0)
    assert np.abs(d[1] - dist.download().reshape(1200, 5)).max() < 2e-6
    # fused observation [x | v | dist] and its adjoint (dd_sim_get_obs / dd_sim_add_obs_grad): bit-identical to the parts, adjoint against
    # the oracle.  1200 particles per environment: blocks and warps straddle environments; body 1 and a range of particles receive
    # exact zeros (the skipped paths of k_obs_grad).
    import torch
    obs = sim.get_obs(S).cpu().numpy()
    assert np.array_equal(obs[..., :3], state["x"]) and np.array_equal(obs[..., 3:6], state["v"]) and np.array_equal(obs[..., 6:], d)
    rng = np.random.default_rng(5)
    gobs = np.float32(rng.normal(size=(E, 1200, 11)))
    gobs[:, :, 6 + 1] = 0.0
    gobs[:, 300:900, 6:] = 0.0
    sim.zero_grad(S)
    sim.add_obs_grad(S, torch.tensor(gobs, device="cuda"))
    g = sim.get_state_grad(S, ("x", "v"))
    gp, gr = sim.get_pose_grads(S, 1)
    dg = array(dtype=float32, length=1200 * 5, library=oracle_lib)
    dg.upload(np.ascontiguousarray(gobs[1, :, 6:]).reshape(-1))
    for k in ("x_grad", "body_pos_grad", "body_rot_grad"):
        a1.states[S][k].zero(a1.stream)
    a1.compute_dist(a1.states[S], dist, dg, 1)
    ref = a1.get(S, "x_grad", "body_pos_grad", "body_rot_grad")
    assert np.abs(g["x"][1] - (ref["x_grad"] + gobs[1, :, :3])).max() < 1e-5
    assert np.array_equal(g["v"][1], gobs[1, :, 3:6])
    assert np.abs(gp[0, 1] - ref["body_pos_grad"]).max() < 2e-4 * max(1.0, np.abs(ref["body_pos_grad"]).max())
    assert np.abs(gr[0, 1] - ref["body_rot_grad"]).max() < 2e-4 * max(1.0, np.abs(ref["body_rot_grad"]).max())
    assert np.all(gp[0, :, 1] == 0.0) and np.all(gr[0, :, 1] == 0.0)
    # the separate entry points share the kernel: same numbers
    sim.zero_grad(S)
    sim.add_state_grad(S, gx=np.ascontiguousarray(gobs[..., :3]), gv=np.ascontiguousarray(gobs[..., 3:6]))
    sim.compute_dist_grad(S, np.ascontiguousarray(gobs[..., 6:]))
    g2 = sim.get_state_grad(S, ("x", "v"))
    assert np.abs(g2["x"] - g["x"]).max() < 1e-6 and np.array_equal(g2["v"], g["v"])
    sim.close()


def test_engine_error_reporting():
    with pytest.raises(EngineError, match="n_bodies"):
        FusedSim(1, 64, 100, (32, 32, 32), 1 / 32, 1e-4, 4)
    sc = make_scene(64, 32, steps=2, nb=0, seed=1)
    sim = FusedSim.from_scene(sc, max_steps=2)
    with pytest.raises(EngineError, match="exceeds max_steps"):
        sim.forward(0, 5)
    with pytest.raises(EngineError, match="no gradient seeded"):
        sim.backward(0, 2)
    sim.close()


@pytest.mark.parametrize("svd_mode", [0, 1])
def test_tiled_path_equals_dense_path_with_drift_and_dense_cells(svd_mode):
    """The shared-memory tile path (conflict serialisation, drift fallback to the grid) against the plain global-reduction
    path on a scene built to stress it: 40 particles per cell (many same-cell lanes per round) moving fast enough that
    most of them change cell (and many leave the tile interior) within the two substeps."""
    S = 2
    sc = make_scene(4000, 32, box_center=(0.5, 0.4, 0.5), box_width=(0.14, 0.14, 0.14), steps=S, perturb=0.02, nb=4, seed=13, ground_friction=0.3)
    # ~0.7 cells per substep: after two substeps most particles sit in another cell than the one they were sorted into
    sc["v"][:] = np.array([450.0, -300.0, 200.0], np.float32) * (1.0 + 0.01 * np.random.default_rng(0).normal(size=(4000, 3)).astype(np.float32))
    seedg = loss_seed(4000, 5)
    outs = {}
    for tile in (False, True):
        outs[tile] = run_engine(sc, S, seedg, svd_mode=svd_mode, tile_mode=tile, use_graphs=False, grid_ckpt=tile)
    a, b = outs[False], outs[True]
    for k in ("x", "v", "F", "C"):
        assert rel_err(b["state"][k], a["state"][k]) < 2e-5, (k, rel_err(b["state"][k], a["state"][k]))
    for k in ("x", "v"):
        assert_close_rows(b["grad"][k][0], a["grad"][k][0], 5e-4, k + "_grad")
    assert np.abs(b["gpos"] - a["gpos"]).max() < 2e-3 * max(np.abs(a["gpos"]).max(), 1.0)


def test_tiled_path_follows_runaway_particles():
    """Particles that out-run the region that was active at the last sort (4.8 cells per substep here) activate the bricks
    they reach on the fly (engine.cu:activate_bricks): same result as the dense path, forward and adjoint, no error."""
    S = 3
    # (perturbed F: at F = I the SVD adjoint multiplies rounding noise by 1e6, integrator.cu:146-157, and no two runs agree)
    sc = make_scene(512, 32, box_center=(0.3, 0.5, 0.5), box_width=(0.05, 0.05, 0.05), steps=S, nb=0, seed=3, ground_friction=0.0, perturb=0.02)
    sc["v"][:] = np.array([1500.0, 200.0, -300.0], np.float32)
    seedg = loss_seed(512, 4)
    seedg["C_grad"][:] = 0   # (gC' is multiplied by (4/dx^2) v': at these speeds it would drown every other term in rounding noise)
    a = run_engine(sc, S, seedg, tile_mode=False, use_graphs=False, grid_ckpt=False)
    sim = FusedSim.from_scene(sc, max_steps=S, tile_mode=True)
    before = sim.segment_info(0)["active_bricks"]
    sim.forward(0, S)
    sim.sync()
    after = sim.segment_info(0)["active_bricks"]
    assert after > before, (before, after)
    sim.zero_grad(S)
    sim.add_state_grad(S, seedg["x_grad"][None], seedg["v_grad"][None], seedg["F_grad"][None], seedg["C_grad"][None])
    sim.backward(0, S)
    st, gr = sim.get_state(S), sim.get_state_grad(0)
    assert np.abs(st["x"][0] - sc["x"]).max() > 4 * sc["dx"]   # they did travel
    # (C is the gradient of a nearly uniform 1500-per-second velocity field: its natural scale is |v| / dx, not its own size)
    c_scale = float(np.abs(a["state"]["v"]).max() * sc["inv_dx"])
    assert np.abs(st["C"] - a["state"]["C"]).max() < 2e-5 * c_scale
    # (F: 1.2e-5 typical, 2.0-2.3e-5 in 2 runs of 80 -- the float-atomic order of the grid sums times dt |v| / dx = 4.8 cells per substep)
    for k, tol in dict(x=2e-5, v=2e-5, F=5e-5).items():
        assert rel_err(st[k], a["state"][k]) < tol, (k, rel_err(st[k], a["state"][k]))
    for k in ("x", "v"):
        assert_close_rows(gr[k][0], a["grad"][k][0], 5e-3, k + "_grad", frac=0.02)   # (summation order differs; |v| dx/dt = 4.8 amplifies rounding)
    # a second rollout from a new initial state reuses the engine: stale grid contents of the earlier run must not leak
    sc["v"][:] = np.array([-900.0, 100.0, 500.0], np.float32)
    sim.set_state(0, sc["x"][None], sc["v"][None], sc["F"][None], sc["C"][None])
    sim.forward(0, S)
    b = run_engine(sc, S, seedg, tile_mode=False, use_graphs=False, grid_ckpt=False)
    st = sim.get_state(S)
    assert np.abs(st["C"] - b["state"]["C"]).max() < 2e-5 * c_scale
    for k, tol in dict(x=2e-5, v=2e-5, F=5e-5).items():
        assert rel_err(st[k], b["state"][k]) < tol, (k, rel_err(st[k], b["state"][k]))
    sim.close()


@pytest.mark.parametrize("interval,graphs", [(4, True), (5, False), (1, True)])
def test_resort_inside_rollout_matches_single_ordering(interval, graphs):
    """Device-side re-sort every `interval` substeps (states at the boundaries stored in both orders, gradient permuted back in
    the adjoint) against the same rollout in one ordering and against the dense path: identical physics, so states, state
    gradients and pose gradients must agree to rounding.  The block moves ~0.5 cells per substep."""
    S = 12
    sc = make_scene(3000, 32, box_center=(0.4, 0.35, 0.5), box_width=(0.14, 0.1, 0.14), steps=S, perturb=0.02, nb=4, seed=17, ground_friction=0.3)
    sc["v"][:] = np.array([160.0, -60.0, 90.0], np.float32) * (1.0 + 0.02 * np.random.default_rng(1).normal(size=(3000, 3)).astype(np.float32))
    seedg = loss_seed(3000, 6)
    a = run_engine(sc, S, seedg, tile_mode=False, use_graphs=False, grid_ckpt=False)
    b = run_engine(sc, S, seedg, tile_mode=True, use_graphs=graphs)
    c = run_engine(sc, S, seedg, tile_mode=True, use_graphs=graphs, resort_interval=interval)
    assert c["info"]["n_segments"] == (S + interval - 1) // interval and b["info"]["n_segments"] == 1
    for other in (a, b):
        for k, tol in dict(x=2e-5, v=2e-5, F=2e-5, C=1e-4).items():   # (C: differences of velocities that are ~50 cells/s apart)
            assert rel_err(c["state"][k], other["state"][k]) < tol, (k, rel_err(c["state"][k], other["state"][k]))
        for k in ("x", "v", "F", "C"):
            assert_close_rows(c["grad"][k][0], other["grad"][k][0], 1e-3, k + "_grad")
        assert np.abs(c["gpos"] - other["gpos"]).max() < 2e-3 * max(np.abs(other["gpos"]).max(), 1.0)
        assert np.abs(c["grot"] - other["grot"]).max() < 2e-3 * max(np.abs(other["grot"]).max(), 1.0)


def test_resort_in_pieces_and_stale_slots():
    """forward / backward called per segment (as GradModel does, one env step at a time) gives the same result as one call over
    the whole range; slots written under an ordering that was replaced are refused instead of returned permuted."""
    S, L = 8, 4
    sc = make_scene(2000, 32, box_width=(0.12, 0.1, 0.12), steps=S, perturb=0.02, vel_scale=0.5, on_floor=True, nb=3, seed=5)
    seedg = loss_seed(2000, 7)
    whole = run_engine(sc, S, seedg, resort_interval=L)
    sim = FusedSim.from_scene(sc, max_steps=S, resort_interval=L)
    sim.forward(0, L)
    x_mid = sim.get_state(L, ("x",))["x"]          # tail copy (the next segment is not built yet)
    sim.forward(L, L)
    assert np.array_equal(sim.get_state(L, ("x",))["x"], x_mid)   # head copy: same particles, caller's order
    sim.zero_grad(S)
    sim.add_state_grad(S, seedg["x_grad"][None], seedg["v_grad"][None], seedg["F_grad"][None], seedg["C_grad"][None])
    sim.backward(L, L)
    g_mid = sim.get_state_grad(L, ("x",))["x"]
    sim.backward(0, L)
    gr = sim.get_state_grad(0)
    for k in ("x", "v", "F", "C"):  # (two runs of the same engine: one yield-branch flip from the float-atomic order moves 5-6 rows by 6e-4, 3 runs in 60)
        assert_close_rows(gr[k][0], whole["grad"][k][0], 1e-4, k + "_grad", frac=1e-2)
    assert np.isfinite(g_mid).all() and np.abs(g_mid).max() > 0
    gp, _ = sim.get_pose_grads(0, S + 1)
    assert np.abs(gp - whole["gpos"]).max() < 1e-3 * max(np.abs(whole["gpos"]).max(), 1.0)
    # a new initial state replaces the ordering of segment 0: its old slots are stale, segment 1 is untouched until re-run
    sim.set_state(0, sc["x"][None], sc["v"][None], sc["F"][None], sc["C"][None])
    with pytest.raises(EngineError, match="not available"):
        sim.get_state(2)
    with pytest.raises(EngineError, match="no gradient seeded|not available"):
        sim.backward(0, L)
    sim.forward(0, L)
    sim.get_state(2)
    # rolling window on the device: state L becomes state 0 (re-sorted), poses included
    xL = sim.get_state(L, ("x", "v", "F", "C"))
    pL = sim.get_poses(L, 1)
    sim.roll(L)
    x0 = sim.get_state(0, ("x", "v", "F", "C"))
    for k in xL:
        assert np.array_equal(x0[k], xL[k]), k
    p0 = sim.get_poses(0, 1)
    assert np.array_equal(p0[0], pL[0]) and np.array_equal(p0[1], pL[1])
    sim.close()


def test_recompute_mode_matches_checkpoint_mode():
    S = 3
    sc = make_scene(3000, 32, box_width=(0.14, 0.1, 0.14), steps=S, perturb=0.02, vel_scale=0.5, on_floor=True, seed=21, nb=6)
    seedg = loss_seed(3000, 2)
    a = run_engine(sc, S, seedg, grid_ckpt=True)
    b = run_engine(sc, S, seedg, grid_ckpt=False)
    for k in ("x", "v", "F", "C"):
        assert_close_rows(b["grad"][k][0], a["grad"][k][0], 1e-4, k + "_grad")
    assert a["launches"] == 6 * S + 1 and b["launches"] == 8 * S + 1


@pytest.mark.parametrize("interval,graphs,E", [(0, True, 1), (4, True, 2), (5, False, 1)])
def test_brick_checkpoints_match_dense_grid_checkpoints(interval, graphs, E):
    """grid_ckpt=2: (mv, m) and v_out of the active bricks only, per substep (what a 64-environment, 400-substep rollout can afford)
    against one dense grid pair per substep: same forward, and an adjoint that differs only in the order of floating-point
    reductions.  The block moves ~0.5 cells per substep, so bricks are activated on the fly inside a segment (the list the
    checkpoints are indexed by grows) and, with an interval, the ordering changes in the middle of the backward sweep."""
    S = 12
    sc = make_scene(3000, 32, box_center=(0.4, 0.35, 0.5), box_width=(0.14, 0.1, 0.14), steps=S, perturb=0.02, nb=4, seed=17, ground_friction=0.3)
    sc["v"][:] = np.array([160.0, -60.0, 90.0], np.float32) * (1.0 + 0.02 * np.random.default_rng(1).normal(size=(3000, 3)).astype(np.float32))
    seedg = loss_seed(3000, 6)
    a = run_engine(sc, S, seedg, E=E, use_graphs=graphs, resort_interval=interval, grid_ckpt=1)
    b = run_engine(sc, S, seedg, E=E, use_graphs=graphs, resort_interval=interval, grid_ckpt=2)
    nseg = (S + interval - 1) // interval if interval else 1
    assert b["launches"] == a["launches"] + nseg   # one stand-alone restore per segment, everything else rides on the grid adjoint
    for k in ("x", "v", "F", "C"):
        assert rel_err(b["state"][k], a["state"][k]) < (1e-4 if k == "C" else 2e-5), k   # (same kernels; the order of the grid reductions varies from run to run)
        assert_close_rows(b["grad"][k].reshape(-1, b["grad"][k].shape[-1]), a["grad"][k].reshape(-1, a["grad"][k].shape[-1]), 1e-3, k + "_grad")
    # (12 substeps at half a cell per substep: two runs of the SAME configuration differ by ~1.5e-4 in the worst rows)
    assert np.abs(b["gpos"] - a["gpos"]).max() < 2e-3 * max(np.abs(a["gpos"]).max(), 1.0)
    assert np.abs(b["grot"] - a["grot"]).max() < 2e-3 * max(np.abs(a["grot"]).max(), 1.0)


def test_brick_checkpoint_overflow_is_reported(monkeypatch):
    monkeypatch.setenv("DD_BRICK_CAP", "8")
    S = 2
    sc = make_scene(3000, 32, box_width=(0.2, 0.2, 0.2), steps=S, perturb=0.02, on_floor=True, seed=3, nb=2)
    sim = FusedSim.from_scene(sc, max_steps=S, grid_ckpt=2)
    sim.forward(0, S)
    with pytest.raises(EngineError, match="more active bricks"):
        sim.sync()
    sim.sync()   # reported once
    sim.close()


@pytest.mark.parametrize("n,E,chunk_max", [(4, 1, 32), (36, 2, 32), (1000, 3, 32), (5000, 1, 96), (5000, 2, 0)])
def test_tiled_rows_small_ragged_and_short_chunks(n, E, chunk_max):
    """Row tables with one short row, chunks much smaller than a brick (many spills into foreign columns, many lanes that
    share a cell within a row and go through the deferred-lane queue), several environments: the production tile path
    (fp32 SVD, staged kernels, ticket scheduling) against the dense path."""
    S = 4
    w = 0.05 + 0.1 * min(1.0, n / 5000.0)
    sc = make_scene(n, 32, box_center=(0.5, 0.3, 0.5), box_width=(w, w, w), steps=S, perturb=0.02, vel_scale=0.5, on_floor=True, nb=4, seed=77)
    seedg = loss_seed(n, 9)
    a = run_engine(sc, S, seedg, E=E, tile_mode=False, use_graphs=False, grid_ckpt=False)
    b = run_engine(sc, S, seedg, E=E, tile_mode=True, chunk_max=chunk_max)
    for k in ("x", "v", "F", "C"):
        assert rel_err(b["state"][k], a["state"][k]) < 2e-5, (k, rel_err(b["state"][k], a["state"][k]))
    for e in range(E):
        for k in ("x", "v"):
            assert_close_rows(b["grad"][k][e], a["grad"][k][e], 1e-3, k + "_grad")
    assert np.abs(b["gpos"] - a["gpos"]).max() < 2e-3 * max(np.abs(a["gpos"]).max(), 1.0)
