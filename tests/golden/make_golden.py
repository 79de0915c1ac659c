"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh from
/root/reference) through the reference's substep / substep_grad call sequence on small seeded scenes.

    python tests/golden/make_golden.py            # reference host build (libmaniskill_mpm_cpu.so), runs anywhere
    python tests/golden/make_golden.py --gpu      # reference CUDA build on a GPU box (files get the suffix _gpu)

Inputs are not stored: they are regenerated from ``scene_kwargs`` by dexdeform_b200.scenes.make_scene (pure numpy,
seeded).  Grids are stored sparsely (indices of nodes with mass)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from abi1_driver import Abi1Sim, loss_seed  # noqa: E402
from dexdeform_b200.scenes import make_scene  # noqa: E402

CASES = {
    # exercises plastic + elastic branches, contact with boxes and capsules, floor friction (Coulomb branch)
    "block_on_floor": dict(n_particles=1500, grid=32, box_width=(0.12, 0.12, 0.12), steps=4, perturb=0.02, vel_scale=0.3,
                           on_floor=True, seed=3),
    # sticky floor (ground_friction >= 99: forward zeroes v, adjoint uses the Coulomb formula), no perturbation (F = I)
    "rest_sticky": dict(n_particles=1200, grid=32, box_width=(0.10, 0.06, 0.10), steps=4, perturb=0.0, vel_scale=0.05,
                        on_floor=True, ground_friction=500.0, yield_stress=130.0, E=4e3, gravity=(0.0, -3.0, 0.0), seed=7),
    # no bodies, frictionless floor, near a wall
    "free_fall_wall": dict(n_particles=800, grid=32, box_center=(0.13, 0.2, 0.5), box_width=(0.06, 0.06, 0.06), steps=4, perturb=0.01,
                           vel_scale=1.0, ground_friction=0.0, nb=0, seed=11),
}
STEPS = 4


def run_case(lib, kw):
    scene = make_scene(**kw)
    sim = Abi1Sim(lib, scene, STEPS)
    out = {}
    for f in range(STEPS):
        sim.substep(f)
        if f == 0:
            t = sim.get_temp("grid_m", "grid_v_in", "grid_v_out", "sig")
            idx = np.nonzero(t["grid_m"] > 0)[0].astype(np.int32)
            out.update(grid_idx=idx, grid_m=t["grid_m"][idx], grid_v_in=t["grid_v_in"][idx], grid_v_out=t["grid_v_out"][idx], sig0=t["sig"])
    for f in (1, STEPS):
        for k, v in sim.get(f).items():
            out[f"s{f}_{k}"] = v
    for k, v in loss_seed(scene["n"]).items():
        sim.states[STEPS][k].upload(v)
    for f in range(STEPS - 1, -1, -1):
        sim.substep_grad(f)
    g = sim.get(0, "x_grad", "v_grad", "F_grad", "C_grad")
    out.update({f"g0_{k}": v for k, v in g.items()})
    if scene["nb"]:
        out["pos_grad"] = np.stack([sim.get(f, "body_pos_grad")["body_pos_grad"] for f in range(STEPS + 1)])
        out["rot_grad"] = np.stack([sim.get(f, "body_rot_grad")["body_rot_grad"] for f in range(STEPS + 1)])
        dist = __import__("dexdeform_b200.types", fromlist=["array"]).array(length=scene["n"] * scene["nb"], library=lib)
        sim.compute_dist(sim.states[STEPS], dist, dist, 0)
        sim.sync()
        out["dist"] = dist.download().reshape(scene["n"], scene["nb"])
    return out


def main():
    from oracle.oracle_lib import load_ref_cpu, load_ref_gpu
    gpu = "--gpu" in sys.argv
    lib = load_ref_gpu() if gpu else load_ref_cpu()
    outdir = os.environ.get("GOLDEN_OUT", HERE)
    os.makedirs(outdir, exist_ok=True)
    for name, kw in CASES.items():
        out = run_case(lib, kw)
        # the reference is not bitwise reproducible (float atomics): run it again with a different
        # summation order and record its own run-to-run spread per field; parity tolerances are tied to it
        if not gpu:
            lib.ref_cpu_set_num_threads(3)
        again = run_case(lib, kw)
        if not gpu:
            lib.ref_cpu_set_num_threads(lib.ref_cpu_num_threads() if False else 8)
        spread = {k: float(np.abs(again[k].astype(np.float64) - out[k]).max() / (np.abs(out[k]).max() + 1e-30))
                  for k in out if k != "grid_idx"}
        path = os.path.join(outdir, f"{name}{'_gpu' if gpu else ''}.npz")
        np.savez_compressed(path, scene_kwargs=json.dumps(kw), ref_spread=json.dumps(spread), **out)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
