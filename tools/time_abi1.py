"""Times every ABI-1 kernel of a library (product or the reference CUDA build) on one scene: wall clock around K
back-to-back launches bracketed by stream syncs.  Usage: python tools/time_abi1.py [n_particles] [grid]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

from abi1_driver import Abi1Sim, loss_seed  # noqa: E402
from dexdeform_b200.scenes import make_scene  # noqa: E402


def bench(sim, K=10):
    cur, nxt = sim.states[0], sim.states[1]
    sim.substep(0)
    for k, v in loss_seed(sim.n).items():
        nxt[k].upload(v)
    sim.substep_grad(0)
    sim.sync()
    out = {}
    for name, fn in [("clear_temp", sim.clear_temp), ("compute_svd", lambda: sim.compute_svd(cur)), ("p2g", lambda: sim.p2g(cur, nxt)),
                     ("grid_op", lambda: sim.grid_op(cur, nxt)), ("g2p", lambda: sim.g2p(cur, nxt)),
                     ("clear_temp_grad", sim.clear_temp_grad), ("g2p_grad", lambda: sim.g2p_grad(cur, nxt)),
                     ("grid_op_grad", lambda: sim.grid_op_grad(cur, nxt)), ("p2g_grad", lambda: sim.p2g_grad(cur, nxt)),
                     ("compute_svd_grad", lambda: sim.compute_svd_grad(cur))]:
        fn(); sim.sync()
        t0 = time.perf_counter()
        for _ in range(K):
            fn()
        sim.sync()
        out[name] = (time.perf_counter() - t0) / K * 1e3
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    grid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    w = 0.4 if n >= 500000 else 0.09 * (n / 10000) ** (1 / 3)
    sc = make_scene(n, grid, box_center=(0.5, 0.3, 0.5), box_width=(w, w, w), steps=1, seed=0, perturb=0.01, vel_scale=0.1,
                    hand_scale=6.0 if n >= 500000 else 1.5)
    from dexdeform_b200.types import load_library
    from oracle.oracle_lib import load_ref_gpu
    res = {}
    for tag, lib in (("reference", load_ref_gpu()), ("product", load_library())):
        sim = Abi1Sim(lib, sc, 1)
        res[tag] = bench(sim)
        del sim
    print(f"n={n} grid={grid}^3   ms per launch")
    fwd = ("clear_temp", "compute_svd", "p2g", "grid_op", "g2p")
    for k in res["reference"]:
        print(f"  {k:18s} reference {res['reference'][k]:9.4f}   product {res['product'][k]:9.4f}")
    for tag in res:
        f = sum(res[tag][k] for k in fwd)
        b = f - res[tag]["g2p"] + sum(res[tag][k] for k in res[tag] if k not in fwd)
        print(f"  {tag}: fwd substep {f:.3f} ms, bwd substep {b:.3f} ms, fwd+bwd {f + b:.3f} ms -> {n / (f + b) / 1e3:.1f} M particle-substeps/s")
