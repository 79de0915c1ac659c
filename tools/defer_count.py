"""Diagnostic (build_variants/lib_cnt.so, -DDD_COUNT_DEFER): share of particle-rows that leave the tile path (same cell as a lower lane of the row, or outside the tile) in k_p2g_tile + k_g2p_grad_tile.  python tools/defer_count.py n grid S E"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
os.environ["DEXDEFORM_B200_LIB"] = os.path.join(ROOT, "build_variants", "lib_cnt.so")
import numpy as np, torch
from dexdeform_b200.engine import FusedSim
from dexdeform_b200.scenes import make_scene
n, grid, S, E = (int(a) for a in sys.argv[1:5])
w = 0.4 if n >= 500000 else 0.09 * (n / 10000) ** (1 / 3)
sc = make_scene(n, grid, box_center=(0.5, 0.3, 0.5), box_width=(w, w, w), steps=S, seed=0, hand_scale=6.0 if n >= 500000 else 1.5)
sim = FusedSim.from_scene(sc, n_envs=E, max_steps=S)
lib = sim.lib
out = (ctypes.c_ulonglong * 2)()
def counts():
    lib.dd_debug_defer_counts(out); return out[0], out[1]
c0 = counts()
sim.forward(0, S); sim.sync()
c1 = counts()
gx = np.zeros((E, n, 3), np.float32); gx[..., 1] = -1.0 / n
sim.zero_grad(S); sim.add_state_grad(S, gx); sim.backward(0, S); sim.sync()
c2 = counts()
rows = E * n * S / 32
print(f"n={n} E={E} S={S}: forward deferred {100 * (c1[0] - c0[0]) / (E * n * S):.2f}% of particles, {100 * (c1[1] - c0[1]) / rows:.1f}% of rows; backward {100 * (c2[0] - c1[0]) / (E * n * S):.2f}% / {100 * (c2[1] - c1[1]) / rows:.1f}%")
