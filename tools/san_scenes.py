"""Small ragged tiled-path scenes forward + backward without CUDA graphs, for compute-sanitizer (profiles/r02_sanitizer.md)."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from dexdeform_b200.engine import FusedSim
from dexdeform_b200.scenes import make_scene
for n, E, cm in ((36, 2, 32), (1500, 2, 32), (4000, 1, 96)):
    S = 3
    w = 0.05 + 0.1 * min(1.0, n / 5000.0)
    sc = make_scene(n, 32, box_center=(0.5, 0.3, 0.5), box_width=(w, w, w), steps=S, perturb=0.02, vel_scale=0.5, on_floor=True, nb=4, seed=77)
    sim = FusedSim.from_scene(sc, n_envs=E, max_steps=S, chunk_max=cm, use_graphs=False)
    sim.forward(0, S)
    sim.zero_grad(S)
    g = np.zeros((E, n, 3), np.float32); g[..., 1] = -1.0 / n
    sim.add_state_grad(S, g)
    sim.backward(0, S)
    x = sim.get_state_grad(0)["x"]
    print("ok", n, E, cm, float(np.abs(x).sum()))
    sim.close()
