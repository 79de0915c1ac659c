"""Device time of dd_sim_get_obs / dd_sim_compute_dist (CUDA events).  python tools/time_obs.py [n] [grid] [E]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from dexdeform_b200.engine import FusedSim
from dexdeform_b200.scenes import make_scene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
E = int(sys.argv[3]) if len(sys.argv) > 3 else 1
w = 0.4 if n >= 500000 else 0.09 * (n / 10000) ** (1 / 3)
sc = make_scene(n, grid, box_center=(0.5, 0.3, 0.5), box_width=(w, w, w), steps=2, seed=0, hand_scale=6.0 if n >= 500000 else 1.5)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sim = FusedSim.from_scene(sc, n_envs=E, max_steps=2, stream=stream.cuda_stream)
res = {}
for name, fn in (("get_obs", lambda: sim.get_obs(0)), ("compute_dist", lambda: sim.compute_dist(0, device=True))):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): fn()
    e1.record(stream); e1.synchronize()
    res[name] = e0.elapsed_time(e1) / 10 * 1e3
obs = sim.get_obs(0); d = sim.compute_dist(0, device=True)
print(f"[{os.environ.get('DEXDEFORM_B200_LIB', 'default')}] n={n} E={E}: " + ", ".join(f"{k} {v:.0f} us" for k, v in res.items()), "| checksum", float(obs.double().sum()), float(d.double().sum()))
