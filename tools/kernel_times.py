"""Per-kernel device times (CUDA events) of one forward+backward substep + graph-launch totals.  Usage: python tools/kernel_times.py [n] [grid] [S] [E]
Select an alternative build with DEXDEFORM_B200_LIB=/path/to/lib.so"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dexdeform_b200.engine import FusedSim  # noqa: E402
from dexdeform_b200.scenes import make_scene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
S = int(sys.argv[3]) if len(sys.argv) > 3 else 10
E = int(sys.argv[4]) if len(sys.argv) > 4 else 1
kw, skw = {}, {}
for a in sys.argv[5:]:
    k, v = a.split("=")
    if k.startswith("scene_"):
        skw[k[6:]] = int(v)  # e.g. scene_nb=0: the same scene without primitives
    else:
        kw[k] = int(v)
w = 0.4 if n >= 500000 else 0.09 * (n / 10000) ** (1 / 3)
sc = make_scene(n, grid, box_center=(0.5, 0.3, 0.5), box_width=(w, w, w), steps=S, seed=0, hand_scale=6.0 if n >= 500000 else 1.5, **skw)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sim = FusedSim.from_scene(sc, n_envs=E, max_steps=S, stream=stream.cuda_stream, **kw)
gx = np.zeros((E, n, 3), np.float32); gx[..., 1] = -1.0 / n
for _ in range(2):
    sim.forward(0, S); sim.zero_grad(S); sim.add_state_grad(S, gx); sim.backward(0, S)
sim.sync()
ts = {}
for name in ("fwd", "bwd"):
    v = []
    for _ in range(5):
        if name == "bwd":
            sim.zero_grad(S); sim.add_state_grad(S, gx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.forward(0, S) if name == "fwd" else sim.backward(0, S)
        e1.record(stream); e1.synchronize()
        v.append(e0.elapsed_time(e1))
    ts[name] = float(np.median(v)) / S * 1e3
sim.forward(0, S); sim.zero_grad(S); sim.add_state_grad(S, gx)
prof = sim.profile_substep(S - 1, reps=10)
tot = ts["fwd"] + ts["bwd"]
print(f"[{os.environ.get('DEXDEFORM_B200_LIB', 'default')}] n={n} E={E} {kw}: fwd {ts['fwd']:.1f} bwd {ts['bwd']:.1f} us/substep -> {E * n / tot:.0f} M p-s/s, {520 * E * n / tot / 1e3 / 6547.5 * 100:.1f}% | " +
      ", ".join(f"{k.split(' ')[0]} {v * 1e3:.0f}" for k, v in prof))
