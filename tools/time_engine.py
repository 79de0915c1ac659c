"""Times the fused engine on one scene: forward and backward ranges as CUDA graphs, CUDA events on the launch stream.
Usage: python tools/time_engine.py [n_particles] [grid] [substeps] [n_envs] [svd_mode]"""
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
S = int(sys.argv[3]) if len(sys.argv) > 3 else 20
E = int(sys.argv[4]) if len(sys.argv) > 4 else 1
svd = int(sys.argv[5]) if len(sys.argv) > 5 else 1
tile = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ckpt = int(sys.argv[7]) if len(sys.argv) > 7 else 1
w = 0.4 if n >= 500000 else 0.09 * (n / 10000) ** (1 / 3)
sc = make_scene(n, grid, box_center=(0.5, 0.3, 0.5), box_width=(w, w, w), steps=S, seed=0, hand_scale=6.0 if n >= 500000 else 1.5)
torch.cuda.init()
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    sim = FusedSim.from_scene(sc, n_envs=E, max_steps=S, svd_mode=svd, tile_mode=tile, grid_ckpt=ckpt, stream=stream.cuda_stream)
    gx = np.zeros((E, n, 3), np.float32); gx[..., 1] = -1.0 / n

    def fwd():
        sim.forward(0, S)

    def bwd():
        sim.zero_grad(S)
        sim.add_state_grad(S, gx)
        sim.backward(0, S)

    for _ in range(2):
        fwd(); bwd()
    sim.sync()
    res = {}
    for name, fn in (("fwd", fwd), ("bwd", None)):
        ts = []
        for _ in range(5):
            if name == "bwd":
                sim.zero_grad(S); sim.add_state_grad(S, gx)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if name == "fwd":
                fwd()
            else:
                sim.backward(0, S)
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name] = float(np.median(ts))
    tot = res["fwd"] + res["bwd"]
    units = E * n * S
    peak = 6547.5
    print(f"n={n} grid={grid}^3 S={S} E={E} svd_mode={svd} tile={tile} grid_ckpt={ckpt}")
    print(f"  forward  {res['fwd'] / S * 1e3:9.1f} us/substep   backward {res['bwd'] / S * 1e3:9.1f} us/substep")
    print(f"  fwd+bwd  {units / tot / 1e3:9.1f} M particle-substeps/s   {520 * units / tot / 1e6:8.1f} GB/s algorithmic = {520 * units / tot / 1e6 / peak * 100:.1f}% of {peak} GB/s")
