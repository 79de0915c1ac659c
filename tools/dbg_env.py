import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from abi1_driver import Abi1Sim, loss_seed
from dexdeform_b200.engine import FusedSim
from dexdeform_b200.scenes import make_scene
from dexdeform_b200.types import load_library
S = 3
sc = make_scene(1200, 32, box_width=(0.12, 0.1, 0.12), steps=S, perturb=0.03, vel_scale=0.4, on_floor=True, seed=40, nb=5)
seedg = loss_seed(1200, 8)
a = Abi1Sim(load_library(), sc, S)
for f in range(S): a.substep(f)
for k, v in seedg.items(): a.states[S][k].upload(v)
for f in range(S - 1, -1, -1): a.substep_grad(f)
ref = np.stack([a.get(f, "body_pos_grad")["body_pos_grad"] for f in range(S + 1)])
print("abi1 gpos absmax per f", np.abs(ref).max(axis=(1, 2)))
for E in (1, 3):
    for svd in (0, 1):
        for g in (False, True):
            sim = FusedSim.from_scene(sc, n_envs=E, max_steps=S, svd_mode=svd, use_graphs=g)
            sim.forward(0, S); sim.zero_grad(S)
            t = lambda x: np.ascontiguousarray(np.broadcast_to(x[None], (E,) + x.shape))
            sim.add_state_grad(S, t(seedg["x_grad"]), t(seedg["v_grad"]), t(seedg["F_grad"]), t(seedg["C_grad"]))
            sim.backward(0, S)
            gp, gr = sim.get_pose_grads(0, S + 1)
            print(f"E={E} svd={svd} graphs={g}: engine gpos absmax per env", [float(np.abs(gp[:, e]).max()) for e in range(E)],
                  "err vs abi1", [float(np.abs(gp[:, e] - ref).max()) for e in range(E)])
            sim.close()
