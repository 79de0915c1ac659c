import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from abi1_driver import Abi1Sim, loss_seed
from dexdeform_b200.engine import FusedSim
from dexdeform_b200.scenes import make_scene
from oracle.oracle_lib import OracleLib
from test_engine_gpu import run_abi1
S, E = 3, 3
scs = [make_scene(1200, 32, box_width=(0.12, 0.1, 0.12), steps=S, perturb=0.03, vel_scale=0.4, on_floor=True, seed=40 + e, nb=5) for e in range(E)]
for sc in scs[1:]:
    sc["tfsr"], sc["args"] = scs[0]["tfsr"], scs[0]["args"]
seedg = loss_seed(1200, 8)
orc = OracleLib()
refs = [run_abi1(orc, sc, S, seedg) for sc in scs]
sc0 = scs[0]
sim = FusedSim(E, 1200, 5, sc0["grid_dim"], sc0["dx"], sc0["dt"], S, sc0["ground_friction"], sc0["ground_height"], sc0["gravity"].reshape(3), svd_mode=0)
st = lambda k: np.ascontiguousarray(np.stack([sc[k] for sc in scs]))
sim.set_material(st("mass"), st("vol"), st("mu_lam_yield"))
sim.set_bodies(sc0["tfsr"], sc0["args"])
sim.set_poses(0, np.ascontiguousarray(np.stack([sc["pos"] for sc in scs], 1)), np.ascontiguousarray(np.stack([sc["rot"] for sc in scs], 1)))
sim.set_state(0, st("x"), st("v"), st("F"), st("C"))
sim.forward(0, S); sim.zero_grad(S)
t = lambda a: np.ascontiguousarray(np.broadcast_to(a[None], (E,) + a.shape))
sim.add_state_grad(S, t(seedg["x_grad"]), t(seedg["v_grad"]), t(seedg["F_grad"]), t(seedg["C_grad"]))
sim.backward(0, S)
gp0, gr0 = sim.get_pose_grads(0, S + 1)
state, grad = sim.get_state(S), sim.get_state_grad(0)
gp, gr = sim.get_pose_grads(0, S + 1)
for e in range(E):
    print(e, "ref max", np.abs(refs[e]["gpos"]).max(axis=(1, 2)), "eng(before get_state)", np.abs(gp0[:, e]).max(axis=(1, 2)), "eng(after)", np.abs(gp[:, e]).max(axis=(1, 2)))
