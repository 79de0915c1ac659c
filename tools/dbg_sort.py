import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from abi1_driver import loss_seed
from conftest import rel_err
from dexdeform_b200.scenes import scene_tutorial
from oracle.oracle_lib import load_ref_gpu
from test_engine_gpu import run_abi1, run_engine
S = 4
sc = scene_tutorial(steps=S, perturb=0.02, vel_scale=0.3, on_floor=True, seed=2)
seedg = loss_seed(sc["n"], 3)
ref = run_abi1(load_ref_gpu(), sc, S, seedg)
for kw in (dict(sort_particles=False, tile_mode=False), dict(sort_particles=True, tile_mode=False), dict(sort_particles=True, tile_mode=True)):
    eng = run_engine(sc, S, seedg, svd_mode=0, use_graphs=False, **kw)
    d = np.abs(eng["grad"]["x"][0] - ref["grad"]["x_grad"]).max(axis=1)
    i = int(np.argmax(d))
    print(kw, "x_grad rel", rel_err(eng["grad"]["x"][0], ref["grad"]["x_grad"]), "worst particle", i, d[i], "x0", sc["x"][i], "n>1e-4:", int((d > 1e-4 * np.abs(ref["grad"]["x_grad"]).max()).sum()),
          "state x err", np.abs(eng["state"]["x"][0] - ref["state"]["x"]).max(), "v", rel_err(eng["state"]["v"][0], ref["state"]["v"]))
