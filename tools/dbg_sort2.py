import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from abi1_driver import loss_seed
from conftest import rel_err
from dexdeform_b200.scenes import scene_tutorial
from oracle.oracle_lib import load_ref_gpu
from test_engine_gpu import run_abi1, run_engine, reference_spread
S = 4
sc = scene_tutorial(steps=S, perturb=0.02, vel_scale=0.3, on_floor=True, seed=2)
seedg = loss_seed(sc["n"], 3)
lib = load_ref_gpu()
ref = run_abi1(lib, sc, S, seedg)
eng = run_engine(sc, S, seedg, svd_mode=0, use_graphs=False, tile_mode=False)
print("before spread: x_grad rel", rel_err(eng["grad"]["x"][0], ref["grad"]["x_grad"]))
ref2 = run_abi1(lib, sc, S, seedg)
print("ref vs ref again", rel_err(ref2["grad"]["x_grad"], ref["grad"]["x_grad"]))
spread = reference_spread(lib, sc, S, seedg, ref)
print("spread", spread["x_grad"])
eng2 = run_engine(sc, S, seedg, svd_mode=0, use_graphs=False, tile_mode=False)
print("after spread: x_grad rel", rel_err(eng2["grad"]["x"][0], ref["grad"]["x_grad"]), "eng vs eng2", rel_err(eng2["grad"]["x"][0], eng["grad"]["x"][0]))
d = np.abs(eng2["grad"]["x"][0] - ref["grad"]["x_grad"]).max(axis=1); i = int(np.argmax(d)); print("worst", i, d[i], sc["x"][i], ref["grad"]["x_grad"][i], eng2["grad"]["x"][0][i])
