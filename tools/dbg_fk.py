import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
from test_hand import tables
from dexdeform_b200.engine import FusedSim
from dexdeform_b200.hand import DeviceFK, read_poses, HandKinematics
t = tables("rh15"); S = 2
eng = FusedSim(1, 64, 19, (32, 32, 32), 1 / 32, 1e-4, S)
fk = DeviceFK(t, np.array([0.33 * 0.002] * 20 + [0.0] * 6))
kin = HandKinematics(t, "cuda")
base = torch.tensor(t.root_frame, dtype=torch.float32, device="cuda")[None]
for name, q0 in (("zero", torch.zeros(1, 1, 24, device="cuda")), ("q3=0.3", torch.zeros(1, 1, 24, device="cuda").index_fill_(2, torch.tensor([1], device="cuda"), 0.3))):
    act = torch.zeros(1, 1, 26, device="cuda")
    fk.run(eng, 0, S, base, q0, act, has_base_action=False)
    pd, rd = read_poses(eng, 1, S)
    pt, rt = kin.forward(base.expand(S, -1, -1, -1), q0.expand(S, -1, -1))
    print(name, "pos err per geom", (pd[0, 0] - pt[0]).norm(dim=-1).cpu().numpy().round(4))
print(t.op_kind[:14], t.op_index[:14], t.op_reset[:14])
