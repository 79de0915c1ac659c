"""Shadow-hand joint / actuator tables (facts of the MJCF model the reference lists in mpm/robots/interface.py:13-145).

24 hinge joints in kinematic order, 20 position actuators (the two distal joints of FF/MF/RF/LF are tendon-coupled to
one actuator), control ranges in radians, the default grasp-ready pose, and the derived joint limits
(actuator range divided by the number of joints it drives, interface.py:122-126)."""
import numpy as np

FINGERTIP_SITES = ["robot0:S_fftip", "robot0:S_mftip", "robot0:S_rftip", "robot0:S_lftip", "robot0:S_thtip"]

_J = "WRJ1 WRJ0 FFJ3 FFJ2 FFJ1 FFJ0 MFJ3 MFJ2 MFJ1 MFJ0 RFJ3 RFJ2 RFJ1 RFJ0 LFJ4 LFJ3 LFJ2 LFJ1 LFJ0 THJ4 THJ3 THJ2 THJ1 THJ0"
JOINTS = ["robot0:" + j for j in _J.split()]

# actuator -> (driven joints, control range)
ACTUATORS = [
    ("A_WRJ1", ["WRJ1"], (-0.4887, 0.1396)), ("A_WRJ0", ["WRJ0"], (-0.6981, 0.4887)),
    ("A_FFJ3", ["FFJ3"], (-0.3491, 0.3491)), ("A_FFJ2", ["FFJ2"], (0.0, 1.5708)), ("A_FFJ1", ["FFJ1", "FFJ0"], (0.0, 3.1416)),
    ("A_MFJ3", ["MFJ3"], (-0.3491, 0.3491)), ("A_MFJ2", ["MFJ2"], (0.0, 1.5708)), ("A_MFJ1", ["MFJ1", "MFJ0"], (0.0, 3.1416)),
    ("A_RFJ3", ["RFJ3"], (-0.3491, 0.3491)), ("A_RFJ2", ["RFJ2"], (0.0, 1.5708)), ("A_RFJ1", ["RFJ1", "RFJ0"], (0.0, 3.1416)),
    ("A_LFJ4", ["LFJ4"], (0.0, 0.7854)), ("A_LFJ3", ["LFJ3"], (-0.3491, 0.3491)), ("A_LFJ2", ["LFJ2"], (0.0, 1.5708)),
    ("A_LFJ1", ["LFJ1", "LFJ0"], (0.0, 3.1416)),
    ("A_THJ4", ["THJ4"], (-1.0472, 1.0472)), ("A_THJ3", ["THJ3"], (0.0, 1.2217)), ("A_THJ2", ["THJ2"], (-0.2094, 0.2094)),
    ("A_THJ1", ["THJ1"], (-0.5236, 0.5236)), ("A_THJ0", ["THJ0"], (-1.5708, 0.0)),
]
N_JOINTS, N_ACTUATORS = len(JOINTS), len(ACTUATORS)

DEFAULT_INITIAL_QPOS = dict(zip(JOINTS, [
    -0.16514339750464327, -0.31973286565062153, 0.14340512546557435, 0.32028208333591573, 0.7126053607727917, 0.6705281001412586,
    0.000246444303701037, 0.3152655251085491, 0.7659800313729842, 0.7323156897425923, 0.00038520700007378114, 0.36743546201985233,
    0.7119514095008576, 0.6699446327514138, 0.0525442258033891, -0.13615534724474673, 0.39872030433433003, 0.7415570009679252,
    0.704096378652974, 0.003673823825070126, 0.5506291436028695, -0.014515151997119306, -0.0015229223564485414, -0.7894883021600622]))


def actuator_of_joint():
    """Index of the actuator driving each of the 24 joints (interface.py:133-145)."""
    out = [None] * N_JOINTS
    for a, (_, joints, _) in enumerate(ACTUATORS):
        for j in joints:
            out[JOINTS.index("robot0:" + j)] = a
    return out


def joint_limits():
    """(24, 2) lower/upper limits: the actuator's control range split evenly over the joints it drives."""
    lim = np.zeros((N_JOINTS, 2))
    for _, joints, rng in ACTUATORS:
        for j in joints:
            lim[JOINTS.index("robot0:" + j)] = np.array(rng) / len(joints)
    return lim
