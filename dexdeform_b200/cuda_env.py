"""Host-side mirror of the reference's ``mpm/cuda_env.py``: ``CudaEnv`` -- tool (primitive) configuration parsing and the
construction of an ``MPMSimulator`` from ``SIMULATOR`` / ``PRIMITIVES`` config sections -- without yacs.

``parse_tools`` reproduces mpm/cuda_env.py:51-103 entry for entry, including what looks accidental but alters results:
softness is 666 for every primitive whatever the config says (cuda_env.py:88), a Box passes its three half extents plus a
zero, a Capsule ``(radius, half length, 0, 0)`` (cuda_env.py:80-85).  The renderer is out of scope:
``renderer`` is ``None`` and the render methods raise.
"""
import copy

import numpy as np

from .simulator import MPMSimulator


class CN(dict):
    """The sliver of ``yacs.config.CfgNode`` the reference's environment code relies on: a dict whose keys are attributes."""

    def __init__(self, init=None, new_allowed=True, **kw):
        super().__init__()
        for k, v in dict(init or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, CN(v) if isinstance(v, dict) and not isinstance(v, CN) else v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)


def merge_inputs(default, **overrides):
    """tools/config/utils.py ``merge_inputs``: ``default`` updated (recursively) with the keys given; unknown keys are added
    (the reference merges with ``new_allowed``: tool entries carry ``size`` / ``round``, which the default lacks)."""
    out = default.clone() if isinstance(default, CN) else CN(default)
    for k, v in overrides.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = merge_inputs(out[k], **v)
        else:
            out[k] = v
    return out


class CudaEnv:
    def __init__(self, cfg=None, cfg_path="configs/plb_cuda.yml", SIMULATOR=None, PRIMITIVES=None, RENDERER=None, SHAPES=None, **engine_kwargs):
        """mpm/cuda_env.py:12-32.  PRIMITIVES: list of tool dicts (``shape``, ``size``, ``round``, ``friction``, ``init_pos``,
        ``init_rot``, ``action.scale``); SIMULATOR: the simulator section (dict)."""
        n_bodies, kwargs = self.parse_tools(PRIMITIVES or [])
        self.simulator = MPMSimulator(n_bodies, cfg=SIMULATOR, **engine_kwargs)
        self.simulator.init_bodies(**kwargs)
        self.renderer = None
        n = self.simulator.n_particles
        self.particle_colors = np.zeros(n) + (((((255) << 8) + 255) << 8) + 255)
        self.simulator.states[0].x.upload(np.random.random(size=(n, 3)) * 0.2 + np.array((0.4, 0.1, 0.4)))

    def default_tool_config(self):
        """mpm/cuda_env.py:34-48."""
        cfg = CN()
        cfg.shape = ""
        cfg.init_pos = (0.3, 0.3, 0.3)
        cfg.init_rot = (1.0, 0.0, 0.0, 0.0)
        cfg.color = (0.3, 0.3, 0.3)
        cfg.lower_bound = (0.0, 0.0, 0.0)
        cfg.upper_bound = (1.0, 1.0, 1.0)
        cfg.friction = 0.9
        cfg.variations = None
        cfg.mass = 1.0
        cfg.stiffness = 0.0
        cfg.action = CN(dim=0, scale=())
        return cfg

    def parse_tools(self, cfgs):
        """mpm/cuda_env.py:51-103 -> (n_bodies, keyword arguments of MPMSimulator.init_bodies)."""
        tools = [merge_inputs(self.default_tool_config(), **dict(i)) for i in cfgs]
        self.action_dims = [0]
        types, softness, pos, rot, mu, round_, args, action_scales = [], [], [], [], [], [], [], []
        for i in tools:
            if i["shape"] == "Box":
                types.append(0)
                args.append([*i["size"], 0])
            elif i["shape"] == "Capsule":
                types.append(1)
                args.append([*i["size"], 0, 0])
            action_scales.append(i["action"]["scale"])
            softness.append(666.0)
            mu.append(i["friction"])
            round_.append(i["round"])
            pos.append(i.init_pos)
            rot.append(i.init_rot)
        return len(tools), {"types": types, "softness": softness, "mu": mu, "round": round_, "args": args, "action_scales": action_scales,
                            "pos": pos, "rot": rot}

    def render_rgb(self, *a, **k):
        raise NotImplementedError("rendering is out of scope for dexdeform_b200 (DESIGN.md, out of scope); use the reference renderer on get_state()")

    render_rgbd = render_rgb
