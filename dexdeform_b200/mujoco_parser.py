"""MJCF loader for the Shadow-hand description files that ship with DexDeform (mpm/assets/robots/shadow/*/shadow_hand.xml).

Produces exactly what the reference's ``mpm/mujoco_parser.py:378-550`` feeds to ``HandSimulator``: body frames (scaled),
the 24 hinge joints (position, axis), collision primitives (``group == 4`` after default-class expansion; boxes and
capsules; the forearm mesh is skipped) attached to the joint of their body, and for each of the five fingertip sites the
kinematic chain from the wrist as an alternating list of constant transforms and joint indices.

Only the standard library is used (xml.etree).  The asset files themselves are data of the DexDeform repository and
are looked up in ``DEXDEFORM_ASSETS`` (or the ``assets_dir`` argument), default ``/root/reference/mpm/assets``."""
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from .robots import FINGERTIP_SITES, JOINTS
from .rotations import mujoco_euler2mat, quat2mat

ROOT_BODY = "robot0:hand mount"


def default_assets_dir():
    return os.environ.get("DEXDEFORM_ASSETS", "/root/reference/mpm/assets")


def _load(path):
    root = ET.parse(path).getroot()
    base = os.path.dirname(os.path.abspath(path))
    for parent in root.findall(".//include/.."):
        new_children = []
        for child in list(parent):
            if child.tag == "include":
                new_children.extend(list(_load(os.path.join(base, child.get("file")))))
            else:
                new_children.append(child)
        for child in list(parent):
            parent.remove(child)
        parent.extend(new_children)
    return root


def _apply_defaults(root):
    """One level of <default class=...> expansion: ``class`` on an element, ``childclass`` on a body (applies to all descendants)."""
    templates = {}
    for d in root.find("default").findall("default"):
        templates[d.get("class")] = {t.tag: dict(t.attrib) for t in d}

    def fill(el, tmpl):
        for k, v in tmpl.get(el.tag, {}).items():
            if k not in el.attrib:
                el.set(k, v)

    def walk(el):
        for x in el:
            if "class" in x.attrib:
                fill(x, templates[x.get("class")])
            elif "childclass" in x.attrib:
                tmpl = templates[x.get("childclass")]
                for y in x.iter():
                    if y is not x:
                        fill(y, tmpl)
            walk(x)

    walk(root.find("worldbody"))


def _vec(s):
    return np.array([float(t) for t in s.split()], np.float64)


def _frame(el, scale):
    m = np.eye(4)
    m[:3, 3] = _vec(el.get("pos")) * scale
    if "euler" in el.attrib:
        m[:3, :3] = mujoco_euler2mat(_vec(el.get("euler")))
    elif "axisangle" in el.attrib:
        aa = _vec(el.get("axisangle"))
        ax = aa[:3] / np.linalg.norm(aa[:3])
        m[:3, :3] = quat2mat(np.r_[np.cos(aa[3] / 2), np.sin(aa[3] / 2) * ax])
    elif "quat" in el.attrib:
        m[:3, :3] = quat2mat(_vec(el.get("quat")))
    return m


@dataclass
class Primitive:
    parent_joint: str          # joint name of the body carrying the geom (or the body name if it has none)
    matrix: np.ndarray         # geom frame in the body
    kind: str                  # "box" | "capsule"
    size: List[float]          # scaled: box half extents, capsule (radius, half length)


@dataclass
class HandModel:
    root_matrix: np.ndarray
    frames: dict                                   # body name -> (4x4 frame in parent, parent name)
    joint_pos: np.ndarray                          # (24, 3) unscaled, as in the reference (mujoco_parser.py:441-444)
    joint_axis: np.ndarray                         # (24, 3)
    chains: List[List[Tuple[str, object]]] = field(default_factory=list)   # per fingertip: ("mat", 4x4) | ("joint", index), wrist -> tip
    primitives: List[Primitive] = field(default_factory=list)

    def frame_of(self, name):
        return self.frames[name][0]


def load_hand(side="right_hand", scale=1.0, assets_dir=None):
    path = os.path.join(assets_dir or default_assets_dir(), "robots", "shadow", side, "shadow_hand.xml")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path}: point DEXDEFORM_ASSETS at DexDeform's mpm/assets directory")
    root = _load(path)
    _apply_defaults(root)
    jidx = {n: i for i, n in enumerate(JOINTS)}
    frames, body_joint, sites = {}, {}, {}
    jpos, jaxis = np.zeros((len(JOINTS), 3)), np.zeros((len(JOINTS), 3))
    prims = []

    def visit(body, parent):
        name = body.get("name", f"noname_body_{len(frames)}")
        frames[name] = (_frame(body, scale), parent)
        j = body.find("joint")
        jname = None
        if j is not None:
            jname = j.get("name", "")
            if jname in jidx:
                if j.get("type", "hinge") != "hinge":
                    raise ValueError("only hinge joints are supported")
                jpos[jidx[jname]], jaxis[jidx[jname]] = _vec(j.get("pos")), _vec(j.get("axis"))
                body_joint[name] = jname
        for g in body.findall("geom"):
            if g.get("group", "") != "4" or g.get("mesh", ""):
                continue
            prims.append(Primitive(jname if jname is not None else name, _frame(g, scale) if g.get("pos", "") else np.eye(4), g.get("type", ""),
                                   [float(t) * scale for t in g.get("size").split()]))
        for s in body.findall("site"):
            if s.get("name", "") in FINGERTIP_SITES:
                sites[s.get("name")] = (_frame(s, scale), name)
        for child in body.findall("body"):
            visit(child, name)

    top = [b for b in root.find("worldbody").findall("body") if b.get("name", "") == ROOT_BODY]
    if not top:
        raise ValueError("no root body found in xml")
    visit(top[0], None)

    chains = []
    for site in FINGERTIP_SITES:
        m, body = sites[site]
        ops = []
        while body is not None:
            pm, parent = frames[body]
            if body in body_joint:
                ops.append(("mat", m))
                ops.append(("joint", jidx[body_joint[body]]))
                m = pm
            else:
                m = pm @ m
            body = parent
        chains.append(list(reversed(ops)))
    root_matrix = np.eye(4)
    root_matrix[:3, :3] = mujoco_euler2mat(np.array([np.pi / 2, 0, np.pi]))
    root_matrix[:3, 3] = [1.0, 1.25, 0.15]
    return HandModel(root_matrix, frames, jpos, jaxis, chains, prims)


@dataclass
class HandTables:
    """Flat arrays for the forward kinematics (host torch mirror and the device kernel share them)."""
    n_hands: int
    root_frame: np.ndarray      # (nh, 4, 4) default wrist pose
    joint_pos: np.ndarray       # (nh, 24, 3)
    joint_axis: np.ndarray      # (nh, 24, 3)
    op_kind: np.ndarray         # (n_ops,) 0 = constant matrix, 1 = joint
    op_index: np.ndarray        # (n_ops,) matrix index or joint index
    op_reset: np.ndarray        # (n_ops,) 1 = first op of a chain (restart from the base pose)
    mats: np.ndarray            # (nh, n_mats, 4, 4)
    geom_joint: np.ndarray      # (n_geoms,) joint whose pose carries the primitive
    geom_local: np.ndarray      # (nh, n_geoms, 4, 4) primitive frame in that joint's body (capsules rotated so their axis is +y)
    prim_type: np.ndarray       # (nh * n_geoms,) 0 box, 1 capsule
    prim_size: np.ndarray       # (nh * n_geoms, 4) args as the kernels expect them


def hand_tables(models):
    from .rotations import axangle2mat
    nh = len(models)
    kinds, index, reset = [], [], []
    mats = [[] for _ in range(nh)]
    for ci, chain in enumerate(models[0].chains):
        for oi, (k, v) in enumerate(chain):
            reset.append(1 if oi == 0 else 0)
            if k == "mat":
                kinds.append(0)
                index.append(len(mats[0]))
                for h in range(nh):
                    mats[h].append(models[h].chains[ci][oi][1])
            else:
                kinds.append(1)
                index.append(v)
    geom_joint, geom_local, ptype, psize = [], [[] for _ in range(nh)], [], []
    for h, m in enumerate(models):
        for p in m.primitives:
            if "forearm" in p.parent_joint:
                continue
            mat = p.matrix.copy()
            if p.kind == "capsule":   # MJCF capsules run along z, the SDF along y (hand.py:161-162)
                mat[:3, :3] = mat[:3, :3] @ axangle2mat([1, 0, 0], np.pi / 2)
                psize.append([p.size[0], p.size[1], 0.0, 0.0])   # (radius, half length) as passed on by cuda_env.py:85
                ptype.append(1)
            else:
                psize.append([p.size[0], p.size[1], p.size[2], 0.0])
                ptype.append(0)
            if h == 0:
                geom_joint.append(JOINTS.index(p.parent_joint))
            geom_local[h].append(mat)
    root = np.stack([m.root_matrix @ m.frame_of(ROOT_BODY) @ m.frame_of("robot0:wrist") for m in models])
    return HandTables(nh, root, np.stack([m.joint_pos for m in models]), np.stack([m.joint_axis for m in models]), np.array(kinds, np.int32),
                      np.array(index, np.int32), np.array(reset, np.int32), np.array(mats), np.array(geom_joint, np.int32), np.array(geom_local),
                      np.array(ptype, np.int32), np.array(psize, np.float32))
