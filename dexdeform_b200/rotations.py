"""Rotation helpers the hand layer needs.  The reference takes them from pytorch3d==0.7.2 and transforms3d==0.4.1
(environment.yml:154,233), neither of which is vendored; these restate their published formulas.

torch (differentiable, batched): ``axis_angle_to_matrix``, ``quaternion_to_matrix``, ``matrix_to_quaternion``
(pytorch3d/transforms/rotation_conversions.py); quaternions are (w, x, y, z).
numpy: ``euler2mat`` (transforms3d.euler, static xyz axes), ``axangle2mat``, ``quat2mat``."""
import numpy as np
import torch


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def axis_angle_to_quaternion(aa):
    angles = torch.norm(aa, p=2, dim=-1, keepdim=True)
    half = angles * 0.5
    small = angles.abs() < 1e-6
    s = torch.where(small, 0.5 - (angles * angles) / 48, torch.sin(half) / torch.where(small, torch.ones_like(angles), angles))
    return torch.cat([torch.cos(half), aa * s], dim=-1)


def axis_angle_to_matrix(aa):
    return quaternion_to_matrix(axis_angle_to_quaternion(aa))


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def matrix_to_quaternion(m):
    batch = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(batch + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].max(torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)))
    pick = torch.nn.functional.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    return cand[pick, :].reshape(batch + (4,))


# ---- numpy
def euler2mat(ai, aj, ak):
    """transforms3d.euler.euler2mat with the default 'sxyz' axes: R = Rz(ak) Ry(aj) Rx(ai)."""
    si, sj, sk = np.sin(ai), np.sin(aj), np.sin(ak)
    ci, cj, ck = np.cos(ai), np.cos(aj), np.cos(ak)
    return np.array([[cj * ck, sj * si * ck - ci * sk, sj * ci * ck + si * sk],
                     [cj * sk, sj * si * sk + ci * ck, sj * ci * sk - si * ck],
                     [-sj, cj * si, cj * ci]])


def axangle2mat(axis, angle):
    x, y, z = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    c, s = np.cos(angle), np.sin(angle)
    C = 1 - c
    return np.array([[x * x * C + c, x * y * C - z * s, x * z * C + y * s],
                     [y * x * C + z * s, y * y * C + c, y * z * C - x * s],
                     [z * x * C - y * s, z * y * C + x * s, z * z * C + c]])


def quat2mat(q):
    w, x, y, z = np.asarray(q, np.float64)
    n = w * w + x * x + y * y + z * z
    if n < np.finfo(np.float64).eps:
        return np.eye(3)
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
                     [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
                     [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)]])


def mujoco_euler2mat(e):
    """MuJoCo's default intrinsic xyz euler convention as the reference's parser evaluates it (mujoco_parser.py:16-38)."""
    ai, aj, ak = -e[2], -e[1], -e[0]
    si, sj, sk = np.sin(ai), np.sin(aj), np.sin(ak)
    ci, cj, ck = np.cos(ai), np.cos(aj), np.cos(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([[cj * ci, cj * si, -sj], [sj * cs - sc, sj * ss + cc, cj * sk], [sj * cc + ss, sj * sc - cs, cj * ck]])
