"""Device-side torch helpers kept for API compatibility (reference: utils/math_utils.py:40-58 and the
``isaacgym.torch_utils`` functions its callers star-import).  The fused kernels carry their own
inlined versions; these exist for the sparse host-side paths (resets) and for user code.
Quaternions are xyzw.
"""
import math

import torch


def normalize(x, eps: float = 1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)


def quat_apply(a, b):
    shape = b.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 3)
    xyz = a[:, :3]
    t = torch.cross(xyz, b, dim=-1) * 2
    return (b + a[:, 3:] * t + torch.cross(xyz, t, dim=-1)).view(shape)


def quat_rotate_inverse(q, v):
    w = q[:, -1:]
    qv = q[:, :3]
    return v * (2.0 * w * w - 1.0) - torch.cross(qv, v, dim=-1) * w * 2.0 + qv * (qv * v).sum(-1, keepdim=True) * 2.0


def quat_rotate(q, v):
    w = q[:, -1:]
    qv = q[:, :3]
    return v * (2.0 * w * w - 1.0) + torch.cross(qv, v, dim=-1) * w * 2.0 + qv * (qv * v).sum(-1, keepdim=True) * 2.0


def quat_mul(a, b):
    shape = a.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 4)
    x1, y1, z1, w1 = a.unbind(-1)
    x2, y2, z2, w2 = b.unbind(-1)
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    return torch.stack([qq - xx + (x1 + w1) * (x2 + w2), qq - yy + (w1 - x1) * (y2 + z2),
                        qq - zz + (z1 + y1) * (w2 - x2), qq - ww + (z1 - y1) * (y2 - z2)], dim=-1).view(shape)


def quat_apply_yaw(quat, vec):
    yaw = quat.clone().view(-1, 4)
    yaw[:, :2] = 0.0
    return quat_apply(normalize(yaw), vec)


def quat_apply_yaw_inverse(quat, vec):
    yaw = quat.clone().view(-1, 4)
    yaw[:, :2] = 0.0
    yaw[:, 2] = -yaw[:, 2]
    return quat_apply(normalize(yaw), vec)


def wrap_to_pi(angles):
    angles %= 2 * math.pi
    angles -= 2 * math.pi * (angles > math.pi)
    return angles


def torch_rand_float(lower, upper, shape, device):
    return (upper - lower) * torch.rand(*shape, device=device) + lower


def torch_rand_sqrt_float(lower, upper, shape, device):
    r = 2 * torch.rand(*shape, device=device) - 1
    r = torch.where(r < 0.0, -torch.sqrt(-r), torch.sqrt(r))
    return (upper - lower) * (r + 1.0) / 2.0 + lower
