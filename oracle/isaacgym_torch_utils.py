"""Restated ``isaacgym.torch_utils`` helpers (oracle; test infrastructure only).

Isaac Gym Preview 3 (``python/isaacgym/torch_utils.py``) is a proprietary download that is not
vendored under /root/reference, so these helpers are restated from the published algorithm
(SURVEY.md Appendix B).  PARITY UNPINNED: no reference test pins them.  The reference reaches them
through ``from isaacgym.torch_utils import *`` in ``bez_isaacgym/utils/torch_jit_utils.py:31`` and
uses them at ``bez_isaacgym/tasks/kick_env.py:164,216-219,405-408,417,786-790,950,1247,1249``.

Quaternions are (x, y, z, w).
"""
from typing import List, Tuple

import numpy as np
import torch
from torch import Tensor


def to_torch(x, dtype=torch.float, device="cpu", requires_grad=False):
    return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)


@torch.jit.script
def tensor_clamp(t: Tensor, min_t: Tensor, max_t: Tensor) -> Tensor:
    return torch.max(torch.min(t, max_t), min_t)


@torch.jit.script
def torch_rand_float(lower: float, upper: float, shape: Tuple[int, int], device: str) -> Tensor:
    return (upper - lower) * torch.rand(shape[0], shape[1], device=device) + lower


@torch.jit.script
def quat_conjugate(a: Tensor) -> Tensor:
    shape = a.shape
    a = a.reshape(-1, 4)
    return torch.cat((-a[:, :3], a[:, -1:]), dim=-1).view(shape)


def get_axis_params(value, axis_idx, x_value=0.0, dtype=float, n_dims=3):
    zs = np.zeros((n_dims,))
    assert axis_idx < n_dims
    zs[axis_idx] = 1.0
    params = np.where(zs == 1.0, value, zs)
    params[0] = x_value
    return list(params.astype(dtype))


@torch.jit.script
def quat_rotate(q: Tensor, v: Tensor) -> Tensor:
    shape = q.shape
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3), v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
    return a + b + c


@torch.jit.script
def quat_rotate_inverse(q: Tensor, v: Tensor) -> Tensor:
    shape = q.shape
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3), v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
    return a - b + c


@torch.jit.script
def get_basis_vector(q: Tensor, v: Tensor) -> Tensor:
    return quat_rotate(q, v)


@torch.jit.script
def normalize_angle(x: Tensor) -> Tensor:
    return torch.atan2(torch.sin(x), torch.cos(x))


@torch.jit.script
def copysign(a: float, b: Tensor) -> Tensor:
    a_t = torch.tensor(a, device=b.device, dtype=torch.float).repeat(b.shape[0])
    return torch.abs(a_t) * torch.sign(b)


@torch.jit.script
def get_euler_xyz(q: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    qx, qy, qz, qw = 0, 1, 2, 3
    sinr_cosp = 2.0 * (q[:, qw] * q[:, qx] + q[:, qy] * q[:, qz])
    cosr_cosp = q[:, qw] * q[:, qw] - q[:, qx] * q[:, qx] - q[:, qy] * q[:, qy] + q[:, qz] * q[:, qz]
    roll = torch.atan2(sinr_cosp, cosr_cosp)

    sinp = 2.0 * (q[:, qw] * q[:, qy] - q[:, qz] * q[:, qx])
    pitch = torch.where(torch.abs(sinp) >= 1, copysign(np.pi / 2.0, sinp), torch.asin(sinp))

    siny_cosp = 2.0 * (q[:, qw] * q[:, qz] + q[:, qx] * q[:, qy])
    cosy_cosp = q[:, qw] * q[:, qw] + q[:, qx] * q[:, qx] - q[:, qy] * q[:, qy] - q[:, qz] * q[:, qz]
    yaw = torch.atan2(siny_cosp, cosy_cosp)

    return roll % (2 * np.pi), pitch % (2 * np.pi), yaw % (2 * np.pi)


# --- helpers that are only *compiled* (never executed on the BezKick path): the reference's
# utils/torch_jit_utils.py:34-181 jit-compiles functions that name them at import time. ---
@torch.jit.script
def quat_mul(a: Tensor, b: Tensor) -> Tensor:
    shape = a.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 4)
    ax, ay, az, aw = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    bx, by, bz, bw = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    x = aw * bx + ax * bw + ay * bz - az * by
    y = aw * by - ax * bz + ay * bw + az * bx
    z = aw * bz + ax * by - ay * bx + az * bw
    w = aw * bw - ax * bx - ay * by - az * bz
    return torch.stack([x, y, z, w], dim=-1).view(shape)


@torch.jit.script
def normalize(x: Tensor, eps: float = 1e-9) -> Tensor:
    return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)


@torch.jit.script
def quat_axis(q: Tensor, axis: int = 0) -> Tensor:
    basis_vec = torch.zeros(q.shape[0], 3, device=q.device)
    basis_vec[:, axis] = 1
    return quat_rotate(q, basis_vec)


__all__ = [
    "quat_mul", "normalize", "quat_axis",
    "to_torch", "tensor_clamp", "torch_rand_float", "quat_conjugate", "get_axis_params",
    "quat_rotate", "quat_rotate_inverse", "get_basis_vector", "normalize_angle", "copysign",
    "get_euler_xyz", "torch", "np", "Tensor", "Tuple", "List",
]
