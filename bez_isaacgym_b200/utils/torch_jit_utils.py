"""The generic helpers of the reference's ``bez_isaacgym/utils/torch_jit_utils.py`` (which star-imports ``isaacgym.torch_utils``)
that the north_star names -- ``quat_rotate_inverse`` (projected gravity), ``quat_rotate``, ``scale_transform`` /
``unscale_transform`` / ``saturate`` -- with the reference's names and argument order, each ONE launch of a libbezk kernel.

``KickEnv`` does not use them (its ``quat_rotate_inverse`` calls are commented out, ``tasks/kick_env.py:905-908``; DOF values are
concatenated unscaled, ``:1398-1417``): they exist so that task code written against the reference's helper module runs on the
B200 path unchanged.  No CPU fallback: CPU tensors raise."""
import torch

from .. import ops


def quat_rotate(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """isaacgym.torch_utils.quat_rotate: rotate ``v`` (N,3) by the xyzw quaternion ``q`` (N,4)."""
    return ops.quat_rotate(q.contiguous(), v.contiguous(), inverse=False)


def quat_rotate_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """isaacgym.torch_utils.quat_rotate_inverse (used by ``compute_rot``, torch_jit_utils.py:52-63)."""
    return ops.quat_rotate(q.contiguous(), v.contiguous(), inverse=True)


def projected_gravity(q: torch.Tensor, gravity_vec: torch.Tensor) -> torch.Tensor:
    """``quat_rotate_inverse(base_quat, gravity_vec)``: the gravity direction in the body frame."""
    return quat_rotate_inverse(q, gravity_vec)


def scale_transform(x: torch.Tensor, lower: torch.Tensor, upper: torch.Tensor) -> torch.Tensor:
    """torch_jit_utils.py:78-96: normalise to [-1, 1]."""
    return ops.scale_transform(x.contiguous(), lower.contiguous(), upper.contiguous(), mode="scale")


def unscale_transform(x: torch.Tensor, lower: torch.Tensor, upper: torch.Tensor) -> torch.Tensor:
    """torch_jit_utils.py:99-117: back from [-1, 1] to (lower, upper)."""
    return ops.scale_transform(x.contiguous(), lower.contiguous(), upper.contiguous(), mode="unscale")


def saturate(x: torch.Tensor, lower: torch.Tensor, upper: torch.Tensor) -> torch.Tensor:
    """torch_jit_utils.py:119-134: ``max(min(x, upper), lower)``."""
    return ops.scale_transform(x.contiguous(), lower.contiguous(), upper.contiguous(), mode="saturate")
