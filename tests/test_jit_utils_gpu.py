"""The generic helpers the north_star names (``quat_rotate_inverse`` / projected gravity, ``scale_transform`` ...) on the B200
kernels against goldens produced by the reference's OWN ``bez_isaacgym/utils/torch_jit_utils.py`` functions
(``tests/golden/fn_jit_utils.npz``, generator ``oracle/make_golden.py --jit-utils-only``): ``scale_transform`` /
``unscale_transform`` / ``saturate`` are defined in that file (pinned); ``quat_rotate[_inverse]`` reach it through the
star-import of ``isaacgym.torch_utils`` (restated, parity unpinned) and are exercised through its ``compute_rot`` / ``quat_axis``."""
import os

import numpy as np
import pytest
import torch

from tests import _util as U

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fn_jit_utils.npz")


def _g():
    return {k: torch.from_numpy(v) for k, v in np.load(GOLDEN).items()}


def _same(a, b):
    return bool(((a == b) | (a.isnan() & b.isnan())).all())


def test_scale_unscale_saturate_match_the_reference_functions():
    from bez_isaacgym_b200.utils import torch_jit_utils as tj
    g = _g()
    x, lo, hi = g["in_x"].cuda(), g["in_lower"].cuda(), g["in_upper"].cuda()
    assert _same(tj.saturate(x, lo, hi).cpu(), g["ref_saturate"]), "saturate is bit-exact (min / max with NaN propagation)"
    U.assert_close(tj.scale_transform(x, lo, hi), g["ref_scale"], what="scale_transform")
    U.assert_close(tj.unscale_transform(x, lo, hi), g["ref_unscale"], what="unscale_transform")
    # bit-exactness where it is defined: same op order, no fused multiply-add
    assert _same(tj.unscale_transform(x, lo, hi).cpu(), g["ref_unscale"])
    # round trip on finite values
    fin = torch.nan_to_num(g["in_x"], nan=0.3, posinf=1.0, neginf=-1.0).cuda()
    back = tj.unscale_transform(tj.scale_transform(fin, lo, hi), lo, hi)
    assert torch.allclose(back, fin, rtol=1e-5, atol=1e-5)


def test_quat_rotate_inverse_and_projected_gravity():
    from bez_isaacgym_b200.utils import torch_jit_utils as tj
    g = _g()
    q, v, w = g["in_q"].cuda(), g["in_v"].cuda(), g["in_w"].cuda()
    scale = (g["in_v"].abs().sum(1, keepdim=True) * (g["in_q"] ** 2).sum(1, keepdim=True) * 2).clamp_min(1.0)
    U.assert_close(tj.quat_rotate_inverse(q, v), g["ref_vel_loc"], scale=scale, what="compute_rot vel_loc")
    scale_w = (g["in_w"].abs().sum(1, keepdim=True) * (g["in_q"] ** 2).sum(1, keepdim=True) * 2).clamp_min(1.0)
    U.assert_close(tj.quat_rotate_inverse(q, w), g["ref_angvel_loc"], scale=scale_w, what="compute_rot angvel_loc")
    z = torch.tensor([[0.0, 0.0, 1.0]], device="cuda").repeat(q.shape[0], 1)
    U.assert_close(tj.quat_rotate(q, z), g["ref_quat_axis2"], scale=(g["in_q"] ** 2).sum(1, keepdim=True) * 2, what="quat_axis(q, 2)")
    # rotate then rotate back (unit quaternions): identity
    unit = q / q.norm(dim=1, keepdim=True)
    assert torch.allclose(tj.quat_rotate(unit, tj.quat_rotate_inverse(unit, v)), v, rtol=1e-4, atol=1e-5)
    # projected gravity of an upright robot is the gravity vector itself
    ident = torch.tensor([[0.0, 0.0, 0.0, 1.0]], device="cuda").repeat(5, 1)
    grav = torch.tensor([[0.0, 0.0, -1.0]], device="cuda").repeat(5, 1)
    assert torch.equal(tj.projected_gravity(ident, grav), grav)


def test_jit_utils_have_no_cpu_fallback():
    from bez_isaacgym_b200.utils import torch_jit_utils as tj
    with pytest.raises(Exception, match="CUDA only"):
        tj.quat_rotate_inverse(torch.zeros(2, 4), torch.zeros(2, 3))
