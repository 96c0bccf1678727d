"""``WalkEnv`` -- drop-in for the reference's walking task (``bez_isaacgym/tasks/walk_env.py:37-620``) on the B200 kernels.

Same skeleton as ``KickEnv`` (K0 + ONE fused post-physics launch); what differs lives in the kernel's ``BEZK_TASK_WALK``
variant (``include/bezk.h``): robot-only root tensor ``(N,13)``, 21 bodies, 52-wide observation
``[dof_pos, dof_vel, imu, off_orn, feet]`` (``walk_env.py:1032-1050``), the walking reward with the up-vector projection, win
state and out-of-bound angle (``:827-997``), 10 s episodes, and the goal redraw on reset -- the reference assigns the FIRST
``U(-2,2)^2`` draw of a reset batch to every env resetting in that step (``:566-574``), which is kept (one draw per step).
"""
from .kick_env import KickEnv


class WalkEnv(KickEnv):
    TASK = "walk"
