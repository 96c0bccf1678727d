"""``OrientEnv`` -- drop-in for the reference's turn-to-heading task (``bez_isaacgym/tasks/orient_env.py:37-620``).

As ``WalkEnv`` but the heading columns are ``compute_off_angle`` = (cos, sin) of ``goal_angle - normalize_angle(yaw)``
(``orient_env.py:719-733``; ``goal_angle`` (N,1) from ``goalState.goal_angle``) and the reward is ``orient_env.py:845-1014``
(angle term, win state on the SIGNED angle < 0.05, out of bound beyond 0.3 m from the start, -5 penalty).
"""
from .kick_env import KickEnv


class OrientEnv(KickEnv):
    TASK = "orient"
