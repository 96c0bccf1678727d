"""Task registry with the reference's name (``bez_isaacgym/tasks/__init__.py:10-16``).  Only BezKick is in scope."""
from .kick_env import KickEnv

isaacgym_task_map = {
    "bez_kick": KickEnv,
}
