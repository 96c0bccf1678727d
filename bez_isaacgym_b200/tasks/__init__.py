"""Task registry with the reference's names (``bez_isaacgym/tasks/__init__.py:10-16``): BezKick (the north-star path) and
its sibling tasks on the same kernel skeleton."""
from .kick_env import KickEnv
from .orient_env import OrientEnv
from .walk_env import WalkEnv

isaacgym_task_map = {
    "bez_kick": KickEnv,
    "bez_walk": WalkEnv,
    "bez_orient": OrientEnv,
}
