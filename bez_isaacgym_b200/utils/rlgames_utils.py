"""rl_games vecenv adapter with the reference's names and signatures (``bez_isaacgym/utils/rlgames_utils.py:39-98,157-181``).

``RLGPUEnv`` is what rl_games' ``vecenv.register('RLGPU', ...)`` instantiates (reference ``train.py:89-94``); it
subclasses ``rl_games.common.vecenv.IVecEnv`` when rl_games is importable and is a plain class otherwise
(rl_games is not installable offline).  Multi-GPU: the reference picks ``cuda:{hvd.rank()}`` through Horovod
(:71-84); here the rank comes from ``torch.distributed`` / ``LOCAL_RANK`` (one process per GPU, NCCL).
"""
import os
from typing import Callable

from ..tasks import isaacgym_task_map

try:                                        # pragma: no cover - rl_games absent offline
    from rl_games.common import env_configurations, vecenv
    _IVecEnv = vecenv.IVecEnv
except Exception:                           # noqa: BLE001
    env_configurations = None
    _IVecEnv = object

#: stand-in for rl_games.common.env_configurations.configurations when rl_games is absent
configurations = {}


def register(name, config):
    if env_configurations is not None:
        env_configurations.register(name, config)
    configurations[name] = config


def get_rlgames_env_creator(task_config: dict, task_name: str, sim_device: str, rl_device: str,
                            graphics_device_id: int, headless: bool, multi_gpu: bool = False,
                            post_create_hook: Callable = None):
    def create_rlgpu_env(_sim_device=sim_device, _rl_device=rl_device, **kwargs):
        if multi_gpu:
            rank = int(os.environ.get("LOCAL_RANK", "0"))
            _sim_device = f"cuda:{rank}"
            _rl_device = f"cuda:{rank}"
            task_config["rank"] = rank
            task_config["rl_device"] = _rl_device
        else:
            _sim_device = sim_device
            _rl_device = rl_device
        env = isaacgym_task_map[task_name](cfg=task_config, sim_device=_sim_device,
                                           graphics_device_id=graphics_device_id, headless=headless)
        if post_create_hook is not None:
            post_create_hook()
        return env
    return create_rlgpu_env


class RLGPUEnv(_IVecEnv):
    def __init__(self, config_name, num_actors, **kwargs):
        table = env_configurations.configurations if env_configurations is not None else configurations
        self.env = table[config_name]["env_creator"](**kwargs)

    def step(self, action):
        return self.env.step(action)

    def reset(self):
        return self.env.reset()

    def get_number_of_agents(self):
        return getattr(self.env, "num_agents", 1)

    def get_env_info(self):
        info = {"action_space": self.env.action_space, "observation_space": self.env.observation_space}
        if self.env.num_states > 0:
            info["state_space"] = self.env.state_space
        return info
