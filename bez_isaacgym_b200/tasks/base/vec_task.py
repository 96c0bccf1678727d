"""``Env`` / ``VecTask`` with the reference's interface (``bez_isaacgym/tasks/base/vec_task.py:50-377``): same
constructor arguments, buffers (``obs_buf``, ``rew_buf``, int64 ``reset_buf`` / ``progress_buf`` / ``timeout_buf`` /
``randomize_buf``), ``step`` / ``reset`` return conventions and ``extras['time_outs']``.

What differs is underneath: ``step`` is two kernel launches (K0 before the simulator, the fused post-physics
kernel after it) instead of ~350 ATen ops, and there is no ``nonzero()`` -> ``len()`` host sync.  Domain
randomisation (vec_task.py:463-725) drives PhysX property setters and is out of scope; the two
noise hooks in ``step`` are kept as optional callables.
"""
import abc
import math
from typing import Any, Dict, Tuple

import numpy as np
import torch


class Box:
    """Minimal stand-in for ``gym.spaces.Box`` (gym is not installable offline); rl_games only reads
    ``shape``, ``low`` and ``high`` from the spaces in ``get_env_info``."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Env(abc.ABC):
    """Mirror of vec_task.py:50-145."""

    def __init__(self, config: Dict[str, Any], sim_device: str, graphics_device_id: int, headless: bool):
        split_device = sim_device.split(":")
        self.device_type = split_device[0]
        self.device_id = int(split_device[1]) if len(split_device) > 1 else 0

        self.device = "cpu"
        if config["sim"]["use_gpu_pipeline"]:
            if self.device_type.lower() in ("cuda", "gpu"):
                self.device = "cuda" + ":" + str(self.device_id)
            else:
                print("GPU Pipeline can only be used with GPU simulation. Forcing CPU Pipeline.")
                config["sim"]["use_gpu_pipeline"] = False

        self.rl_device = config.get("rl_device", "cuda:0")
        self.headless = headless
        enable_camera_sensors = config.get("enableCameraSensors", False)
        self.graphics_device_id = graphics_device_id
        if enable_camera_sensors is False and self.headless is True:
            self.graphics_device_id = -1

        self.num_environments = config["env"]["numEnvs"]
        self.num_agents = config["env"].get("numAgents", 1)
        self.num_observations = config["env"]["numObservations"]
        self.num_states = config["env"].get("numStates", 0)
        self.num_actions = config["env"]["numActions"]
        self.control_freq_inv = config["env"].get("controlFrequencyInv", 1)

        self.obs_space = Box(np.ones(self.num_obs) * -np.inf, np.ones(self.num_obs) * np.inf)
        self.state_space = Box(np.ones(self.num_states) * -np.inf, np.ones(self.num_states) * np.inf)
        self.act_space = Box(np.ones(self.num_actions) * -1.0, np.ones(self.num_actions) * 1.0)

        self.clip_obs = config["env"].get("clipObservations", np.inf)
        self.clip_actions = config["env"].get("clipActions", np.inf)

    @abc.abstractmethod
    def allocate_buffers(self):
        ...

    @abc.abstractmethod
    def step(self, actions: torch.Tensor) -> Tuple[Dict[str, torch.Tensor], torch.Tensor, torch.Tensor, Dict[str, Any]]:
        ...

    @abc.abstractmethod
    def reset(self) -> Dict[str, torch.Tensor]:
        ...

    @property
    def observation_space(self):
        return self.obs_space

    @property
    def action_space(self):
        return self.act_space

    @property
    def num_envs(self) -> int:
        return self.num_environments

    @property
    def num_acts(self) -> int:
        return self.num_actions

    @property
    def num_obs(self) -> int:
        return self.num_observations


class VecTask(Env):
    """Mirror of vec_task.py:148-377.  ``compute_device`` is where the kernels run: ``self.device`` when the GPU
    pipeline is on; with ``use_gpu_pipeline: False`` the simulator tensors live in (pinned) host memory and are
    staged to ``cuda:<device_id>`` every step -- the task math itself never runs on the CPU."""

    def __init__(self, config, sim_device, graphics_device_id, headless):
        super().__init__(config, sim_device, graphics_device_id, headless)
        self.cfg = config
        if self.cfg["physics_engine"] not in ("physx", "flex"):
            raise ValueError(f"Invalid physics engine backend: {self.cfg['physics_engine']}")
        self.physics_engine = self.cfg["physics_engine"]
        if self.cfg["sim"].get("up_axis", "z") not in ("z", "y"):
            raise ValueError(f"Invalid physics up-axis: {self.cfg['sim']['up_axis']}")
        self.compute_device = torch.device(self.device if self.device != "cpu" else f"cuda:{self.device_id}")
        self.host_staged = self.device == "cpu"
        self.dr_randomizations = {}
        self.first_randomization = True
        self.last_step = self.last_rand_step = -1
        self._dr_seed = int(config.get("seed", 42)) ^ 0x5DEECE66D
        self.sim_initialized = False
        self.create_sim()
        self.sim_initialized = True
        self.viewer = None
        self.enable_viewer_sync = True
        self.allocate_buffers()
        self.obs_dict = {}

    # buffers live on the compute device; in host-staged mode `step` returns host copies
    def allocate_buffers(self):
        dev = self.compute_device
        self.obs_buf = torch.zeros((self.num_envs, self.num_obs), device=dev, dtype=torch.float)
        self.states_buf = torch.zeros((self.num_envs, self.num_states), device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.float)
        self.reset_buf = torch.ones(self.num_envs, device=dev, dtype=torch.long)
        self.timeout_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.long)
        self.progress_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.long)
        self.randomize_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.long)
        self.extras = {}

    @abc.abstractmethod
    def create_sim(self):
        ...

    @abc.abstractmethod
    def pre_physics_step(self, actions: torch.Tensor):
        ...

    @abc.abstractmethod
    def post_physics_step(self):
        ...

    def get_state(self):
        return torch.clamp(self.states_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)

    def render(self):
        """Viewer / keyboard handling is out of scope (headless only)."""

    def _observations_out(self):
        """vec_task.py:343.  clip_obs = inf -> clamp is the identity; the kernel writes the clipped copy
        otherwise (``obs_clipped_buf``)."""
        return self.obs_buf

    def step(self, actions: torch.Tensor):
        if self.dr_randomizations.get("actions", None):
            actions = self.dr_randomizations["actions"]["noise_lambda"](actions)
        # action clamp (vec_task.py:317) is folded into the K0 kernel
        self.pre_physics_step(actions)
        for _ in range(self.control_freq_inv):
            self.render()
            self.sim.simulate()
        # timeout_buf (vec_task.py:331-332) is produced by the fused post-physics kernel from the
        # pre-increment progress_buf
        self.post_physics_step()
        self._apply_obs_randomization()
        self.extras["time_outs"] = self.timeout_buf.to(self.rl_device)
        self.obs_dict["obs"] = self._observations_out().to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict, self.rew_buf.to(self.rl_device), self.reset_buf.to(self.rl_device), self.extras

    def _apply_obs_randomization(self):
        """vec_task.py:338-339: the observation noise lambda after post_physics_step, then the clamp of :343."""
        if self.dr_randomizations.get("observations", None):
            # the step kernel clipped the noise-free rows; the reference clamps AFTER the noise (:343): the noise kernel refreshes
            # the clipped copy in the same pass
            self.obs_buf = self.dr_randomizations["observations"]["noise_lambda"](
                self.obs_buf, out_clipped=getattr(self, "obs_clipped_buf", None))

    # ------------------------------------------------------------------ domain randomisation: the tensor-path part
    def apply_randomizations(self, dr_params):
        """The "observations" / "actions" half of the reference's ``apply_randomizations`` (tasks/base/vec_task.py:505-618):
        builds ``dr_randomizations[name]`` = parameters + ``noise_lambda`` with the reference's frequency gate, schedule
        scaling ('linear' / 'constant') and additive / scaling operation for the gaussian and uniform distributions.  The
        lambdas are ONE kernel each (``bezk_dr_noise``: Philox white noise generated in registers, the persistent correlated
        draw read from memory).  The "sim_params" / "actor_params" half drives PhysX property setters and is out of scope:
        such keys raise."""
        from ... import ops
        for key in ("sim_params", "actor_params"):
            if dr_params.get(key):
                raise NotImplementedError(f"randomization_params.{key} drives PhysX property setters: out of scope (SURVEY 8)")
        rand_freq = dr_params.get("frequency", 1)
        self.last_step = int(getattr(self.sim, "frame", 0))           # gym.get_frame_count(sim)
        if self.first_randomization:
            do_nonenv_randomize = True
        else:
            do_nonenv_randomize = (self.last_step - self.last_rand_step) >= rand_freq
            rand_envs = (self.randomize_buf >= rand_freq) & (self.reset_buf != 0)
            self.randomize_buf[rand_envs] = 0
        if do_nonenv_randomize:
            self.last_rand_step = self.last_step
        for name in ("observations", "actions"):
            if name not in dr_params or not do_nonenv_randomize:
                continue
            p = dr_params[name]
            dist, op_type = p["distribution"], p["operation"]
            sched_type = p.get("schedule")
            sched_step = p.get("schedule_steps") if "schedule" in p else None
            if sched_type == "linear":
                s = 1.0 / sched_step * min(self.last_step, sched_step)
            elif sched_type == "constant":
                s = 0 if self.last_step < sched_step else 1
            else:
                s = 1
            x0, x1 = p["range"]
            c0, c1 = p.get("range_correlated", [0.0, 0.0])
            if dist == "gaussian":
                if op_type == "additive":
                    x0, x1, c0, c1 = x0 * s, x1 * s, c0 * s, c1 * s
                elif op_type == "scaling":
                    x1, x0 = x1 * s, x0 * s + 1.0 * (1.0 - s)
                    c1, c0 = c1 * s, c0 * s + 1.0 * (1.0 - s)
                params = {"mu": x0, "var": x1, "mu_corr": c0, "var_corr": c1}
                kcfg = ops.make_noise_cfg("gaussian", op_type, a=x1, b=x0, a_corr=c1, b_corr=c0)
            elif dist == "uniform":
                if op_type == "additive":
                    x0, x1, c0, c1 = x0 * s, x1 * s, c0 * s, c1 * s
                elif op_type == "scaling":
                    x0, x1 = x0 * s + 1.0 * (1.0 - s), x1 * s + 1.0 * (1.0 - s)
                    c0, c1 = c0 * s + 1.0 * (1.0 - s), c1 * s + 1.0 * (1.0 - s)
                params = {"lo": x0, "hi": x1, "lo_corr": c0, "hi_corr": c1}
                kcfg = ops.make_noise_cfg("uniform", op_type, a=x1 - x0, b=x0, a_corr=c1 - c0, b_corr=c0)
            else:
                raise ValueError(f"unknown distribution {dist!r}")
            if op_type not in ("additive", "scaling"):
                raise ValueError(f"unknown operation {op_type!r}")
            self._dr_period = getattr(self, "_dr_period", 0) + 1
            stream_id = (self._dr_period << 1) | (name == "actions")

            def noise_lambda(tensor, param_name=name, kcfg=kcfg, stream_id=stream_id, out_clipped=None):
                prm = self.dr_randomizations[param_name]
                if tensor.device != self.compute_device:
                    tensor = tensor.to(self.compute_device)
                tensor = tensor.contiguous()
                corr = prm.get("corr")
                if corr is None:                                      # drawn once per randomisation period, like upstream
                    corr = ops.dr_fill(self._dr_seed, stream_id << 32, torch.empty_like(tensor))
                    prm["corr"] = corr
                prm["calls"] = prm.get("calls", 0) + 1
                out = torch.empty_like(tensor) if param_name == "actions" else tensor      # obs_buf is replaced in place
                return ops.dr_noise(tensor, kcfg, corr=corr, seed=self._dr_seed, step=(stream_id << 32) + prm["calls"], out=out,
                                    out_clipped=out_clipped, clip=float(self.clip_obs) if out_clipped is not None else None)

            params["noise_lambda"] = noise_lambda
            self.dr_randomizations[name] = params
        self.first_randomization = False

    def zero_actions(self) -> torch.Tensor:
        return torch.zeros([self.num_envs, self.num_actions], dtype=torch.float32, device=self.rl_device)

    def reset(self):
        self.step(self.zero_actions())
        self.obs_dict["obs"] = self._observations_out().to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict
