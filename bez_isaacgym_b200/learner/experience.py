"""Rollout storage with rl_games' names (rl_games/common/experience.py ``ExperienceBuffer``,
rl_games/common/datasets.py ``PPODataset``, v1.1.3; driven by ``horizon_length`` / ``minibatch_size`` of
``cfg/train/bez_kickPPO.yaml:73-75``).

Storage is time-major exactly like rl_games' -- ``tensor_dict[name]`` is ``(horizon, num_actors, ...)`` -- so slot ``t`` of
every tensor is one contiguous block that the step kernels write IN PLACE (``slot()`` hands the fused post-physics
kernel its ``obs`` pointer and ``bezk_policy_head`` its actions / neglogpacs / values / mus / sigmas pointers;
``update_data`` remains for callers that hold the value elsewhere).

Two ways out of the buffer:

* ``get_transformed_list(swap_and_flatten01, names)`` + ``PPODataset`` -- rl_games' path, bit-identical rows in
  env-major order, one tiled transposition kernel per tensor (``bezk_swap_and_flatten01``);
* ``SlabDataset`` -- no flattening pass at all: minibatch ``i`` is the dict of VIEWS ``tensor[:, e0:e0+E]`` that the
  ``*_slabs`` kernels read in place (same sample set as rl_games' minibatch ``i``; rows ordered time-major inside
  the batch, see ``include/bezk.h``).
"""
import torch

from .. import ops

_NAMES_F32 = ("obses", "rewards", "values", "neglogpacs", "actions", "mus", "sigmas")


def swap_and_flatten01(arr: torch.Tensor) -> torch.Tensor:
    """(T, N, ...) -> (N*T, ...) env-major, as rl_games' helper of the same name: one transposition kernel."""
    return ops.swap_and_flatten01(arr.contiguous())


class ExperienceBuffer:
    def __init__(self, env_info, algo_info, device, aux_tensor_dict=None):
        self.env_info, self.algo_info, self.device = env_info, algo_info, torch.device(device)
        self.num_actors = int(algo_info["num_actors"])
        self.horizon_length = int(algo_info["horizon_length"])
        if algo_info.get("has_central_value") or algo_info.get("use_action_masks"):
            raise NotImplementedError("central value / action masks are not part of the BezKick configuration")
        obs_dim = int(env_info["observation_space"].shape[0])
        act_dim = int(env_info["action_space"].shape[0])
        base = (self.horizon_length, self.num_actors)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.tensor_dict = {
            "obses": torch.zeros(base + (obs_dim,), **f32),
            "rewards": torch.zeros(base + (1,), **f32),
            "values": torch.zeros(base + (1,), **f32),
            "neglogpacs": torch.zeros(base, **f32),
            "dones": torch.zeros(base, dtype=torch.uint8, device=self.device),
            "actions": torch.zeros(base + (act_dim,), **f32),
            "mus": torch.zeros(base + (act_dim,), **f32),
            "sigmas": torch.zeros(base + (act_dim,), **f32),
        }
        if aux_tensor_dict:
            for k, shape in aux_tensor_dict.items():
                self.tensor_dict[k] = torch.zeros(base + tuple(shape), **f32)

    # ---- rl_games API
    def update_data(self, name, index, val):
        self.tensor_dict[name][index, :] = val

    def update_data_rnn(self, name, indices, play_mask, val):
        raise NotImplementedError("the BezKick policy is not recurrent")

    def get_transformed(self, transform_op):
        return {k: transform_op(v) for k, v in self.tensor_dict.items()}

    def get_transformed_list(self, transform_op, tensor_list):
        return {k: transform_op(self.tensor_dict[k]) for k in tensor_list if self.tensor_dict.get(k) is not None}

    # ---- in-place access for the kernels
    def slot(self, name, index):
        """Contiguous view of slot ``index`` of ``name``: hand it to a kernel as its output instead of calling update_data."""
        return self.tensor_dict[name][index]


class PPODataset:
    """rl_games/common/datasets.py: minibatch ``idx`` = rows ``[idx*mb, (idx+1)*mb)`` of every flattened tensor."""

    def __init__(self, batch_size, minibatch_size, is_discrete=False, is_rnn=False, device="cuda", seq_len=1):
        if is_rnn or is_discrete:
            raise NotImplementedError("the BezKick policy is continuous and not recurrent")
        if batch_size % minibatch_size:
            raise ValueError("batch_size must be a multiple of minibatch_size")   # rl_games asserts the same
        self.batch_size, self.minibatch_size, self.device = batch_size, minibatch_size, device
        self.length = batch_size // minibatch_size
        self.values_dict = None
        self.last_range = (0, 0)

    def update_values_dict(self, values_dict):
        self.values_dict = values_dict

    def update_mu_sigma(self, mu, sigma):
        start, end = self.last_range
        self.values_dict["mu"][start:end] = mu
        self.values_dict["sigma"][start:end] = sigma

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        start, end = idx * self.minibatch_size, (idx + 1) * self.minibatch_size
        self.last_range = (start, end)
        return {k: (v[start:end] if v is not None else None) for k, v in self.values_dict.items()}


class SlabDataset:
    """Minibatches as slab views of the time-major rollout: ``ds[i][name]`` is ``tensor[:, e0:e0+E]`` with
    ``E = minibatch_size // horizon`` and ``e0 = i * E`` -- the sample set of rl_games' minibatch ``i`` (which is rows
    ``[i*mb, (i+1)*mb)`` of the env-major flattening, i.e. exactly envs ``[e0, e0+E)`` at every timestep), with no copy.
    ``extra`` holds further time-major (T, N, ...) tensors produced after the rollout (returns, advantages, ...)."""

    def __init__(self, experience: ExperienceBuffer, minibatch_size: int, extra=None):
        T, N = experience.horizon_length, experience.num_actors
        if minibatch_size % T or (N * T) % minibatch_size:
            raise ValueError("minibatch_size must be a multiple of horizon_length and divide num_actors * horizon_length")
        self.horizon, self.num_actors = T, N
        self.envs_per_batch = minibatch_size // T
        self.minibatch_size = minibatch_size
        self.length = (N * T) // minibatch_size
        self.tensors = dict(experience.tensor_dict)
        if extra:
            for k, v in extra.items():
                if v.shape[0] != T or v.shape[1] != N or not v.is_contiguous():
                    raise ValueError(f"{k} must be a contiguous (horizon, num_actors, ...) tensor")
                self.tensors[k] = v

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        if not 0 <= idx < self.length:
            raise IndexError(idx)
        e0 = idx * self.envs_per_batch
        return {k: v[:, e0:e0 + self.envs_per_batch] for k, v in self.tensors.items()}

    def permutation_to_env_major(self, device=None):
        """Index tensor p with ``env_major_rows = slab_rows[p]``: row ``(e-e0)*T + t`` of rl_games' minibatch is row
        ``t*E + (e-e0)`` here.  For checkers and for callers that need per-sample outputs in rl_games' order."""
        T, E = self.horizon, self.envs_per_batch
        e = torch.arange(E, device=device).repeat_interleave(T)
        t = torch.arange(T, device=device).repeat(E)
        return t * E + e
