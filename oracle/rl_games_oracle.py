"""torch-CPU restatement of the rl_games==1.1.3 learner math BezKick feeds (oracle; TEST INFRASTRUCTURE ONLY).

rl_games is a third-party dependency that is NOT vendored under /root/reference (pinned by
``setup.py:22`` ``rl-games==1.1.3`` / ``README.md:10``); its published algorithm is restated here from
``rl_games/algos_torch/running_mean_std.py``, ``rl_games/common/a2c_common.py`` (``discount_values``,
``play_steps``, ``prepare_dataset``), ``rl_games/algos_torch/a2c_continuous.py`` (``calc_gradients``,
``bound_loss``), ``rl_games/common/common_losses.py``, ``rl_games/algos_torch/torch_ext.py``
(``policy_kl``) and ``rl_games/algos_torch/models.py`` (``ModelA2CContinuousLogStd.neglogp``) at tag v1.1.3
(SURVEY.md App. C).

PARITY UNPINNED: the reference holds no test or golden vector for this half.  What *is* pinned are
the reference's own call sites and hyper-parameters (``cfg/train/bez_kickPPO.yaml:45-79``) and the
shipped checkpoint's identities (``results/Bez_Kick/Normal/Bez_Kick_33.pth``: fp64 ``running_mean``/
``running_var``/``count`` buffers, ``count_obs = 1 + 5*frame``, ``count_val = 1 + 2*frame``), which
``tests/test_rl_games_oracle.py`` (``test_checkpoint_state_layout``, ``test_checkpoint_update_cadence_identities``) checks against
this restatement's update cadence.  Because this half is unpinned, the CUDA learner kernels are ALSO checked without it:
``tests/test_learner_properties_gpu.py`` (closed-form GAE, two-pass fp64 moments, fp64 autograd of the published loss formulas).
"""
import math
from typing import Dict, Tuple

import torch
from torch import Tensor


class RunningMeanStd:
    """rl_games/algos_torch/running_mean_std.py (C.1): fp64 running stats, batch moments merged by the
    parallel-variance formula; normalise + clip to +-5."""

    def __init__(self, insize, epsilon=1e-05):
        self.insize = insize
        self.epsilon = epsilon
        shape = (insize,) if isinstance(insize, int) else tuple(insize)
        self.running_mean = torch.zeros(shape, dtype=torch.float64)
        self.running_var = torch.ones(shape, dtype=torch.float64)
        self.count = torch.ones((), dtype=torch.float64)
        self.training = True

    @staticmethod
    def merge(mean, var, count, batch_mean, batch_var, batch_count):
        delta = batch_mean - mean
        tot = count + batch_count
        new_mean = mean + delta * batch_count / tot
        m2 = var * count + batch_var * batch_count + delta ** 2 * count * batch_count / tot
        return new_mean, m2 / tot, tot

    def __call__(self, x: Tensor, unnorm: bool = False) -> Tensor:
        if self.training:
            mean = x.mean(0)
            var = x.var(0)          # unbiased
            self.running_mean, self.running_var, self.count = self.merge(
                self.running_mean, self.running_var, self.count, mean, var, x.size(0))
        cur_mean, cur_var = self.running_mean, self.running_var
        if unnorm:
            y = torch.clamp(x, min=-5.0, max=5.0)
            return torch.sqrt(cur_var.float() + self.epsilon) * y + cur_mean.float()
        y = (x - cur_mean.float()) / torch.sqrt(cur_var.float() + self.epsilon)
        return torch.clamp(y, min=-5.0, max=5.0)

    def state_dict(self) -> Dict[str, Tensor]:
        return {"running_mean": self.running_mean, "running_var": self.running_var, "count": self.count}


def shape_rewards(rewards: Tensor, values: Tensor, time_outs: Tensor, gamma: float,
                  scale: float = 0.01, shift: float = 0.0) -> Tensor:
    """play_steps reward path (C.3): DefaultRewardsShaper then value bootstrap on time-outs.
    ``rewards`` (N,), ``values`` (N,1) un-normalised, ``time_outs`` (N,) int64 -> (N,1)."""
    shaped = (rewards.unsqueeze(1) + shift) * scale
    return shaped + gamma * values * time_outs.unsqueeze(1).float()


def discount_values(fdones: Tensor, last_extrinsic_values: Tensor, mb_fdones: Tensor,
                    mb_extrinsic_values: Tensor, mb_rewards: Tensor, gamma: float, tau: float) -> Tensor:
    """A2CBase.discount_values (C.2): GAE(gamma, tau) reverse scan.  Shapes: fdones (N,), last values
    (N,1), mb_fdones (T,N), mb values / rewards (T,N,1)."""
    horizon = mb_rewards.shape[0]
    lastgaelam = 0
    mb_advs = torch.zeros_like(mb_rewards)
    for t in reversed(range(horizon)):
        if t == horizon - 1:
            nextnonterminal = 1.0 - fdones
            nextvalues = last_extrinsic_values
        else:
            nextnonterminal = 1.0 - mb_fdones[t + 1]
            nextvalues = mb_extrinsic_values[t + 1]
        nextnonterminal = nextnonterminal.unsqueeze(1)
        delta = mb_rewards[t] + gamma * nextvalues * nextnonterminal - mb_extrinsic_values[t]
        mb_advs[t] = lastgaelam = delta + gamma * tau * nextnonterminal * lastgaelam
    return mb_advs


def swap_and_flatten01(arr: Tensor) -> Tensor:
    """(T, N, ...) -> (N*T, ...), env-major (index env*T + t)."""
    s = arr.size()
    return arr.transpose(0, 1).reshape(s[0] * s[1], *s[2:])


def prepare_dataset(returns: Tensor, values: Tensor, value_mean_std: RunningMeanStd,
                    normalize_advantage: bool = True) -> Tuple[Tensor, Tensor, Tensor]:
    """prepare_dataset core (C.4) on flattened (M,1) tensors: returns (advantages (M,), values, returns)."""
    advantages = returns - values
    values = value_mean_std(values)
    returns = value_mean_std(returns)
    advantages = torch.sum(advantages, axis=1)
    if normalize_advantage:
        advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return advantages, values, returns


def neglogp(x: Tensor, mean: Tensor, std: Tensor, logstd: Tensor) -> Tensor:
    """ModelA2CContinuousLogStd.neglogp."""
    return (0.5 * (((x - mean) / std) ** 2).sum(dim=-1)
            + 0.5 * math.log(2.0 * math.pi) * x.size()[-1]
            + logstd.sum(dim=-1))


def actor_loss(old_neglogp, new_neglogp, advantage, e_clip: float) -> Tensor:
    """common_losses.actor_loss (ppo=True)."""
    ratio = torch.exp(old_neglogp - new_neglogp)
    surr1 = advantage * ratio
    surr2 = advantage * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip)
    return torch.max(-surr1, -surr2)


def critic_loss(value_preds_batch, values, e_clip: float, return_batch, clip_value: bool = True) -> Tensor:
    """common_losses.critic_loss."""
    if clip_value:
        clipped = value_preds_batch + (values - value_preds_batch).clamp(-e_clip, e_clip)
        return torch.max((values - return_batch) ** 2, (clipped - return_batch) ** 2)
    return (return_batch - values) ** 2


def bound_loss(mu: Tensor, soft_bound: float = 1.1, form: str = "v1.1.3") -> Tensor:
    """A2CAgent.bound_loss.  ``form="v1.1.3"``: the pinned release as recalled --
    ``clamp_max(mu - b, 0)^2 + clamp_max(-mu + b, 0)^2`` (penalises the |mu| < b side; fixed upstream
    later).  ``form="outside"``: the later releases' ``clamp_min(mu - b, 0)^2 + clamp_max(mu + b, 0)^2``.
    Un-verifiable offline, so both exist on both sides of the parity test."""
    if form == "v1.1.3":
        mu_loss_high = torch.clamp_max(mu - soft_bound, 0.0) ** 2
        mu_loss_low = torch.clamp_max(-mu + soft_bound, 0.0) ** 2
    elif form == "outside":
        mu_loss_high = torch.clamp_min(mu - soft_bound, 0.0) ** 2
        mu_loss_low = torch.clamp_max(mu + soft_bound, 0.0) ** 2
    else:
        raise ValueError(form)
    return (mu_loss_low + mu_loss_high).sum(axis=-1)


def entropy(logstd_rows: Tensor) -> Tensor:
    """Normal(mu, sigma).entropy().sum(-1) = sum(0.5 + 0.5*log(2*pi) + log sigma)."""
    return (0.5 + 0.5 * math.log(2.0 * math.pi) + logstd_rows).sum(dim=-1)


def policy_kl(p0_mu, p0_sigma, p1_mu, p1_sigma) -> Tensor:
    """torch_ext.policy_kl(reduce=True): mean over the batch of the summed per-dimension KL."""
    c1 = torch.log(p1_sigma / p0_sigma + 1e-5)
    c2 = (p0_sigma ** 2 + (p1_mu - p0_mu) ** 2) / (2.0 * (p1_sigma ** 2 + 1e-5))
    c3 = -1.0 / 2.0
    kl = (c1 + c2 + c3).sum(dim=-1)
    return kl.mean()


def ppo_loss(mb: Dict[str, Tensor], e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=0.001,
             clip_value=True, bound_form="v1.1.3") -> Dict[str, Tensor]:
    """calc_gradients loss block (C.5), no rnn masks.  ``mb`` holds mu (M,18) and values (M,1) (model
    outputs, may require grad), logstd (18,), and the stored old_* / actions / returns / advantages."""
    mu, logstd = mb["mu"], mb["logstd"]
    sigma = torch.exp(mu * 0.0 + logstd)
    logstd_rows = mu * 0.0 + logstd
    new_neglogp = neglogp(mb["actions"], mu, sigma, logstd_rows)
    a_loss = actor_loss(mb["old_neglogp"], new_neglogp, mb["advantages"], e_clip)
    c_loss = critic_loss(mb["old_values"], mb["values"], e_clip, mb["returns"], clip_value)
    ent = entropy(logstd_rows)
    b_loss = bound_loss(mu, form=bound_form)
    a_m, c_m, e_m, b_m = a_loss.mean(), c_loss.mean(), ent.mean(), b_loss.mean()
    loss = a_m + 0.5 * c_m * critic_coef - e_m * entropy_coef + b_m * bounds_loss_coef
    with torch.no_grad():
        kl = policy_kl(mu.detach(), sigma.detach(), mb["old_mu"], mb["old_sigma"])
    return dict(loss=loss, a_loss=a_m, c_loss=c_m, entropy=e_m, b_loss=b_m, kl=kl, neglogp=new_neglogp)


def adaptive_lr(lr: float, kl: float, kl_threshold: float = 0.008, min_lr=1e-6, max_lr=1e-2) -> float:
    """schedulers.AdaptiveScheduler.update."""
    if kl > 2.0 * kl_threshold:
        lr = max(lr / 1.5, min_lr)
    if kl < 0.5 * kl_threshold:
        lr = min(lr * 1.5, max_lr)
    return lr


# ------------------------------------------------------------------------------------------------
# Rollout storage either side of the task step.  Restated from rl_games v1.1.3
# rl_games/common/experience.py (ExperienceBuffer), rl_games/common/datasets.py (PPODataset),
# rl_games/algos_torch/models.py (ModelA2CContinuousLogStd.forward), rl_games/common/a2c_common.py
# (get_action_values, play_steps, preprocess_actions) and rl_games/algos_torch/torch_ext.py (rescale_actions;
# the reference's in-tree fork has the same function at utils/players.py:11-15).  PARITY UNPINNED.
# ------------------------------------------------------------------------------------------------
class ExperienceBuffer:
    """rl_games/common/experience.py for the continuous-action, non-RNN case BezKick uses: time-major
    ``tensor_dict[name]`` of shape (horizon, num_actors, ...); ``update_data(name, index, val)`` writes slot
    ``index``; ``get_transformed_list(fn, names)`` applies ``fn`` (swap_and_flatten01) to the named tensors."""

    def __init__(self, num_actors: int, horizon: int, obs_dim: int = 54, act_dim: int = 18, device="cpu"):
        self.num_actors, self.horizon_length = num_actors, horizon
        base = (horizon, num_actors)
        f32 = dict(dtype=torch.float32, device=device)
        self.tensor_dict = {
            "obses": torch.zeros(base + (obs_dim,), **f32),
            "rewards": torch.zeros(base + (1,), **f32),
            "values": torch.zeros(base + (1,), **f32),
            "neglogpacs": torch.zeros(base, **f32),
            "dones": torch.zeros(base, dtype=torch.uint8, device=device),
            "actions": torch.zeros(base + (act_dim,), **f32),
            "mus": torch.zeros(base + (act_dim,), **f32),
            "sigmas": torch.zeros(base + (act_dim,), **f32),
        }

    def update_data(self, name: str, index: int, val: Tensor):
        self.tensor_dict[name][index, :] = val

    def get_transformed_list(self, transform_op, tensor_list):
        return {k: transform_op(self.tensor_dict[k]) for k in tensor_list if self.tensor_dict.get(k) is not None}


class PPODataset:
    """rl_games/common/datasets.py: minibatch ``idx`` = rows [idx*mb, (idx+1)*mb) of every flattened tensor
    (no shuffling for the non-RNN path)."""

    def __init__(self, batch_size: int, minibatch_size: int):
        assert batch_size % minibatch_size == 0
        self.batch_size, self.minibatch_size = batch_size, minibatch_size
        self.length = batch_size // minibatch_size
        self.values_dict = None

    def update_values_dict(self, values_dict):
        self.values_dict = values_dict

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        start, end = idx * self.minibatch_size, (idx + 1) * self.minibatch_size
        return {k: (v[start:end] if v is not None else None) for k, v in self.values_dict.items()}


def rescale_actions(low, high, action):
    """torch_ext.rescale_actions (= utils/players.py:11-15 of the reference's fork)."""
    d = (high - low) / 2.0
    m = (high + low) / 2.0
    return action * d + m


def preprocess_actions(actions: Tensor, low: float = -1.0, high: float = 1.0) -> Tensor:
    """A2CBase.preprocess_actions with clip_actions=True: clamp to [-1, 1] then rescale to the action space."""
    return rescale_actions(low, high, torch.clamp(actions, -1.0, 1.0))


def policy_head(mu: Tensor, logstd: Tensor, value_norm: Tensor, value_rms: RunningMeanStd, noise: Tensor):
    """ModelA2CContinuousLogStd.forward (is_train=False) + get_action_values' value un-normalisation, with the
    Normal sample written as ``mu + sigma * noise`` (``noise`` ~ N(0,1) supplied by the caller so that both sides
    of a parity test see the same draws).  mu (N,18), logstd (18,), value_norm (N,1)."""
    logstd_rows = mu * 0.0 + logstd
    sigma = torch.exp(logstd_rows)
    actions = mu + sigma * noise
    nlp = neglogp(actions, mu, sigma, logstd_rows)
    was_training = value_rms.training
    value_rms.training = False
    values = value_rms(value_norm, unnorm=True)
    value_rms.training = was_training
    return dict(actions=actions, neglogpacs=nlp, values=values, mus=mu, sigmas=sigma)


# ------------------------------------------------------------------------------------------------
# Domain-randomisation noise lambdas.  NOT rl_games: restated from the reference's own
# bez_isaacgym/tasks/base/vec_task.py:562-618 (in tree), kept here with the other elementwise rollout helpers.
# ------------------------------------------------------------------------------------------------
def dr_noise_lambda(tensor: Tensor, corr: Tensor, white: Tensor, distribution: str, operation: str, p0: float, p1: float,
                    c0: float = 0.0, c1: float = 0.0) -> Tensor:
    """``noise_lambda`` of vec_task.py:586-593 (gaussian: p0 = mu, p1 = var, c0 = mu_corr, c1 = var_corr) and :609-616
    (uniform: p0 = lo, p1 = hi, c0 = lo_corr, c1 = hi_corr) with the two random draws passed in (``corr`` ~ N(0,1) persistent,
    ``white`` ~ N(0,1) resp. U[0,1) fresh)."""
    import operator
    op = operator.add if operation == "additive" else operator.mul
    if distribution == "gaussian":
        corr = corr * c1 + c0
        return op(tensor, corr + white * p1 + p0)
    corr = corr * (c1 - c0) + c0
    return op(tensor, corr + white * (p1 - p0) + p0)
