"""``policy_head`` -- what rl_games does between the policy MLP and ``env.step`` inside ``play_steps``
(rl_games/algos_torch/models.py ``ModelA2CContinuousLogStd.forward`` with ``is_train=False``,
rl_games/common/a2c_common.py ``get_action_values`` / ``play_steps`` / ``preprocess_actions``), as ONE kernel:

    sigma = exp(logstd);  actions = Normal(mu, sigma).sample();  neglogpacs = neglogp(actions, mu, sigma, logstd)
    values = value_mean_std(value, unnorm=True)
    experience_buffer.update_data('actions' | 'neglogpacs' | 'values' | 'mus' | 'sigmas', t, ...)      (written in place)
    env actions = rescale_actions(-1, 1, clamp(actions, -1, 1))  ->  KickEnv.pre_physics_step's PD targets

Documented deviation: the N(0,1) draws come from Philox4x32-10 keyed (seed, step, GLOBAL env id) + Box-Muller instead of
torch's global generator; the global id is ``env_base + row`` (``env_base`` = the rank's shard offset, taken from
``env.env_base`` when an env is given), which is what makes the noise invariant to sharding: ranks that share a seed draw
different noise and their union equals the un-sharded draw.  Pass ``noise=`` to supply the draws."""
import torch

from .. import ops


def policy_head(mu, logstd, value, value_mean_std=None, experience=None, t=None, noise=None, seed=0, step=0, env=None,
                env_base=None):
    """mu (N,18), logstd (18,), value (N,1)/(N,) normalised critic output.  ``experience``/``t``: ExperienceBuffer slot to
    write in place (otherwise fresh tensors).  ``env``: a KickEnv -- its PD ``targets`` are produced by the same launch
    (the caller then skips ``pre_physics_step``'s K0).  Returns rl_games' ``res_dict`` (+ ``env_actions``)."""
    n, dev = mu.shape[0], mu.device
    f32 = dict(dtype=torch.float32, device=dev)
    if experience is not None:
        out = {k: experience.slot(k, t) for k in ("actions", "neglogpacs", "values", "mus", "sigmas")}
    else:
        out = dict(actions=torch.empty(n, 18, **f32), neglogpacs=torch.empty(n, **f32), values=torch.empty(n, 1, **f32),
                   mus=torch.empty(n, 18, **f32), sigmas=torch.empty(n, 18, **f32))
    env_actions = torch.empty(n, 18, **f32)
    mean = var = None
    eps = 1e-5
    if value_mean_std is not None:
        mean, var, eps = value_mean_std.running_mean.view(1), value_mean_std.running_var.view(1), value_mean_std.epsilon
    ops.policy_head(mu.detach().contiguous(), logstd.detach().contiguous(), value.detach().reshape(-1).contiguous(), mean, var, eps,
                    noise=noise, seed=seed, step=step, actions=out["actions"], neglogp=out["neglogpacs"],
                    values=out["values"].view(-1), mus=out["mus"], sigmas=out["sigmas"],
                    task_cfg=env._kcfg if env is not None else None, env_actions=env_actions,
                    targets=env.targets if env is not None else None,
                    env_base=(getattr(env, "env_base", 0) if env_base is None else env_base))
    res = dict(out)
    res["env_actions"] = env_actions
    return res
