"""Fused PPO loss (rl_games/algos_torch/a2c_continuous.py ``calc_gradients`` loss block + common_losses +
``bound_loss`` + ``policy_kl`` + ``ModelA2CContinuousLogStd.neglogp``) as ONE kernel pass that produces the loss
terms AND their gradients; exposed as a ``torch.autograd.Function`` so ``loss.backward()`` flows into the policy
MLP (the GEMMs stay in torch)."""
from dataclasses import dataclass

import torch

from .. import ops


@dataclass
class PPOLossConfig:
    e_clip: float = 0.2
    critic_coef: float = 2.0
    entropy_coef: float = 0.0
    bounds_loss_coef: float = 0.001
    soft_bound: float = 1.1
    clip_value: bool = True
    bound_form: str = "v1.1.3"          # the pinned release as recalled; "outside" = later releases (see oracle)


class _PPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, values, logstd, actions, old_mu, old_sigma, old_values, returns, old_neglogp, advantages, cfg):
        m = mu.shape[0]
        dev = mu.device
        f = lambda t, shape: t.detach().float().reshape(shape).contiguous()
        slabs = not actions.is_contiguous()          # rollout-side tensors given as slab views of time-major storage
        kcfg = ops.make_ppo_cfg(cfg.e_clip, cfg.critic_coef, cfg.entropy_coef, cfg.bounds_loss_coef, cfg.soft_bound,
                                cfg.clip_value, cfg.bound_form)
        stats = torch.empty(8, dtype=torch.float64, device=dev)
        partials = _scratch(dev)
        g_mu = torch.empty(m, 18, dtype=torch.float32, device=dev)
        g_v = torch.empty(m, dtype=torch.float32, device=dev)
        g_ls = torch.empty(18, dtype=torch.float32, device=dev)
        nlp = torch.empty(m, dtype=torch.float32, device=dev)
        if slabs:
            ops.ppo_loss_slabs(actions.detach(), f(mu, (m, 18)), f(logstd, (18,)), old_mu.detach(), old_sigma.detach(),
                               f(values, (m,)), old_values.detach(), returns.detach(), old_neglogp.detach(),
                               advantages.detach(), kcfg, stats, partials, grad_mu=g_mu, grad_values=g_v, grad_logstd=g_ls,
                               neglogp_out=nlp)
        else:
            ops.ppo_loss(f(actions, (m, 18)), f(mu, (m, 18)), f(logstd, (18,)), f(old_mu, (m, 18)), f(old_sigma, (m, 18)),
                         f(values, (m,)), f(old_values, (m,)), f(returns, (m,)), f(old_neglogp, (m,)), f(advantages, (m,)),
                         kcfg, stats, partials, grad_mu=g_mu, grad_values=g_v, grad_logstd=g_ls, neglogp_out=nlp)
        ctx.save_for_backward(g_mu, g_v, g_ls)
        ctx.shapes = (mu.shape, values.shape, logstd.shape, mu.dtype, values.dtype, logstd.dtype)
        ctx.mark_non_differentiable(stats, nlp)
        return stats[0].float(), stats, nlp

    @staticmethod
    def backward(ctx, grad_loss, _gs, _gn):
        g_mu, g_v, g_ls = ctx.saved_tensors
        mu_shape, v_shape, ls_shape, mu_dt, v_dt, ls_dt = ctx.shapes
        return ((g_mu * grad_loss).to(mu_dt).view(mu_shape), (g_v * grad_loss).to(v_dt).view(v_shape),
                (g_ls * grad_loss).to(ls_dt).view(ls_shape), None, None, None, None, None, None, None, None)


_SCRATCH = {}


def _scratch(device):
    key = str(device)
    if key not in _SCRATCH:
        _SCRATCH[key] = torch.empty(ops.ppo_scratch_doubles(), dtype=torch.float64, device=device)
    return _SCRATCH[key]


def ppo_loss(mu, values, logstd, actions, old_mu, old_sigma, old_values, returns, old_neglogp, advantages,
             cfg: PPOLossConfig = PPOLossConfig()):
    """``actions, old_mu, old_sigma, old_values, returns, old_neglogp, advantages`` may be contiguous minibatch tensors or
    slab views ``storage[:, e0:e0+E]`` of time-major rollout tensors (``SlabDataset``); ``mu`` / ``values`` are the network
    outputs for the same batch rows.

    Returns ``(loss, info)``; ``loss`` is differentiable w.r.t. ``mu`` (M,18), ``values`` (M,1)/(M,) and ``logstd``
    (18,).  ``info``: a_loss, c_loss, entropy, b_loss, kl, clip_frac (fp64 scalars on device, no host sync) and
    ``neglogp`` (M,)."""
    loss, stats, nlp = _PPOLoss.apply(mu, values, logstd, actions, old_mu, old_sigma, old_values, returns, old_neglogp,
                                      advantages, cfg)
    info = dict(a_loss=stats[1], c_loss=stats[2], entropy=stats[3], b_loss=stats[4], kl=stats[5], clip_frac=stats[6],
                neglogp=nlp)
    return loss, info
