"""Rollout math with rl_games' names (rl_games/common/a2c_common.py, v1.1.3): ``discount_values`` (GAE reverse
scan) and the ``prepare_dataset`` advantage normalisation.  The ``play_steps`` reward shaping / value bootstrap is not a
function here: it rides in the epilogue of the fused step kernel (``KickEnv.set_rollout_targets`` ->
``bezk_post_physics_rollout``)."""
import torch

from .. import dist as bdist
from .. import ops
from .experience import swap_and_flatten01  # noqa: F401  (rl_games keeps the helper in a2c_common; the ONE definition is experience's)


def discount_values(fdones, last_extrinsic_values, mb_fdones, mb_extrinsic_values, mb_rewards, gamma, tau,
                    out_advs=None, out_returns=None, return_returns=False):
    """``A2CBase.discount_values`` (+ ``mb_returns = mb_advs + mb_values``).  Shapes as in rl_games: ``fdones`` (N,),
    ``last_extrinsic_values`` (N,1), ``mb_fdones`` (T,N) float32 or uint8, values / rewards (T,N,1).  One kernel:
    one thread per env keeps the recurrence in registers over the horizon."""
    if mb_fdones.dtype not in (torch.uint8, torch.float32):
        mb_fdones = mb_fdones.float()
    if fdones.dtype != mb_fdones.dtype:
        fdones = fdones.to(mb_fdones.dtype)
    advs = out_advs if out_advs is not None else torch.empty_like(mb_rewards)
    rets = out_returns if out_returns is not None else torch.empty_like(mb_rewards)
    ops.gae(mb_rewards.contiguous(), mb_extrinsic_values.contiguous(), mb_fdones.contiguous(),
            last_extrinsic_values.contiguous(), fdones.contiguous(), gamma, tau, advs, rets)
    return (advs, rets) if return_returns else advs


class _AdvWorkspace:
    acc = None
    scratch = None


def normalize_advantages(returns, values, normalize=True, process_group=None, out=None):
    """``prepare_dataset``: ``advantages = returns - values`` then ``(adv - mean) / (std + 1e-8)`` (unbiased std).
    returns / values: (M,1) or (M,).  With ``process_group`` the moments are summed over ranks (global
    normalisation, as the north_star asks; the reference normalises per rank -> pass ``process_group=None``)."""
    r = returns.detach().reshape(-1).contiguous()
    v = values.detach().reshape(-1).contiguous()
    adv = out if out is not None else torch.empty_like(r)
    ws = _AdvWorkspace
    if ws.acc is None or ws.acc.device != r.device:
        ws.acc = torch.empty(3, dtype=torch.float64, device=r.device)
        ws.scratch = torch.empty(ops.rms_scratch_doubles(1), dtype=torch.float64, device=r.device)
    if not (normalize and process_group is not None and bdist.is_distributed(process_group)):
        return ops.adv_normalize_fused(r, v, adv, ws.scratch, normalize=normalize)      # one call, one kernel at rollout sizes
    acc = ops.adv_moments(r, v, ws.acc, ws.scratch)
    bdist.allreduce_sum_(acc, process_group)
    return ops.adv_normalize(r, v, acc, adv, normalize=True)
