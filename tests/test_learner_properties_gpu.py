"""Property tests of the learner kernels that do NOT go through ``oracle/rl_games_oracle.py`` (the learner half
of the oracle is a restatement of un-vendored rl_games, "parity unpinned"; these properties pin the kernels to the published
definitions independently):

* GAE against the closed-form discounted sum  A_t = sum_{k>=t} (gamma*tau)^(k-t) * prod_{j=t+1..k} (1 - d_j) * delta_k,
  delta_k = r_k + gamma * V_{k+1} * (1 - d_{k+1}) - V_k, evaluated in fp64 with explicit loops (Schulman et al. 2016, eq. 16);
* RunningMeanStd after k train-mode batches against two-pass fp64 moments of the batches: the pooled mean is exactly
  sum(x) / (1 + N) (the (0, 1, 1) start is one pseudo-sample at 0) and the pooled variance follows from the per-batch two-pass
  sums of squares through the parallel-variance identity with rl_games' unbiased batch variance;
* the fused PPO loss backward against fp64 autograd of the loss written out here from the published formulas (clipped surrogate,
  clipped value loss, bound loss, Normal neglogp), and a central finite difference of the kernel's own forward.
"""
import math

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

pytestmark = pytest.mark.gpu

SET = dict(max_examples=20, deadline=None, derandomize=True)


def _ops():
    from bez_isaacgym_b200 import ops
    return ops


# ----------------------------------------------------------------------------------------------- GAE
@settings(**SET)
@given(n=st.integers(1, 300), T=st.integers(1, 40), gamma=st.floats(0.8, 0.999), tau=st.floats(0.0, 1.0),
       p_done=st.floats(0.0, 0.5), seed=st.integers(0, 2 ** 20))
def test_gae_equals_closed_form_discounted_sum(n, T, gamma, tau, p_done, seed):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    r = torch.randn(T, n, 1, generator=g) * 0.5
    v = torch.randn(T, n, 1, generator=g)
    d = (torch.rand(T, n, generator=g) < p_done).to(torch.uint8)
    lv = torch.randn(n, 1, generator=g)
    ld = (torch.rand(n, generator=g) < p_done).to(torch.uint8)
    advs = torch.empty(T, n, 1, device="cuda"); rets = torch.empty_like(advs)
    ops.gae(r.cuda(), v.cuda(), d.cuda(), lv.cuda(), ld.cuda(), gamma, tau, advs, rets)
    # closed form in fp64, explicit loops (no recurrence): uses the fp32-rounded gamma and gamma*tau the kernel is specified with
    g32, gt32 = float(np.float32(gamma)), float(np.float32(gamma * tau))
    R, V = r[..., 0].double().numpy(), v[..., 0].double().numpy()
    D = np.concatenate([d.double().numpy(), ld.double().numpy()[None]], 0)           # D[t] = done flag entering step t; D[T] = final
    Vn = np.concatenate([V, lv[:, 0].double().numpy()[None]], 0)
    delta = R + g32 * Vn[1:] * (1.0 - D[1:]) - V
    want = np.zeros((T, n))
    for t in range(T):
        w = np.ones(n)
        for k in range(t, T):
            if k > t:
                w = w * gt32 * (1.0 - D[k])
            want[t] += w * delta[k]
    got = advs[..., 0].cpu().double().numpy()
    scale = np.abs(V).max() + np.abs(R).max() + 1.0
    assert np.allclose(got, want, rtol=1e-5, atol=2e-5 * scale * min(T, 1.0 / max(1e-3, 1.0 - gt32)))
    assert torch.equal(rets.cpu(), advs.cpu() + v)


# ----------------------------------------------------------------------------------------------- RunningMeanStd
@settings(**SET)
@given(k=st.integers(1, 5), c=st.sampled_from([1, 2, 7, 54]), seed=st.integers(0, 2 ** 20),
       sizes=st.lists(st.integers(2, 3000), min_size=5, max_size=5), shift=st.floats(-50, 50), spread=st.floats(0.01, 30))
def test_running_mean_std_equals_two_pass_fp64_moments(k, c, seed, sizes, shift, spread):
    from bez_isaacgym_b200.learner import RunningMeanStd
    g = torch.Generator().manual_seed(seed)
    rms = RunningMeanStd(c).cuda()
    rms.train()
    mean, m2, count = np.zeros(c), np.ones(c), 1.0                     # running_var * count with the (0, 1, 1) start
    total_sum, total_n = np.zeros(c), 0
    for i in range(k):
        x = torch.randn(sizes[i], c, generator=g) * spread + shift + i
        rms(x.cuda())
        xd = x.double().numpy()
        nb = xd.shape[0]
        bm = xd.mean(0)
        ss = ((xd - bm) ** 2).sum(0)                                   # two-pass sum of squares of the batch
        bvar_unbiased = ss / (nb - 1)                                  # rl_games feeds torch.var (unbiased) into the merge
        delta = bm - mean
        tot = count + nb
        m2 = m2 + bvar_unbiased * nb + delta ** 2 * count * nb / tot
        mean = mean + delta * nb / tot
        count = tot
        total_sum += xd.sum(0); total_n += nb
    assert rms.count.item() == 1.0 + total_n
    # the mean needs no merge formula at all: one pseudo-sample at 0 plus every sample seen
    assert np.allclose(rms.running_mean.cpu().numpy(), total_sum / (1.0 + total_n), rtol=1e-9, atol=1e-9 * (abs(shift) + spread))
    assert np.allclose(rms.running_mean.cpu().numpy(), mean, rtol=1e-9, atol=1e-9 * (abs(shift) + spread))
    # fp32 inputs with |mean| >> spread lose digits in ANY fp32-input variance; the kernel accumulates pivoted sums in fp64
    assert np.allclose(rms.running_var.cpu().numpy(), m2 / count, rtol=1e-6, atol=1e-12)


# ----------------------------------------------------------------------------------------------- PPO loss
def _loss_fp64(mu, values, logstd, mb, e_clip, critic_coef, bounds_coef, bound_form):
    """The published formulas, fp64, written out independently of oracle/."""
    sigma = torch.exp(logstd)
    nlp = 0.5 * (((mb["actions"] - mu) / sigma) ** 2).sum(-1) + 0.5 * math.log(2 * math.pi) * 18 + logstd.sum()
    ratio = torch.exp(mb["old_neglogp"] - nlp)
    a = torch.max(-mb["advantages"] * ratio, -mb["advantages"] * ratio.clamp(1 - e_clip, 1 + e_clip))
    vclip = mb["old_values"] + (values - mb["old_values"]).clamp(-e_clip, e_clip)
    c = torch.max((values - mb["returns"]) ** 2, (vclip - mb["returns"]) ** 2)
    if bound_form == "v1.1.3":
        b = (torch.clamp_max(mu - 1.1, 0) ** 2 + torch.clamp_max(-mu + 1.1, 0) ** 2).sum(-1)
    else:
        b = (torch.clamp_min(mu - 1.1, 0) ** 2 + torch.clamp_max(mu + 1.1, 0) ** 2).sum(-1)
    return a.mean() + 0.5 * critic_coef * c.mean() + bounds_coef * b.mean()


@settings(**SET)
@given(m=st.integers(1, 700), seed=st.integers(0, 2 ** 20), e_clip=st.floats(0.05, 0.4), critic_coef=st.floats(0.5, 4.0),
       bounds_coef=st.floats(0.0, 0.01), bound_form=st.sampled_from(["v1.1.3", "outside"]), mu_scale=st.floats(0.2, 3.0))
def test_ppo_loss_gradients_equal_fp64_autograd(m, seed, e_clip, critic_coef, bounds_coef, bound_form, mu_scale):
    from bez_isaacgym_b200 import synthetic_gym as sg
    ops = _ops()
    mb = sg.make_minibatch(m, seed=seed)
    mb["mu"] = mb["mu"] * mu_scale
    mb["advantages"][::3] *= -1.0
    d64 = {k: v.double() for k, v in mb.items()}
    d64["values"], d64["old_values"], d64["returns"] = d64["values"].view(-1), d64["old_values"].view(-1), d64["returns"].view(-1)
    mu = d64["mu"].clone().requires_grad_(True); val = d64["values"].clone().requires_grad_(True)
    ls = d64["logstd"].clone().requires_grad_(True)
    loss = _loss_fp64(mu, val, ls, d64, e_clip, critic_coef, bounds_coef, bound_form)
    loss.backward()
    dv = {k: v.cuda().contiguous() for k, v in mb.items()}
    cfg = ops.make_ppo_cfg(e_clip=e_clip, critic_coef=critic_coef, entropy_coef=0.0, bounds_loss_coef=bounds_coef, bound_form=bound_form)
    stats = torch.empty(8, dtype=torch.float64, device="cuda")
    pp = torch.empty(ops.ppo_scratch_doubles(), dtype=torch.float64, device="cuda")
    g_mu = torch.empty(m, 18, device="cuda"); g_v = torch.empty(m, device="cuda"); g_ls = torch.empty(18, device="cuda")
    ops.ppo_loss(dv["actions"], dv["mu"], dv["logstd"], dv["old_mu"], dv["old_sigma"], dv["values"].view(-1), dv["old_values"].view(-1),
                 dv["returns"].view(-1), dv["old_neglogp"], dv["advantages"], cfg, stats, pp, grad_mu=g_mu, grad_values=g_v,
                 grad_logstd=g_ls)
    assert math.isclose(stats[0].item(), loss.item(), rel_tol=3e-5, abs_tol=3e-6)
    # gradients: fp32 kernel vs fp64 autograd.  A sample sitting within fp32 rounding of a clip boundary / max() tie may pick the
    # other branch: compare through a robust norm and allow a vanishing fraction of boundary samples.
    def close(got, want, what):
        got, want = got.cpu().double(), want
        tol = 2e-4 * want.abs().max().clamp_min(1e-12) + 2e-4 * want.abs()
        bad = ((got - want).abs() > tol).double().mean().item()
        assert bad <= 2.0 / max(m, 1) + 1e-3, (what, bad)
    close(g_mu, mu.grad, "d loss / d mu")
    close(g_v, val.grad, "d loss / d value")
    assert torch.allclose(g_ls.cpu().double(), ls.grad, rtol=2e-3, atol=2e-4 * ls.grad.abs().max().item() + 1e-7), "d loss / d logstd"
