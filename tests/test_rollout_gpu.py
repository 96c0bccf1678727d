"""GPU parity of the rollout-storage kernels through the C ABI against ``oracle.rl_games_oracle``:
``swap_and_flatten01`` and the slab-addressed minibatch path (bit-exact data movement; statistics / loss within the
learner tolerances), and the policy-head epilogue (sampling with supplied noise, neglogp, value un-normalisation,
clamp, PD targets).  rl_games is un-vendored: parity unpinned beyond the restatement (see the oracle header)."""
import math

import pytest
import torch

from bez_isaacgym_b200 import bez_model as bm
from bez_isaacgym_b200 import synthetic_gym as sg
from tests import _util as U

pytestmark = pytest.mark.gpu


def _ops():
    from bez_isaacgym_b200 import ops
    return ops


# ----------------------------------------------------------------------------------------------- swap_and_flatten01
@pytest.mark.parametrize("T,N,tail,dtype", [
    (32, 4096, (54,), torch.float32), (32, 4096, (18,), torch.float32), (32, 4096, (1,), torch.float32),
    (32, 4096, (), torch.float32), (32, 4096, (), torch.uint8), (5, 37, (54,), torch.float32), (1, 1, (3,), torch.float32),
    (40, 130, (7,), torch.float32), (33, 77, (), torch.uint8), (32, 1000, (5,), torch.float64), (3, 9, (54,), torch.int64),
])
def test_swap_and_flatten01_bit_exact(T, N, tail, dtype):
    from oracle import rl_games_oracle as rg
    ops = _ops()
    g = torch.Generator().manual_seed(T * 1000 + N)
    if dtype.is_floating_point:
        src = torch.randn((T, N) + tail, generator=g, dtype=dtype)
    else:
        src = torch.randint(0, 200, (T, N) + tail, generator=g, dtype=dtype)
    want = rg.swap_and_flatten01(src)
    got = ops.swap_and_flatten01(src.cuda())
    assert got.shape == want.shape and got.dtype == want.dtype
    assert torch.equal(got.cpu(), want)


def test_swap_and_flatten01_env_range_is_the_dataset_slice():
    """rows [i*mb, (i+1)*mb) of the flattened tensor == flattening envs [i*E, (i+1)*E) only (PPODataset slice)."""
    from oracle import rl_games_oracle as rg
    ops = _ops()
    T, N, mb = 32, 2048, 8192
    E = mb // T
    src = torch.randn(T, N, 54)
    flat = rg.swap_and_flatten01(src)
    d = src.cuda()
    for i in (0, 3, N * T // mb - 1):
        got = ops.swap_and_flatten01(d, env0=i * E, envs=E)
        assert torch.equal(got.cpu(), flat[i * mb:(i + 1) * mb])


def test_swap_and_flatten01_involution_at_full_size():
    """Size-independent property at 262144 envs x 32: flattening (T,N,c) then flattening the result viewed as (N,T,c)
    gives the original back (transposition is an involution)."""
    ops = _ops()
    T, N = 32, 262144
    src = torch.randn(T, N, 18, device="cuda")
    once = ops.swap_and_flatten01(src)
    twice = ops.swap_and_flatten01(once.view(N, T, 18))
    assert torch.equal(twice.view(T, N, 18), src)


# ----------------------------------------------------------------------------------------------- slab minibatches
def _rollout(T, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    return dict(obses=torch.randn(T, N, 54, generator=g) * 2 + 0.3, actions=torch.randn(T, N, 18, generator=g),
                mus=torch.randn(T, N, 18, generator=g) * 0.5, sigmas=torch.rand(T, N, 18, generator=g) + 0.5,
                values=torch.randn(T, N, 1, generator=g), returns=torch.randn(T, N, 1, generator=g),
                neglogpacs=torch.randn(T, N, generator=g) + 20, advantages=torch.randn(T, N, generator=g))


@pytest.mark.parametrize("T,N,mb", [(32, 4096, 32768), (32, 1024, 4096), (8, 96, 64), (5, 35, 35)])
def test_slab_normalize_and_moments_match_flattened_minibatch(T, N, mb):
    """RunningMeanStd train forward on minibatch i: slab path == rl_games path (flatten, slice, update, normalise)."""
    from oracle import rl_games_oracle as rg
    ops = _ops()
    ro = _rollout(T, N, seed=T + N)
    E = mb // T
    flat = rg.swap_and_flatten01(ro["obses"])
    d_obs = ro["obses"].cuda()
    perm = torch.arange(E).repeat_interleave(T) + torch.arange(T).repeat(E) * E       # env-major row -> slab row
    for i in (0, (N * T) // mb - 1):
        rms = rg.RunningMeanStd(54)
        want = rms(flat[i * mb:(i + 1) * mb])
        mean = torch.zeros(54, dtype=torch.float64, device="cuda"); var = torch.ones(54, dtype=torch.float64, device="cuda")
        count = torch.ones(1, dtype=torch.float64, device="cuda")
        acc = torch.empty(109, dtype=torch.float64, device="cuda")
        scratch = torch.empty(ops.rms_scratch_doubles(54), dtype=torch.float64, device="cuda")
        view = d_obs[:, i * E:(i + 1) * E]
        ops.rms_moments_slabs(view, mean.clone(), acc, scratch)
        ops.rms_merge(acc, mean.clone(), mean, var, count)
        y = torch.empty(mb, 54, device="cuda")
        ops.rms_normalize_slabs(view, mean, var, y)
        torch.testing.assert_close(mean.cpu(), rms.running_mean, rtol=2e-6, atol=1e-7)
        torch.testing.assert_close(var.cpu(), rms.running_var, rtol=2e-5, atol=1e-7)
        assert float(count) == float(rms.count)
        U.assert_close(y.cpu()[perm], want, rtol=2e-5, atol=2e-6, what="normalised minibatch (slab order -> env-major)")
        # pure data movement check: normalising with mean 0 / var 1-eps... instead compare against the kernel's own
        # contiguous path on the permuted rows -> bit-exact
        y2 = torch.empty(mb, 54, device="cuda")
        ops.rms_normalize(flat[i * mb:(i + 1) * mb].cuda(), mean, var, y2)
        assert torch.equal(y.cpu()[perm], y2.cpu())


@pytest.mark.parametrize("T,N,mb", [(32, 4096, 32768), (32, 512, 4096), (4, 24, 16)])
def test_slab_ppo_loss_matches_flattened_minibatch(T, N, mb):
    from oracle import rl_games_oracle as rg
    ops = _ops()
    ro = _rollout(T, N, seed=7 + N)
    E = mb // T
    i = (N * T) // mb - 1
    sl = slice(i * mb, (i + 1) * mb)
    fl = {k: rg.swap_and_flatten01(v)[sl] for k, v in ro.items()}
    g = torch.Generator().manual_seed(5)
    logstd = torch.randn(18, generator=g) * 0.1
    mu_new = (fl["mus"] + 0.05 * torch.randn(mb, 18, generator=g)).requires_grad_(True)        # env-major network outputs
    val_new = (fl["values"] + 0.1 * torch.randn(mb, 1, generator=g)).requires_grad_(True)
    logstd_p = logstd.clone().requires_grad_(True)
    out = rg.ppo_loss(dict(mu=mu_new, logstd=logstd_p, values=val_new, actions=fl["actions"], old_mu=fl["mus"],
                           old_sigma=fl["sigmas"], old_values=fl["values"], returns=fl["returns"],
                           old_neglogp=fl["neglogpacs"], advantages=fl["advantages"]), bound_form="outside")
    # critic loss in the oracle works on (M,1); sum over the trailing dim is a no-op
    out["loss"].backward()

    perm = torch.arange(E).repeat_interleave(T) + torch.arange(T).repeat(E) * E       # env-major row r -> slab row perm[r]
    inv = torch.empty_like(perm); inv[perm] = torch.arange(mb)
    dev = {k: v.cuda() for k, v in ro.items()}
    v = {k: t[:, i * E:(i + 1) * E] for k, t in dev.items()}
    mu_slab = mu_new.detach()[inv].contiguous().cuda()          # network outputs in slab (time-major-within-batch) order
    val_slab = val_new.detach().view(-1)[inv].contiguous().cuda()
    cfg = ops.make_ppo_cfg(bound_form="outside")
    stats = torch.empty(8, dtype=torch.float64, device="cuda")
    part = torch.empty(ops.ppo_scratch_doubles(), dtype=torch.float64, device="cuda")
    gmu = torch.empty(mb, 18, device="cuda"); gv = torch.empty(mb, device="cuda"); gls = torch.empty(18, device="cuda")
    nlp = torch.empty(mb, device="cuda")
    ops.ppo_loss_slabs(v["actions"], mu_slab, logstd.cuda(), v["mus"], v["sigmas"], val_slab, v["values"], v["returns"],
                       v["neglogpacs"], v["advantages"], cfg, stats, part, grad_mu=gmu, grad_values=gv, grad_logstd=gls,
                       neglogp_out=nlp)
    s = stats.cpu()
    for k, idx in (("loss", 0), ("a_loss", 1), ("c_loss", 2), ("entropy", 3), ("b_loss", 4), ("kl", 5)):
        assert math.isclose(float(s[idx]), float(out[k]), rel_tol=2e-5, abs_tol=2e-6), (k, float(s[idx]), float(out[k]))
    U.assert_close(nlp.cpu()[perm], out["neglogp"].detach(), rtol=2e-5, atol=2e-5, what="neglogp")
    U.assert_close(gmu.cpu()[perm], mu_new.grad, rtol=2e-4, atol=2e-9, what="grad_mu")
    U.assert_close(gv.cpu()[perm], val_new.grad.view(-1), rtol=2e-4, atol=2e-9, what="grad_values")
    U.assert_close(gls.cpu(), logstd_p.grad, rtol=2e-4, atol=2e-7, what="grad_logstd")
    # and the slab path equals the kernel's own contiguous path on the same rows, bit for bit (pure addressing)
    stats2 = torch.empty_like(stats); gmu2 = torch.empty_like(gmu)
    slab_rows = {k: t.reshape(T * E, *t.shape[2:]).contiguous() for k, t in v.items()}
    ops.ppo_loss(slab_rows["actions"], mu_slab, logstd.cuda(), slab_rows["mus"], slab_rows["sigmas"], val_slab,
                 slab_rows["values"].view(-1), slab_rows["returns"].view(-1), slab_rows["neglogpacs"], slab_rows["advantages"],
                 cfg, stats2, part, grad_mu=gmu2)
    assert torch.equal(stats.cpu(), stats2.cpu()) and torch.equal(gmu.cpu(), gmu2.cpu())


def test_experience_buffer_and_datasets_agree():
    """ExperienceBuffer + get_transformed_list(swap_and_flatten01) + PPODataset == oracle's; SlabDataset minibatch i holds
    the same sample set as PPODataset minibatch i."""
    from oracle import rl_games_oracle as rg
    from bez_isaacgym_b200.learner import experience as ex

    class Box:
        def __init__(self, n):
            self.shape = (n,)
    T, N, mb = 8, 256, 512
    buf = ex.ExperienceBuffer(dict(observation_space=Box(54), action_space=Box(18)), dict(num_actors=N, horizon_length=T), "cuda")
    orc = rg.ExperienceBuffer(N, T)
    g = torch.Generator().manual_seed(0)
    for t in range(T):
        for name, shape in (("obses", (N, 54)), ("rewards", (N, 1)), ("values", (N, 1)), ("neglogpacs", (N,)),
                            ("actions", (N, 18)), ("mus", (N, 18)), ("sigmas", (N, 18))):
            val = torch.randn(shape, generator=g)
            orc.update_data(name, t, val)
            if name in ("obses", "actions"):
                buf.slot(name, t).copy_(val)                 # the in-place route the kernels use
            else:
                buf.update_data(name, t, val.cuda())
        d = (torch.rand(N, generator=g) < 0.1).to(torch.uint8)
        orc.update_data("dones", t, d); buf.update_data("dones", t, d.cuda())
    names = ["obses", "actions", "values", "neglogpacs", "mus", "sigmas", "dones", "rewards"]
    got = buf.get_transformed_list(ex.swap_and_flatten01, names)
    want = orc.get_transformed_list(rg.swap_and_flatten01, names)
    for k in names:
        assert torch.equal(got[k].cpu(), want[k]), k
    ds, ods = ex.PPODataset(N * T, mb), rg.PPODataset(N * T, mb)
    ds.update_values_dict(got); ods.update_values_dict(want)
    sds = ex.SlabDataset(buf, mb)
    assert len(ds) == len(ods) == len(sds) == N * T // mb
    perm = sds.permutation_to_env_major()
    for i in range(len(ds)):
        a, b, s = ds[i], ods[i], sds[i]
        for k in names:
            assert torch.equal(a[k].cpu(), b[k]), (i, k)
            rows = s[k].reshape(mb, *s[k].shape[2:]).cpu()
            assert torch.equal(rows[perm], b[k]), (i, k)


# ----------------------------------------------------------------------------------------------- policy head
@pytest.mark.parametrize("n", [1, 127, 128, 4096, 4099])
def test_policy_head_matches_oracle(n):
    from oracle import rl_games_oracle as rg
    from oracle import task_oracle as to
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    mu = torch.randn(n, 18, generator=g) * 0.8
    logstd = torch.randn(18, generator=g) * 0.3
    vnorm = torch.randn(n, 1, generator=g) * 3          # some beyond the +-5 clamp
    noise = torch.randn(n, 18, generator=g)
    vr = rg.RunningMeanStd(1)
    vr.running_mean = torch.tensor([0.37], dtype=torch.float64); vr.running_var = torch.tensor([2.5], dtype=torch.float64)
    want = rg.policy_head(mu, logstd, vnorm, vr, noise)
    env_act = rg.preprocess_actions(want["actions"])
    goal, ball_init, default, lower, upper = sg.make_constants(n)
    _, want_targets = to.pre_physics(env_act, default, lower, upper, clip_actions=3.9)

    cfg = ops.make_task_cfg()
    out = {k: torch.full((n, 18), float("nan"), device="cuda") for k in ("actions", "mus", "sigmas", "env_actions", "targets")}
    nlp = torch.empty(n, device="cuda"); vals = torch.empty(n, device="cuda")
    ops.policy_head(mu.cuda(), logstd.cuda(), vnorm.view(-1).cuda(), vr.running_mean.cuda(), vr.running_var.cuda(), 1e-5,
                    noise=noise.cuda(), actions=out["actions"], neglogp=nlp, values=vals, mus=out["mus"], sigmas=out["sigmas"],
                    task_cfg=cfg, env_actions=out["env_actions"], targets=out["targets"])
    assert torch.equal(out["mus"].cpu(), mu)
    U.assert_close(out["sigmas"], want["sigmas"], what="sigmas")
    U.assert_close(out["actions"], want["actions"], what="actions")
    U.assert_close(nlp, want["neglogpacs"], rtol=2e-5, atol=2e-5, what="neglogp")
    U.assert_close(vals, want["values"].view(-1), what="values")
    U.assert_close(out["env_actions"], env_act, what="env actions")
    U.assert_close(out["targets"], want_targets, what="targets")


def test_policy_head_philox_noise_is_standard_normal_and_reproducible():
    from oracle import philox_ref
    ops = _ops()
    n = 65536
    z = ops.normal_noise(11, 5, torch.empty(n, 18, device="cuda"))
    z2 = ops.normal_noise(11, 5, torch.empty(n, 18, device="cuda"))
    assert torch.equal(z, z2)
    assert not torch.equal(z, ops.normal_noise(11, 6, torch.empty(n, 18, device="cuda")))
    zc = z.double().cpu()
    assert abs(float(zc.mean())) < 5e-3 and abs(float(zc.var()) - 1.0) < 1e-2
    assert abs(float((zc ** 4).mean()) - 3.0) < 0.1                      # kurtosis of a Gaussian
    col = zc - zc.mean(0)
    corr = (col.T @ col) / n
    assert float((corr - torch.diag(torch.diag(corr))).abs().max()) < 2e-2   # columns uncorrelated
    # the numpy restatement of Philox + Box-Muller reproduces the device draws: the STREAM (counters, keys, pairing) is what is
    # pinned; the device evaluates lg2 / sin / cos / sqrt on the special-function unit (absolute error of a draw up to ~1e-5,
    # reached next to the zeros of sin / cos where the fast range reduction loses its low bits)
    want = philox_ref.normals18(seed=11, step=5, envs=range(64))
    U.assert_close(z[:64].cpu(), torch.from_numpy(want).float(), rtol=1e-5, atol=3e-5, what="philox normals")
    # sampling through the head with noise=None uses exactly these draws
    mu = torch.zeros(n, 18, device="cuda"); logstd = torch.zeros(18, device="cuda")
    act = torch.empty(n, 18, device="cuda")
    ops.policy_head(mu, logstd, noise=None, seed=11, step=5, actions=act)
    assert torch.equal(act, z)


def test_policy_head_rejects_bad_arguments():
    from bez_isaacgym_b200._lib import BezkError
    ops = _ops()
    mu = torch.zeros(4, 18, device="cuda"); logstd = torch.zeros(18, device="cuda")
    with pytest.raises(BezkError):
        ops.policy_head(mu, logstd, values=torch.empty(4, device="cuda"))              # values without value_norm
    with pytest.raises(BezkError):
        ops.policy_head(mu, logstd, targets=torch.empty(4, 18, device="cuda"))         # targets without a task cfg
    with pytest.raises(BezkError):
        ops.policy_head(mu.cpu(), logstd)                                              # no CPU path


# ----------------------------------------------------------------------------------------------- end to end through the API
def test_play_steps_and_mini_epoch_through_the_api():
    """A 4-step rl_games-shaped rollout on KickEnv with everything written in place (obs slot by the step kernel;
    actions / neglogpacs / values / mus / sigmas slots and the PD targets by the policy-head kernel), then one mini-epoch
    of minibatches taken two ways -- rl_games' flatten + PPODataset and the copy-free SlabDataset -- must give the same
    running statistics and the same losses."""
    from bez_isaacgym_b200 import bez_model as bmod
    from bez_isaacgym_b200 import learner as L
    from bez_isaacgym_b200.learner import experience as ex
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    from bez_isaacgym_b200.tasks.kick_env import KickEnv

    T, N, mb = 4, 1024, 1024
    dev = torch.device("cuda")
    cfg = bmod.default_task_cfg(N, rl_device="cuda:0")
    env = KickEnv(cfg, "cuda:0", 0, True, sim=SyntheticGym(N, device="cuda:0", seed=3))
    buf = ex.ExperienceBuffer(dict(observation_space=env.observation_space, action_space=env.action_space),
                              dict(num_actors=N, horizon_length=T), dev)
    torch.manual_seed(0)
    actor = torch.nn.Linear(54, 18).to(dev); critic = torch.nn.Linear(54, 1).to(dev)
    logstd = torch.zeros(18, device=dev)
    obs_rms, val_rms = L.RunningMeanStd(54).to(dev), L.RunningMeanStd(1).to(dev)
    obs_rms.eval(); val_rms.eval()
    rewards = torch.empty(T, N, 1, device=dev)
    obs = env.reset()["obs"].clone()
    dones = torch.zeros(N, dtype=torch.uint8, device=dev)
    for t in range(T):
        buf.update_data("obses", t, obs)
        buf.update_data("dones", t, dones)
        with torch.no_grad():
            x = obs_rms(obs)
            mu, value = actor(x), critic(x)
        res = L.policy_head(mu, logstd, value, val_rms, experience=buf, t=t, seed=7, step=t, env=env)
        # the same targets K0 would have produced from the head's env actions
        want_targets = torch.empty(N, 18, device=dev)
        _ops().pre_physics(res["env_actions"], want_targets, env._kcfg)
        assert torch.equal(env.targets, want_targets)
        if t + 1 < T:
            env.set_obs_target(buf.slot("obses", t + 1))        # the step kernel writes the next slot directly
        # reward shaping + value bootstrap + uint8 dones ride in the step kernel's epilogue
        env.set_rollout_targets(values=res["values"], shaped_rewards=rewards[t], dones_u8=dones, gamma=0.99, scale_value=0.01)
        o, rew, done, info = env.step_precomputed_targets(res["env_actions"])
        if t + 1 < T:
            assert o["obs"].data_ptr() == buf.slot("obses", t + 1).data_ptr()
        obs = o["obs"]
        from oracle import rl_games_oracle as rgo
        want_shaped = rgo.shape_rewards(rew.cpu(), res["values"].cpu(), info["time_outs"].cpu(), 0.99)
        assert torch.equal(rewards[t].cpu(), want_shaped), "play_steps reward shaping / value bootstrap (bit-exact)"
        assert torch.equal(dones.cpu(), done.cpu().to(torch.uint8))
        dones = dones.clone()
    last_values = torch.randn(N, 1, device=dev)
    advs, rets = L.discount_values(dones, last_values, buf.tensor_dict["dones"], buf.tensor_dict["values"], rewards, 0.99, 0.95,
                                   return_returns=True)
    adv_n = L.normalize_advantages(rets, buf.tensor_dict["values"]).view(T, N)

    # ---- mini-epoch, rl_games' way: flatten + slice
    names = ["obses", "actions", "mus", "sigmas", "values", "neglogpacs"]
    flat = buf.get_transformed_list(L.swap_and_flatten01, names)
    flat["returns"] = L.swap_and_flatten01(rets); flat["advantages"] = L.swap_and_flatten01(adv_n)
    ds = ex.PPODataset(N * T, mb); ds.update_values_dict(flat)
    sds = ex.SlabDataset(buf, mb, extra=dict(returns=rets, advantages=adv_n))
    perm = sds.permutation_to_env_major(dev)
    inv = torch.empty_like(perm); inv[perm] = torch.arange(mb, device=dev)
    rms_a, rms_b = L.RunningMeanStd(54).to(dev), L.RunningMeanStd(54).to(dev)
    for i in range(len(ds)):
        a, s = ds[i], sds[i]
        xa = rms_a(a["obses"])                                    # train mode: update + normalise
        xs = rms_b(s["obses"])                                    # slab view, read in place
        assert torch.equal(xs[perm], xa)
        with torch.no_grad():
            mu_a, v_a = actor(xa), critic(xa)
        loss_a, info_a = L.ppo_loss(mu_a, v_a, logstd, a["actions"], a["mus"], a["sigmas"], a["values"], a["returns"],
                                    a["neglogpacs"], a["advantages"])
        loss_s, info_s = L.ppo_loss(mu_a[inv].contiguous(), v_a[inv].contiguous(), logstd, s["actions"], s["mus"], s["sigmas"],
                                    s["values"], s["returns"], s["neglogpacs"], s["advantages"])
        assert math.isclose(float(loss_a), float(loss_s), rel_tol=1e-6, abs_tol=1e-7)
        assert math.isclose(float(info_a["kl"]), float(info_s["kl"]), rel_tol=1e-6, abs_tol=1e-9)
        assert torch.equal(info_s["neglogp"][perm], info_a["neglogp"])
    torch.testing.assert_close(rms_a.running_mean, rms_b.running_mean, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(rms_a.running_var, rms_b.running_var, rtol=1e-10, atol=1e-12)
