"""CPU: the restated rl_games==1.1.3 math (oracle.rl_games_oracle).  rl_games is not vendored in the reference, so
this half is PARITY UNPINNED except for what the reference's shipped checkpoint pins (tests/golden/
checkpoint_facts.json, extracted by oracle/make_golden.py from results/Bez_Kick/Normal/Bez_Kick_33.pth):
state layout (fp64 running_mean / running_var / count), update cadence (count identities) and -- through the
constant observation columns -- the exact form of the parallel-variance merge with its (mean 0, var 1, count 1) start.
The remaining tests are closed-form / property checks of the restatement itself."""
import json
import math
import os

import pytest
import torch

from oracle import rl_games_oracle as rg

FACTS = os.path.join(os.path.dirname(__file__), "golden", "checkpoint_facts.json")


@pytest.fixture(scope="module")
def facts():
    with open(FACTS) as f:
        return json.load(f)


def test_checkpoint_state_layout(facts):
    for key in ("running_mean_std", "reward_mean_std", "model", "optimizer", "scaler", "epoch", "frame"):
        assert key in facts["keys"]
    assert facts["obs_rms"]["running_mean"] == {"shape": [54], "dtype": "torch.float64"}
    assert facts["obs_rms"]["running_var"] == {"shape": [54], "dtype": "torch.float64"}
    assert facts["obs_rms"]["count"] == {"shape": [], "dtype": "torch.float64"}
    assert facts["val_rms"]["running_mean"]["shape"] == [1]
    orc = rg.RunningMeanStd(54).state_dict()
    assert {k: (list(v.shape), str(v.dtype)) for k, v in orc.items()} == \
        {k: (v["shape"], v["dtype"]) for k, v in facts["obs_rms"].items()}


def test_checkpoint_update_cadence_identities(facts):
    frame, epoch = facts["frame"], facts["epoch"]
    assert frame == epoch * 4096 * 32                                  # horizon 32 x 4096 envs per epoch
    assert facts["obs_count"] == 1 + 5 * frame                         # obs RMS updated in every mini-epoch (5x)
    assert facts["val_count"] == 1 + 2 * frame                         # value RMS: values + returns per epoch
    assert facts["adam_steps"] == epoch * 5 * 4 - 56                   # 20 minibatch steps / epoch minus AMP skips
    assert math.isclose(facts["lr"], 3e-4 / 1.5 ** 4, rel_tol=1e-12)   # adaptive-KL schedule, factor 1.5
    assert sum(math.prod(s) for s in facts["model_shapes"].values()) == 124237


def test_checkpoint_pins_the_merge_formula(facts):
    """obs[52:54] = ball_init = (0.175, 0) is constant, so its running stats are a closed form of the merge:
    start (mean 0, var 1, count 1); first minibatch of 32768 samples with batch var 0:
        var_N = (1 + delta^2 * B/(B+1)) / N,   mean_N = c * (N-1)/N   with delta = c."""
    n = facts["obs_count"]
    b = 32768.0
    c = float(torch.tensor(0.175, dtype=torch.float32))
    want_var52 = (1.0 + c * c * b / (b + 1.0)) / n
    assert math.isclose(facts["obs_running_var"][52], want_var52, rel_tol=1e-6)
    assert math.isclose(facts["obs_running_var"][53], 1.0 / n, rel_tol=1e-9)
    # fp32 batch means (input.mean(0) is fp32) leave ~1e-7 of accumulated rounding in the fp64 running mean
    assert math.isclose(facts["obs_running_mean"][52], c * (n - 1.0) / n, rel_tol=5e-7)
    assert facts["obs_running_mean"][53] == 0.0
    # and the restatement reproduces that closed form when fed constant batches
    orc = rg.RunningMeanStd(2)
    x = torch.tensor([[0.175, 0.0]]).repeat(32768, 1)
    for _ in range(7):
        orc(x)
    cnt = 1.0 + 7 * b
    assert math.isclose(orc.running_var[0].item(), (1.0 + c * c * b / (b + 1.0)) / cnt, rel_tol=1e-4)
    assert math.isclose(orc.running_var[1].item(), 1.0 / cnt, rel_tol=1e-12)
    assert math.isclose(orc.running_mean[0].item(), c * (cnt - 1.0) / cnt, rel_tol=5e-7)


def test_checkpoint_lin_acc_statistics_confirm_prev_lin_vel_aliasing(facts):
    """obs[36:39] has mean ~(-0.106, -0.085, 0.982) and tiny variance: the third column of the mis-convention
    matrix times the UNIT gravity vector, i.e. lin_acc == -g every step (prev_lin_vel aliases the velocity view)."""
    m, v = facts["obs_running_mean"][36:39], facts["obs_running_var"][36:39]
    assert 0.95 < m[2] < 1.0 and abs(m[0]) < 0.2 and abs(m[1]) < 0.2 and max(v) < 0.02
    assert abs(sum(x * x for x in m) - 1.0) < 0.05


def test_running_mean_std_matches_whole_history_statistics():
    g = torch.Generator().manual_seed(0)
    orc = rg.RunningMeanStd(5)
    chunks = [torch.randn(100 + 37 * i, 5, generator=g) * (i + 1) + i for i in range(4)]
    for c in chunks:
        orc(c)
    # the merge is exact for (count-weighted) mean; the variance mixes unbiased batch variances by design
    allx = torch.cat(chunks).double()
    tot = 1 + allx.shape[0]
    assert torch.allclose(orc.running_mean, allx.sum(0) / tot, rtol=1e-6, atol=1e-7)
    assert orc.count.item() == tot
    y = orc(chunks[0], unnorm=False)
    assert y.abs().max() <= 5.0
    orc.training = False
    z = orc(orc(chunks[1]), unnorm=True)
    inside = (orc(chunks[1]).abs() < 5.0)
    assert torch.allclose(z[inside], chunks[1][inside], rtol=1e-4, atol=1e-4)


def test_discount_values_closed_form():
    """T=3, one env, no dones: adv_t = sum_k (gamma*tau)^k delta_{t+k}."""
    gamma, tau = 0.99, 0.95
    r = torch.tensor([[[1.0]], [[2.0]], [[3.0]]]); v = torch.tensor([[[0.5]], [[0.25]], [[0.125]]])
    last_v = torch.tensor([[4.0]])
    dones = torch.zeros(3, 1); last_d = torch.zeros(1)
    adv = rg.discount_values(last_d, last_v, dones, v, r, gamma, tau)
    d2 = 3.0 + gamma * 4.0 - 0.125
    d1 = 2.0 + gamma * 0.125 - 0.25
    d0 = 1.0 + gamma * 0.25 - 0.5
    want = [d0 + gamma * tau * (d1 + gamma * tau * d2), d1 + gamma * tau * d2, d2]
    assert torch.allclose(adv.view(-1), torch.tensor(want), rtol=1e-6)
    # a done observed BEFORE step 2 cuts the bootstrap from step 1
    dones[2] = 1.0
    adv = rg.discount_values(last_d, last_v, dones, v, r, gamma, tau)
    assert math.isclose(adv[1].item(), 2.0 - 0.25, rel_tol=1e-6)
    assert torch.equal(rg.swap_and_flatten01(torch.arange(6).view(3, 2, 1)).view(-1), torch.tensor([0, 2, 4, 1, 3, 5]))


def test_ppo_loss_pieces():
    m = 64
    g = torch.Generator().manual_seed(1)
    mu = torch.randn(m, 18, generator=g); logstd = 0.1 * torch.randn(18, generator=g)
    sigma = logstd.exp().expand(m, 18)
    x = mu + sigma * torch.randn(m, 18, generator=g)
    nlp = rg.neglogp(x, mu, sigma, logstd.expand(m, 18))
    want = -torch.distributions.Normal(mu, sigma).log_prob(x).sum(-1)
    assert torch.allclose(nlp, want, rtol=1e-5, atol=1e-5)
    assert torch.allclose(rg.entropy(logstd.expand(m, 18)), torch.distributions.Normal(mu, sigma).entropy().sum(-1), rtol=1e-5)
    assert rg.policy_kl(mu, sigma, mu, sigma).abs() < 1e-3                      # the +1e-5 terms leave a small bias
    adv = torch.randn(m, generator=g)
    assert torch.allclose(rg.actor_loss(nlp, nlp, adv, 0.2), -adv)             # ratio 1
    ratio_big = rg.actor_loss(nlp + 1.0, nlp, torch.ones(m), 0.2)               # ratio e > 1.2, A > 0 -> clipped
    assert torch.allclose(ratio_big, torch.full((m,), -1.2))
    mu_in = torch.tensor([[0.5] * 18]); mu_out = torch.tensor([[2.0] * 18])
    assert rg.bound_loss(mu_in, form="outside").item() == 0.0
    assert math.isclose(rg.bound_loss(mu_out, form="outside").item(), 18 * 0.9 ** 2, rel_tol=1e-5)
    assert rg.bound_loss(mu_in, form="v1.1.3").item() > 0.0                     # the recalled 1.1.3 form penalises inside
    assert rg.adaptive_lr(3e-4, 0.02) == 3e-4 / 1.5 and rg.adaptive_lr(3e-4, 0.001) == 3e-4 * 1.5
    assert rg.adaptive_lr(3e-4, 0.008) == 3e-4


def test_prepare_dataset_updates_value_rms_twice():
    orc = rg.RunningMeanStd(1)
    ret = torch.randn(4096, 1); val = torch.randn(4096, 1)
    adv, v, r = rg.prepare_dataset(ret, val, orc)
    assert orc.count.item() == 1 + 2 * 4096
    assert abs(adv.mean().item()) < 1e-5 and abs(adv.std().item() - 1.0) < 1e-4
    shaped = rg.shape_rewards(torch.ones(3), torch.full((3, 1), 2.0), torch.tensor([0, 1, 0]), 0.99)
    assert torch.allclose(shaped.view(-1), torch.tensor([0.01, 0.01 + 0.99 * 2.0, 0.01]))


# ----------------------------------------------------------------------------------------------- rollout storage (8f rows 1-2)
def test_slab_dataset_is_the_ppo_dataset_sample_set_cpu():
    """Host logic only (no kernels): SlabDataset minibatch i == PPODataset minibatch i as a sample set, and
    permutation_to_env_major() maps one row order onto the other."""
    from bez_isaacgym_b200.learner import experience as ex

    class Box:
        def __init__(self, n):
            self.shape = (n,)
    T, N, mb = 4, 24, 16
    buf = ex.ExperienceBuffer(dict(observation_space=Box(54), action_space=Box(18)), dict(num_actors=N, horizon_length=T), "cpu")
    orc = rg.ExperienceBuffer(N, T)
    g = torch.Generator().manual_seed(1)
    for t in range(T):
        for name, shape in (("obses", (N, 54)), ("values", (N, 1)), ("neglogpacs", (N,)), ("actions", (N, 18))):
            val = torch.randn(shape, generator=g)
            buf.update_data(name, t, val); orc.update_data(name, t, val)
    names = ["obses", "values", "neglogpacs", "actions"]
    flat = orc.get_transformed_list(rg.swap_and_flatten01, names)
    ods = rg.PPODataset(N * T, mb); ods.update_values_dict(flat)
    sds = ex.SlabDataset(buf, mb)
    perm = sds.permutation_to_env_major()
    assert len(sds) == len(ods) == 6
    for i in range(len(sds)):
        for k in names:
            rows = sds[i][k].reshape(mb, *sds[i][k].shape[2:])
            assert torch.equal(rows[perm], ods[i][k]), (i, k)
    with pytest.raises(ValueError):
        ex.SlabDataset(buf, 6)              # not a multiple of the horizon
    with pytest.raises(ValueError):
        ex.PPODataset(96, 36)


def test_philox_normals_reference_is_gaussian():
    from oracle import philox_ref
    z = philox_ref.normals18(seed=3, step=9, envs=range(20000)).astype("float64")
    assert z.shape == (20000, 18)
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1.0) < 1e-2 and abs((z ** 4).mean() - 3.0) < 0.1
    again = philox_ref.normals18(seed=3, step=9, envs=range(100, 110))
    assert (again == philox_ref.normals18(seed=3, step=9, envs=range(20000))[100:110]).all()     # keyed by env id


def test_policy_head_oracle_identities():
    """neglogp of the sampled action reduces to 0.5*|noise|^2 + const + sum(logstd); un-normalisation inverts normalisation
    inside the clamp."""
    g = torch.Generator().manual_seed(0)
    mu, logstd, noise = torch.randn(64, 18, generator=g), torch.randn(18, generator=g) * 0.2, torch.randn(64, 18, generator=g)
    vr = rg.RunningMeanStd(1)
    vr.running_mean = torch.tensor([1.5], dtype=torch.float64); vr.running_var = torch.tensor([4.0], dtype=torch.float64)
    vr.training = False
    raw = torch.randn(64, 1, generator=g) * 2 + 1.5
    out = rg.policy_head(mu, logstd, vr(raw), vr, noise)
    want = 0.5 * (noise ** 2).sum(-1) + 0.5 * math.log(2 * math.pi) * 18 + logstd.sum()
    torch.testing.assert_close(out["neglogpacs"], want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out["values"], raw, rtol=1e-5, atol=1e-5)
    assert torch.equal(rg.preprocess_actions(torch.tensor([-3.0, -0.5, 0.25, 7.0])), torch.tensor([-1.0, -0.5, 0.25, 1.0]))


def test_network_and_normaliser_state_match_the_shipped_checkpoint_layout():
    """``A2CNetwork`` / ``RunningMeanStd`` state dicts carry exactly the names, shapes and dtypes of the reference's shipped
    checkpoint (tests/golden/checkpoint_facts.json, App. G), so ``A2CAgent.set_full_state_weights`` can load it."""
    from bez_isaacgym_b200.learner.agent import A2CNetwork, AdaptiveScheduler
    from bez_isaacgym_b200.learner.running_mean_std import RunningMeanStd as RMS
    with open(os.path.join(os.path.dirname(__file__), "golden", "checkpoint_facts.json")) as f:
        facts = json.load(f)
    net = A2CNetwork()
    assert {"a2c_network." + k: list(v.shape) for k, v in net.state_dict().items()} == facts["model_shapes"]
    for mod, key in ((RMS(54), "obs_rms"), (RMS(1), "val_rms")):
        sd = mod.state_dict()
        assert {k: {"shape": list(v.shape), "dtype": str(v.dtype)} for k, v in sd.items()} == facts[key]
    # the checkpoint's learning rate is 3e-4 / 1.5^4: four net down-steps of the adaptive schedule
    sched, lr = AdaptiveScheduler(0.008), 3e-4
    for _ in range(4):
        lr, _ = sched.update(lr, 0.0, 0, 0, 0.02)
    assert lr == pytest.approx(facts["lr"])
    assert sched.update(lr, 0.0, 0, 0, 0.001)[0] == pytest.approx(lr * 1.5) and sched.update(lr, 0.0, 0, 0, 0.008)[0] == lr


def test_shipped_reference_checkpoint_loads_strictly():
    """Where /root/reference exists: the reference's own ``results/Bez_Kick/Normal/Bez_Kick_33.pth`` loads, strict, into
    ``A2CNetwork`` and the two ``RunningMeanStd`` modules (the layout ``A2CAgent.set_full_state_weights`` relies on), and the
    restored statistics are the checkpoint's (count = 1 + 5*frame / 1 + 2*frame)."""
    import numpy
    from bez_isaacgym_b200.learner.agent import A2CNetwork
    from bez_isaacgym_b200.learner.running_mean_std import RunningMeanStd as RMS
    from oracle import reference_loader as rl
    path = os.path.join(rl.REFERENCE_ROOT, "bez_isaacgym", "results", "Bez_Kick", "Normal", "Bez_Kick_33.pth")
    if not os.path.isfile(path):
        pytest.skip("needs /root/reference (authoring container)")
    allow = [(numpy._core.multiarray.scalar, "numpy.core.multiarray.scalar"), (numpy.dtype, "numpy.dtype")]
    allow += [getattr(numpy.dtypes, n) for n in dir(numpy.dtypes) if n.endswith("DType")]
    with torch.serialization.safe_globals(allow):
        ck = torch.load(path, map_location="cpu", weights_only=True)
    net, obs_rms, val_rms = A2CNetwork(), RMS(54), RMS(1)
    net.load_state_dict({k[len("a2c_network."):]: v for k, v in ck["model"].items()}, strict=True)
    obs_rms.load_state_dict(ck["running_mean_std"], strict=True)
    val_rms.load_state_dict(ck["reward_mean_std"], strict=True)
    frame = int(ck["frame"])
    assert float(obs_rms.count) == 1 + 5 * frame and float(val_rms.count) == 1 + 2 * frame
    assert obs_rms.running_mean.dtype == torch.float64 and tuple(obs_rms.running_var.shape) == (54,)
    # the policy runs: ready-pose observation -> finite 18-dim mean action, scalar value
    x = torch.zeros(1, 54)
    mu, value = net(x)
    assert mu.shape == (1, 18) and value.shape == (1, 1) and torch.isfinite(mu).all() and torch.isfinite(value).all()
    # constant observation columns 52:54 (ball_init) pin the normaliser: mean (0.175, 0), variance -> 0
    assert abs(float(obs_rms.running_mean[52]) - 0.175) < 1e-6 and float(obs_rms.running_var[52]) < 1e-6
