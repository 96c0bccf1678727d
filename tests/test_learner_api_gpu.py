"""GPU: the rl_games-shaped host API (bez_isaacgym_b200.learner) against oracle.rl_games_oracle -- the calls an
rl_games A2CAgent would make: RunningMeanStd module (train / eval / unnorm, checkpoint layout), discount_values,
normalize_advantages and the fused PPO loss as an autograd Function feeding a policy MLP."""
import json
import math
import os

import pytest
import torch

from bez_isaacgym_b200 import synthetic_gym as sg
from tests import _util as U

pytestmark = pytest.mark.gpu
FACTS = os.path.join(os.path.dirname(__file__), "golden", "checkpoint_facts.json")


def test_running_mean_std_module_matches_oracle_and_loads_reference_checkpoint_layout():
    from bez_isaacgym_b200.learner import RunningMeanStd
    from oracle import rl_games_oracle as rg
    mod = RunningMeanStd(54).cuda()
    orc = rg.RunningMeanStd(54)
    g = torch.Generator().manual_seed(0)
    for it in range(4):                                        # 5 mini-epochs x minibatches in the real loop
        x = torch.randn(32768, 54, generator=g) * 2 + it
        want = orc(x)
        got = mod(x.cuda())
        assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-4)
    assert mod.count.item() == orc.count.item() == 1 + 4 * 32768
    assert torch.allclose(mod.running_mean.cpu(), orc.running_mean, rtol=2e-6, atol=1e-7)
    assert torch.allclose(mod.running_var.cpu(), orc.running_var, rtol=2e-5, atol=1e-7)
    mod.eval(); orc.training = False
    x = torch.randn(4096, 54, generator=g)
    before = mod.running_mean.clone()
    U.assert_close(mod(x.cuda()), torch.clamp((x - mod.running_mean.cpu().float()) /
                                              torch.sqrt(mod.running_var.cpu().float() + 1e-5), -5, 5), what="eval")
    assert torch.equal(mod.running_mean, before)
    # state-dict layout of the reference's shipped checkpoint loads as is
    with open(FACTS) as f:
        facts = json.load(f)
    sd = {"running_mean": torch.tensor(facts["obs_running_mean"], dtype=torch.float64),
          "running_var": torch.tensor(facts["obs_running_var"], dtype=torch.float64),
          "count": torch.tensor(facts["obs_count"], dtype=torch.float64)}
    mod.load_state_dict(sd)
    assert mod.count.item() == 1 + 5 * facts["frame"]
    y = mod(torch.zeros(2, 54, device="cuda"))
    assert torch.isfinite(y).all() and y.abs().max() <= 5.0
    val = RunningMeanStd((1,)).cuda()
    v = torch.randn(1000, 1, generator=g)
    out = val(v.cuda())
    un = val(out, unnorm=True) if not val.eval() else None
    val.eval()
    assert torch.allclose(val(val(v.cuda()), unnorm=True).cpu(), v, rtol=1e-4, atol=1e-4)


def test_discount_values_and_advantage_normalisation_api():
    from bez_isaacgym_b200 import learner as L
    from oracle import rl_games_oracle as rg
    n, T = 4096, 32
    rewards, values, dones, last_values, last_dones = sg.make_rollout(n, T, seed=5, p_done=0.02)
    want_adv = rg.discount_values(last_dones.float(), last_values, dones.float(), values, rewards, 0.99, 0.95)
    want_ret = want_adv + values
    adv, ret = L.discount_values(last_dones.float().cuda(), last_values.cuda(), dones.float().cuda(), values.cuda(),
                                 rewards.cuda(), 0.99, 0.95, return_returns=True)
    scale = values.abs() + 1.0
    U.assert_close(adv, want_adv, scale=scale, what="mb_advs")
    U.assert_close(ret, want_ret, scale=scale, what="mb_returns")
    flat_r, flat_v = rg.swap_and_flatten01(want_ret), rg.swap_and_flatten01(values)
    assert torch.equal(L.swap_and_flatten01(ret).cpu(), rg.swap_and_flatten01(ret.cpu()))
    want, _, _ = rg.prepare_dataset(flat_r, flat_v, rg.RunningMeanStd(1))
    got = L.normalize_advantages(flat_r.cuda(), flat_v.cuda())
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("bound_form", ["v1.1.3", "outside"])
def test_ppo_loss_autograd_through_policy_mlp(bound_form):
    """Gradients w.r.t. the MLP parameters (124 237, the BezKick policy shape) equal the oracle's autograd ones."""
    from bez_isaacgym_b200.learner import PPOLossConfig, ppo_loss
    from oracle import rl_games_oracle as rg
    m = 4096
    torch.manual_seed(0)
    def make():
        torch.manual_seed(3)
        trunk = torch.nn.Sequential(torch.nn.Linear(54, 400), torch.nn.ELU(), torch.nn.Linear(400, 200), torch.nn.ELU(),
                                    torch.nn.Linear(200, 100), torch.nn.ELU())
        return trunk, torch.nn.Linear(100, 18), torch.nn.Linear(100, 1), torch.nn.Parameter(torch.zeros(18) - 0.5)
    mb = sg.make_minibatch(m, seed=9)
    obs = torch.randn(m, 54, generator=torch.Generator().manual_seed(4))

    trunk, mu_h, v_h, sigma = make()
    h = trunk(obs)
    o = rg.ppo_loss(dict(mb, mu=mu_h(h), values=v_h(h), logstd=sigma), bound_form=bound_form)
    o["loss"].backward()
    ref_grads = [p.grad.clone() for p in list(trunk.parameters()) + list(mu_h.parameters()) + list(v_h.parameters()) + [sigma]]
    assert sum(g.numel() for g in ref_grads) == 124237

    trunk, mu_h, v_h, sigma = make()
    for mod in (trunk, mu_h, v_h):
        mod.cuda()
    sigma = torch.nn.Parameter(sigma.detach().cuda())
    d = {k: v.cuda() for k, v in mb.items()}
    h = trunk(obs.cuda())
    loss, info = ppo_loss(mu_h(h), v_h(h), sigma, d["actions"], d["old_mu"], d["old_sigma"], d["old_values"], d["returns"],
                          d["old_neglogp"], d["advantages"], PPOLossConfig(bound_form=bound_form))
    loss.backward()
    got = [p.grad for p in list(trunk.parameters()) + list(mu_h.parameters()) + list(v_h.parameters()) + [sigma]]
    assert abs(loss.item() - o["loss"].item()) < 2e-5 * max(1.0, abs(o["loss"].item()))
    assert abs(info["kl"].item() - o["kl"].item()) < 1e-5 and abs(info["a_loss"].item() - o["a_loss"].item()) < 1e-5
    for a, b in zip(got, ref_grads):
        assert torch.allclose(a.cpu(), b, rtol=2e-3, atol=2e-6), (a.cpu() - b).abs().max()


def test_a2c_agent_epoch_reproduces_the_checkpoint_cadence():
    """Two epochs of ``A2CAgent`` (rollout + dataset + mini-epochs through the kernels): the normaliser counts obey the
    identities the reference's shipped checkpoint pins -- ``count_obs = 1 + mini_epochs * frame``, ``count_val = 1 + 2 * frame``
     --, losses are finite, parameters move, the adaptive LR stays inside its bounds, and a checkpoint written
    with rl_games' keys restores the agent."""
    from bez_isaacgym_b200 import bez_model as bm
    from bez_isaacgym_b200 import learner as L
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    from bez_isaacgym_b200.tasks import KickEnv
    n = 1024
    env = KickEnv(bm.default_task_cfg(n), "cuda:0", 0, True, sim=SyntheticGym(n, device="cuda:0", seed=2))
    agent = L.A2CAgent(env, dict(horizon_length=8, minibatch_size=2048, mini_epochs=5), seed=1)
    before = [p.detach().clone() for p in agent.model.parameters()]
    for _ in range(2):
        info = agent.train_epoch()
    frame = 2 * n * 8
    assert agent.frame == frame and agent.epoch_num == 2
    assert float(agent.running_mean_std.count) == 1 + 5 * frame
    assert float(agent.value_mean_std.count) == 1 + 2 * frame
    assert all(math.isfinite(float(info[k])) for k in ("a_loss", "c_loss", "kl")) and 1e-6 <= info["lr"] <= 1e-2
    assert any(not torch.equal(a, b) for a, b in zip(before, agent.model.parameters()))
    assert torch.isfinite(agent.running_mean_std.running_var).all() and float(agent.running_mean_std.running_var.min()) >= 0
    # obs column 52:54 is the constant ball_init (0.175, 0): its running mean converges there and its variance towards 0
    assert abs(float(agent.running_mean_std.running_mean[52]) - 0.175) < 1e-3
    ck = agent.get_full_state_weights()
    assert set(ck) >= {"model", "running_mean_std", "reward_mean_std", "optimizer", "epoch", "frame", "last_mean_rewards"}
    assert all(k.startswith("a2c_network.") for k in ck["model"])
    env2 = KickEnv(bm.default_task_cfg(n), "cuda:0", 0, True, sim=SyntheticGym(n, device="cuda:0", seed=2))
    other = L.A2CAgent(env2, dict(horizon_length=8, minibatch_size=2048, mini_epochs=5), seed=9)
    other.set_full_state_weights(ck)
    assert other.frame == frame and float(other.running_mean_std.count) == 1 + 5 * frame
    for a, b in zip(agent.model.parameters(), other.model.parameters()):
        assert torch.equal(a, b)


def test_a2c_agent_cuda_graph_learner_phase_equals_eager():
    """``cuda_graph: True``: the mini-epoch loop (obs RunningMeanStd, MLP forward / backward, fused PPO loss, clip, Adam, adaptive-KL
    schedule evaluated on the device) replayed as ONE CUDA graph from the third epoch on.  Same kernels in the same order as the
    eager loop: parameters, normaliser statistics and the learning rate agree after four epochs."""
    from bez_isaacgym_b200 import bez_model as bm
    from bez_isaacgym_b200 import learner as L
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    from bez_isaacgym_b200.tasks import KickEnv
    n = 1024
    agents = []
    for graph in (False, True):
        env = KickEnv(bm.default_task_cfg(n), "cuda:0", 0, True, sim=SyntheticGym(n, device="cuda:0", seed=2))
        agents.append(L.A2CAgent(env, dict(horizon_length=8, minibatch_size=2048, mini_epochs=3, mixed_precision=False,
                                           cuda_graph=graph), seed=1))
    infos = [[ag.train_epoch() for _ in range(4)] for ag in agents]
    eager, graphed = agents
    assert graphed._learn_graph is not None, "the learner phase was captured"
    assert float(eager.running_mean_std.count) == float(graphed.running_mean_std.count) == 1 + 3 * 4 * n * 8
    torch.testing.assert_close(eager.running_mean_std.running_mean, graphed.running_mean_std.running_mean, rtol=1e-6, atol=1e-9)
    for a, b in zip(eager.model.parameters(), graphed.model.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)
    assert infos[0][-1]["lr"] == pytest.approx(infos[1][-1]["lr"], rel=1e-5)
    assert float(infos[0][-1]["kl"]) == pytest.approx(float(infos[1][-1]["kl"]), rel=1e-3, abs=1e-7)


@pytest.mark.parametrize("slabs", [False, True])
def test_planned_running_mean_std_equals_the_per_minibatch_updates(slabs):
    """``RunningMeanStd.plan`` / ``planned`` (``bezk_rms_merge_sequence``): ONE moments pass per distinct minibatch and one merge
    kernel for the epoch's 5 x 4 train-mode updates give the statistics -- after EVERY update -- and the normalised minibatches of
    the per-minibatch ``forward`` calls rl_games makes (same arithmetic; only the moments' pivot differs, at fp64 rounding level),
    also against the fp64 oracle, and keep the checkpoint identity count = 1 + 5 * frame."""
    from bez_isaacgym_b200.learner import RunningMeanStd
    from oracle import rl_games_oracle as rg
    T, n, c, nmb, mini_epochs = 32, 1024, 54, 4, 5
    E = n // nmb
    g = torch.Generator().manual_seed(3)
    obses = (torch.randn(T, n, c, generator=g) * torch.linspace(0.5, 3, c) + torch.linspace(-2, 2, c)).cuda()
    if slabs:
        batches = [obses[:, i * E:(i + 1) * E] for i in range(nmb)]                 # read in place from time-major storage
    else:
        batches = [obses[:, i * E:(i + 1) * E].reshape(-1, c).contiguous() for i in range(nmb)]
    order = list(range(nmb)) * mini_epochs
    seq_mod, plan_mod, orc = RunningMeanStd(c).cuda(), RunningMeanStd(c).cuda(), rg.RunningMeanStd(c)
    for m in (seq_mod, plan_mod):                                                   # a warm state, not the (0, 1, 1) start
        m.running_mean.copy_(torch.linspace(-1, 1, c)); m.running_var.fill_(2.0); m.count.fill_(1000.0)
    orc.running_mean.copy_(torch.linspace(-1, 1, c).double()); orc.running_var.fill_(2.0); orc.count.fill_(1000.0)
    seq = plan_mod.plan(batches, order)
    assert seq.shape == (len(order), 2, c)
    for u, b in enumerate(order):
        want = seq_mod(batches[b])                                                  # the per-minibatch train forward
        got = plan_mod.planned(u, batches[b])
        assert torch.allclose(seq[u, 0], seq_mod.running_mean, rtol=1e-12, atol=1e-13), u
        assert torch.allclose(seq[u, 1], seq_mod.running_var, rtol=1e-11, atol=1e-13), u
        U.assert_close(got, want, rtol=1e-6, atol=1e-6, what=f"normalised minibatch of update {u}")
        flat = batches[b].reshape(-1, c).cpu()
        U.assert_close(got, orc(flat), rtol=1e-4, atol=1e-4, what=f"oracle, update {u}")
    assert torch.allclose(plan_mod.running_mean, seq_mod.running_mean, rtol=1e-12, atol=1e-13)
    assert torch.allclose(plan_mod.running_var, seq_mod.running_var, rtol=1e-11, atol=1e-13)
    assert plan_mod.count.item() == seq_mod.count.item() == 1000 + mini_epochs * T * n
    # one launch per mini-epoch: the four minibatches of mini-epoch `me`, each with the statistics of ITS update
    for me in range(mini_epochs):
        grp = plan_mod.planned_group(me * nmb, batches)
        assert grp.shape == (nmb, T * E, c)
        for i in range(nmb):
            assert torch.equal(grp[i], plan_mod.planned(me * nmb + i, batches[i]).view(T * E, c)), (me, i)
    with pytest.raises(Exception):
        plan_mod.planned_group(len(order) - 1, batches)                            # runs past the planned updates
    # a second epoch reuses the plan's buffers; eval mode refuses to plan
    plan_mod.plan(batches, order)
    assert plan_mod.count.item() == 1000 + 2 * mini_epochs * T * n
    plan_mod.eval()
    with pytest.raises(RuntimeError):
        plan_mod.plan(batches, order)
