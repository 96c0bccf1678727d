"""SURVEY 8 a16 -- rl_games ``play_steps`` reward path inside the fused step kernel (``bezk_post_physics_rollout``), and the
``env_base`` (global env id) keying of the Philox reset / exploration noise.

The epilogue is elementwise on the kernel's own ``rew`` / ``timeout`` / ``reset`` outputs, so it is checked BIT-EXACTLY against
``oracle.rl_games_oracle.shape_rewards`` (the restated ``DefaultRewardsShaper`` + value bootstrap) fed with those outputs."""
import pytest
import torch

from bez_isaacgym_b200 import bez_model as bm
from bez_isaacgym_b200 import synthetic_gym as sg
from tests import _util as U

pytestmark = pytest.mark.gpu


def _ops():
    from bez_isaacgym_b200 import ops
    return ops


def _run(task, n, seed, rollout_kw=None, values=None, env_base=0, want_shaped=True, want_dones=True, step=3):
    ops = _ops()
    st = sg.make_state(n, seed=seed, task=task).to("cuda")
    actors, nb, width = bm.task_dims(task)
    cfg = ops.make_task_cfg(num_bodies=nb)
    goal = torch.tensor([[1.5, 0.0]] if task == "kick" else [[2.0, 0.0]], device="cuda").repeat(n, 1)
    ball_init = torch.tensor([[0.175, 0.0]], device="cuda").repeat(n, 1) if task == "kick" else None
    goal_angle = torch.full((n,), 1.5708, device="cuda") if task == "orient" else None
    init_root = sg.make_initial_root_states(n, "cuda", task=task)
    progress, reset = sg.make_bookkeeping(n, seed=seed + 1, device="cuda", p_reset=0.1)
    progress[: min(n, 4)] = torch.tensor([899, 900, 897, 898], device="cuda")[: min(n, 4)]
    timeout = torch.empty(n, dtype=torch.long, device="cuda")
    obs = torch.empty(n, width, device="cuda"); rew = torch.empty(n, device="cuda")
    prev = torch.zeros(n, 3, device="cuda")
    shaped = torch.full((n,), float("nan"), device="cuda") if want_shaped else None
    dones = torch.full((n,), 7, dtype=torch.uint8, device="cuda") if want_dones else None
    rcfg = ops.make_rollout_cfg(**rollout_kw) if rollout_kw is not None else None
    ops.post_physics_rollout(task, st.dof_state, st.rigid_body, st.root_states, st.net_contact, goal, init_root, reset, progress,
                             timeout, cfg, obs, rew, rollout_cfg=rcfg, values=values, shaped_rewards=shaped, dones_u8=dones,
                             goal_angle=goal_angle, ball_init=ball_init, prev_lin_vel=prev, seed=5, step=step, env_base=env_base)
    torch.cuda.synchronize()
    return dict(st=st, obs=obs, rew=rew, reset=reset, progress=progress, timeout=timeout, shaped=shaped, dones=dones)


@pytest.mark.parametrize("task", ["kick", "walk", "orient"])
@pytest.mark.parametrize("n", [1, 31, 4099, 70001])
@pytest.mark.parametrize("kw", [dict(gamma=0.99, scale_value=0.01, shift_value=0.0, value_bootstrap=True),
                                dict(gamma=0.9, scale_value=0.5, shift_value=-1.25, value_bootstrap=True),
                                dict(gamma=0.99, scale_value=0.01, shift_value=0.0, value_bootstrap=False)])
def test_reward_shaping_epilogue_is_bit_exact(task, n, kw):
    from oracle import rl_games_oracle as rg
    values = torch.randn(n, generator=torch.Generator().manual_seed(n)).cuda() * 3
    r = _run(task, n, seed=40 + n, rollout_kw=kw, values=values if kw["value_bootstrap"] else None)
    rew, timeout, reset = r["rew"].cpu(), r["timeout"].cpu(), r["reset"].cpu()
    assert int(timeout.sum()) >= min(n, 2), "the timeout rows (progress 899 / 900) must be exercised"
    want = rg.shape_rewards(rew, values.cpu().view(n, 1), timeout if kw["value_bootstrap"] else torch.zeros_like(timeout),
                            kw["gamma"], scale=kw["scale_value"], shift=kw["shift_value"]).view(-1)
    got = r["shaped"].cpu()
    same = (got == want) | (got.isnan() & want.isnan())
    assert bool(same.all()), f"{int((~same).sum())} of {n} shaped rewards differ"
    assert torch.equal(r["dones"].cpu(), (reset != 0).to(torch.uint8))
    # and the step itself is what bezk_post_physics_task computes
    ops = _ops()
    st = sg.make_state(n, seed=40 + n, task=task).to("cuda")
    actors, nb, width = bm.task_dims(task)
    cfg = ops.make_task_cfg(num_bodies=nb)
    goal = torch.tensor([[1.5, 0.0]] if task == "kick" else [[2.0, 0.0]], device="cuda").repeat(n, 1)
    progress, reset2 = sg.make_bookkeeping(n, seed=41 + n, device="cuda", p_reset=0.1)
    progress[: min(n, 4)] = torch.tensor([899, 900, 897, 898], device="cuda")[: min(n, 4)]
    timeout2 = torch.empty(n, dtype=torch.long, device="cuda")
    obs2 = torch.empty(n, width, device="cuda"); rew2 = torch.empty(n, device="cuda")
    ops.post_physics_task(task, st.dof_state, st.rigid_body, st.root_states, st.net_contact, goal,
                          sg.make_initial_root_states(n, "cuda", task=task), reset2, progress, timeout2, cfg, obs2, rew2,
                          goal_angle=torch.full((n,), 1.5708, device="cuda") if task == "orient" else None,
                          ball_init=torch.tensor([[0.175, 0.0]], device="cuda").repeat(n, 1) if task == "kick" else None,
                          prev_lin_vel=torch.zeros(n, 3, device="cuda"), seed=5, step=3)
    torch.cuda.synchronize()
    eq = lambda a, b: bool(((a == b) | (a.isnan() & b.isnan())).all())      # noqa: E731
    assert eq(r["obs"], obs2) and eq(r["rew"], rew2) and torch.equal(r["reset"], reset2) and torch.equal(r["timeout"], timeout2)
    assert torch.equal(r["st"].dof_state, st.dof_state)


def test_epilogue_outputs_are_optional_and_checked():
    from bez_isaacgym_b200._lib import BezkError
    n = 257
    values = torch.zeros(n, device="cuda")
    kw = dict(gamma=0.99, scale_value=0.01, shift_value=0.0, value_bootstrap=True)
    r = _run("kick", n, 3, rollout_kw=kw, values=values, want_dones=False)
    assert r["dones"] is None and not bool(r["shaped"].isnan().any())
    r = _run("kick", n, 3, rollout_kw=None, values=None, want_shaped=False)
    assert r["shaped"] is None and int(r["dones"].max()) <= 1
    with pytest.raises(BezkError, match="BezkRolloutCfg"):
        _run("kick", n, 3, rollout_kw=None, values=values)
    with pytest.raises(BezkError, match="value_bootstrap needs values"):
        _run("kick", n, 3, rollout_kw=kw, values=None)


def test_reset_noise_is_keyed_by_global_env_id():
    """Two 'ranks' holding envs [0, 2048) and [2048, 4099) of one task, sharing one seed, reproduce the un-sharded step bit for
    bit when each passes its shard offset as env_base -- and do NOT when they pass 0 (the round-1 behaviour)."""
    ops = _ops()
    n, cut = 4099, 2048
    whole = _run("kick", n, 9, env_base=0, want_shaped=False)
    assert int((whole["progress"] == 0).sum()) > 100, "resets must fire"
    for lo, hi, base, same in ((cut, n, cut, True), (cut, n, 0, False)):
        m = hi - lo
        full = sg.make_state(n, seed=9).to("cuda")
        part = sg.SimState(full.root_states.view(n, -1)[lo:hi].reshape(-1, 13).contiguous(),
                           full.dof_state.view(n, -1)[lo:hi].reshape(-1, 2).contiguous(),
                           full.rigid_body.view(n, -1)[lo:hi].reshape(-1, 13).contiguous(),
                           full.net_contact.view(n, -1)[lo:hi].reshape(-1, 3).contiguous(), m, full.num_bodies)
        cfg = ops.make_task_cfg()
        goal, ball_init, _, _, _ = U.constants(m, "cuda")
        progress, reset = sg.make_bookkeeping(n, seed=10, device="cuda", p_reset=0.1)
        progress[:4] = torch.tensor([899, 900, 897, 898], device="cuda")
        progress, reset = progress[lo:hi].contiguous(), reset[lo:hi].contiguous()
        timeout = torch.empty(m, dtype=torch.long, device="cuda")
        obs = torch.empty(m, 54, device="cuda"); rew = torch.empty(m, device="cuda")
        ops.post_physics_rollout("kick", part.dof_state, part.rigid_body, part.root_states, part.net_contact, goal,
                                 sg.make_initial_root_states(m, "cuda"), reset, progress, timeout, cfg, obs, rew, ball_init=ball_init,
                                 prev_lin_vel=torch.zeros(m, 3, device="cuda"), seed=5, step=3, env_base=base)
        torch.cuda.synchronize()
        eq = torch.equal(part.dof_state, whole["st"].dof_state.view(n, -1)[lo:hi].reshape(-1, 2))
        assert eq == same


def test_explicit_reset_idx_is_keyed_by_global_env_id():
    ops = _ops()
    from oracle.philox_ref import reset_uniforms
    n, base = 300, 1000
    st = sg.make_state(n, seed=1).to("cuda")
    cfg = ops.make_task_cfg()
    progress = torch.ones(n, dtype=torch.long, device="cuda"); reset = torch.ones(n, dtype=torch.long, device="cuda")
    ids = torch.arange(n, device="cuda")
    ops.reset_idx_task("kick", ids, st.dof_state, st.root_states, sg.make_initial_root_states(n, "cuda"), None, progress, reset, cfg,
                       seed=77, step=2, env_base=base)
    from oracle import task_oracle as to
    u = torch.from_numpy(reset_uniforms(77, 2, n, first_env=base))
    _, _, default, lower, upper = U.constants(n)
    pos, vel = to.reset_idx_dof(default, lower, upper, u[:, 0:18], u[:, 18:36])
    got = st.dof_state.cpu().view(n, 18, 2)
    assert torch.equal(got[..., 0], pos) and torch.equal(got[..., 1], vel)
    assert int(progress.sum()) == 0 and int(reset.sum()) == 0


def test_policy_head_noise_is_keyed_by_global_env_id():
    """two ranks with the same seed must not draw identical exploration noise; the union of their draws equals the
    1-rank draw."""
    ops = _ops()
    n, cut = 1000, 384
    mu = torch.zeros(n, 18, device="cuda"); logstd = torch.zeros(18, device="cuda")
    whole = torch.empty(n, 18, device="cuda")
    ops.policy_head(mu, logstd, noise=None, seed=3, step=8, actions=whole)
    a = torch.empty(cut, 18, device="cuda"); b = torch.empty(n - cut, 18, device="cuda")
    ops.policy_head(mu[:cut].contiguous(), logstd, noise=None, seed=3, step=8, actions=a, env_base=0)
    ops.policy_head(mu[cut:].contiguous(), logstd, noise=None, seed=3, step=8, actions=b, env_base=cut)
    assert torch.equal(torch.cat((a, b)), whole)
    b0 = torch.empty(n - cut, 18, device="cuda")
    ops.policy_head(mu[cut:].contiguous(), logstd, noise=None, seed=3, step=8, actions=b0, env_base=0)
    assert not torch.equal(b0[:cut], whole[cut:2 * cut]) and torch.equal(b0[:cut], a[: min(cut, n - cut)])
    z = ops.normal_noise(3, 8, torch.empty(n - cut, 18, device="cuda"), env_base=cut)
    assert torch.equal(z, b)


def test_agent_writes_shaped_rewards_and_dones_from_the_step_kernel():
    """A2CAgent.play_steps: mb_rewards / dones slots come out of the step kernel; compared with the oracle's shaping of the
    env's own rewards, and the dones slot t+1 equals the reset mask of step t."""
    from bez_isaacgym_b200 import learner as L
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    from bez_isaacgym_b200.tasks import KickEnv
    from oracle import rl_games_oracle as rg
    n, T = 2048, 8
    log = []

    class Spy(KickEnv):
        def post_physics_step(self):
            super().post_physics_step()
            log.append((self.rew_buf.clone(), self.timeout_buf.clone(), self.reset_buf.clone()))

    env = Spy(bm.default_task_cfg(n), "cuda:0", 0, True, sim=SyntheticGym(n, device="cuda:0", seed=2))
    agent = L.A2CAgent(env, dict(horizon_length=T, minibatch_size=n * T, mixed_precision=False), seed=1)
    env.progress_buf.copy_(torch.randint(880, 900, (n,), device="cuda"))       # time-outs inside the horizon
    del log[:]
    agent.play_steps()
    assert len(log) == T
    vals = agent.experience_buffer.tensor_dict["values"]
    for t, (rew, timeout, reset) in enumerate(log):
        want = rg.shape_rewards(rew.cpu(), vals[t].cpu(), timeout.cpu(), 0.99)
        assert torch.equal(agent.mb_rewards[t].cpu(), want), f"mb_rewards[{t}]"
        nxt = agent.experience_buffer.tensor_dict["dones"][t + 1] if t + 1 < T else agent.dones
        assert torch.equal(nxt.cpu(), (reset != 0).to(torch.uint8).cpu())
    assert int(torch.stack([l[1] for l in log]).sum()) > 0
