"""GPU parity of the WalkEnv / OrientEnv kernel variants through the C ABI:

* against the committed goldens = outputs of the reference's OWN ``tasks/walk_env.py`` / ``tasks/orient_env.py`` functions
  (``fn_{walk,orient}_*.npz``) and 8-step traces of the UNMODIFIED reference ``WalkEnv`` / ``OrientEnv`` stepped over the fake
  gym (``step_trace_{walk,orient}_n64.npz``; generator: ``oracle/make_golden.py``);
* against ``oracle.task_oracle`` (pinned bit-exactly to those goldens by tests/test_oracle_pinning.py) on seeded random states.

Tolerances as for BezKick: fp32 rtol 1e-5 / atol 1e-6 (condition-aware for the IMU mat-vec and the reward sum); copies, foot
bits, bookkeeping, reset write-backs, goal redraws bit-exact; reset masks bit-exact outside the documented tie band around
the computed thresholds (n_goal 0.05, pos 0.15, velocities 0.1, up_proj 0.7, out-of-bound 0.3 / 1.5708, signed angle 0.05).
Nothing here reads /root/reference."""
import numpy as np
import pytest
import torch

from bez_isaacgym_b200 import bez_model as bm
from bez_isaacgym_b200 import synthetic_gym as sg
from tests import _util as U
from tests.test_oracle_pinning import _eq, _load, _sibling_state, sibling_oracle, sibling_views

pytestmark = pytest.mark.gpu


def _band(task, st, goal):
    """Envs whose mask hangs on a computed quantity within a relative 2e-6 (norm-fed) / absolute 2e-6 (angle-fed) band."""
    from oracle import task_oracle as to
    from oracle.isaacgym_torch_utils import get_basis_vector, get_euler_xyz, normalize_angle
    v = sibling_views(st)
    n = st.num_envs
    default = torch.tensor(sg.READY_POSE).repeat(n, 1)
    up = get_basis_vector(v["quat"], torch.tensor([[0.0, 0.0, 1.0]]).repeat(n, 1))[:, 2]
    pos = torch.linalg.norm(default - v["dof_pos"], dim=1)
    vl, va = torch.linalg.norm(v["lin"], dim=1), torch.linalg.norm(v["ang"], dim=1)
    near = lambda x, thr, tol=2e-6: (x - thr).abs() <= tol * max(abs(thr), 1.0)      # noqa: E731
    band = near(up, 0.7) | near(pos, 0.15) | near(va, 0.1) | near(vl, 0.1)
    if task == "walk":
        d = goal - v["bez_pos"][:, 0:2]
        n_goal = torch.linalg.norm(d, dim=1)
        u = d / n_goal.unsqueeze(1)
        ui = goal / torch.linalg.norm(goal, dim=1, keepdim=True)
        ang = (torch.atan2(ui[:, 1], ui[:, 0]) - torch.atan2(u[:, 1], u[:, 0])).abs()
        band |= near(n_goal, 0.05) | near(ang, 1.5708)
    else:
        _, _, yaw = get_euler_xyz(v["quat"])
        a = 1.5708 - normalize_angle(yaw)
        band |= near(a, 0.05) | near(torch.linalg.norm(v["bez_pos"][:, 0:2], dim=1), 0.3)
    return band


def _reward_scale(st):
    v = sibling_views(st)
    default = torch.tensor(sg.READY_POSE).repeat(st.num_envs, 1)
    return (10.0 * torch.linalg.norm(v["lin"][:, 0:2], dim=1) + torch.linalg.norm(torch.cat((v["lin"], v["ang"]), 1), dim=1)
            + torch.linalg.norm(default - v["dof_pos"], dim=1) + 4.0)


def _run_kernels(task, st, prev, goal, progress, reset_in):
    from bez_isaacgym_b200 import ops
    n = st.num_envs
    d = st.to("cuda")
    cfg = ops.make_task_cfg(num_bodies=st.num_bodies, max_episode_length=600, cleats=st.num_bodies > 21)
    obs = torch.full((n, 52), float("nan"), device="cuda")
    rew = torch.empty(n, device="cuda")
    g = goal.cuda().contiguous()
    ga = torch.full((n,), 1.5708, device="cuda")
    p = prev.cuda().contiguous()
    # the observation kernel (parts = 2) then the reward kernel (parts = 4); the fused step (7) is covered by the traces below
    reset = reset_in.cuda().clone()
    ops.post_physics_task(task, d.dof_state, d.rigid_body, d.root_states, d.net_contact, g, None, None, None, None, cfg, obs, None,
                          goal_angle=ga, prev_lin_vel=p, parts=2)
    ops.post_physics_task(task, d.dof_state, d.rigid_body, d.root_states, None, g, None, reset, progress.cuda(), None, cfg, None, rew,
                          goal_angle=ga, parts=4)
    return obs.cpu(), rew.cpu(), reset.cpu(), d.net_contact.cpu(), p.cpu()


def _check(task, st, prev, goal, got, want_obs, want_rew, want_reset, want_cf):
    obs, rew, reset, cf, prev_after = got
    assert _eq(obs[:, 0:36], want_obs[:, 0:36]) and _eq(obs[:, 39:42], want_obs[:, 39:42]), "dof / angular velocity columns"
    if st.num_bodies > 21:               # cleats: bits hang on ||f|| > 1 (norm-fed: 2-ulp tie band)
        v = sibling_views(st)
        norms = torch.cat((torch.linalg.norm(v["left_c"], dim=-1), torch.linalg.norm(v["right_c"], dim=-1)), 1)
        tie = (norms - 1.0).abs() <= 2 * U.ulp(1.0)
        assert not bool(((obs[:, 44:52] != want_obs[:, 44:52]) & ~tie).any()), "cleat bits"
    else:
        assert _eq(obs[:, 44:52], want_obs[:, 44:52]), "foot pressure bits"
    assert _eq(cf, want_cf), "in-place contact filter"
    n = st.num_envs
    lin = st.rigid_body.view(n, -1, 13)[:, bm.IMU_BODY, 7:10]
    a = (lin - prev) / 0.01667 - torch.tensor([0.0, 0.0, -1.0])
    U.assert_close(obs[:, 36:39], want_obs[:, 36:39], scale=a.abs().sum(1, keepdim=True), what="imu lin_acc")
    assert _eq(prev_after, lin), "prev_lin_vel <- current velocity"
    U.assert_close(obs[:, 42:44], want_obs[:, 42:44], rtol=2e-5, atol=2e-6, what="heading columns")
    band = _band(task, st, goal)
    assert not bool(((reset != want_reset) & ~band).any()), "reset mask outside the tie band"
    keep = ~band & (reset == want_reset)
    U.assert_close(rew[keep], want_rew[keep], scale=_reward_scale(st)[keep], rtol=2e-5, what="reward")
    return int(band.sum())


@pytest.mark.parametrize("task", ["walk", "orient"])
@pytest.mark.parametrize("which", ["n31", "n257", "edges", "cleats_n64"])
def test_sibling_kernels_match_reference_function_goldens(task, which):
    g = _load(f"fn_{task}_{which}.npz")
    st = _sibling_state(g)
    got = _run_kernels(task, st, g["in_prev_lin_vel"], g["in_goal"], g["in_progress"], g["in_reset"])
    _check(task, st, g["in_prev_lin_vel"], g["in_goal"], got, g["ref_obs"], g["ref_rew"], g["ref_reset"], g["ref_net_contact_after"])


@pytest.mark.parametrize("task", ["walk", "orient"])
@pytest.mark.parametrize("n", [1, 33, 4096, 70001])
def test_sibling_kernels_match_oracle(task, n):
    st = sg.make_state(n, seed=900 + n, task=task)
    prev = 0.2 * torch.randn(n, 3, generator=torch.Generator().manual_seed(n))
    progress, reset_in = sg.make_bookkeeping(n, seed=n, max_episode_length=602)
    goal = torch.tensor([[2.0, 0.0]]).repeat(n, 1)
    goal[n // 2:] = 4.0 * torch.rand(n - n // 2, 2, generator=torch.Generator().manual_seed(3)) - 2.0
    want = sibling_oracle(task, st, prev, goal, progress, reset_in)
    got = _run_kernels(task, st, prev, goal, progress, reset_in)
    in_band = _check(task, st, prev, goal, got, want[0], want[1], want[2], want[3])
    assert in_band <= max(2, n // 20000), f"{in_band} envs inside the tie band"
    if n >= 4096:
        assert int(want[2].sum()) > n // 100 and int((want[1] > 100).sum()) > 0       # terminations and win states present


@pytest.mark.parametrize("task,cls", [("walk", "WalkEnv"), ("orient", "OrientEnv")])
def test_sibling_env_replays_unmodified_reference_trace(task, cls):
    from bez_isaacgym_b200 import ops, tasks
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    g = _load(f"step_trace_{task}_n64.npz")
    steps, n = g["in_actions"].shape[0], g["in_actions"].shape[1]
    st = _sibling_state(g, "init_")
    counter = {"k": 0}

    def on_simulate(sim):
        k = counter["k"]
        sim.root_states.copy_(g["sim_root_states"][k].cuda()); sim.rigid_body.copy_(g["sim_rigid_body"][k].cuda())
        sim.net_contact.copy_(g["sim_net_contact"][k].cuda()); sim.dof_state.add_(g["sim_dof_drift"][k].cuda())
        counter["k"] += 1

    sim = SyntheticGym(n, device="cuda:0", state=st.to("cuda:0"), on_simulate=on_simulate, task=task)
    cfg = bm.default_task_cfg(n, task=task)
    cfg["seed"] = int(g["meta_seed"])
    env = getattr(tasks, cls)(cfg, "cuda:0", 0, True, sim=sim)
    assert env.num_obs == 52 and env.max_episode_length == 600 and env.obs_buf.shape == (n, 52)
    assert _eq(env.dof_state.cpu(), g["init_dof_state_after_ctor"]), "constructor reset_idx(arange(N))"
    # the goal draws: the golden trace injected its own table; hand the same per-step uniforms to the kernels
    gu = g["in_goal_uniforms"].cuda()
    env.goal.copy_((4.0 * gu[0] + -2.0).expand(n, 2))
    assert _eq(env.goal.cpu(), g["init_goal_after_ctor"])
    env.progress_buf.copy_(g["init_progress"].cuda())
    mid = env._post_fixed["mid_task"]
    for k in range(steps):
        mid[5] = ops._p(gu[k + 1].contiguous(), torch.float32, "goal_uniforms", 2)
        obs_dict, rew, reset, extras = env.step(g["in_actions"][k].cuda())
        torch.cuda.synchronize()
        assert _eq(env.targets.cpu(), g["ref_targets"][k]), f"step {k}: PD targets"
        assert _eq(extras["time_outs"].cpu(), g["ref_timeout"][k]) and _eq(env.progress_buf.cpu(), g["ref_progress"][k])
        assert _eq(env.dof_state.cpu(), g["ref_dof_state"][k]), f"step {k}: dof_state after masked reset"
        assert _eq(env.root_states.cpu(), g["ref_root_states"][k]) and _eq(env.net_contact.cpu(), g["ref_net_contact"][k])
        assert _eq(env.goal.cpu(), g["ref_goal"][k]), f"step {k}: goal redraw (first draw of the batch for every reset env)"
        after = sg.SimState(g["ref_root_states"][k], g["ref_dof_state"][k], g["sim_rigid_body"][k], g["ref_net_contact"][k],
                            n, st.num_bodies)
        band = _band(task, after, g["ref_goal"][k])
        assert not bool(band.any()), "a tie-band env would fork the trajectories; regenerate the trace with another seed"
        assert _eq(reset.cpu(), g["ref_reset"][k]), f"step {k}: reset mask"
        got, want = obs_dict["obs"].cpu(), g["ref_obs"][k]
        assert _eq(got[:, 0:36], want[:, 0:36]) and _eq(got[:, 44:52], want[:, 44:52])
        prev = torch.zeros(n, 3) if k == 0 else None
        lin = after.rigid_body.view(n, -1, 13)[:, bm.IMU_BODY, 7:10]
        a = ((lin - prev) / 0.01667 if prev is not None else torch.zeros(n, 3)) - torch.tensor([0.0, 0.0, -1.0])
        U.assert_close(got[:, 36:42], want[:, 36:42], scale=a.abs().sum(1, keepdim=True), what=f"imu step {k}")
        U.assert_close(got[:, 42:44], want[:, 42:44], rtol=2e-5, atol=2e-6, what=f"heading step {k}")
        U.assert_close(rew.cpu(), g["ref_rew"][k], scale=_reward_scale(after), rtol=2e-5, what=f"reward step {k}")
    assert int(g["meta_resets_per_step"].sum()) > 0


@pytest.mark.parametrize("task,cls", [("walk", "WalkEnv"), ("orient", "OrientEnv")])
def test_sibling_env_philox_goal_is_one_draw_per_step(task, cls):
    """With the Philox path every env that resets in a step receives the SAME goal, equal to bezk_goal_uniforms(seed, step)."""
    from bez_isaacgym_b200 import ops, tasks
    n = 4096
    cfg = bm.default_task_cfg(n, task=task)
    cfg["seed"] = 5
    env = getattr(tasks, cls)(cfg, "cuda:0", 0, True)
    ctor_goal = 4.0 * ops.goal_uniforms(5, 0) + -2.0
    assert torch.equal(env.goal, ctor_goal.expand(n, 2))
    act = torch.zeros(n, 18, device="cuda")
    env.step(act)
    flagged = env.reset_buf.clone().bool()
    assert 0 < int(flagged.sum()) < n
    goal_before = env.goal.clone()
    env.step(act)
    want = 4.0 * ops.goal_uniforms(5, env._rng_step) + -2.0
    assert torch.equal(env.goal[flagged], want.expand(int(flagged.sum()), 2))
    assert torch.equal(env.goal[~flagged], goal_before[~flagged])
    assert int(env.progress_buf[flagged].max()) == 0


def test_sibling_rejects_unsupported_parts_and_serves_the_host_pipeline():
    from bez_isaacgym_b200 import ops, tasks
    from bez_isaacgym_b200._lib import BezkError
    n = 8
    st = sg.make_state(n, task="walk").to("cuda")
    cfg = ops.make_task_cfg(num_bodies=21)
    goal = torch.zeros(n, 2, device="cuda")
    with pytest.raises(BezkError):
        ops.post_physics_task("walk", st.dof_state, st.rigid_body, st.root_states, st.net_contact, goal, None,
                              torch.zeros(n, dtype=torch.long, device="cuda"), torch.zeros(n, dtype=torch.long, device="cuda"),
                              torch.zeros(n, dtype=torch.long, device="cuda"), cfg, torch.empty(n, 52, device="cuda"),
                              torch.empty(n, device="cuda"), parts=5)
    with pytest.raises(BezkError):          # orient without goal_angle
        ops.post_physics_task("orient", st.dof_state, st.rigid_body, st.root_states, st.net_contact, goal, None, None, None, None,
                              cfg, torch.empty(n, 52, device="cuda"), None, parts=2)
    # sim_device="cpu" (the reference's device selection, tasks/base/vec_task.py:51-98) serves every task: host pipeline
    env = tasks.WalkEnv(bm.default_task_cfg(n, task="walk", use_gpu_pipeline=False, rl_device="cpu"), "cpu", 0, True)
    o, r, d, e = env.step(torch.zeros(n, 18))
    assert o["obs"].shape == (n, 52) and o["obs"].device.type == "cpu" and r.device.type == "cpu" and d.dtype == torch.long
