"""GPU: the product ``KickEnv`` (CUDA kernels through the C ABI) against the committed golden vectors, i.e. against
outputs of the reference's OWN code (tests/golden, made by oracle/make_golden.py from /root/reference).

* ``fn_*.npz``: the reference's jit functions on seeded / edge-case states -> ``bezk_compute_observations`` and
  ``bezk_compute_reward``.
* ``step_trace_n64.npz``: 8 ``VecTask.step`` calls of the UNMODIFIED reference ``KickEnv`` -> the drop-in
  ``bez_isaacgym_b200.tasks.KickEnv`` driven through the same public API with the same actions, simulator
  refreshes and reset draws.
Nothing here reads /root/reference."""
import os

import numpy as np
import pytest
import torch

from bez_isaacgym_b200 import bez_model as bm
from bez_isaacgym_b200 import synthetic_gym as sg
from tests import _util as U
from tests.test_oracle_pinning import _eq, _load, _state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cleats", [("fn_n1.npz", False), ("fn_n31.npz", False), ("fn_n64.npz", False),
                                         ("fn_n257.npz", False), ("fn_edges.npz", False), ("fn_cleats_n64.npz", True)])
def test_kernels_match_reference_function_goldens(name, cleats):
    from bez_isaacgym_b200 import ops
    g = _load(name)
    st = _state(g, cleats=cleats)
    n = st.num_envs
    goal, ball_init, default, _, _ = U.constants(n)
    cfg = ops.make_task_cfg(num_bodies=st.num_bodies, cleats=cleats)
    d = st.to("cuda")
    obs = torch.empty(n, 54, device="cuda")
    prev = g["in_prev_lin_vel"].cuda().contiguous()
    ops.compute_observations(d.dof_state, d.rigid_body, d.root_states, d.net_contact, goal.cuda(), ball_init.cuda(), cfg, obs,
                             prev_lin_vel=prev)
    rew = torch.empty(n, device="cuda"); reset = torch.empty(n, dtype=torch.long, device="cuda")
    ops.compute_reward(d.dof_state, d.rigid_body, d.root_states, goal.cuda(), ball_init.cuda(), g["in_reset"].cuda(),
                       g["in_progress"].cuda(), cfg, rew, reset)
    got, want = obs.cpu(), g["ref_obs"]
    assert _eq(got[:, 0:36], want[:, 0:36]) and _eq(got[:, 39:42], want[:, 39:42]) and _eq(got[:, 52:54], want[:, 52:54])
    if cleats:
        v = U.views(st, cleats=True)
        norms = torch.cat((torch.linalg.norm(v["left_c"], dim=-1), torch.linalg.norm(v["right_c"], dim=-1)), 1)
        tie = (norms - 1.0).abs() <= 2 * U.ulp(1.0)
        assert not bool(((got[:, 44:52] != want[:, 44:52]) & ~tie).any())
    else:
        assert _eq(got[:, 44:52], want[:, 44:52]), "foot pressure bits"
    assert _eq(d.net_contact.cpu(), g["ref_net_contact_after"]), "in-place contact filter"
    U.assert_close(got[:, 36:39], want[:, 36:39], scale=U.imu_term_scale(st, g["in_prev_lin_vel"]), what="imu lin_acc")
    U.assert_close(got[:, 42:44], want[:, 42:44], what="off_orn")
    band = U.reward_tie_band(st, goal, ball_init)
    assert not bool(((reset.cpu() != g["ref_reset"]) & ~band).any()), "reset mask outside the tie band"
    keep = ~band
    U.assert_close(rew.cpu()[keep], g["ref_rew"][keep], scale=U.reward_scale(st, goal, ball_init, default)[keep], what="reward")


@pytest.mark.parametrize("fusion", ["fused", "split"])
def test_kickenv_replays_unmodified_reference_trace(fusion):
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    from bez_isaacgym_b200.tasks import KickEnv
    g = _load("step_trace_n64.npz")
    n = g["in_actions"].shape[1]
    steps = g["in_actions"].shape[0]
    st = _state(g, "init_")
    counter = {"k": 0}

    def on_simulate(sim):
        k = counter["k"]
        sim.root_states.copy_(g["sim_root_states"][k].cuda()); sim.rigid_body.copy_(g["sim_rigid_body"][k].cuda())
        sim.net_contact.copy_(g["sim_net_contact"][k].cuda()); sim.dof_state.add_(g["sim_dof_drift"][k].cuda())
        counter["k"] += 1

    sim = SyntheticGym(n, device="cuda:0", state=st.to("cuda:0"), on_simulate=on_simulate)
    cfg = bm.default_task_cfg(n)
    cfg["seed"] = int(g["meta_seed"])
    env = KickEnv(cfg, "cuda:0", 0, True, sim=sim, fusion=fusion)
    assert _eq(env.dof_state.cpu(), g["init_dof_state_after_ctor"]), "constructor reset_idx(arange(N))"
    assert int(env.reset_buf.sum()) == 0 and int(env.progress_buf.sum()) == 0
    env.progress_buf.copy_(g["init_progress"].cuda())
    goal, ball_init, default, _, _ = U.constants(n)
    for k in range(steps):
        prev_before = torch.zeros(n, 3) if k == 0 else None           # aliasing from the second observation on
        obs_dict, rew, reset, extras = env.step(g["in_actions"][k].cuda())
        torch.cuda.synchronize()
        assert _eq(env.targets.cpu(), g["ref_targets"][k]), f"step {k}: PD targets"
        assert _eq(env.actions.cpu(), g["ref_actions_attr"][k]), f"step {k}: self.actions"
        assert _eq(extras["time_outs"].cpu(), g["ref_timeout"][k]) and _eq(env.progress_buf.cpu(), g["ref_progress"][k])
        assert _eq(env.dof_state.cpu(), g["ref_dof_state"][k]), f"step {k}: dof_state after masked reset"
        assert _eq(env.root_states.cpu(), g["ref_root_states"][k]) and _eq(env.net_contact.cpu(), g["ref_net_contact"][k])
        after = sg.SimState(g["ref_root_states"][k], g["ref_dof_state"][k], g["sim_rigid_body"][k], g["ref_net_contact"][k],
                            n, st.num_bodies)
        band = U.reward_tie_band(after, goal, ball_init)
        assert not bool(band.any()), "a tie-band env would fork the trajectories; regenerate the trace with another seed"
        assert _eq(reset.cpu(), g["ref_reset"][k]), f"step {k}: reset mask"
        got, want = obs_dict["obs"].cpu(), g["ref_obs"][k]
        assert _eq(got[:, 0:36], want[:, 0:36]) and _eq(got[:, 44:54], want[:, 44:54])
        U.assert_close(got[:, 36:42], want[:, 36:42], scale=U.imu_term_scale(after, prev_before), what=f"imu step {k}")
        U.assert_close(got[:, 42:44], want[:, 42:44], what=f"off_orn step {k}")
        U.assert_close(rew.cpu(), g["ref_rew"][k], scale=U.reward_scale(after, goal, ball_init, default), what=f"reward step {k}")
    assert env.randomize_buf.cpu().tolist() == [steps] * n
