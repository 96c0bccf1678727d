#!/usr/bin/env python
"""Generate ``tests/golden/*.npz`` from the REAL reference (oracle tooling; runs only where /root/reference exists).

    python -m oracle.make_golden

Everything stored under ``ref_*`` keys is the output of the reference's own code executed from /root/reference
(``tasks/kick_env.py`` jit functions, and the unmodified ``KickEnv.step`` over ``oracle.fake_isaacgym``); the
inputs are stored next to it so the fixtures are self-contained on the GPU box, where /root/reference does not
exist.  The only restated arithmetic on the reference side is ``isaacgym.torch_utils`` (not vendored).
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bez_isaacgym_b200 import bez_model as bm            # noqa: E402
from bez_isaacgym_b200 import synthetic_gym as sg        # noqa: E402
from oracle import reference_loader as rl                # noqa: E402
from oracle.philox_ref import reset_uniforms             # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TRACE_SEED = 1234


def _np(t):
    return t.detach().cpu().numpy().copy()


def state_arrays(st, prefix="in_"):
    return {prefix + "root_states": _np(st.root_states), prefix + "dof_state": _np(st.dof_state),
            prefix + "rigid_body": _np(st.rigid_body), prefix + "net_contact": _np(st.net_contact)}


def function_level(ref, st, prev, progress, reset_in, cleats=False):
    """Run the reference's own jit functions on the state; returns dict of reference outputs."""
    n = st.num_envs
    st = st.clone()
    root, rb = st.root_states.view(n, 2, 13), st.rigid_body.view(n, -1, 13)
    cf, dof = st.net_contact.view(n, -1, 3), st.dof_state.view(n, 18, 2)
    quat, lin, ang = rb[:, 1, 3:7], rb[:, 1, 7:10], rb[:, 1, 10:13]
    bez_pos, ball_pos, ball_vel = root[:, 0, 0:3], root[:, 1, 0:3], root[:, 1, 7:10]
    goal, ball_init, default, _, _ = sg.make_constants(n)
    gravity = torch.tensor([[0.0, 0.0, -1.0]]).repeat(n, 1)
    inv_rot = torch.tensor([[0.0, 0.0, 0.0, 1.0]]).repeat(n, 1)
    imu6, _ = ref.compute_imu(quat, lin, ang, prev, gravity, inv_rot, 2.0 * 9.81, 8.7266, 0.01667, n)
    orn = ref.compute_off_orn(bez_pos, quat, goal)
    if cleats:
        feet = ref.compute_feet_sensors_cleats(cf[:, 13:17, :], cf[:, 25:29, :], torch.tensor([[-1.0] * 8]).repeat(n, 1),
                                               torch.ones(n, 8))
    else:
        rows = [[1., -1., -1., -1.], [-1., -1., 1., -1.], [1., -1., 1., -1.], [-1., 1., -1., -1.], [-1., -1., -1., 1.],
                [-1., 1., -1., 1.], [1., 1., -1., -1.], [-1., -1., 1., 1.], [1., 1., 1., 1.], [-1.] * 4]
        args = [torch.tensor([[-1.0] * 4]).repeat(n, 1), torch.ones(1), torch.zeros(1), torch.zeros(3)] + \
               [torch.tensor(r) for r in rows]
        left = ref.compute_feet_sensors_no_cleats(cf[:, 12, :], *args)
        right = ref.compute_feet_sensors_no_cleats(cf[:, 20, :], *args)
        feet = torch.cat((left, right), 1)
    obs = ref.compute_bez_observations(dof[..., 0], dof[..., 1], imu6, orn, feet, ball_init)
    rew, reset = ref.compute_bez_reward(dof[..., 0], dof[..., 1], default, lin, ang, bez_pos, quat,
                                        torch.tensor([[0.0, 0.0, 1.0]]).repeat(n, 1), ball_pos, ball_vel, goal, ball_init,
                                        torch.tensor([0.0, 0.0]), reset_in, progress, feet, 900, n)
    return dict(ref_obs=_np(obs), ref_rew=_np(rew), ref_reset=_np(reset), ref_net_contact_after=_np(st.net_contact))


def edge_state(n=64, seed=21):
    """Contact-filter thresholds, NaN / inf forces, fz around 1 N, zero distance to goal, odd quaternions,
    termination thresholds and progress in {898, 899, 900, 901}."""
    import math
    st = sg.make_state(n, seed=seed)
    goal, ball_init, *_ = sg.make_constants(n)
    cf = st.net_contact.view(n, -1, 3)
    f32 = lambda x: float(torch.tensor(x, dtype=torch.float32))
    vals = [0.0, -0.0, f32(0.01), -f32(0.01), float(np.nextafter(np.float32(0.01), np.float32(1.0))), 0.005, -0.005, 2.0, -2.0, float("nan"),
            float("inf"), -float("inf")]
    fz = [0.5, f32(0.999), 1.0, float(np.nextafter(np.float32(1.0), np.float32(2.0))), 2.0, 0.005, float("nan"), f32(0.99), f32(1.01)]
    k = 0
    for e in range(n):
        for body in (bm.LEFT_FOOT_BODY, bm.RIGHT_FOOT_BODY):
            cf[e, body, 0] = vals[k % len(vals)]
            cf[e, body, 1] = vals[(k // len(vals) + k) % len(vals)]
            cf[e, body, 2] = fz[k % len(fz)]
            k += 1
    root = st.root_states.view(n, 2, 13)
    rb = st.rigid_body.view(n, -1, 13)
    root[0, 0, 0:2] = goal[0]                                        # ||goal - bez|| = 0 -> NaN heading
    rb[1, bm.IMU_BODY, 3:7] = torch.tensor([0.0, 0.0, 0.0, 1.0])     # identity
    rb[2, bm.IMU_BODY, 3:7] = torch.tensor([0.0, 0.0, 1.0, 0.0])     # 180 deg about z
    rb[3, bm.IMU_BODY, 3:7] = torch.tensor([0.3, -0.2, 0.1, 2.5])    # non-unit
    rb[4, bm.IMU_BODY, 10:13] = torch.tensor([100.0, -100.0, float("nan")])
    rb[5, bm.IMU_BODY, 7:10] = torch.tensor([1e3, -1e3, 0.0])        # lin-acc clamp
    root[6, 0, 2] = f32(0.275)                                       # rule 1: z == threshold -> no reset
    root[7, 0, 2] = float(np.nextafter(np.float32(0.275), np.float32(0.0)))                  # just below -> reset
    root[8, 1, 0:2] = goal[8]                                        # ball on the goal: division by zero, rule 4
    root[9, 0, 2] = float("nan")
    root[10, 1, 0:2] = torch.tensor([1.6, 1.0])                      # rule 3
    root[11, 0, 0:2] = torch.tensor([0.4, 0.4])                      # rule 2
    progress = torch.full((n,), 10, dtype=torch.long)
    progress[12:16] = torch.tensor([898, 899, 900, 901])
    root[16, 1, 0:2] = torch.tensor([1.5, 0.02]); progress[16] = 450     # rule 4 reward 50
    root[17, 1, 0:2] = torch.tensor([1.5, 0.02]); progress[17] = 900     # rules 4 + 5 -> 0
    reset_in = torch.zeros(n, dtype=torch.long)
    reset_in[18] = 1
    return st, torch.zeros(n, 3), progress, reset_in


def step_trace(n=64, steps=8, cleats=False):
    """The unmodified reference KickEnv stepped over the fake gym; reset draws injected from the Philox table."""
    nb = bm.BODIES_CLEATS if cleats else bm.BODIES_NO_CLEATS
    st0 = sg.make_state(n, seed=4000, cleats=cleats)
    init = state_arrays(st0, "init_")
    fresh = [sg.make_state(n, seed=4100 + k, cleats=cleats) for k in range(steps)]
    drift = [0.01 * torch.randn(n * 18, 2, generator=torch.Generator().manual_seed(50 + k)) for k in range(steps)]
    pending = []
    counter = {"step": 0, "sim": 0}

    def rand_source(shape):
        return pending.pop(0)

    def on_simulate(gym):
        k = counter["sim"]
        gym.root_states.copy_(fresh[k].root_states)
        gym.rigid_body.copy_(fresh[k].rigid_body)
        gym.net_contact.copy_(fresh[k].net_contact)
        gym.dof_state.add_(drift[k])
        counter["sim"] += 1

    def queue(env_ids, rng_step):
        u = torch.from_numpy(reset_uniforms(TRACE_SEED, rng_step, n))[env_ids]
        pending.extend([u[:, 0:18].clone(), u[:, 18:36].clone()])

    queue(torch.arange(n), 0)                               # KickEnv.__init__ -> reset_idx(arange(N)), kick_env.py:238
    env = rl.make_reference_env(st0, on_simulate=on_simulate, cleats=cleats, rand_source=rand_source)
    assert not pending
    out = dict(init)
    out["init_dof_state_after_ctor"] = _np(st0.dof_state)
    progress0 = torch.randint(0, 890, (n,), generator=torch.Generator().manual_seed(9))
    progress0[0:4] = torch.tensor([896, 897, 898, 899])
    env.progress_buf[:] = progress0
    out["init_progress"] = _np(progress0)
    actions = [sg.make_actions(n, seed=70 + k) * (4.5 if k == 2 else 1.0) for k in range(steps)]
    rows = {k: [] for k in ("obs", "rew", "reset", "timeout", "progress", "dof_state", "root_states", "net_contact",
                            "targets", "actions_attr")}
    for k in range(steps):
        env_ids = env.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(env_ids) > 0:
            queue(env_ids, k + 1)
        obs_dict, rew, reset, extras = env.step(actions[k].clone())
        assert not pending
        rows["obs"].append(_np(obs_dict["obs"])); rows["rew"].append(_np(rew)); rows["reset"].append(_np(reset))
        rows["timeout"].append(_np(extras["time_outs"])); rows["progress"].append(_np(env.progress_buf))
        rows["dof_state"].append(_np(env.dof_state)); rows["root_states"].append(_np(env.root_states))
        rows["net_contact"].append(_np(st0.net_contact)); rows["targets"].append(_np(env._fake_gym.targets))
        rows["actions_attr"].append(_np(env.actions))
    for key, v in rows.items():
        out["ref_" + key] = np.stack(v)
    out["in_actions"] = np.stack([_np(a) for a in actions])
    for name in ("root_states", "rigid_body", "net_contact"):
        out["sim_" + name] = np.stack([_np(getattr(f, name)) for f in fresh])
    out["sim_dof_drift"] = np.stack([_np(d) for d in drift])
    out["meta_seed"] = np.int64(TRACE_SEED)
    out["meta_num_bodies"] = np.int64(nb)
    out["meta_resets_per_step"] = np.array([int(r.sum()) for r in rows["reset"]])
    return out


# ----------------------------------------------------------------------------------------------- sibling tasks
_FEET_ROWS = [[1., -1., -1., -1.], [-1., -1., 1., -1.], [1., -1., 1., -1.], [-1., 1., -1., -1.], [-1., -1., -1., 1.],
              [-1., 1., -1., 1.], [1., 1., -1., -1.], [-1., -1., 1., 1.], [1., 1., 1., 1.], [-1.] * 4]


def sibling_constants(task, n):
    """goal (N,2), goal_angle (N,1), default (N,18) as WalkEnv / OrientEnv build them (walk_env.py:143, orient_env.py:145)."""
    goal = torch.tensor([[2.0, 0.0]]).repeat(n, 1)
    goal_angle = torch.tensor([[1.5708]]).repeat(n, 1)
    default = torch.tensor(sg.READY_POSE).repeat(n, 1)
    return goal, goal_angle, default


def function_level_sibling(task, ref, st, prev, progress, reset_in, goal, cleats=False):
    """The reference walk_env / orient_env jit functions on the state."""
    n = st.num_envs
    st = st.clone()
    root, rb = st.root_states.view(n, 1, 13), st.rigid_body.view(n, -1, 13)
    cf, dof = st.net_contact.view(n, -1, 3), st.dof_state.view(n, 18, 2)
    quat, lin, ang = rb[:, 1, 3:7], rb[:, 1, 7:10], rb[:, 1, 10:13]
    bez_pos = root[:, 0, 0:3]
    _, goal_angle, default = sibling_constants(task, n)
    gravity = torch.tensor([[0.0, 0.0, -1.0]]).repeat(n, 1)
    imu6, _ = ref.compute_imu(quat, lin, ang, prev, gravity, 2.0 * 9.81, 8.7266, 0.01667, n)
    heading = ref.compute_off_orn(bez_pos, quat, goal) if task == "walk" else ref.compute_off_angle(quat, goal_angle)
    if cleats:                          # walk_env.py:176-181, 388-416
        feet = ref.compute_feet_sensors_cleats(cf[:, 13:17, :], cf[:, 25:29, :], torch.tensor([[-1.0] * 8]).repeat(n, 1),
                                               torch.ones(n, 8))
    else:
        args = [torch.tensor([[-1.0] * 4]).repeat(n, 1), torch.ones(1), torch.zeros(1), torch.zeros(3)] + \
               [torch.tensor(r) for r in _FEET_ROWS]
        feet = torch.cat((ref.compute_feet_sensors_no_cleats(cf[:, 12, :], *args),
                          ref.compute_feet_sensors_no_cleats(cf[:, 20, :], *args)), 1)
    obs = ref.compute_bez_observations(dof[..., 0], dof[..., 1], imu6, heading, feet)
    up = torch.tensor([[0.0, 0.0, 1.0]]).repeat(n, 1)
    rew, reset = ref.compute_bez_reward(dof[..., 0], default, lin, ang, bez_pos, quat, up, goal if task == "walk" else goal_angle,
                                        reset_in, progress, feet, torch.tensor([0.0, 0.0]), 600, n, 0.01667, False)
    return dict(ref_obs=_np(obs), ref_rew=_np(rew), ref_reset=_np(reset), ref_net_contact_after=_np(st.net_contact),
                in_goal=_np(goal))


def sibling_edge_state(task, n=64, seed=31):
    """Fall threshold (up_proj around 0.7), win state on / just off each of its four conditions, out-of-bound rules,
    progress in {598..601}, zero distance to the goal."""
    st = sg.make_state(n, seed=seed, task=task)
    root, rb, dof = st.root_states.view(n, 1, 13), st.rigid_body.view(n, -1, 13), st.dof_state.view(n, 18, 2)
    goal = torch.tensor([[2.0, 0.0]]).repeat(n, 1)
    goal[20:30] = torch.tensor([0.3, -1.1])
    if task == "orient":
        root[:, 0, 0:2] *= 0.25           # keep the edge envs inside orient's 0.3 m out-of-bound circle unless a rule moves them
    ready = torch.tensor(sg.READY_POSE)
    import math

    def yaw_quat(yaw, tilt=0.0):
        return torch.tensor([math.sin(tilt / 2), 0.0, math.sin(yaw / 2) * math.cos(tilt / 2), math.cos(yaw / 2) * math.cos(tilt / 2)])
    # win state: at the goal (walk) / facing the goal angle (orient), ready pose, at rest
    for e in range(0, 6):
        # walk: 3 cm short of the goal on the start->goal line; orient: next to the start (its out-of-bound rule is 0.3 m)
        root[e, 0, 0:2] = goal[e] * (1.0 - 0.015) if task == "walk" else torch.tensor([0.05, 0.02])
        rb[e, bm.IMU_BODY, 3:7] = yaw_quat(1.5708 + 0.01)
        rb[e, bm.IMU_BODY, 7:13] = 0.01
        dof[e, :, 0] = ready + 0.005
    dof[1, :, 0] = ready + 0.05            # pos_reward = 0.05*sqrt(18) = 0.212 > 0.15: one condition short
    rb[2, bm.IMU_BODY, 10:13] = torch.tensor([0.2, 0.0, 0.0])          # angular velocity too high
    rb[3, bm.IMU_BODY, 7:10] = torch.tensor([0.0, 0.2, 0.0])           # linear velocity too high
    root[4, 0, 0:2] = goal[4] * 0.9 if task == "walk" else torch.tensor([0.05, 0.02])    # walk: 20 cm short -> not close
    rb[5, bm.IMU_BODY, 3:7] = yaw_quat(1.5708 - 0.2)                   # orient: signed angle 0.2 > 0.05
    # fall rule: tilt so that up_proj = cos(tilt) straddles 0.7
    for e, c in zip(range(6, 10), (0.69, 0.7, 0.71, -0.5)):
        rb[e, bm.IMU_BODY, 3:7] = yaw_quat(0.3, math.acos(c))
    root[10, 0, 0:2] = goal[10]                                        # zero distance: NaN heading
    root[11, 0, 0:2] = torch.tensor([3.0, 0.5])                        # walk: beyond the goal -> angle rule
    root[12, 0, 0:2] = torch.tensor([0.29, 0.0]); root[13, 0, 0:2] = torch.tensor([0.31, 0.0])     # orient: 0.3 m rule
    rb[14, bm.IMU_BODY, 3:7] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    rb[15, bm.IMU_BODY, 3:7] = torch.tensor([0.3, -0.2, 0.1, 2.5])     # non-unit
    progress = torch.full((n,), 10, dtype=torch.long)
    progress[16:20] = torch.tensor([598, 599, 600, 601])
    progress[0] = 300                                                  # win reward 500
    reset_in = torch.zeros(n, dtype=torch.long)
    reset_in[30] = 1
    return st, torch.zeros(n, 3), progress, reset_in, goal


def step_trace_sibling(task, n=64, steps=8):
    """The unmodified reference WalkEnv / OrientEnv stepped over the fake gym; DOF reset draws from the Philox table, goal
    draws (the reference uses element [0] of each (k,1) draw, walk_env.py:570-574) from a per-step table."""
    st0 = sg.make_state(n, seed=5000, task=task)
    init = state_arrays(st0, "init_")
    fresh = [sg.make_state(n, seed=5100 + k, task=task) for k in range(steps)]
    drift = [0.01 * torch.randn(n * 18, 2, generator=torch.Generator().manual_seed(60 + k)) for k in range(steps)]
    goal_u = torch.rand(steps + 1, 2, generator=torch.Generator().manual_seed(77))
    pending = []
    counter = {"sim": 0}

    def rand_source(shape):
        return pending.pop(0)

    def on_simulate(gym):
        k = counter["sim"]
        # .detach(): the legacy jit executor the reference selects (vec_task.py:170-172) conservatively marks tensors it
        # wrote in place as requiring grad; the simulator refresh is not part of any graph
        gym.root_states.detach().copy_(fresh[k].root_states); gym.rigid_body.detach().copy_(fresh[k].rigid_body)
        gym.net_contact.detach().copy_(fresh[k].net_contact); gym.dof_state.detach().add_(drift[k])
        counter["sim"] += 1

    def queue(env_ids, rng_step):
        u = torch.from_numpy(reset_uniforms(TRACE_SEED, rng_step, n))[env_ids]
        k = len(env_ids)
        pending.extend([u[:, 0:18].clone(), u[:, 18:36].clone(), goal_u[rng_step, 0].repeat(k, 1), goal_u[rng_step, 1].repeat(k, 1)])

    queue(torch.arange(n), 0)
    env = rl.make_reference_env(st0, on_simulate=on_simulate, rand_source=rand_source, task=task)
    assert not pending
    out = dict(init)
    out["init_dof_state_after_ctor"] = _np(st0.dof_state)
    out["init_goal_after_ctor"] = _np(env.goal)
    progress0 = torch.randint(0, 590, (n,), generator=torch.Generator().manual_seed(9))
    progress0[0:4] = torch.tensor([596, 597, 598, 599])
    env.progress_buf[:] = progress0
    out["init_progress"] = _np(progress0)
    actions = [sg.make_actions(n, seed=80 + k) * (4.5 if k == 2 else 1.0) for k in range(steps)]
    rows = {k: [] for k in ("obs", "rew", "reset", "timeout", "progress", "dof_state", "root_states", "net_contact",
                            "targets", "goal")}
    for k in range(steps):
        env_ids = env.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(env_ids) > 0:
            queue(env_ids, k + 1)
        obs_dict, rew, reset, extras = env.step(actions[k].clone())
        assert not pending
        rows["obs"].append(_np(obs_dict["obs"])); rows["rew"].append(_np(rew)); rows["reset"].append(_np(reset))
        rows["timeout"].append(_np(extras["time_outs"])); rows["progress"].append(_np(env.progress_buf))
        rows["dof_state"].append(_np(env.dof_state)); rows["root_states"].append(_np(env.root_states))
        rows["net_contact"].append(_np(st0.net_contact)); rows["targets"].append(_np(env._fake_gym.targets))
        rows["goal"].append(_np(env.goal))
    for key, v in rows.items():
        out["ref_" + key] = np.stack(v)
    out["in_actions"] = np.stack([_np(a) for a in actions])
    out["in_goal_uniforms"] = _np(goal_u)
    for name in ("root_states", "rigid_body", "net_contact"):
        out["sim_" + name] = np.stack([_np(getattr(f, name)) for f in fresh])
    out["sim_dof_drift"] = np.stack([_np(d) for d in drift])
    out["meta_seed"] = np.int64(TRACE_SEED)
    out["meta_resets_per_step"] = np.array([int(r.sum()) for r in rows["reset"]])
    return out


def siblings():
    for task in ("walk", "orient"):
        ref = rl.load_reference_task(task)
        for n in (31, 257):
            st = sg.make_state(n, seed=3000 + n, task=task)
            prev = 0.3 * torch.randn(n, 3, generator=torch.Generator().manual_seed(n))
            progress, reset_in = sg.make_bookkeeping(n, seed=n, p_reset=0.1, max_episode_length=600)
            progress[:4] = torch.tensor([598, 599, 600, 601])
            goal = torch.tensor([[2.0, 0.0]]).repeat(n, 1)
            goal[n // 2:] = 4.0 * torch.rand(n - n // 2, 2, generator=torch.Generator().manual_seed(5)) - 2.0
            d = dict(state_arrays(st), in_prev_lin_vel=_np(prev), in_progress=_np(progress), in_reset=_np(reset_in))
            d.update(function_level_sibling(task, ref, st, prev, progress, reset_in, goal))
            np.savez_compressed(os.path.join(OUT, f"fn_{task}_n{n}.npz"), **d)
        st, prev, progress, reset_in, goal = sibling_edge_state(task)
        d = dict(state_arrays(st), in_prev_lin_vel=_np(prev), in_progress=_np(progress), in_reset=_np(reset_in))
        d.update(function_level_sibling(task, ref, st, prev, progress, reset_in, goal))
        np.savez_compressed(os.path.join(OUT, f"fn_{task}_edges.npz"), **d)
        n = 64                                                       # cleats variant: 29 bodies, 4 + 4 cleat force rows
        st = sg.make_state(n, seed=3300, task=task, cleats=True)
        progress, reset_in = sg.make_bookkeeping(n, seed=3, max_episode_length=600)
        goal = torch.tensor([[2.0, 0.0]]).repeat(n, 1)
        d = dict(state_arrays(st), in_prev_lin_vel=np.zeros((n, 3), np.float32), in_progress=_np(progress), in_reset=_np(reset_in))
        d.update(function_level_sibling(task, ref, st, torch.zeros(n, 3), progress, reset_in, goal, cleats=True))
        np.savez_compressed(os.path.join(OUT, f"fn_{task}_cleats_n64.npz"), **d)
        print(task, "edge resets:", d["ref_reset"][:20], "rew:", np.round(d["ref_rew"][:8], 3))
        tr = step_trace_sibling(task)
        np.savez_compressed(os.path.join(OUT, f"step_trace_{task}_n64.npz"), **tr)
        print(task, "resets per step in the trace:", tr["meta_resets_per_step"])


def checkpoint_facts():
    """Facts of the shipped checkpoint that pin the rl_games state layout and update cadence."""
    import json
    import numpy
    path = os.path.join(rl.REFERENCE_ROOT, "bez_isaacgym", "results", "Bez_Kick", "Normal", "Bez_Kick_33.pth")
    allow = [(numpy._core.multiarray.scalar, "numpy.core.multiarray.scalar"), (numpy.dtype, "numpy.dtype")]
    allow += [getattr(numpy.dtypes, n) for n in dir(numpy.dtypes) if n.endswith("DType")]
    with torch.serialization.safe_globals(allow):
        ck = torch.load(path, map_location="cpu", weights_only=True)
    obs, val = ck["running_mean_std"], ck["reward_mean_std"]
    facts = {
        "keys": sorted(ck.keys()),
        "frame": int(ck["frame"]), "epoch": int(ck["epoch"]),
        "obs_rms": {k: {"shape": list(v.shape), "dtype": str(v.dtype)} for k, v in obs.items()},
        "val_rms": {k: {"shape": list(v.shape), "dtype": str(v.dtype)} for k, v in val.items()},
        "obs_count": float(obs["count"]), "val_count": float(val["count"]),
        "obs_running_mean": [float(x) for x in obs["running_mean"]],
        "obs_running_var": [float(x) for x in obs["running_var"]],
        "val_running_mean": float(val["running_mean"][0]), "val_running_var": float(val["running_var"][0]),
        "model_shapes": {k: list(v.shape) for k, v in ck["model"].items()},
        "sigma": [float(x) for x in ck["model"]["a2c_network.sigma"]],
        "adam_steps": int(ck["optimizer"]["state"][0]["step"]), "lr": float(ck["optimizer"]["param_groups"][0]["lr"]),
    }
    with open(os.path.join(OUT, "checkpoint_facts.json"), "w") as f:
        json.dump(facts, f, indent=1)


def jit_utils():
    """Outputs of the reference's OWN ``utils/torch_jit_utils.py`` helpers (executed from /root/reference over the isaacgym stubs):
    ``scale_transform`` / ``unscale_transform`` / ``saturate`` are defined in that file; ``quat_rotate`` / ``quat_rotate_inverse``
    reach it through ``from isaacgym.torch_utils import *`` (restated, parity unpinned) and are exercised through its
    ``compute_rot`` (``vel_loc = quat_rotate_inverse(torso_quat, velocity)``)."""
    import importlib
    rl.install_stubs()
    pkg_dir = os.path.join(rl.REFERENCE_ROOT, "bez_isaacgym")
    if pkg_dir not in sys.path:
        sys.path.insert(0, pkg_dir)
    tj = importlib.import_module("utils.torch_jit_utils")
    g = torch.Generator().manual_seed(77)
    n, dims = 257, 18
    x = 3.0 * torch.randn(n, dims, generator=g)
    x.view(-1)[::31] = float("nan"); x.view(-1)[5::37] = float("inf")
    lower = -1.0 - torch.rand(dims, generator=g) * 2
    upper = 0.5 + torch.rand(dims, generator=g) * 3
    q = torch.randn(n, 4, generator=g); q = q / q.norm(dim=1, keepdim=True)
    q[0] = torch.tensor([0.0, 0.0, 0.0, 1.0]); q[1] = torch.tensor([1.0, 0.0, 0.0, 0.0]); q[2] = 2.5 * q[2]       # identity, 180 deg, non-unit
    v = torch.randn(n, 3, generator=g); w = torch.randn(n, 3, generator=g)
    targets, pos = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    vel_loc, angvel_loc, *_ = tj.compute_rot(q, v, w, targets, pos)
    np.savez_compressed(os.path.join(OUT, "fn_jit_utils.npz"), in_x=_np(x), in_lower=_np(lower), in_upper=_np(upper), in_q=_np(q),
                        in_v=_np(v), in_w=_np(w), ref_scale=_np(tj.scale_transform(x, lower, upper)),
                        ref_unscale=_np(tj.unscale_transform(x, lower, upper)), ref_saturate=_np(tj.saturate(x, lower, upper)),
                        ref_vel_loc=_np(vel_loc), ref_angvel_loc=_np(angvel_loc), ref_quat_axis2=_np(tj.quat_axis(q, 2)))


def main():
    warnings.simplefilter("ignore")
    if not rl.reference_available():
        raise SystemExit("needs /root/reference")
    os.makedirs(OUT, exist_ok=True)
    if "--jit-utils-only" in sys.argv:
        jit_utils()
        return
    if "--siblings-only" in sys.argv:
        siblings()
        return
    ref = rl.load_reference_kick_env()
    for n in (1, 31, 64, 257):
        st = sg.make_state(n, seed=1000 + n)
        prev = 0.3 * torch.randn(n, 3, generator=torch.Generator().manual_seed(n))
        progress, reset_in = sg.make_bookkeeping(n, seed=n, p_reset=0.1)
        progress[: min(n, 4)] = torch.tensor([898, 899, 900, 901])[: min(n, 4)]
        d = dict(state_arrays(st), in_prev_lin_vel=_np(prev), in_progress=_np(progress), in_reset=_np(reset_in))
        d.update(function_level(ref, st, prev, progress, reset_in))
        np.savez_compressed(os.path.join(OUT, f"fn_n{n}.npz"), **d)
    st, prev, progress, reset_in = edge_state()
    d = dict(state_arrays(st), in_prev_lin_vel=_np(prev), in_progress=_np(progress), in_reset=_np(reset_in))
    d.update(function_level(ref, st, prev, progress, reset_in))
    np.savez_compressed(os.path.join(OUT, "fn_edges.npz"), **d)
    n = 64
    st = sg.make_state(n, seed=2064, cleats=True)
    progress, reset_in = sg.make_bookkeeping(n, seed=3)
    d = dict(state_arrays(st), in_prev_lin_vel=np.zeros((n, 3), np.float32), in_progress=_np(progress), in_reset=_np(reset_in))
    d.update(function_level(ref, st, torch.zeros(n, 3), progress, reset_in, cleats=True))
    np.savez_compressed(os.path.join(OUT, "fn_cleats_n64.npz"), **d)
    checkpoint_facts()
    tr = step_trace()
    np.savez_compressed(os.path.join(OUT, "step_trace_n64.npz"), **tr)
    print("resets per step in the trace:", tr["meta_resets_per_step"])
    siblings()
    jit_utils()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
