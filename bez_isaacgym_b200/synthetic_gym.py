"""Synthetic Isaac Gym state tensors for BezKick (stands in for PhysX, which is out of scope).

Isaac Gym Preview is not installable offline, so benchmarks and tests feed state tensors with the
exact flat layouts ``gym.acquire_*_tensor`` hands out (reference ``bez_isaacgym/tasks/kick_env.py:143-196``,
SURVEY.md App. D):

    actor_root_state   (N*2, 13)   [pos3, quat_xyzw4, linvel3, angvel3]; actor 0 = bez, 1 = ball
    dof_state          (N*18, 2)   [pos, vel]
    rigid_body_state   (N*NB, 13)  same row format per body, NB = 22 (30 with cleats)
    net_contact_force  (N*NB, 3)

The distributions follow SURVEY.md §8(d) so that both reward branches, all five termination rules,
the contact-noise filter and the IMU clamps are exercised.
"""
import math
from dataclasses import dataclass

import torch

from . import bez_model as bm


@dataclass
class SimState:
    root_states: torch.Tensor
    dof_state: torch.Tensor
    rigid_body: torch.Tensor
    net_contact: torch.Tensor
    num_envs: int
    num_bodies: int

    def to(self, device):
        return SimState(self.root_states.to(device), self.dof_state.to(device), self.rigid_body.to(device),
                        self.net_contact.to(device), self.num_envs, self.num_bodies)

    def clone(self):
        return SimState(self.root_states.clone(), self.dof_state.clone(), self.rigid_body.clone(),
                        self.net_contact.clone(), self.num_envs, self.num_bodies)


READY_POSE = tuple(bm.default_task_cfg(1)["env"]["readyJointAngles"][n] for n in bm.DOF_NAMES)


def make_state(num_envs, seed=1234, device="cpu", cleats=False, filler=True, task="kick") -> SimState:
    """Seeded synthetic simulator state.  ``filler=False`` leaves the rigid-body / contact rows the
    task never reads at zero (cheaper to build at 1M envs; the kernels never touch them).
    ``task="walk"`` / ``"orient"``: one actor per env (root_states (N,13)) and no ball body; quaternions are drawn near
    upright so that the up-vector rules (up_proj < 0.7) and the win state both occur."""
    n = int(num_envs)
    actors, nb, _ = bm.task_dims(task, cleats)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f32 = dict(dtype=torch.float32, device=device)

    def randn(*shape):
        return torch.randn(*shape, generator=g, **f32)

    def rand(*shape):
        return torch.rand(*shape, generator=g, **f32)

    root = torch.zeros(n, 2, 13, **f32)
    root[:, 0, 0:3] = torch.tensor([0.0, 0.0, 0.34], **f32) + 0.03 * randn(n, 3)
    outlier = rand(n) < 0.01                                  # ~1 % strayed beyond 0.5 m
    root[:, 0, 0:2] += outlier.unsqueeze(1) * randn(n, 2)
    root[:, 0, 6] = 1.0
    root[:, 1, 0:3] = torch.tensor([0.175, 0.0, 0.1], **f32)
    root[:, 1, 0:2] += 0.25 * randn(n, 2)
    near_goal = rand(n) < 0.002                               # rule 4 (ball within 5 cm of the goal)
    root[:, 1, 0:2] = torch.where(near_goal.unsqueeze(1),
                                  torch.tensor([1.5, 0.0], **f32) + 0.03 * randn(n, 2), root[:, 1, 0:2])
    root[:, 1, 6] = 1.0
    root[:, 1, 7:10] = randn(n, 3)

    rb = randn(n, nb, 13) if filler else torch.zeros(n, nb, 13, **f32)
    q = randn(n, 4)
    rb[:, bm.IMU_BODY, 3:7] = q / q.norm(dim=1, keepdim=True)
    rb[:, bm.IMU_BODY, 7:10] = 0.3 * randn(n, 3)
    rb[:, bm.IMU_BODY, 10:13] = 3.0 * randn(n, 3)

    dof = torch.empty(n, 18, 2, **f32)
    dof[..., 0] = torch.tensor(READY_POSE, **f32) + 0.2 * randn(n, 18)
    dof[..., 1] = 3.3 * randn(n, 18)
    if actors == 1:
        # walk / orient: mostly upright robots (small tilt about x/y, any yaw); a few percent at rest in the ready pose near
        # the goal / goal heading so that the win state (4 conditions at once) fires; positions spread over +-0.4 m
        tilt = 0.15 * randn(n, 2)                               # up_proj = 1 - 2(qx^2+qy^2) ~ 0.91; ~4 % below the 0.7 fall rule
        yaw = 6.283185307179586 * rand(n)
        q = torch.stack((tilt[:, 0], tilt[:, 1], torch.sin(yaw / 2), torch.cos(yaw / 2)), 1)
        rb[:, bm.IMU_BODY, 3:7] = q / q.norm(dim=1, keepdim=True)
        root = root[:, 0:1].clone()
        root[:, 0, 0:2] = 0.1 * randn(n, 2)                     # ~1 % beyond orient's 0.3 m out-of-bound circle
        calm = rand(n) < 0.03
        rb[:, bm.IMU_BODY, 7:13] = torch.where(calm.unsqueeze(1), 0.02 * randn(n, 6), rb[:, bm.IMU_BODY, 7:13])
        dof[..., 0] = torch.where(calm.view(n, 1), torch.tensor(READY_POSE, **f32) + 0.01 * randn(n, 18), dof[..., 0])
        at_goal = calm & (rand(n) < 0.5)
        if task == "walk":      # just short of the goal ON the start->goal line (elsewhere within 5 cm the angle rule fires first)
            near = torch.tensor([2.0, 0.0], **f32) * (1.0 - 0.02 * rand(n, 1)) + 0.0005 * randn(n, 2)
        else:                   # orient: within 0.3 m of the start
            near = 0.05 * randn(n, 2)
        root[:, 0, 0:2] = torch.where(at_goal.unsqueeze(1), near, root[:, 0, 0:2])
        facing = 1.5708 + 0.02 * randn(n)                       # orient: yaw at the goal angle (signed angle_to_goal < 0.05)
        qf = torch.stack((torch.zeros(n, **f32), torch.zeros(n, **f32), torch.sin(facing / 2), torch.cos(facing / 2)), 1)
        rb[:, bm.IMU_BODY, 3:7] = torch.where(at_goal.unsqueeze(1), qf, rb[:, bm.IMU_BODY, 3:7])

    cf = randn(n, nb, 3) if filler else torch.zeros(n, nb, 3, **f32)

    def foot(k):
        f = torch.empty(n, k, 3, **f32)
        xy = 0.5 * randn(n, k, 2)
        xy = torch.where(rand(n, k, 2) < 0.5, torch.zeros_like(xy), xy)
        tiny = rand(n, k, 2) < 0.05                            # 5 % inside the (0, 0.01) noise band
        xy = torch.where(tiny, 0.01 * rand(n, k, 2), xy)
        f[..., 0:2] = xy
        f[..., 2] = torch.clamp(1.5 + 2.0 * randn(n, k), min=0.0)
        return f

    if cleats:
        cf[:, bm.LEFT_CLEATS[0]:bm.LEFT_CLEATS[1]] = foot(4)
        cf[:, bm.RIGHT_CLEATS[0]:bm.RIGHT_CLEATS[1]] = foot(4)
    else:
        cf[:, bm.LEFT_FOOT_BODY] = foot(1)[:, 0]
        cf[:, bm.RIGHT_FOOT_BODY] = foot(1)[:, 0]

    return SimState(root.reshape(n * actors, 13), dof.reshape(n * 18, 2), rb.reshape(n * nb, 13),
                    cf.reshape(n * nb, 3), n, nb)


def make_actions(num_envs, seed=4321, device="cpu"):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    a = torch.randn(num_envs, 18, generator=g, dtype=torch.float32, device=device)
    return a.clamp_(-1.0, 1.0)                                 # rl_games preprocess_actions clamp


def make_bookkeeping(num_envs, seed=99, device="cpu", max_episode_length=900, p_reset=1.0 / 300):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    progress = torch.randint(0, max_episode_length, (num_envs,), generator=g, device=device, dtype=torch.long)
    reset = (torch.rand(num_envs, generator=g, device=device) < p_reset).long()
    return progress, reset


def make_rollout(num_envs, horizon=32, seed=7, device="cpu", p_done=1.0 / 300):
    """Synthetic rl_games rollout tensors in the experience-buffer layout (T, N[,1])."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f32 = dict(dtype=torch.float32, device=device)
    rewards = 0.01 * torch.randn(horizon, num_envs, 1, generator=g, **f32)
    values = torch.randn(horizon, num_envs, 1, generator=g, **f32)
    dones = (torch.rand(horizon, num_envs, generator=g, device=device) < p_done).to(torch.uint8)
    last_values = torch.randn(num_envs, 1, generator=g, **f32)
    last_dones = (torch.rand(num_envs, generator=g, device=device) < p_done).to(torch.uint8)
    return rewards, values, dones, last_values, last_dones


def make_minibatch(m, seed=11, device="cpu"):
    """Synthetic PPO minibatch: model outputs and stored rollout quantities."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f32 = dict(dtype=torch.float32, device=device)
    mu = 0.5 * torch.randn(m, 18, generator=g, **f32)
    logstd = 0.1 * torch.randn(18, generator=g, **f32)
    old_mu = mu + 0.05 * torch.randn(m, 18, generator=g, **f32)
    old_logstd = logstd + 0.02 * torch.randn(18, generator=g, **f32)
    old_sigma = old_logstd.exp().expand(m, 18).contiguous()
    actions = old_mu + old_sigma * torch.randn(m, 18, generator=g, **f32)
    values = torch.randn(m, 1, generator=g, **f32)
    old_values = values + 0.3 * torch.randn(m, 1, generator=g, **f32)
    returns = torch.randn(m, 1, generator=g, **f32)
    adv = torch.randn(m, generator=g, **f32)
    old_neglogp = (0.5 * (((actions - old_mu) / old_sigma) ** 2).sum(-1)
                   + 0.5 * math.log(2.0 * math.pi) * 18 + old_logstd.sum())
    return dict(mu=mu, logstd=logstd, old_mu=old_mu, old_sigma=old_sigma, actions=actions, values=values,
                old_values=old_values, returns=returns, advantages=adv, old_neglogp=old_neglogp)


def make_constants(num_envs, device="cpu"):
    """goal (N,2), ball_init (N,2), default_dof_pos (N,18), lower (18,), upper (18,) with the BezKick defaults."""
    goal = torch.tensor([[1.5, 0.0]], device=device).repeat(num_envs, 1)
    ball_init = torch.tensor([[0.175, 0.0]], device=device).repeat(num_envs, 1)
    default = torch.tensor(READY_POSE, device=device).repeat(num_envs, 1)
    lower = torch.tensor(bm.DOF_LOWER, dtype=torch.float32, device=device)
    upper = torch.tensor(bm.DOF_UPPER, dtype=torch.float32, device=device)
    return goal, ball_init, default, lower, upper


def make_initial_root_states(num_envs, device="cpu", task="kick"):
    """(N*2,13) rows the reference restores on reset (kick_env.py:163-166); (N,13) for walk / orient (walk_env.py:146-149)."""
    if task != "kick":
        r = torch.zeros(num_envs, 13, device=device)
        r[:, 0:3] = torch.tensor([0.0, 0.0, 0.34], device=device)
        r[:, 6] = 1.0
        return r
    r = torch.zeros(num_envs, 2, 13, device=device)
    r[:, 0, 0:3] = torch.tensor([0.0, 0.0, 0.34], device=device)
    r[:, 0, 6] = 1.0
    r[:, 1, 0:3] = torch.tensor([0.175, 0.0, 0.1], device=device)
    r[:, 1, 6] = 1.0
    return r.view(num_envs * 2, 13)
