"""torch-CPU restatement of the BezKick task-side step (oracle; TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows (paths relative to
/root/reference/bez_isaacgym).  The restatement uses the same ATen ops in the same association
order as the reference so that on CPU it is *bit-identical* to the reference's own
``@torch.jit.script`` functions; ``tests/test_oracle_pinning.py`` asserts that (where
/root/reference exists) and ``tests/golden/*.npz`` carries reference outputs to the GPU box.

PINNED: against the reference's own functions (imported, ``oracle/reference_loader.py``) and the
golden vectors.  The helpers from ``isaacgym.torch_utils`` are restated (parity unpinned, see
``oracle/isaacgym_torch_utils.py``).
"""
import math
from typing import Optional, Tuple

import torch
from torch import Tensor

from oracle.isaacgym_torch_utils import get_basis_vector, get_euler_xyz, normalize_angle, tensor_clamp

IMU_MAX_LIN_ACC = 2.0 * 9.81     # tasks/kick_env.py:100
IMU_MAX_ANG_VEL = 8.7266         # tasks/kick_env.py:99


# --------------------------------------------------------------------------- pre-physics
def pre_physics(actions: Tensor, default_dof_pos: Tensor, lower: Tensor, upper: Tensor,
                clip_actions: float) -> Tuple[Tensor, Tensor]:
    """tasks/base/vec_task.py:317 (clip) + tasks/kick_env.py:410-419 (head zeroing, PD targets).

    Returns (stored_actions, targets)."""
    a = torch.clamp(actions, -clip_actions, clip_actions).clone()
    a[..., 0:2] = 0.0
    return a, tensor_clamp(a + default_dof_pos, lower, upper)


# --------------------------------------------------------------------------- observation parts
def wxyz_matrix(q: Tensor) -> Tensor:
    """tasks/kick_env.py:857-885: rotation matrix of a REAL-FIRST quaternion, no normalisation assumed."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    rows = (1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j))
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


def imu(quat_xyzw: Tensor, lin_vel: Tensor, ang_vel: Tensor, prev_lin_vel: Tensor,
        gravity_vec: Tensor, dt: float) -> Tuple[Tensor, Tensor]:
    """tasks/kick_env.py:918-930.  The xyzw quaternion is fed UNCHANGED to the real-first matrix
    formula (bug-for-bug).  Second return value is the input ``lin_vel`` object itself (aliasing,
    kick_env.py:930)."""
    n = quat_xyzw.shape[0]
    acc = torch.sub(lin_vel, prev_lin_vel)
    acc = torch.div(acc, dt)
    acc = torch.sub(acc, gravity_vec)
    acc_t = torch.matmul(wxyz_matrix(quat_xyzw), acc.reshape((n, -1, 1))).reshape((n, -1))
    out = torch.cat([torch.clamp(acc_t, -IMU_MAX_LIN_ACC, IMU_MAX_LIN_ACC),
                     torch.clamp(ang_vel, -IMU_MAX_ANG_VEL, IMU_MAX_ANG_VEL)], 1)
    return out, lin_vel


def off_orn(bez_pos: Tensor, quat_xyzw: Tensor, goal: Tensor) -> Tensor:
    """tasks/kick_env.py:941-960: (|sin|, -cos) of the heading relative to the robot->goal direction."""
    d = torch.sub(goal, bez_pos[..., 0:2])
    u = torch.div(d, torch.linalg.norm(d, dim=1).reshape(-1, 1))
    _, _, yaw = get_euler_xyz(quat_xyzw[..., 0:4])
    h = torch.cat((torch.cos(yaw).reshape(-1, 1), torch.sin(yaw).reshape(-1, 1)), dim=-1)
    c = torch.sum(h * u, dim=-1)
    u3 = torch.nn.functional.pad(u, (0, 1, 0, 0), value=0.0)
    h3 = torch.nn.functional.pad(h, (0, 1, 0, 0), value=0.0)
    s = torch.linalg.norm(torch.cross(u3, h3, dim=1), dim=1)
    return torch.cat((s.reshape(-1, 1), -c.reshape(-1, 1)), dim=1)


_FOOT_ROWS = {  # value of (y + sensor) -> sensor pattern, tasks/kick_env.py:518-536,1010-1036
    1.0: (1., -1., -1., -1.), 2.0: (-1., -1., 1., -1.), 3.0: (1., -1., 1., -1.),
    5.0: (-1., 1., -1., -1.), 6.0: (-1., -1., -1., 1.), 7.0: (-1., 1., -1., 1.),
    9.0: (1., 1., -1., -1.), 10.0: (-1., -1., 1., 1.), 11.0: (1., 1., 1., 1.),
}


def feet_no_cleats(force: Tensor) -> Tensor:
    """tasks/kick_env.py:987-1038 for ONE foot.  ``force`` (N,3) is filtered IN PLACE (the reference
    writes the noise filter back into the simulator's contact buffer, :987-990)."""
    force[..., 0:3] = torch.where(torch.abs(force[..., 0:3]) > 0.01, force[..., 0:3],
                                  force.new_zeros(3))
    one, zero = force.new_ones(1), force.new_zeros(1)
    x = torch.where(torch.abs(force[..., 0]) > 0.0, one, zero)
    x = torch.where(force[..., 0] == 0, 2.0 * one, x)
    y = torch.where(torch.abs(force[..., 1]) > 0.0, one, 3.0 * one)
    y = torch.where(force[..., 1] == 0, 3.0 * one, y)
    sensor = torch.where(x == 1.0, zero, 4.0 * one)
    sensor = torch.where(x == 2.0, 8.0 * one, sensor)
    case = torch.add(y, sensor).reshape(-1, 1)
    dev = force.device
    out = torch.tensor([[-1.0] * 4], device=dev).repeat(force.shape[0], 1)
    for code, row in _FOOT_ROWS.items():
        out = torch.where(case == code, torch.tensor(row, device=dev), out)
    return torch.where(force[..., 2].reshape(-1, 1) < 1, torch.tensor([-1.0] * 4, device=dev), out)


def feet_cleats(left: Tensor, right: Tensor) -> Tensor:
    """tasks/kick_env.py:1053-1061: 8 bits from the 2-norm of the 4+4 cleat forces (> 1 N)."""
    pts = torch.cat((torch.linalg.norm(left, dim=-1), torch.linalg.norm(right, dim=-1)), 1)
    return torch.where(pts > 1.0, torch.ones_like(pts), -torch.ones_like(pts))


def observations(dof_pos, dof_vel, imu6, orn2, feet8, ball_init) -> Tensor:
    """tasks/kick_env.py:1409-1415."""
    return torch.cat((dof_pos, dof_vel, imu6, orn2, feet8, ball_init), dim=-1)


# --------------------------------------------------------------------------- reward / termination
def reward(dof_pos, default_dof_pos, imu_lin, imu_ang, bez_pos, ball_pos, ball_vel, goal, ball_init,
           bez_init_xy, reset_buf, progress_buf, max_episode_length: int) -> Tuple[Tensor, Tensor]:
    """tasks/kick_env.py:1224-1395 with the dead computations (up_proj, euler angles, feet and
    dof-velocity terms, :1247-1249,1266,1271-1280) removed."""
    d_ball = torch.sub(ball_pos[..., 0:2], bez_pos[..., 0:2])
    u_bb = torch.div(d_ball, torch.linalg.norm(d_ball, dim=1).reshape(-1, 1))
    vel_fwd = torch.sum(torch.mul(u_bb, imu_lin[..., 0:2]), dim=1)

    d_goal = torch.sub(goal, ball_pos[..., 0:2])
    n_goal = torch.linalg.norm(d_goal, dim=1).reshape(-1, 1)
    u_bg = torch.div(d_goal, n_goal)
    ball_fwd = torch.sum(torch.mul(u_bg, ball_vel[..., 0:2]), dim=1)

    d_init = torch.sub(goal, ball_init)
    u_ig = torch.div(d_init, torch.linalg.norm(d_init, dim=1).reshape(-1, 1))
    ang_now = torch.atan2(u_bg[..., 1], u_bg[..., 0])
    ang_init = torch.atan2(u_ig[..., 1], u_ig[..., 0])
    angle_diff = torch.abs(ang_init - ang_now).reshape(-1)

    vel_r = torch.mul(torch.linalg.norm(torch.cat((imu_lin, imu_ang), dim=1), dim=1), 0.05)
    pos_r = torch.mul(torch.linalg.norm(default_dof_pos - dof_pos, dim=1), 0.05)
    height = torch.mul(torch.abs(0.325 - bez_pos[..., 2]), 1)
    kicked = torch.linalg.norm(torch.sub(ball_pos[..., 0:2], ball_init), dim=1)

    far = torch.sub(torch.mul(ball_fwd, 0.1), torch.add(height, torch.add(vel_r, pos_r)))
    near = torch.add(torch.mul(ball_fwd, 0.1), torch.sub(torch.mul(vel_fwd, 0.05), height))
    rew = torch.where(kicked > 0.3, far, near)

    ones = torch.ones_like(reset_buf)
    minus = torch.ones_like(rew) * -1.0
    fell = bez_pos[..., 2] < 0.275                                              # rule 1 (:1331)
    reset = torch.where(fell, ones, reset_buf)
    rew = torch.where(fell, minus, rew)
    strayed = torch.linalg.norm(torch.sub(bez_pos[..., 0:2], bez_init_xy), dim=1).reshape(-1) > 0.5
    reset = torch.where(strayed, ones, reset)                                   # rule 2 (:1340-1349)
    rew = torch.where(strayed, minus, rew)
    wide = angle_diff > 1.5708                                                  # rule 3 (:1370-1377)
    reset = torch.where(wide, ones, reset)
    rew = torch.where(wide, minus, rew)
    scored = n_goal.reshape(-1) < 0.05                                          # rule 4 (:1380-1385)
    reset = torch.where(scored, ones, reset)
    rew = torch.where(scored,
                      torch.ones_like(rew) * (100.0 - 100.0 * (progress_buf / max_episode_length)), rew)
    over = progress_buf >= max_episode_length                                   # rule 5 (:1388-1391)
    reset = torch.where(over, ones, reset)
    rew = torch.where(over, torch.zeros_like(rew), rew)
    return rew, reset


# --------------------------------------------------------------------------- sibling tasks
def off_angle(quat_xyzw: Tensor, goal_angle: Tensor) -> Tensor:
    """tasks/orient_env.py:720-733 (compute_off_angle): (cos, sin) of goal_angle - normalize_angle(yaw); goal_angle (N,1)."""
    _, _, yaw = get_euler_xyz(quat_xyzw[..., 0:4])
    d = goal_angle - normalize_angle(yaw).unsqueeze(-1)
    return torch.cat((torch.cos(d).reshape(-1, 1), torch.sin(d).reshape(-1, 1)), dim=-1)


def observations_walk(dof_pos, dof_vel, imu6, heading2, feet8) -> Tensor:
    """tasks/walk_env.py:1032-1050 / orient_env.py: the 52-wide row."""
    return torch.cat((dof_pos, dof_vel, imu6, heading2, feet8), dim=-1)


def _walk_terms(dof_pos, default_dof_pos, imu_lin, imu_ang, quat, num_envs):
    up = get_basis_vector(quat[..., 0:4], torch.tensor([[0.0, 0.0, 1.0]], device=quat.device).repeat(num_envs, 1)).view(num_envs, 3)
    up_proj = up[:, 2]
    vel6 = torch.linalg.norm(torch.cat((imu_lin, imu_ang), dim=1), dim=1)
    vel_lin = torch.linalg.norm(imu_lin, dim=1)
    vel_ang = torch.linalg.norm(imu_ang, dim=1)
    pos = torch.linalg.norm(default_dof_pos - dof_pos, dim=1)
    return up_proj, vel6, vel_lin, vel_ang, pos


def _walk_tail(rew, reset_buf, close, up_proj, pos, vel_ang, vel_lin, out, out_value, progress_buf, max_episode_length):
    ones = torch.ones_like(reset_buf)
    reset = torch.where(up_proj < 0.7, ones, reset_buf)
    rew = torch.where(up_proj < 0.7, torch.ones_like(rew) * -100.0, rew)
    state = torch.where(close, ones, torch.zeros_like(rew))
    state = torch.where(pos < 0.15, state + torch.ones_like(rew), state)
    state = torch.where(vel_ang < 0.1, state + torch.ones_like(rew), state)
    state = torch.where(vel_lin < 0.1, state + torch.ones_like(rew), state)
    reset = torch.where(state == 4.0, ones, reset)
    rew = torch.where(state == 4.0, torch.ones_like(rew) * (1000.0 - 1000.0 * (progress_buf / max_episode_length)), rew)
    reset = torch.where(out, ones, reset)
    rew = torch.where(out, torch.ones_like(rew) * out_value, rew)
    reset = torch.where(progress_buf >= max_episode_length, ones, reset)
    rew = torch.where(progress_buf >= max_episode_length, torch.zeros_like(rew), rew)
    return rew, reset


def reward_walk(dof_pos, default_dof_pos, imu_lin, imu_ang, bez_pos, quat, goal, reset_buf, progress_buf,
                max_episode_length: int) -> Tuple[Tensor, Tensor]:
    """tasks/walk_env.py:848-997 with the dead computations (feet term, legacy branch, debug prints) removed;
    ``bez_init_state`` is (0, 0) (zeroed in place at :966-967)."""
    n = dof_pos.shape[0]
    d = torch.sub(goal, bez_pos[..., 0:2])
    n_goal = torch.linalg.norm(d, dim=1).reshape(-1, 1)
    u = torch.div(d, n_goal)
    vel_fwd = torch.sum(torch.mul(u, imu_lin[..., 0:2]), dim=-1)
    up_proj, vel6, vel_lin, vel_ang, pos = _walk_terms(dof_pos, default_dof_pos, imu_lin, imu_ang, quat, n)
    dist_h = torch.abs(1 - up_proj)
    vel_s, pos_s = torch.mul(vel6, 0.05), torch.mul(pos, 0.05)
    height_vel_pos = -torch.add(torch.add(vel_s, pos_s), dist_h)
    vel_height = torch.mul(torch.sub(torch.mul(vel_fwd, 10), torch.add(dist_h, 5 * pos_s)), 1)
    n_goal = n_goal.reshape(-1)
    close = n_goal < 0.05
    rew = torch.where(close, height_vel_pos, vel_height)
    init = torch.sub(goal, torch.zeros(2, device=goal.device))
    ui = torch.div(init, torch.linalg.norm(init, dim=1).reshape(-1, 1))
    ang_now = torch.atan2(u[..., 1], u[..., 0])
    ang_init = torch.atan2(ui[..., 1], ui[..., 0])
    out = torch.abs(ang_init - ang_now).reshape(-1) > 1.5708
    return _walk_tail(rew, reset_buf, close, up_proj, pos, vel_ang, vel_lin, out, -100.0, progress_buf, max_episode_length)


def reward_orient(dof_pos, default_dof_pos, imu_lin, imu_ang, bez_pos, quat, goal_angle, reset_buf, progress_buf,
                  bez_init_xy, max_episode_length: int) -> Tuple[Tensor, Tensor]:
    """tasks/orient_env.py:866-1014 with the dead computations removed; goal_angle (N,1)."""
    n = dof_pos.shape[0]
    _, _, yaw = get_euler_xyz(quat[..., 0:4])
    ang = torch.sub(goal_angle, normalize_angle(yaw).unsqueeze(-1)).reshape(-1)
    up_proj, vel6, vel_lin, vel_ang, pos = _walk_terms(dof_pos, default_dof_pos, imu_lin, imu_ang, quat, n)
    dist_h = torch.abs(1 - up_proj)
    vel_s, pos_s = torch.mul(vel6, 0.05), torch.mul(pos, 0.05)
    height_vel_pos = -torch.add(torch.add(vel_s, pos_s), dist_h)
    vel_height = torch.mul(torch.sub(torch.mul(torch.abs(ang), -0.5), torch.add(dist_h, 0.05 * pos_s)), 1)
    close = ang < 0.05
    rew = torch.where(close, height_vel_pos, vel_height)
    out = torch.linalg.norm(torch.sub(bez_pos[..., 0:2], bez_init_xy), dim=1).reshape(-1) > 0.3
    return _walk_tail(rew, reset_buf, close, up_proj, pos, vel_ang, vel_lin, out, -5.0, progress_buf, max_episode_length)


# --------------------------------------------------------------------------- reset
def reset_idx_dof(default_dof_pos_rows: Tensor, lower: Tensor, upper: Tensor, u_pos: Tensor,
                  u_vel: Tensor) -> Tuple[Tensor, Tensor]:
    """tasks/kick_env.py:786-791 with the two ``torch.rand`` draws passed in explicitly (``u_pos``,
    ``u_vel`` uniform in [0,1), shape (k,18)); ``torch_rand_float`` = (hi-lo)*u + lo."""
    pos = tensor_clamp(default_dof_pos_rows + ((0.15 - -0.15) * u_pos + -0.15), lower, upper)
    vel = (0.1 - -0.1) * u_vel + -0.1
    return pos, vel


# --------------------------------------------------------------------------- whole-step oracle
class KickStepOracle:
    """CPU model of ``VecTask.step`` (tasks/base/vec_task.py:303-349) + ``KickEnv.pre/post_physics_step``
    (tasks/kick_env.py:410-438) over Isaac-Gym-layout AoS state tensors it is handed (it never
    simulates: the caller mutates the state tensors between ``pre`` and ``post`` the way
    ``gym.simulate`` + ``refresh_*`` would).

    State tensors: root_states (N*2,13), dof_state (N*18,2), rigid_body (N*NB,13),
    net_contact (N*NB,3).  ``reset_uniforms(step) -> (N,36) in [0,1)`` supplies, per ENV ID, the draws a
    reset of that env consumes (cols 0:18 positions, 18:36 velocities) -- the reference draws from
    torch's global generator in env_ids order (:786-787); keying by env id makes the oracle and the
    CUDA path comparable without sharing a generator."""

    def __init__(self, num_envs, root_states, dof_state, rigid_body, net_contact, default_dof_pos,
                 lower, upper, goal, ball_init, bez_init_xy, initial_root_states, dt=0.01667,
                 max_episode_length=900, clip_actions=3.9, clip_obs=math.inf, cleats=False,
                 imu_body=1, left_foot=12, right_foot=20, left_cleats=(13, 17), right_cleats=(25, 29),
                 alias_prev_lin_vel=True, reset_uniforms=None):
        n = self.n = num_envs
        self.root, self.dof, self.rb, self.cf = root_states, dof_state, rigid_body, net_contact
        self.default_dof_pos = default_dof_pos.reshape(1, 18).repeat(n, 1) if default_dof_pos.dim() == 1 \
            else default_dof_pos
        self.lower, self.upper = lower, upper
        self.goal, self.ball_init, self.bez_init_xy = goal, ball_init, bez_init_xy
        self.initial_root_states = initial_root_states
        self.dt, self.max_len = dt, max_episode_length
        self.clip_actions, self.clip_obs, self.cleats = clip_actions, clip_obs, cleats
        self.alias = alias_prev_lin_vel
        self.reset_uniforms = reset_uniforms
        # views exactly as tasks/kick_env.py:168-196
        self.dof_pos = self.dof.view(n, 18, 2)[..., 0]
        self.dof_vel = self.dof.view(n, 18, 2)[..., 1]
        self.bez_pos = self.root.view(n, 2, 13)[..., 0, 0:3]
        self.ball_pos = self.root.view(n, 2, 13)[..., 1, 0:3]
        self.ball_vel = self.root.view(n, 2, 13)[..., 1, 7:10]
        rb = self.rb.view(n, -1, 13)
        self.quat, self.lin, self.ang = rb[..., imu_body, 3:7], rb[..., imu_body, 7:10], rb[..., imu_body, 10:13]
        cf = self.cf.view(n, -1, 3)
        if cleats:
            self.left_c = cf[..., left_cleats[0]:left_cleats[1], 0:3]
            self.right_c = cf[..., right_cleats[0]:right_cleats[1], 0:3]
        else:
            self.left_f, self.right_f = cf[..., left_foot, 0:3], cf[..., right_foot, 0:3]
        dev = root_states.device              # CPU for the oracle proper; a CUDA device = "the reference's torch ops on the GPU"
        self.gravity_vec = torch.tensor([[0.0, 0.0, -1.0]], device=dev).repeat(n, 1)       # :217
        self.prev_lin_vel = torch.tensor([[0, 0, 0]], device=dev).repeat(n, 1)             # int64 zeros, :183
        # vec_task.py:226-249 (reset_buf starts at ones, KickEnv.__init__ then resets everything, :238)
        self.obs_buf = torch.zeros(n, 54, device=dev)
        self.rew_buf = torch.zeros(n, device=dev)
        self.reset_buf = torch.ones(n, dtype=torch.long, device=dev)
        self.timeout_buf = torch.zeros(n, dtype=torch.long, device=dev)
        self.progress_buf = torch.zeros(n, dtype=torch.long, device=dev)
        self.randomize_buf = torch.zeros(n, dtype=torch.long, device=dev)
        self.step_count = 0
        self.targets = None
        self.actions = torch.zeros(n, 18, device=dev)

    def reset_idx(self, env_ids: Tensor, uniforms: Optional[Tensor] = None):
        """tasks/kick_env.py:779-850; the four indexed setters are modelled as row copies of
        ``initial_root_states`` (what a simulator does with them)."""
        if uniforms is None and self.reset_uniforms is None:
            # the reference's own draw order: positions then velocities, (k,18) each (:786-787)
            dev = self.dof.device
            u = torch.cat((torch.rand(len(env_ids), 18, device=dev), torch.rand(len(env_ids), 18, device=dev)), 1)
        else:
            if uniforms is None:
                uniforms = self.reset_uniforms(self.step_count)
            u = uniforms[env_ids]
        pos, vel = reset_idx_dof(self.default_dof_pos[env_ids], self.lower, self.upper, u[:, 0:18], u[:, 18:36])
        self.dof_pos[env_ids] = pos
        self.dof_vel[env_ids] = vel
        for a in (env_ids * 2, env_ids * 2 + 1):
            self.root[a] = self.initial_root_states[a]
        self.progress_buf[env_ids] = 0
        self.reset_buf[env_ids] = 0

    def pre_physics_step(self, actions: Tensor):
        self.actions, self.targets = pre_physics(actions, self.default_dof_pos, self.lower, self.upper,
                                                 self.clip_actions)
        return self.targets

    def compute_observations(self):
        """tasks/kick_env.py:749-777."""
        imu6, new_prev = imu(self.quat, self.lin, self.ang, self.prev_lin_vel, self.gravity_vec, self.dt)
        self.prev_lin_vel = new_prev if self.alias else new_prev.clone()
        orn2 = off_orn(self.bez_pos, self.quat, self.goal)
        if self.cleats:
            self.feet = feet_cleats(self.left_c, self.right_c)
        else:
            self.feet = torch.cat((feet_no_cleats(self.left_f), feet_no_cleats(self.right_f)), 1)
        self.obs_buf[:] = observations(self.dof_pos, self.dof_vel, imu6, orn2, self.feet, self.ball_init)

    def compute_reward(self):
        self.rew_buf[:], self.reset_buf[:] = reward(
            self.dof_pos, self.default_dof_pos, self.lin, self.ang, self.bez_pos, self.ball_pos,
            self.ball_vel, self.goal, self.ball_init, self.bez_init_xy, self.reset_buf,
            self.progress_buf, self.max_len)

    def post_physics_step(self):
        """tasks/base/vec_task.py:331-332 + tasks/kick_env.py:426-438."""
        self.timeout_buf = torch.where(self.progress_buf >= self.max_len - 1,
                                       torch.ones_like(self.timeout_buf), torch.zeros_like(self.timeout_buf))
        self.progress_buf += 1
        self.randomize_buf += 1
        env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(env_ids) > 0:
            self.reset_idx(env_ids)
        self.compute_observations()
        self.compute_reward()
        self.step_count += 1
        return (torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs), self.rew_buf, self.reset_buf,
                self.timeout_buf)
