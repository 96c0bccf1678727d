"""Shape, index and limit constants of the Bez robot as Isaac Gym lays them out.

Isaac Gym orders bodies and DOFs by a depth-first traversal of the URDF with children sorted by
joint name; applying that to the reference's ``resources/assets/bez/model/soccerbot_stl.urdf``
reproduces the hard-coded indices in ``bez_isaacgym/tasks/kick_env.py:23-41`` (``Joints`` enum),
``:175-177`` (body 1 = imu_link) and ``:193-196`` (bodies 12 / 20 = feet).  The limits are the
``<limit lower= upper=>`` attributes at ``soccerbot_stl.urdf:100,108,156,164,293-333,462-502,550,558``
(checked against the URDF by ``tests/test_model_constants.py`` where /root/reference exists).
"""
import math

NUM_DOF = 18
NUM_OBS = 54
NUM_ACTIONS = 18

#: DOF order = kick_env.py:23-41
DOF_NAMES = (
    "head_motor_0", "head_motor_1",
    "left_arm_motor_0", "left_arm_motor_1",
    "left_leg_motor_0", "left_leg_motor_1", "left_leg_motor_2",
    "left_leg_motor_3", "left_leg_motor_4", "left_leg_motor_5",
    "right_arm_motor_0", "right_arm_motor_1",
    "right_leg_motor_0", "right_leg_motor_1", "right_leg_motor_2",
    "right_leg_motor_3", "right_leg_motor_4", "right_leg_motor_5",
)

_ARM = ((-math.pi / 2, 5 * math.pi / 4), (0.0, math.pi))
_LEG = ((-1.309, 0.524), (-math.pi / 4, math.pi / 2), (-math.pi / 4, 3 * math.pi / 4),
        (-2.793, 0.0), (-math.pi / 4, math.pi / 2), (-math.pi / 4, math.pi / 4))
_HEAD = ((-math.pi / 2, math.pi / 2), (-3 * math.pi / 4, 3 * math.pi / 4))
_LIMITS = _HEAD + _ARM + _LEG + _ARM + _LEG
DOF_LOWER = tuple(lo for lo, _ in _LIMITS)
DOF_UPPER = tuple(hi for _, hi in _LIMITS)

#: bodies per env: robot links + the ball actor's single body
BODIES_NO_CLEATS = 21 + 1      # soccerbot_stl.urdf
BODIES_CLEATS = 29 + 1         # soccerbot_stl_sensor.urdf
IMU_BODY = 1                   # kick_env.py:175-177
LEFT_FOOT_BODY = 12            # kick_env.py:193
RIGHT_FOOT_BODY = 20           # kick_env.py:195
LEFT_CLEATS = (13, 17)         # kick_env.py:188  (slice 13:17)
RIGHT_CLEATS = (25, 29)        # kick_env.py:190  (slice 25:29)
ACTORS_PER_ENV = 2             # bez, ball (kick_env.py:365-378)

#: obs row layout (kick_env.py:1409-1415)
OBS_DOF_POS = slice(0, 18)
OBS_DOF_VEL = slice(18, 36)
OBS_IMU = slice(36, 42)
OBS_OFF_ORN = slice(42, 44)
OBS_FEET = slice(44, 52)
OBS_BALL_INIT = slice(52, 54)

#: sibling tasks (tasks/walk_env.py, tasks/orient_env.py): the robot only, 52-wide observation (walk_env.py:104)
TASKS = ("kick", "walk", "orient")
NUM_OBS_WALK = 52
BODIES_NO_CLEATS_WALK = 21
BODIES_CLEATS_WALK = 29


def task_dims(task="kick", cleats=False):
    """(actors per env, bodies per env, observation width) of a task."""
    if task == "kick":
        return 2, (BODIES_CLEATS if cleats else BODIES_NO_CLEATS), NUM_OBS
    if task in ("walk", "orient"):
        return 1, (BODIES_CLEATS_WALK if cleats else BODIES_NO_CLEATS_WALK), NUM_OBS_WALK
    raise ValueError(f"unknown task {task!r}")


IMU_MAX_ANG_VEL = 8.7266       # kick_env.py:99
IMU_MAX_LIN_ACC = 2.0 * 9.81   # kick_env.py:100


def default_task_cfg(num_envs=4096, cleats=False, use_gpu_pipeline=True, rl_device="cuda:0", task="kick"):
    """The subset of ``cfg/task/bez_kick.yaml`` (+ the three top-level keys ``VecTask`` reads) that the
    per-step path consumes, with the Hydra interpolations resolved by hand.  ``task="walk"`` / ``"orient"``:
    the same for ``cfg/task/bez_walk.yaml`` / ``bez_orient.yaml`` (goal (2, 0), 10 s episodes, no ball; orient adds
    ``goal_angle`` 1.5708)."""
    if task != "kick":
        cfg = default_task_cfg(num_envs, cleats, use_gpu_pipeline, rl_device)
        cfg["name"] = f"bez_{task}"
        env = cfg["env"]
        del env["ballInitState"]
        env["envSpacing"] = 5
        env["goalState"] = {"goal": [2.0, 0.0]}
        if task == "orient":
            env["goalState"]["goal_angle"] = 1.5708
        elif task != "walk":
            raise ValueError(f"unknown task {task!r}")
        env["learn"]["episodeLength_s"] = 10
        return cfg
    ready = {n: 0.0 for n in DOF_NAMES}
    for side in ("left", "right"):
        ready[f"{side}_leg_motor_2"] = 0.564
        ready[f"{side}_leg_motor_3"] = -1.176
        ready[f"{side}_leg_motor_4"] = 0.613
        ready[f"{side}_arm_motor_1"] = 1.5
    zero3 = [0.0, 0.0, 0.0]
    return {
        "name": "bez_kick",
        "physics_engine": "physx",
        "rl_device": rl_device,
        "env": {
            "numEnvs": int(num_envs),
            "envSpacing": 4,
            "clipActions": 3.9,
            "controlFrequencyInv": 1,
            "plane": {"staticFriction": 1, "dynamicFriction": 1, "restitution": 0.0},
            "bezInitState": {"pos": [0.0, 0.0, 0.34], "rot": [0.0, 0.0, 0.0, 1.0],
                             "vLinear": list(zero3), "vAngular": list(zero3)},
            "ballInitState": {"pos": [0.175, 0, 0.1], "rot": [0.0, 0.0, 0.0, 1.0],
                              "vLinear": list(zero3), "vAngular": list(zero3)},
            "goalState": {"goal": [1.5, 0.0]},
            "control": {"stiffness": 100, "damping": 7.5, "actionScale": 0.5},
            "readyJointAngles": ready,
            "learn": {"episodeLength_s": 15},
            "asset": {"cleats": bool(cleats), "stl": True},
            "debug": {"rewards": False},
        },
        "sim": {"dt": 0.01667, "substeps": 2, "up_axis": "z",
                "use_gpu_pipeline": bool(use_gpu_pipeline), "gravity": [0.0, 0.0, -9.81]},
        "task": {"randomize": False, "randomization_params": {}},
    }
