"""Import the reference's OWN ``bez_isaacgym/tasks/kick_env.py`` on CPU (oracle; test infrastructure).

Only usable where /root/reference exists (the authoring container).  It is used to (a) validate
``oracle.task_oracle`` against the real reference functions and (b) generate the golden vectors
under ``tests/golden`` (``oracle/make_golden.py``).  Nothing that runs on the GPU box imports this.

The reference imports packages that are not installable offline (``isaacgym``, ``gym``,
``matplotlib``); they are replaced by inert ``sys.modules`` stubs.  The only stub that carries
arithmetic is ``isaacgym.torch_utils`` -> ``oracle.isaacgym_torch_utils`` (restated, see there).
No reference source is copied: the module is executed from where it lies.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("BEZ_REFERENCE_ROOT", "/root/reference")
_PKG_DIR = os.path.join(REFERENCE_ROOT, "bez_isaacgym")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_PKG_DIR, "tasks", "kick_env.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    """Attribute sink used for gymapi enums/classes that are only touched at sim-creation time."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, item):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def install_stubs():
    if "isaacgym" in sys.modules and getattr(sys.modules["isaacgym"], "_bez_stub", False):
        return
    if not hasattr(np, "Inf"):          # removed in numpy 2; tasks/base/vec_task.py:92-97 uses it
        np.Inf = np.inf
    from oracle import isaacgym_torch_utils as tu

    from oracle import fake_isaacgym as fg
    gymapi = _stub("isaacgym.gymapi", SimParams=fg._Bag, Vec3=fg.Vec3, Quat=fg.Quat,
                   Transform=fg.Transform, PlaneParams=fg._Bag, AssetOptions=fg._Bag,
                   ContactCollection=lambda v: v, CameraProperties=fg._Bag,
                   SIM_PHYSX=1, SIM_FLEX=0, UP_AXIS_Z=1, UP_AXIS_Y=0, DOMAIN_SIM=0,
                   acquire_gym=lambda: _current_gym[0] if _current_gym[0] is not None else _Anything())
    gymtorch = _stub("isaacgym.gymtorch", wrap_tensor=lambda t: t, unwrap_tensor=lambda t: t)
    noop = lambda *a, **k: None
    gymutil = _stub("isaacgym.gymutil", get_property_setter_map=noop, get_property_getter_map=noop,
                    get_default_setter_args=noop, apply_random_samples=noop, check_buckets=noop,
                    generate_random_samples=noop)
    torch_utils = _stub("isaacgym.torch_utils", **{k: getattr(tu, k) for k in tu.__all__})
    torch_utils.__all__ = list(tu.__all__)
    _stub("isaacgym", gymapi=gymapi, gymtorch=gymtorch, gymutil=gymutil, torch_utils=torch_utils,
          _bez_stub=True)

    pyplot = _stub("matplotlib.pyplot", subplots=noop)
    _stub("matplotlib", use=noop, pyplot=pyplot)

    class Box:                            # gym.spaces.Box as used by tasks/base/vec_task.py:92-95
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high = np.asarray(low), np.asarray(high)
            self.shape = self.low.shape if shape is None else shape

    spaces = _stub("gym.spaces", Box=Box)
    _stub("gym", spaces=spaces, Space=object)


_cached = {}
_current_gym = [None]      # the FakeGym instance gymapi.acquire_gym() hands to the next reference KickEnv


def load_reference_kick_env():
    """Return the reference module ``tasks.kick_env`` (executed from /root/reference)."""
    return load_reference_task("kick")


def load_reference_task(task="kick"):
    """Return the reference module ``tasks.{kick,walk,orient}_env`` (executed from /root/reference)."""
    if task in _cached:
        return _cached[task]
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    install_stubs()
    if _PKG_DIR not in sys.path:
        sys.path.insert(0, _PKG_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # tasks/__init__.py also imports walk_env/orient_env; import the module file directly
        # under the package name so its relative import (.base.vec_task) resolves.
        pkg = types.ModuleType("tasks")
        pkg.__path__ = [os.path.join(_PKG_DIR, "tasks")]
        sys.modules.setdefault("tasks", pkg)
        _cached[task] = importlib.import_module(f"tasks.{task}_env")
    return _cached[task]


def reference_task_cfg(num_envs, cleats=False, task="kick"):
    """The reference's OWN task config (cfg/task/bez_kick_test.yaml is the interpolation-free twin of
    bez_kick.yaml, SURVEY 5 'Config'; bez_walk_test.yaml likewise for bez_walk.yaml; bez_orient.yaml = bez_walk.yaml +
    ``goalState.goal_angle: 1.5708`` and has no interpolation-free twin, so the walk twin is used with that key added),
    with numEnvs set, the CPU pipeline selected and the debug printing switched off."""
    import yaml
    with open(os.path.join(_PKG_DIR, "cfg", "task", "bez_kick_test.yaml" if task == "kick" else "bez_walk_test.yaml")) as f:
        cfg = yaml.safe_load(f)
    if task != "kick":
        cfg["env"]["debug"]["rewards"] = False
        cfg["env"]["envSpacing"] = 5
        if task == "orient":
            with open(os.path.join(_PKG_DIR, "cfg", "task", "bez_orient.yaml")) as f:
                cfg["env"]["goalState"]["goal_angle"] = yaml.safe_load(f)["env"]["goalState"]["goal_angle"]
    cfg["env"]["numEnvs"] = int(num_envs)
    cfg["env"]["asset"]["cleats"] = bool(cleats)
    cfg["sim"]["use_gpu_pipeline"] = False
    cfg["sim"]["physx"]["use_gpu"] = False
    cfg["rl_device"] = "cpu"
    return cfg


def make_reference_env(state, on_simulate=None, cleats=False, rand_source=None, task="kick"):
    """Instantiate the UNMODIFIED reference ``KickEnv`` over the four tensors of ``state`` (a
    ``bez_isaacgym_b200.synthetic_gym.SimState`` on CPU) through ``oracle.fake_isaacgym.FakeGym``.

    ``rand_source(shape) -> Tensor in [0,1)`` replaces ``torch.rand`` inside the restated
    ``torch_rand_float`` so a checker can feed the same reset draws to another implementation."""
    from oracle import fake_isaacgym as fg
    from oracle import isaacgym_torch_utils as tu
    mod = load_reference_task(task)
    gym = fg.FakeGym(state.root_states, state.dof_state, state.rigid_body, state.net_contact, on_simulate,
                     actors_per_env=2 if task == "kick" else 1)
    _current_gym[0] = gym
    if rand_source is not None:
        def torch_rand_float(lower, upper, shape, device):
            return (upper - lower) * rand_source(shape) + lower
        mod.torch_rand_float = torch_rand_float
    else:
        mod.torch_rand_float = tu.torch_rand_float
    cwd = os.getcwd()
    try:
        os.chdir(_PKG_DIR)        # assetRoot is relative to bez_isaacgym/ (README.md:43-46)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cls = {"kick": "KickEnv", "walk": "WalkEnv", "orient": "OrientEnv"}[task]
            env = getattr(mod, cls)(reference_task_cfg(state.num_envs, cleats, task), "cpu", 0, True)
    finally:
        os.chdir(cwd)
        _current_gym[0] = None
    env._fake_gym = gym
    return env
