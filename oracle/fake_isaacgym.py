"""A headless fake of the Isaac Gym API surface the UNMODIFIED reference ``KickEnv`` touches (oracle; TEST
INFRASTRUCTURE ONLY; only usable where /root/reference exists).

It lets ``/root/reference/bez_isaacgym/tasks/kick_env.py`` + ``tasks/base/vec_task.py`` run ``KickEnv.step()`` on
CPU from their own source text: ``acquire_*_tensor`` hands out four caller-provided torch tensors with the Isaac
Gym layouts, ``simulate`` calls a hook that refreshes them, the indexed setters do what a
simulator does with them (root-state rows are copied from the tensor that is passed in), and asset queries are
answered from the reference's own URDF (``oracle/urdf_layout.py``).  PhysX itself is out of scope.
API list: SURVEY.md App. F (``kick_env.py:143-157,240-408,419,750-753,831-847``; ``vec_task.py:189,280,324,328``).
"""
import os

import numpy as np
import torch

from oracle import urdf_layout


class Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z


class Quat:
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = x, y, z, w


class Transform:
    def __init__(self):
        self.p, self.r = Vec3(), Quat()


class _Bag:
    """Attribute bag for SimParams / AssetOptions / PlaneParams (attributes are only stored)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = _Bag()
        object.__setattr__(self, name, v)
        return v


DOF_PROP_DTYPE = np.dtype([("hasLimits", "?"), ("lower", "f4"), ("upper", "f4"), ("driveMode", "i4"),
                           ("velocity", "f4"), ("effort", "f4"), ("stiffness", "f4"), ("damping", "f4"),
                           ("friction", "f4"), ("armature", "f4")])


class FakeGym:
    def __init__(self, root_states, dof_state, rigid_body, net_contact, on_simulate=None, actors_per_env=2):
        self.actors_per_env = actors_per_env        # 2 for BezKick (robot + ball), 1 for walk / orient
        self.root_states, self.dof_state = root_states, dof_state
        self.rigid_body, self.net_contact = rigid_body, net_contact
        self.on_simulate = on_simulate
        self.num_envs_created = 0
        self.targets = None
        self.calls = []

    # --- sim / scene construction (kick_env.py:240-408) ---
    def create_sim(self, *a):
        return "sim"

    def prepare_sim(self, sim):
        self.calls.append("prepare_sim")

    def add_ground(self, sim, params):
        pass

    def load_asset(self, sim, root, file, options):
        path = os.path.join(root, file)
        if "ball" in os.path.basename(file):
            return dict(kind="ball", layout=urdf_layout.Layout(["ball"], [], [], [], 0))
        # the reference resolves assetRoot relative to bez_isaacgym/ (bez_kick.yaml:116)
        if not os.path.isabs(path):
            from oracle.reference_loader import REFERENCE_ROOT
            path = os.path.normpath(os.path.join(REFERENCE_ROOT, "bez_isaacgym", path))
        return dict(kind="bez", layout=urdf_layout.parse(path))

    def get_asset_dof_count(self, asset):
        return len(asset["layout"].dof_names)

    def get_asset_rigid_body_count(self, asset):
        return len(asset["layout"].bodies)

    def get_asset_rigid_shape_count(self, asset):
        return len(asset["layout"].bodies)

    def get_asset_joint_count(self, asset):
        return asset["layout"].num_joints

    def get_asset_dof_names(self, asset):
        return list(asset["layout"].dof_names)

    def find_asset_dof_index(self, asset, name):
        return asset["layout"].dof_names.index(name)

    def _dof_props(self, asset):
        lay = asset["layout"]
        props = np.zeros(len(lay.dof_names), dtype=DOF_PROP_DTYPE)
        props["hasLimits"] = True
        props["lower"] = np.asarray(lay.lower, dtype=np.float32)
        props["upper"] = np.asarray(lay.upper, dtype=np.float32)
        return props

    def get_asset_dof_properties(self, asset):
        return self._dof_props(asset)

    def create_env(self, sim, lower, upper, per_row):
        self.num_envs_created += 1
        return self.num_envs_created - 1

    def begin_aggregate(self, *a):
        pass

    def end_aggregate(self, *a):
        pass

    def create_actor(self, env, asset, pose, name, group, filt, seg=0):
        self._last_bez_asset = asset if asset["kind"] == "bez" else getattr(self, "_last_bez_asset", None)
        return 0 if asset["kind"] == "bez" else 1

    def set_actor_dof_properties(self, env, handle, props):
        pass

    def enable_actor_dof_force_sensors(self, env, handle):
        pass

    def get_actor_index(self, env, handle, domain):
        return env * self.actors_per_env + handle

    def get_actor_dof_properties(self, env, handle):
        return self._dof_props(self._last_bez_asset)

    # --- tensor API (kick_env.py:143-157,750-753) ---
    def acquire_actor_root_state_tensor(self, sim):
        return self.root_states

    def acquire_dof_state_tensor(self, sim):
        return self.dof_state

    def acquire_rigid_body_state_tensor(self, sim):
        return self.rigid_body

    def acquire_net_contact_force_tensor(self, sim):
        return self.net_contact

    def refresh_dof_state_tensor(self, sim):
        pass

    def refresh_actor_root_state_tensor(self, sim):
        pass

    def refresh_rigid_body_state_tensor(self, sim):
        pass

    def refresh_net_contact_force_tensor(self, sim):
        pass

    def get_sim_dof_count(self, sim):
        return self.dof_state.shape[0]

    # --- stepping ---
    def simulate(self, sim):
        self.calls.append("simulate")
        if self.on_simulate is not None:
            self.on_simulate(self)

    def fetch_results(self, sim, wait):
        pass

    # --- setters ---
    def set_dof_position_target_tensor(self, sim, targets):
        self.targets = targets.clone()

    def set_actor_root_state_tensor_indexed(self, sim, states, indices, count):
        idx = indices.long()[:count]
        self.root_states[idx] = states[idx]

    def set_dof_position_target_tensor_indexed(self, sim, targets, indices, count):
        pass

    def set_dof_state_tensor_indexed(self, sim, states, indices, count):
        # the reference already wrote the new DOF rows into this very tensor (kick_env.py:789-791)
        assert states.data_ptr() == self.dof_state.data_ptr()
