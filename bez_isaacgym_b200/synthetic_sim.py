"""Simulator backend protocol + the synthetic backend that stands in for Isaac Gym (PhysX is out of scope).

A backend owns the four Isaac-Gym-layout state tensors (see ``synthetic_gym``) and is what
``gym.acquire_*_tensor`` + ``gymtorch.wrap_tensor`` give the reference task
(``bez_isaacgym/tasks/kick_env.py:143-157``): tensors that are *borrowed* by the task, refreshed in place by the
simulator and written in place by the task on reset.  A real Isaac Gym adapter would implement the same
five members.
"""
import torch

from . import synthetic_gym as sg


class SimBackend:
    """What ``KickEnv`` needs from a simulator."""
    root_states: torch.Tensor      # (N*2, 13)
    dof_state: torch.Tensor        # (N*18, 2)
    rigid_body: torch.Tensor       # (N*NB, 13)
    net_contact: torch.Tensor      # (N*NB, 3)
    num_bodies: int
    #: True when the simulator itself restores actor root states on reset (the reference's
    #: set_actor_root_state_tensor_indexed path, kick_env.py:831-837); False lets the reset kernel copy
    #: ``initial_root_states`` rows.
    owns_root_reset = False

    def set_dof_position_targets(self, targets: torch.Tensor):   # gym.set_dof_position_target_tensor
        raise NotImplementedError

    def simulate(self):                                          # gym.simulate (+ fetch_results / refresh_*)
        raise NotImplementedError


class SyntheticGym(SimBackend):
    """Seeded synthetic state; ``simulate()`` calls ``on_simulate(self)`` if given (tests use it to move the
    state between steps) and is otherwise a no-op.  ``host=True`` keeps the tensors in pinned host memory
    (the reference's ``sim_device=cpu pipeline=cpu`` configuration)."""

    def __init__(self, num_envs, device="cuda:0", cleats=False, seed=1234, host=False, on_simulate=None,
                 filler=True, state=None, task="kick"):
        st = state if state is not None else sg.make_state(num_envs, seed=seed, device="cpu" if host else device,
                                                           cleats=cleats, filler=filler, task=task)
        if host:
            st = sg.SimState(*(t.pin_memory() if torch.cuda.is_available() else t for t in
                               (st.root_states, st.dof_state, st.rigid_body, st.net_contact)), st.num_envs, st.num_bodies)
        elif str(st.root_states.device) != str(torch.device(device)):
            st = st.to(device)
        self.state = st
        self.root_states, self.dof_state = st.root_states, st.dof_state
        self.rigid_body, self.net_contact = st.rigid_body, st.net_contact
        self.num_envs, self.num_bodies = st.num_envs, st.num_bodies
        self.targets = None
        self.frame = 0
        self.on_simulate = on_simulate

    def set_dof_position_targets(self, targets):
        self.targets = targets

    def simulate(self):
        self.frame += 1
        if self.on_simulate is not None:
            self.on_simulate(self)
