"""Demonstration recording and reference-state initialisation on the batched env.

Reference: `env/wrappers/get_demonstration_wrapper.py:7-70` (row layout, `read_demo`, `save_demo`),
`env/wrappers/save_demo_wrapper.py:7-19`, `env/wrappers/reference_state_initialization_wrapper.py:10-43`.
The reference ships no demonstration files (its `*_DEMO` tasks load one at construction, `tasks/task_base.py:169-176`):
here a user records demonstrations from the batched env in the reference's `.npy` layout, hands one to the imitation
tasks (`env.set_demo`, kernels: `QS_TASK_*_DEMO`) and starts episodes from demonstration rows.  Host-side glue over the
C ABI (`qs_set_demo`, `qs_reset_to_state`), not part of the measured path.
"""
import os
import random

import numpy as np
import torch

# [action(A), q(12), qd(12), base_pos(3), base_quat(4), base_lin_vel(3), base_ang_vel(3), landing_flag(1)]
DEMO_FIELDS = ("action", "joint_position", "joint_velocity", "base_position", "base_orientation", "base_linear_velocity",
               "base_angular_velocity", "landing_started")


def read_demo(demo, action_dim=6, num_joints=12):
    """GetDemonstrationWrapper.read_demo (get_demonstration_wrapper.py:59-70): split one row (or [..., D] rows)"""
    bounds = np.cumsum([0, action_dim, num_joints, num_joints, 3, 4, 3, 3, 1])
    return [demo[..., a:b] for a, b in zip(bounds[:-1], bounds[1:])]


def demo_rows_to_states(rows, action_dim=6):
    """demonstration rows [..., D] -> physical states [..., 37] in qs_set_state order (pos, quat, lin, ang, q, qd)"""
    _, q, qd, pos, quat, lin, ang, _ = read_demo(rows, action_dim)
    cat = torch.cat if torch.is_tensor(rows) else np.concatenate
    return cat([pos, quat, lin, ang, q, qd], -1)


class DemonstrationRecorder:
    """GetDemonstrationWrapper for every env of the batch at once: after each step one row per env is appended on the
    device; `demo(i)` / `save_demo(i)` give env i's current episode in the reference's layout (the last step is dropped
    like `save_demo`, :29-31).  `landing_started` latches when the task has switched and the base moves down (:45-47)."""

    def __init__(self, env, path=None, name="demo_list"):
        if getattr(env, "_auto_reset", False):
            # an auto-reset env starts the next episode inside step(): rows would run over episode boundaries and the
            # terminal state would be lost.  The reference wrapper records ONE episode between two reset() calls.
            raise ValueError("DemonstrationRecorder records per episode: build the env with auto_reset=False")
        self.env = env
        self.name = f"{name}.npy"
        self.save_path = path
        if path is not None:
            os.makedirs(path, exist_ok=True)
        self.rows = []
        self.start = [0] * env.num_envs       # first row of each env's current episode
        self.landing_started = torch.zeros(env.num_envs, dtype=torch.bool, device=env.device)

    def __getattr__(self, k):      # gym.Wrapper forwarding
        return getattr(self.env, k)

    def reset(self, mask=None, **k):
        if mask is None:
            self.rows = []
            self.start = [0] * self.env.num_envs
            self.landing_started.zero_()
        else:   # partial reset: only the selected envs start a new episode (and a new demonstration)
            m = torch.as_tensor(mask, device=self.env.device).bool()
            for i in m.nonzero().flatten().tolist():
                self.start[i] = len(self.rows)
            self.landing_started &= ~m
        return self.env.reset(mask=mask, **k)

    def step(self, action):
        out = self.env.step(action)
        self.rows.append(self._get_demo())
        return out

    def _get_demo(self):
        e, r = self.env, self.env.robot
        lin = r.GetBaseLinearVelocity()
        switched = e._views["task"][0] != 0          # task.is_switched_controller()
        self.landing_started |= switched & (lin[:, 2] <= 0.0)
        return torch.cat([e.get_last_filtered_action(), r.GetMotorAngles(), r.GetMotorVelocities(), r.GetBasePosition(),
                          r.GetBaseOrientation(), lin, r.GetBaseAngularVelocity(),
                          self.landing_started[:, None].to(torch.float32)], dim=1).clone()

    def demo(self, env_index=0):
        rows = self.rows[self.start[env_index]:]
        if len(rows) < 2:
            return np.zeros((0, self.env.action_dim + 38), np.float32)
        return torch.stack([r[env_index] for r in rows[:-1]]).cpu().numpy()

    def save_demo(self, env_index=0):
        d = self.demo(env_index)
        np.save(os.path.join(self.save_path, self.name), d)
        return d


class ReferenceStateInitialization:
    """ReferenceStateInitializationWrapper (:10-43) on the batch: every reset of an env draws an element of the
    demonstration (`compute_random_el`: anywhere but the last five rows, and every sixth reset within the first fifth)
    and starts the episode from that row's state, without the settle (`env.set_robot_desired_state` + `reset`)."""

    def __init__(self, env, demo_list, seed=None):
        self.env = env
        self.demo_list = torch.as_tensor(np.asarray(demo_list), dtype=torch.float32, device=env.device)
        self.demo_length = int(self.demo_list.shape[0])
        self._rng = random.Random(seed)
        self.counter = 0
        self.counter_reset_period = 5
        self.random_el = torch.zeros(env.num_envs, dtype=torch.long, device=env.device)

    def __getattr__(self, k):
        return getattr(self.env, k)

    def compute_random_el(self):
        limit = self.demo_length - 5
        if self.counter == self.counter_reset_period:
            self.counter = 0
            limit = self.demo_length // 5
        else:
            self.counter += 1
        return self._rng.randint(0, limit - 1)

    def reset(self, mask=None):
        n = self.env.num_envs
        m = torch.ones(n, dtype=torch.bool, device=self.env.device) if mask is None else \
            torch.as_tensor(mask, device=self.env.device).bool()
        idx = m.nonzero().flatten().tolist()
        els = torch.as_tensor([self.compute_random_el() for _ in idx], dtype=torch.long, device=self.env.device)
        self.random_el[m] = els
        if self.env.task_env.endswith("_DEMO"):      # env.task.set_demo_counter(value=self.random_el), :31
            self.env.task.set_demo_counter(els.to(torch.float32), m)
        states = self.env.get_state()
        states[m] = demo_rows_to_states(self.demo_list[els], self.env.action_dim)
        return self.env.reset_to_state(states, mask=m)

    def step(self, action):
        return self.env.step(action)
