"""Batched Hopf CPG (quadruped_spring/hopf_network.py:26-173) and the joint-PD +
Cartesian-impedance torque law of its __main__ loop (hopf_network.py:241-289),
evaluated by kernel K4 for N robots at once."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class HopfNetwork:
    """Same constructor keywords as the reference; state X is [N, 2, 4]
    (amplitude row 0, phase row 1).  Foot order FR, FL, RR, RL."""

    def __init__(self, num_envs=1, device="cuda:0", mu=2, omega_swing=1 * 2 * np.pi, omega_stance=1 * 2 * np.pi,
                 gait="TROT", coupling_strength=1, couple=True, time_step=0.001, ground_clearance=0.05,
                 ground_penetration=0.01, robot_height=0.25, des_step_len=0.04, seed=None):
        if not torch.cuda.is_available():
            raise RuntimeError("HopfNetwork runs on a CUDA device only (no CPU fallback)")
        self._L = _lib.lib()
        self.device = torch.device(device)
        self.num_envs = num_envs
        self._mu, self._omega_swing, self._omega_stance = mu, omega_swing, omega_stance
        self._couple, self._coupling_strength, self._dt = couple, coupling_strength, time_step
        self._set_gait(gait)
        g = torch.Generator(device="cpu")
        if seed is not None:
            g.manual_seed(seed)
        # float64 state: the phases start exactly on the swing/stance switch (see csrc K4)
        X = torch.zeros(num_envs, 2, 4, dtype=torch.float64)
        X[:, 0, :] = torch.rand(num_envs, 4, generator=g, dtype=torch.float64) * 0.1      # hopf_network.py:63-64
        X[:, 1, :] = torch.as_tensor(self.PHI[0, :], dtype=torch.float64)
        self.X = X.to(self.device).contiguous()
        self._ground_clearance, self._ground_penetration = ground_clearance, ground_penetration
        self._robot_height, self._des_step_len = robot_height, des_step_len

    def _set_gait(self, gait):                                       # hopf_network.py:74-115
        pi = np.pi
        table = {
            "TROT": [[0, -pi, -pi, 0], [pi, 0, 0, pi], [pi, 0, 0, pi], [0, -pi, -pi, 0]],
            "WALK": [[0, -pi, -pi / 2, pi / 2], [pi, 0, pi / 2, 3 * pi / 2], [pi / 2, -pi / 2, 0, pi],
                     [-pi / 2, -3 * pi / 2, -pi, 0]],
            "BOUND": [[0, 0, -pi, -pi], [0, 0, -pi, -pi], [pi, pi, 0, 0], [pi, pi, 0, 0]],
            "PACE": [[0, -pi, 0, -pi], [pi, 0, pi, 0], [0, -pi, 0, -pi], [pi, 0, pi, 0]],
        }
        if gait not in table:
            raise ValueError(gait + "not implemented.")
        self.PHI = np.array(table[gait], dtype=np.float64)

    def _params(self):
        coupling = self._coupling_strength if self._couple else 0.0
        p = np.array([self._mu, self._omega_swing, self._omega_stance, coupling, self._dt, self._des_step_len,
                      self._robot_height, self._ground_clearance, self._ground_penetration], dtype=np.float64)
        phi = np.ascontiguousarray(self.PHI.reshape(16), dtype=np.float64)
        return p, phi

    def update(self, q=None, qd=None, kp=(150, 70, 70), kd=(2, 0.5, 0.5), kp_cartesian=2500.0, kd_cartesian=40.0,
               foot_y=0.0838):
        """One Euler step of the oscillators -> desired foot x, z [N,4] (hopf_network.py:117-135).
        With q, qd [N,12] also returns the torques of hopf_network.py:241-289."""
        p, phi = self._params()
        n = self.num_envs
        xs = torch.empty(n, 4, device=self.device)
        zs = torch.empty(n, 4, device=self.device)
        tau = None
        gains_p = None
        if q is not None:
            q = q.to(torch.float32).contiguous()
            qd = qd.to(torch.float32).contiguous()
            tau = torch.empty(n, 12, device=self.device)
            gains = np.array(list(kp) + list(kd) + [kp_cartesian, kd_cartesian], dtype=np.float32)
            gains_p = gains.ctypes.data_as(C.POINTER(C.c_float))
        Xf = self.X.view(n, 8)
        _lib.check(self._L.qs_cpg_update(
            _p(Xf), p.ctypes.data_as(C.POINTER(C.c_double)), phi.ctypes.data_as(C.POINTER(C.c_double)), _p(q), _p(qd),
            gains_p, float(foot_y), _p(xs), _p(zs), _p(tau), n,
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        if tau is None:
            return xs, zs
        return xs, zs, tau

    def drive(self, env, n_ticks=1, kp=(150, 70, 70), kd=(2, 0.5, 0.5), kp_cartesian=2500.0, kd_cartesian=40.0, foot_y=0.0838):
        """n_ticks turns of the reference's loop `tau = law(cpg.update(), q, qd); env.step(tau)` (hopf_network.py:241-289)
        in one library call (qs_cpg_steps): the torque law reads the env's own joint state, nothing is staged through
        torch.  Returns env.step's tuple of the last tick."""
        p, phi = self._params()
        gains = np.array(list(kp) + list(kd) + [kp_cartesian, kd_cartesian], dtype=np.float32)
        _lib.check(self._L.qs_cpg_steps(
            env._h, _p(self.X.view(self.num_envs, 8)), p.ctypes.data_as(C.POINTER(C.c_double)),
            phi.ctypes.data_as(C.POINTER(C.c_double)), gains.ctypes.data_as(C.POINTER(C.c_float)), float(foot_y), int(n_ticks),
            _p(env._obs), _p(env._reward), _p(env._done), _p(env._trunc),
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        env._last_host_obs = None
        return env._obs, env._reward, env._done.bool(), {"TimeLimit.truncated": env._trunc.bool()}
