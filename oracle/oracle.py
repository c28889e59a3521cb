"""ctypes binding of the CPU oracle (oracle/libqso.so).

TEST INFRASTRUCTURE, not the product: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TASKS = {
    "NO_TASK": 0, "JUMPING_IN_PLACE": 1, "JUMPING_FORWARD": 2, "BACKFLIP": 3,
    "JUMPING_IN_PLACE_PPO": 4, "JUMPING_FORWARD_PPO": 5, "BACKFLIP_PPO": 6,
    "JUMPING_IN_PLACE_PPO_HP": 7, "JUMPING_FORWARD_PPO_HP": 8, "CONTINUOUS_JUMPING_FORWARD": 9,
    "CONTINUOUS_JUMPING_FORWARD2": 10, "CONTINUOUS_JUMPING_FORWARD3": 11, "CONTINUOUS_JUMPING_FORWARD_PPO": 12,
    "JUMPING_IN_PLACE_DEMO": 13, "JUMPING_FORWARD_DEMO": 14, "BACKFLIP_DEMO": 15,
    "CONTINUOUS_JUMPING_FORWARD_DEMO": 16,
}
CONTROL = {"PD": 0, "CARTESIAN_PD": 1, "TORQUE": 2}
ACTION = {"DEFAULT": 0, "SYMMETRIC": 1, "SYMMETRIC_NO_HIP": 2}
OBS = {
    "ENCODER": 0, "ENCODER_2": 1, "CARTESIAN_NO_IMU": 2, "ARS_BASIC": 3, "ARS_SENSOR": 4,
    "LANDING_SENSOR": 5, "PPO_BASIC": 6, "PPO_BASIC_X": 7, "PPO_BASIC_CONTACT": 8,
    "ARS_BACKFLIP": 9, "PPO_BACKFLIP": 10, "PPO_CONTINUOUS_JUMPING_FORWARD": 11,
}


class WorldParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("num_iterations", C.c_int), ("gravity_z", C.c_double),
        ("mu_ground", C.c_double), ("mu_link", C.c_double), ("contact_erp", C.c_double),
        ("limit_erp", C.c_double), ("linear_slop", C.c_double), ("warmstart", C.c_double),
        ("residual_threshold", C.c_double), ("max_coord_vel", C.c_double),
        ("breaking_threshold", C.c_double), ("enable_limits", C.c_int),
        ("body_contact_response", C.c_int), ("self_collision", C.c_int),
    ]


class EnvConfig(C.Structure):
    _fields_ = [
        ("enable_springs", C.c_int), ("control_mode", C.c_int), ("action_mode", C.c_int),
        ("task", C.c_int), ("obs_mode", C.c_int), ("action_repeat", C.c_int),
        ("is_rl_interface", C.c_int), ("enable_action_interpolation", C.c_int),
        ("enable_action_filter", C.c_int), ("settling_steps", C.c_int), ("landing_mode", C.c_int),
        ("rest_mode", C.c_int),
        ("time_step", C.c_double),
    ]


def build(force=False):
    so = os.path.join(_HERE, "libqso.so")
    srcs = [os.path.join(_HERE, f) for f in ("qso_physics.c", "qso_env.c", "qso.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libqso.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    vp = C.c_void_p
    sig = {
        "qso_default_params": (None, [C.POINTER(WorldParams)]),
        "qso_world_create": (vp, []),
        "qso_world_destroy": (None, [vp]),
        "qso_world_set_params": (None, [vp, C.POINTER(WorldParams)]),
        "qso_world_get_params": (None, [vp, C.POINTER(WorldParams)]),
        "qso_world_set_state": (None, [vp, dp]),
        "qso_world_get_state": (None, [vp, dp]),
        "qso_world_add_torque": (None, [vp, dp]),
        "qso_world_step": (None, [vp]),
        "qso_world_num_contacts": (C.c_int, [vp]),
        "qso_world_get_contact": (None, [vp, C.c_int, ip, dp, dp, dp]),
        "qso_world_num_self_contacts": (C.c_int, [vp]),
        "qso_world_get_self_contact": (None, [vp, C.c_int, ip, ip, dp]),
        "qso_world_detect_self_contacts": (C.c_int, [vp]),
        "qso_world_get_dynamics": (None, [vp, C.c_int, dp, dp, dp]),
        "qso_world_set_mass": (None, [vp, C.c_int, C.c_double]),
        "qso_world_set_payload": (None, [vp, C.c_double, dp]),
        "qso_world_last_iterations": (C.c_int, [vp]),
        "qso_world_cone_clamped": (C.c_int, [vp]),
        "qso_world_last_rows": (C.c_int, [vp, ip, ip]),
        "qso_world_mass_matrix": (None, [vp, dp]),
        "qso_world_bias": (None, [vp, dp, dp]),
        "qso_world_link_pose": (None, [vp, C.c_int, dp, dp]),
        "qso_world_energy": (C.c_double, [vp, dp, dp]),
        "qso_env_default_config": (None, [C.POINTER(EnvConfig)]),
        "qso_env_create": (vp, [C.POINTER(EnvConfig)]),
        "qso_env_destroy": (None, [vp]),
        "qso_env_world": (vp, [vp]),
        "qso_env_action_dim": (C.c_int, [vp]),
        "qso_env_obs_dim": (C.c_int, [vp]),
        "qso_env_reset": (None, [vp, C.c_double, dp]),
        "qso_env_reset_to_state": (None, [vp, C.c_double, dp, dp]),
        "qso_env_set_demo": (None, [vp, dp, C.c_int]),
        "qso_env_set_demo_counter": (None, [vp, C.c_int]),
        "qso_env_get_demo_counter": (C.c_int, [vp]),
        "qso_env_step": (None, [vp, dp, dp, dp, ip, ip]),
        "qso_env_get_task_state": (None, [vp, dp]),
        "qso_env_get_jump_arrays": (None, [vp, dp]),
        "qso_env_get_landing_state": (None, [vp, dp]),
        "qso_env_get_rest_state": (None, [vp, dp]),
        "qso_env_get_torques": (None, [vp, dp, dp]),
        "qso_env_get_last_action": (None, [vp, dp]),
        "qso_env_set_gains": (None, [vp, dp, dp]),
        "qso_env_set_springs": (None, [vp, dp, dp, dp]),
        "qso_action_to_command": (None, [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp]),
        "qso_pd_torque": (None, [dp, dp, dp, dp, dp, dp, C.c_int, dp]),
        "qso_spring_torque": (None, [dp, dp, dp, dp, dp, dp]),
        "qso_fk_jacobian": (None, [dp, C.c_int, dp, dp]),
        "qso_ik": (None, [dp, C.c_int, dp]),
        "qso_rpy_from_quat": (None, [dp, dp]),
        "qso_backflip_pitch": (C.c_double, [dp, C.c_int]),
        "qso_cpg_step": (None, [dp, dp] + [C.c_double] * 9 + [dp, dp]),
        "qso_cpg_torque": (None, [dp, dp, dp, dp, C.c_double, dp, dp, C.c_double, C.c_double, C.c_int, dp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _out(n):
    a = np.zeros(n, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class World:
    """The part of the path pybullet provides: one Go1 on a plane."""

    def __init__(self, handle=None, **params):
        self.L = lib()
        self._own = handle is None
        self.h = self.L.qso_world_create() if handle is None else handle
        if params:
            self.set_params(**params)

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            self.L.qso_world_destroy(self.h)
            self.h = None

    def params(self):
        p = WorldParams()
        self.L.qso_world_get_params(self.h, C.byref(p))
        return p

    def set_params(self, **kw):
        p = self.params()
        for k, v in kw.items():
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, v)
        self.L.qso_world_set_params(self.h, C.byref(p))

    def set_state(self, s):
        a, p = _d(s)
        assert a.shape == (37,)
        self.L.qso_world_set_state(self.h, p)

    def get_state(self):
        a, p = _out(37)
        self.L.qso_world_get_state(self.h, p)
        return a

    def add_torque(self, tau):
        a, p = _d(tau)
        self.L.qso_world_add_torque(self.h, p)

    def step(self, tau=None):
        if tau is not None:
            self.add_torque(tau)
        self.L.qso_world_step(self.h)

    def contacts(self):
        out = []
        for i in range(self.L.qso_world_num_contacts(self.h)):
            link = C.c_int()
            nf = C.c_double()
            dist = C.c_double()
            pos, pp = _out(3)
            self.L.qso_world_get_contact(self.h, i, C.byref(link), C.byref(nf), C.byref(dist), pp)
            out.append((link.value, nf.value, dist.value, pos))
        return out

    def self_contacts(self, detect=False):
        """[(pyb link a, pyb link b, distance)] of the calf-involved self contacts of the last step's collision phase
        (detect=True: of the current state)"""
        if detect:
            self.L.qso_world_detect_self_contacts(self.h)
        out = []
        for i in range(self.L.qso_world_num_self_contacts(self.h)):
            a, b, d = C.c_int(), C.c_int(), C.c_double()
            self.L.qso_world_get_self_contact(self.h, i, C.byref(a), C.byref(b), C.byref(d))
            out.append((a.value, b.value, d.value))
        return out

    def set_mass(self, pyb_link, mass):
        """changeDynamics(robot, link, mass=) (quadruped.py:744-776)"""
        self.L.qso_world_set_mass(self.h, int(pyb_link), float(mass))

    def set_payload(self, mass, pos):
        """Quadruped._add_base_mass_offset (quadruped.py:778-819): block welded to the trunk at pos (base frame)"""
        p = np.ascontiguousarray(pos, dtype=np.float64)
        self.L.qso_world_set_payload(self.h, float(mass), p.ctypes.data_as(C.POINTER(C.c_double)))

    def dynamics(self, pyb_link):
        m = C.c_double()
        I, ip_ = _out(3)
        c, cp = _out(3)
        self.L.qso_world_get_dynamics(self.h, pyb_link, C.byref(m), ip_, cp)
        return m.value, I, c

    def mass_matrix(self):
        M, p = _out(18 * 18)
        self.L.qso_world_mass_matrix(self.h, p)
        return M.reshape(18, 18)

    def accel(self, tau=None):
        t, tp = _d(np.zeros(12) if tau is None else tau)
        a, ap = _out(18)
        self.L.qso_world_bias(self.h, tp, ap)
        return a

    def link_pose(self, link):
        R, rp = _out(9)
        p, pp = _out(3)
        self.L.qso_world_link_pose(self.h, link, rp, pp)
        return R.reshape(3, 3), p

    def energy(self):
        P, pp = _out(3)
        Lm, lp = _out(3)
        E = self.L.qso_world_energy(self.h, pp, lp)
        return E, P, Lm

    @property
    def last_iterations(self):
        return self.L.qso_world_last_iterations(self.h)

    @property
    def last_rows(self):
        a = C.c_int()
        b = C.c_int()
        self.L.qso_world_last_rows(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    @property
    def cone_clamped(self):
        return self.L.qso_world_cone_clamped(self.h)


class Env:
    """One-env restatement of QuadrupedGymEnv (quadruped_gym_env.py:41-256)."""

    def __init__(self, enable_springs=False, motor_control_mode="PD", action_space_mode="SYMMETRIC",
                 task_env="NO_TASK", observation_space_mode="ENCODER", action_repeat=10,
                 isRLGymInterface=True, enable_action_filter=False, enable_action_interpolation=False,
                 time_step=0.001, settling_steps=2500, landing_mode=0, rest_mode=0, **world_params):
        self.L = lib()
        c = EnvConfig()
        self.L.qso_env_default_config(C.byref(c))
        c.enable_springs = int(enable_springs)
        c.control_mode = CONTROL[motor_control_mode]
        c.action_mode = ACTION[action_space_mode]
        c.task = TASKS[task_env]
        c.obs_mode = OBS[observation_space_mode]
        c.action_repeat = action_repeat
        c.is_rl_interface = int(isRLGymInterface)
        c.enable_action_filter = int(enable_action_filter)
        c.enable_action_interpolation = int(enable_action_interpolation)
        c.time_step = time_step
        c.settling_steps = settling_steps
        c.landing_mode = int(landing_mode)
        c.rest_mode = int(rest_mode)
        self.h = self.L.qso_env_create(C.byref(c))
        self.world = World(handle=self.L.qso_env_world(self.h))
        if world_params:
            self.world.set_params(**world_params)
        self.action_dim = self.L.qso_env_action_dim(self.h)
        self.obs_dim = self.L.qso_env_obs_dim(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.qso_env_destroy(self.h)
            self.h = None

    def reset(self, mu=1.0):
        o, op = _out(self.obs_dim)
        self.L.qso_env_reset(self.h, float(mu), op)
        return o

    def reset_to_state(self, state37, mu=1.0):
        """reset() after set_robot_desired_state (quadruped_gym_env.py:288-289): no settle"""
        s, sp = _d(state37)
        assert s.shape == (37,)
        o, op = _out(self.obs_dim)
        self.L.qso_env_reset_to_state(self.h, float(mu), sp, op)
        return o

    def set_demo(self, actions):
        """demonstration actions [L, A] of the *_DEMO tasks (task_base.py:169-176)"""
        a = np.ascontiguousarray(actions, dtype=np.float64)
        assert a.ndim == 2 and a.shape[1] == self.action_dim
        self.L.qso_env_set_demo(self.h, a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0])

    def set_demo_counter(self, value):
        self.L.qso_env_set_demo_counter(self.h, int(value))

    def demo_counter(self):
        return int(self.L.qso_env_get_demo_counter(self.h))

    def step(self, action):
        a, ap = _d(action)
        assert a.shape == (self.action_dim,)
        o, op = _out(self.obs_dim)
        r = C.c_double()
        d = C.c_int()
        t = C.c_int()
        self.L.qso_env_step(self.h, ap, op, C.byref(r), C.byref(d), C.byref(t))
        return o, r.value, bool(d.value), bool(t.value)

    def task_state(self):
        o, op = _out(32)
        self.L.qso_env_get_task_state(self.h, op)
        return o

    def landing_state(self):
        """(mode, timer, end) of the landing controller (0 policy, 1 take-off hold, 2 landing, 3 spent)."""
        o, op = _out(3)
        self.L.qso_env_get_landing_state(self.h, op)
        return int(o[0]), o[1], o[2]

    def rest_state(self):
        """(active, h_actual, t_start) of the go-to-rest controller"""
        o, op = _out(3)
        self.L.qso_env_get_rest_state(self.h, op)
        return int(o[0]), o[1], o[2]

    def jump_arrays(self):
        """(fwd[n], perf[n], [jump_counter, good_jump_counter, first_jump, max_jump_height, end_jump]) of the continuous-jumping tasks (task_base.py:340-353)."""
        o, op = _out(1030)
        self.L.qso_env_get_jump_arrays(self.h, op)
        n = int(o[0])
        return o[6:6 + n].copy(), o[518:518 + n].copy(), o[1:6].copy()

    def last_action(self):
        o, op = _out(12)
        self.L.qso_env_get_last_action(self.h, op)
        return o[:self.action_dim]

    def torques(self):
        a, ap = _out(12)
        b, bp = _out(12)
        self.L.qso_env_get_torques(self.h, ap, bp)
        return a, b

    def set_springs(self, k3, b3, rest3):
        """nominal spring stiffness / damping / rest angle of hip, thigh, calf (env/springs.py:28-52)"""
        a, ap = _d(k3)
        b, bp = _d(b3)
        c, cp = _d(rest3)
        self.L.qso_env_set_springs(self.h, ap, bp, cp)

    def set_gains(self, kp, kd):
        a, ap = _d(np.broadcast_to(kp, (12,)).copy())
        b, bp = _d(np.broadcast_to(kd, (12,)).copy())
        self.L.qso_env_set_gains(self.h, ap, bp)


# --- thin functional wrappers over the analytic entry points -----------------
def action_to_command(action, enable_springs=True, control="PD", action_mode="SYMMETRIC", task="NO_TASK"):
    a, ap = _d(action)
    o, op = _out(12)
    lib().qso_action_to_command(int(enable_springs), CONTROL[control], ACTION[action_mode], TASKS[task], ap, op)
    return o


def pd_torque(kp, kd, tau_max, cmd, q, qd, torque_mode=False):
    args = [_d(np.broadcast_to(x, (12,)).copy()) for x in (kp, kd, tau_max, cmd, q, qd)]
    o, op = _out(12)
    lib().qso_pd_torque(*[a[1] for a in args], int(torque_mode), op)
    return o


def spring_torque(k3, b3, rest3, q, qd):
    args = [_d(x) for x in (k3, b3, rest3, q, qd)]
    o, op = _out(12)
    lib().qso_spring_torque(*[a[1] for a in args], op)
    return o


def fk_jacobian(q3, leg):
    a, ap = _d(q3)
    pos, pp = _out(3)
    J, jp = _out(9)
    lib().qso_fk_jacobian(ap, leg, pp, jp)
    return J.reshape(3, 3), pos


def ik(xyz, leg):
    a, ap = _d(xyz)
    q, qp = _out(3)
    lib().qso_ik(ap, leg, qp)
    return q


def rpy_from_quat(quat):
    a, ap = _d(quat)
    o, op = _out(3)
    lib().qso_rpy_from_quat(ap, op)
    return o


def backflip_pitch(quat, switched):
    a, ap = _d(quat)
    return lib().qso_backflip_pitch(ap, int(switched))


def cpg_step(X, PHI, mu, omega_swing, omega_stance, coupling, dt, des_step_len, robot_height,
             ground_clearance, ground_penetration):
    Xc, xp = _d(np.array(X, dtype=np.float64).reshape(8))
    P, pp = _d(np.array(PHI, dtype=np.float64).reshape(16))
    xs, xsp = _out(4)
    zs, zsp = _out(4)
    lib().qso_cpg_step(xp, pp, mu, omega_swing, omega_stance, coupling, dt, des_step_len, robot_height,
                       ground_clearance, ground_penetration, xsp, zsp)
    return Xc.reshape(2, 4), xs, zs


def cpg_torque(xs, zs, q, qd, foot_y, kp3, kd3, kpc, kdc, add_cartesian=True):
    args = [_d(x) for x in (xs, zs, q, qd)]
    k, kp_ = _d(kp3)
    d, kd_ = _d(kd3)
    o, op = _out(12)
    lib().qso_cpg_torque(args[0][1], args[1][1], args[2][1], args[3][1], foot_y, kp_, kd_, kpc, kdc,
                         int(add_cartesian), op)
    return o
