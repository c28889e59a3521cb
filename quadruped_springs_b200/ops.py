"""Stand-alone batched kernels (K3) behind the Quadruped / motor-model accessor
surface: every function takes and returns CUDA tensors and calls the C ABI."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _s(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _f32(x, n=None):
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float32), (n,)) if n else np.asarray(x, dtype=np.float32))
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _cuda(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("quadruped_springs_b200.ops works on CUDA tensors only (no CPU fallback)")
    return t.to(torch.float32).contiguous()


def pd_pea_torque(cmd, q, qd, kp, kd, tau_max, springs=None, torque_mode=False):
    """QuadrupedMotorModel.convert_to_torque + compute_spring_torques
    (quadruped_motor.py:45-104).  springs = (k3, b3, rest3) or None."""
    cmd, q, qd = _cuda(cmd), _cuda(q), _cuda(qd)
    n = q.shape[0]
    kp_a, kp_p = _f32(kp, 12)
    kd_a, kd_p = _f32(kd, 12)
    tm_a, tm_p = _f32(tau_max, 12)
    sp_p = None
    if springs is not None:
        sp_a, sp_p = _f32(np.concatenate([np.asarray(s, dtype=np.float32).reshape(3) for s in springs]))
    tau_m = torch.empty_like(q)
    tau_s = torch.empty_like(q)
    _lib.check(_lib.lib().qs_pd_pea_torque(_p(cmd), _p(q), _p(qd), kp_p, kd_p, tm_p, sp_p, int(torque_mode),
                                           _p(tau_m), _p(tau_s), n, _s(q)))
    return tau_m, tau_s


def fk_jacobian(q, qd=None):
    """ComputeJacobianAndPosition / ComputeFeetPosAndVel (quadruped.py:348-397,440-449):
    q [N,12] -> pos [N,12], J [N,4,3,3], vel [N,12] (None without qd)."""
    q = _cuda(q)
    n = q.shape[0]
    qd = _cuda(qd) if qd is not None else None
    pos = torch.empty(n, 12, device=q.device)
    jac = torch.empty(n, 4, 3, 3, device=q.device)
    vel = torch.empty(n, 12, device=q.device) if qd is not None else None
    _lib.check(_lib.lib().qs_fk_jacobian(_p(q), _p(qd), _p(pos), _p(jac), _p(vel), n, _s(q)))
    return pos, jac, vel


def inverse_kinematics(xyz):
    """ComputeInverseKinematics (quadruped.py:399-438): xyz [N,12] (leg frame) -> q [N,12]."""
    xyz = _cuda(xyz)
    out = torch.empty_like(xyz)
    _lib.check(_lib.lib().qs_ik(_p(xyz), _p(out), xyz.shape[0], _s(xyz)))
    return out


def make_config(enable_springs=True, motor_control_mode="PD", action_space_mode="SYMMETRIC", task_env="NO_TASK",
                observation_space_mode="ENCODER", **kw):
    from .env import ActionInterfaceCollection, MotorInterfaceCollection, SensorCollection, TaskCollection
    cfg = _lib.QsConfig()
    _lib.lib().qs_default_config(C.byref(cfg))
    cfg.enable_springs = int(enable_springs)
    cfg.control_mode = MotorInterfaceCollection().get_el(motor_control_mode)
    cfg.action_mode = ActionInterfaceCollection().get_el(action_space_mode)
    cfg.task = TaskCollection().get_el(task_env)
    cfg.obs_mode = SensorCollection().get_el(observation_space_mode)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def action_to_command(action, **cfg_kw):
    """ActionWrapper._transform_action_to_motor_command (interface_base.py:162-164)."""
    cfg = make_config(**cfg_kw)
    a = _cuda(action)
    out = torch.empty(a.shape[0], 12, device=a.device)
    _lib.check(_lib.lib().qs_action_to_command(C.byref(cfg), _p(a), _p(out), a.shape[0], _s(a)))
    return out


def obs_noise_std(**cfg_kw):
    """host-only: per-element sensor noise of an observation mode"""
    cfg = make_config(**cfg_kw)
    out = (C.c_float * _lib.QS_MAX_OBS)()
    _lib.check(_lib.lib().qs_obs_noise_std(C.byref(cfg), out))
    n = _lib.lib().qs_config_obs_dim(C.byref(cfg))
    return np.array(out[:n], dtype=np.float32)
