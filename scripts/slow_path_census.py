"""Who takes the general-solver path in the bench workload: joint limits only, or body contacts?"""
import ctypes as C
import numpy as np
import torch
import quadruped_springs_b200 as qs
from quadruped_springs_b200 import _lib

n = 65536
env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=1, enable_springs=True, task_env="JUMPING_FORWARD",
                                motor_control_mode="CARTESIAN_PD", action_space_mode="SYMMETRIC",
                                observation_space_mode="ARS_BASIC")
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
# URDF joint limits (go1.urdf, csrc/qs_model_host.h)
lo = torch.tensor([-1.0471975512, -0.663225115758, -2.72271363311] * 4, device="cuda")
hi = torch.tensor([1.0471975512, 2.96705972839, -0.837758040957] * 4, device="cuda")
print("limits", lo[:3].tolist(), hi[:3].tolist())
acc = np.zeros(5)
cnt = (C.c_int32 * 4)()
for t in range(400):
    a = torch.rand(n, env.action_dim, device="cuda", generator=g) * 2 - 1
    env.step(a)
    if t < 100:
        continue
    _lib.check(env._L.qs_debug_counters(env._h, cnt, None))
    q = env.robot.GetMotorAngles()
    at_lim = ((q <= lo) | (q >= hi)).any(1)
    ninv = env.robot.GetContactInfo()[1] > 0
    acc += [cnt[0], at_lim.sum().item(), ninv.sum().item(), (at_lim & ~ninv).sum().item(), 1]
acc /= acc[4]
print(f"per step: slow-path envs {acc[0]:.0f}; at a joint limit at step end {acc[1]:.0f}; body contact at step end {acc[2]:.0f}; "
      f"limit without body contact {acc[3]:.0f}")
