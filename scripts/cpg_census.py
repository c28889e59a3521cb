"""Who is in the general solver under the CPG workload (BASELINE config 4)?"""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import quadruped_springs_b200 as qs
from quadruped_springs_b200 import _lib

n = 16384
lo = torch.tensor([-1.0471975512, -0.663225115758, -2.72271363311] * 4, device="cuda")
hi = torch.tensor([1.0471975512, 2.96705972839, -0.837758040957] * 4, device="cuda")
for gait in ("TROT", "BOUND", "WALK", "PACE"):
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, isRLGymInterface=False, time_step=0.001, action_repeat=1,
                                    motor_control_mode="TORQUE", enable_springs=True, auto_reset=True, seed=0)
    env.reset()
    cpg = qs.HopfNetwork(num_envs=n, gait=gait, omega_swing=16 * np.pi, omega_stance=4 * np.pi, time_step=0.001, seed=0)
    cnt = (C.c_int32 * 4)()
    for t in (500, 1000, 2000, 4000):
        cpg.drive(env, 500 if t == 500 else t - prev)
        prev = t
        _lib.check(env._L.qs_debug_counters(env._h, cnt, None))
        q = env.robot.GetMotorAngles()
        at_lim = ((q <= lo) | (q >= hi)).any(1).float().mean().item()
        ninv = (env.robot.GetContactInfo()[1] > 0).float().mean().item()
        z = env.robot.GetBasePosition()[:, 2]
        x = env.robot.GetBasePosition()[:, 0]
        up = env.robot.GetBaseOrientationMatrix()[:, 2, 2] if hasattr(env.robot, "GetBaseOrientationMatrix") else None
        print(f"{gait} t={t}: slow-list {cnt[0] / n:.3f}  at-limit {at_lim:.3f}  invalid-contact {ninv:.3f}  z mean {z.mean():.3f} min {z.min():.3f}  "
              f"x mean {x.mean():.3f}  upright(R22>0.85) {(up > 0.85).float().mean().item() if up is not None else -1:.3f}", flush=True)
    env.close()
