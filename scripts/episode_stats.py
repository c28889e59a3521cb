import os, sys, time, ctypes as C
import torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadruped_springs_b200 as qs
from quadruped_springs_b200 import _lib
W = dict(enable_springs=True, task_env="JUMPING_FORWARD", motor_control_mode="CARTESIAN_PD",
         action_space_mode="SYMMETRIC", observation_space_mode="ARS_BASIC")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = qs.BatchedQuadrupedGymEnv(num_envs=N, auto_reset=True, **W)
env.reset()
L = _lib.lib()
age = torch.zeros(N, dtype=torch.int32, device="cuda")
lens = []
cnt = (C.c_int32 * 4)()
log = []
for t in range(400):
    a = torch.rand(N, 6, device="cuda") * 2 - 1
    torch.cuda.synchronize(); t0 = time.time()
    obs, r, d, info = env.step(a)
    torch.cuda.synchronize(); dt = time.time() - t0
    L.qs_debug_counters(env._h, cnt, None)
    age += 1
    lens.append(age[d].cpu().numpy())
    age[d] = 0
    log.append((t, int(d.sum()), cnt[0], cnt[1], cnt[2], cnt[3], dt * 1e3))
lens = np.concatenate(lens)
print("episodes", len(lens), "mean", lens.mean(), "min", lens.min(), "quantiles 0.1%,1%,5%,25%,50%:", np.quantile(lens, [0.001, 0.01, 0.05, 0.25, 0.5]))
print("hist <=5,<=10,<=15,<=20:", [(lens <= k).mean() for k in (5, 10, 15, 20)])
for row in log:
    if row[0] < 60 or row[0] % 10 == 0 or row[3] > 0:
        print("t=%d done=%d slow=%d urgent=%d conveyor=%d slice=%d ms=%.2f" % row)
