"""Steady-state trace of the settle conveyor: entries pending and ticks of the late slice per step (QS_FILL=... to vary the target)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
import quadruped_springs_b200 as qs
from quadruped_springs_b200 import _lib

n = 65536
env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=0, auto_reset=True, enable_springs=True, task_env="JUMPING_FORWARD",
                                motor_control_mode="CARTESIAN_PD", observation_space_mode="ARS_BASIC")
L = _lib.lib()
env.reset()
g = torch.Generator(device="cuda").manual_seed(1)
pend, sl, slow = [], [], []
for i in range(400):
    env.step((torch.rand(n, 6, device="cuda", generator=g) * 2 - 1).contiguous())
    if i >= 250:
        cnt = (C.c_int32 * 4)()
        _lib.check(L.qs_debug_counters(env._h, cnt, None))
        pend.append(cnt[2]); sl.append(cnt[3]); slow.append(cnt[0])
import statistics as st
print("QS_FILL", os.environ.get("QS_FILL"), "pending mean/min/max", st.mean(pend), min(pend), max(pend), "late slice ticks mean/min/max", st.mean(sl), min(sl), max(sl),
      "slow envs mean", st.mean(slow), "window", 296 * 128)
