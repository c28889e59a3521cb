"""Why are steps 1-3 after a host synchronisation ~1 ms slower?  Per-step, per-kernel device times around a sync."""
import ctypes as C
import sys
import time

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
act = lambda: (torch.rand(n, 6, device="cuda", generator=g) * 2 - 1).contiguous()
for _ in range(250):
    env.step(act())
torch.cuda.synchronize()


def series(k):
    out = []
    for name in ("qs_step_kernel_time", "qs_settle_kernel_time", "qs_slow_kernel_time"):
        f = getattr(L, name)
        prev, row = 0.0, []
        for j in range(1, k + 1):
            v = C.c_float()
            _lib.check(f(env._h, j, C.byref(v)))
            row.append(v.value - prev)
            prev = v.value
        out.append([round(x, 2) for x in row[::-1]])
    return out


def run(label, k=10, pre=None, flush=None):
    if pre:
        pre()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
    for i in range(k):
        if flush is not None:
            flush.fill_(float(i))
        a = act()
        ev[i][0].record()
        env.step(a)
        ev[i][1].record()
    torch.cuda.synchronize()
    cnt = (C.c_int32 * 4)()
    _lib.check(L.qs_debug_counters(env._h, cnt, None))
    s = series(k)
    print(label, "step", [round(a.elapsed_time(b), 2) for a, b in ev])
    print("   k_step+contact", s[0]); print("   settle", s[1]); print("   slow", s[2], "counters", list(cnt))


flush = torch.empty(192 * 1024 * 1024 // 4, device="cuda")
run("after sync, no flush")
run("after sync + 20 ms sleep", pre=lambda: time.sleep(0.02))
run("after sync, with flush", flush=flush)
w = (C.c_uint64 * 3)()
run("after qs_work_counters", pre=lambda: _lib.check(L.qs_work_counters(env._h, w, None)))
run("after qs_settle_work_counters", pre=lambda: _lib.check(L.qs_settle_work_counters(env._h, w, None)))
run("after .item()", pre=lambda: env._done.sum().item())
run("plain again")
