"""Workload for compute-sanitizer (racecheck / memcheck / synccheck): 1024 envs, auto-reset with the settle conveyor, the
urgent path forced by a short slice cap, device steps (CUDA graph and direct) and host-buffer steps (qs_step_host)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import quadruped_springs_b200 as qs

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n = int(os.environ.get("QS_SAN_ENVS", "1024"))
settle = int(os.environ.get("QS_SAN_SETTLE", "2500"))   # racecheck is ~1000x slower than native: shorten reset()'s settle there
env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=3, auto_reset=True, enable_springs=True, task_env="JUMPING_FORWARD",
                                motor_control_mode="CARTESIAN_PD", observation_space_mode="ARS_BASIC", solver=dict(settling_steps=settle))
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
dones = 0
for t in range(steps):
    o, r, d, _ = env.step(torch.rand(n, 6, device="cuda", generator=g) * 2 - 1)
    dones += int(d.sum())
rng = np.random.default_rng(0)
for t in range(steps // 3):
    o, r, d, tr = env.step_host(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
    dones += int(d.sum())
mask = np.zeros(n, np.uint8); mask[::5] = 1
env.reset_host(mask)
st = qs.stats.gather_rollout_stats(env.rollout_stats())
torch.cuda.synchronize()
print(f"sanitize_run ok: {steps} device steps + {steps // 3} host steps, {dones} episode ends, episodes {st['episodes']:.0f}, "
      f"graph={'off' if os.environ.get('QS_GRAPH') == '0' else 'on'}, slice max {os.environ.get('QS_SETTLE_SLICE_MAX', 'default')}")
