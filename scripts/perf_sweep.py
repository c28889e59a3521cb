"""Quick perf sweep on a GPU box: step-only, reset-only and steady-state (auto-reset) timings per block size."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadruped_springs_b200 as qs

W = dict(enable_springs=True, task_env="JUMPING_FORWARD", motor_control_mode="CARTESIAN_PD",
         action_space_mode="SYMMETRIC", observation_space_mode="ARS_BASIC")
N = int(os.environ.get("N", 65536))
blocks = [int(b) for b in os.environ.get("BLOCKS", "64,128,256").split(",")]
for B in blocks:
    env = qs.BatchedQuadrupedGymEnv(num_envs=N, auto_reset=False, block_size=B, **W)
    torch.cuda.synchronize(); t = time.time(); env.reset(); torch.cuda.synchronize(); t_reset = time.time() - t
    acts = [torch.rand(N, 6, device="cuda") * 2 - 1 for _ in range(8)]
    for i in range(5): env.step(acts[i % 8])
    torch.cuda.synchronize(); t = time.time()
    for i in range(30): env.step(acts[i % 8])
    torch.cuda.synchronize(); t_step = (time.time() - t) / 30
    env.close()
    env = qs.BatchedQuadrupedGymEnv(num_envs=N, auto_reset=True, block_size=B, **W)
    env.reset()
    for i in range(40): env.step(acts[i % 8])
    torch.cuda.synchronize(); t = time.time()
    for i in range(100): env.step(acts[i % 8])
    torch.cuda.synchronize(); t_auto = (time.time() - t) / 100
    st = env.rollout_stats().cpu()
    print(f"block={B}: reset_all {t_reset*1e3:.1f} ms ({N*2500/t_reset/1e6:.0f} M ticks/s) | step(no reset) {t_step*1e3:.3f} ms "
          f"({N/t_step/1e6:.1f} M steps/s) | steady auto-reset {t_auto*1e3:.3f} ms ({N/t_auto/1e6:.2f} M steps/s) "
          f"episodes={int(st[1])} mean_len={float(st[10]/max(st[1],1)):.1f}", flush=True)
    env.close()
