"""Per control step: how many envs touch the ground at all, and for how many of the 10 ticks."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadruped_springs_b200 as qs
n = 65536
env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=1, enable_springs=True, task_env="JUMPING_FORWARD",
                                motor_control_mode="CARTESIAN_PD", action_space_mode="SYMMETRIC",
                                observation_space_mode="ARS_BASIC")
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
hist = np.zeros(42)
start_mask_vs_any = np.zeros((2, 2))
for t in range(300):
    a = torch.rand(n, env.action_dim, device="cuda", generator=g) * 2 - 1
    w0 = env._views["work"].clone()
    m0 = (env._views["contact"] & 15) != 0
    env.step(a)
    if t < 100:
        continue
    dw = (env._views["work"] - w0)
    ticks, contacts = dw[0], dw[1]
    ok = ticks == 10                      # envs that ran the whole step in k_step
    c = contacts[ok].clamp(max=41).cpu().numpy()
    hist += np.bincount(c, minlength=42)
    anyc = contacts[ok] > 0
    mm = m0[ok]
    for i in (0, 1):
        for j in (0, 1):
            start_mask_vs_any[i, j] += ((mm == bool(i)) & (anyc == bool(j))).sum().item()
hist /= hist.sum()
print("P(no foot contact in the whole step) = %.3f" % hist[0])
print("contact-ticks per step (sum over feet and ticks) distribution:", np.round(hist[:41], 3).tolist())
print("rows: contact mask at step start (0/1); cols: any contact during the step (0/1)")
print(start_mask_vs_any / start_mask_vs_any.sum())
