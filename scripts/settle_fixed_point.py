"""Does the fp32 settle reach an exact fixed point (or 2-cycle) before tick 2500?
For a range of T: reset with settling_steps = T, T+1, T+2 (same seed -> same friction draw)
and count the envs whose complete state (pose, velocities, contact impulses) repeats."""
import sys
import torch
import quadruped_springs_b200 as qs

n = 4096
cfgs = {"pea_cart": dict(enable_springs=True, task_env="JUMPING_FORWARD", motor_control_mode="CARTESIAN_PD"),
        "pea_pd": dict(enable_springs=True, task_env="JUMPING_IN_PLACE", motor_control_mode="PD"),
        "nosprings_pd": dict(enable_springs=False, task_env="JUMPING_IN_PLACE", motor_control_mode="PD")}


def settled(T, cfg):
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=3, enable_noise=False, auto_reset=False,
                                    action_space_mode="SYMMETRIC", observation_space_mode="ARS_BASIC",
                                    solver=dict(settling_steps=T), **cfg)
    env.reset()
    s = torch.cat([env.get_state(), env._views["foot_force"].t().clone(), env._views["contact"][:, None].float()], 1).clone()
    env.close()
    return s


for name, cfg in cfgs.items():
    for T in (400, 600, 800, 1000, 1250, 1500, 2000, 2400):
        a, b, c = settled(T, cfg), settled(T + 1, cfg), settled(T + 2, cfg)
        fix = (a == b).all(1)
        cyc2 = (a == c).all(1)
        dmax = (a - b).abs().max().item()
        print(f"{name} T={T}: fixed {fix.float().mean().item():.3f}  2-cycle {cyc2.float().mean().item():.3f}  max|s(T+1)-s(T)| {dmax:.3e}", flush=True)
