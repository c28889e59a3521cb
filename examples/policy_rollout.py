"""BASELINE config 5 shape: backflip task with an SB3-shaped MlpPolicy (2 x 64 tanh) evaluated on the
device every step -- observations never leave the GPU.

    python examples/policy_rollout.py [N] [steps]"""
import sys
import time

import torch

import quadruped_springs_b200 as qs
from quadruped_springs_b200 import stats


class MlpPolicy(torch.nn.Module):
    """actor of stable-baselines3's MlpPolicy: obs -> 64 tanh -> 64 tanh -> action mean"""

    def __init__(self, obs_dim, act_dim):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(obs_dim, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64),
                                       torch.nn.Tanh(), torch.nn.Linear(64, act_dim))

    @torch.no_grad()
    def forward(self, obs):
        return self.net(obs).clamp(-1, 1)


def main(n=32768, steps=200):
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, enable_springs=True, task_env="BACKFLIP", motor_control_mode="PD",
                                    action_space_mode="SYMMETRIC", observation_space_mode="ARS_BACKFLIP")
    torch.manual_seed(0)
    policy = MlpPolicy(env.obs_dim, env.action_dim).cuda()
    obs = env.reset()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        obs, reward, done, infos = env.step(policy(obs))
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"{n} envs x {steps} steps with policy inference in {dt:.2f} s -> {n * steps / dt / 1e6:.2f} M env-steps/s")
    print({k: round(v, 4) for k, v in stats.gather_rollout_stats(env.rollout_stats()).items() if k.startswith("mean")})


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:3]))
