"""BASELINE config 5 shape: backflip task with an SB3-shaped MlpPolicy (2 x 64 tanh, separate value net) and VecNormalize
evaluated on the device every step -- observations never leave the GPU (load_model.py:109-134 on tensors).

    python examples/policy_rollout.py [N] [steps] [best_model.zip]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import quadruped_springs_b200 as qs
from quadruped_springs_b200 import stats


def main(n=32768, steps=200, model_zip=None):
    venv = qs.BatchedVecEnv(num_envs=n, enable_springs=True, task_env="BACKFLIP", motor_control_mode="PD",
                            action_space_mode="SYMMETRIC", observation_space_mode="ARS_BACKFLIP",
                            landing_wrapper="LandingWrapperBackflip")
    env = qs.VecNormalizeTorch(venv, training=True, norm_reward=False)
    torch.manual_seed(0)
    if model_zip:   # a stable-baselines3 PPO.save() archive; SB3 itself is not needed
        policy = qs.MlpPolicyTorch.from_sb3_zip(model_zip, device="cuda")
    else:
        policy = qs.MlpPolicyTorch(venv.env.obs_dim, venv.env.action_dim).cuda()
    obs = env.reset()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        obs, reward, done, infos = env.step(policy.predict(obs, deterministic=True))
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"{n} envs x {steps} steps with policy inference + normalisation in {dt:.2f} s -> {n * steps / dt / 1e6:.2f} M env-steps/s")
    print({k: round(v, 4) for k, v in stats.gather_rollout_stats(venv.env.rollout_stats()).items() if k.startswith("mean")})


if __name__ == "__main__":
    a = sys.argv[1:]
    main(*(int(x) for x in a[:2]), *(a[2:3]))
