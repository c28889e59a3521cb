"""BASELINE config 4: Go1 + PEA driven by the Hopf CPG (hopf_network.py:176-289) for N robots.

    python examples/cpg_trot.py [N] [ticks]

Per 1 ms tick: cpg.update -> desired foot xz -> IK + joint PD + Cartesian impedance (kernel K4) ->
torques -> env.step (TORQUE mode, action_repeat = 1), exactly the reference's __main__ loop."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import quadruped_springs_b200 as qs


def main(n=4096, ticks=2000, gait="TROT"):
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, isRLGymInterface=False, time_step=0.001, action_repeat=1,
                                    motor_control_mode="TORQUE", enable_springs=True, auto_reset=False)
    env.reset()
    cpg = qs.HopfNetwork(num_envs=n, gait=gait, omega_swing=16 * np.pi, omega_stance=4 * np.pi, time_step=0.001, seed=0)
    x0 = env.robot.GetBasePosition()[:, 0].clone()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(ticks):
        xs, zs, tau = cpg.update(env.robot.GetMotorAngles(), env.robot.GetMotorVelocities())
        env.step(tau)
    torch.cuda.synchronize()
    dt = time.time() - t0
    pos = env.robot.GetBasePosition()
    print(f"{n} robots x {ticks} ticks in {dt:.2f} s -> {n * ticks / dt / 1e6:.2f} M ticks/s "
          f"({n * ticks / dt / 10 / 1e6:.2f} M control-step equivalents/s)")
    print(f"mean forward distance {float((pos[:, 0] - x0).mean()):.3f} m, mean height {float(pos[:, 2].mean()):.3f} m")
    return env, cpg


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:3]))
