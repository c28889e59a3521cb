"""First-contact diagnostics on a GPU box: error magnitudes of the CUDA path vs the CPU oracle."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from quadruped_springs_b200 import BatchedQuadrupedGymEnv, ops

GOLDEN = os.path.join(ROOT, "tests", "golden")


def random_states(n, rng, contact=True):
    S = np.zeros((n, 37))
    for i in range(n):
        q = np.array([0, np.pi / 4, -np.pi / 2] * 4) + rng.normal(size=12) * 0.25
        quat = np.array([0, 0, 0, 1.0]) + rng.normal(size=4) * (0.05 if contact else 0.5)
        quat /= np.linalg.norm(quat)
        S[i, 3:7] = quat
        S[i, 13:25] = q
        S[i, 7:13] = rng.normal(size=6) * 0.5
        S[i, 25:37] = rng.normal(size=12) * 2
        S[i, 0:2] = rng.normal(size=2) * 0.1
        S[i, 2] = 1.0
    if contact:
        # drop every robot so that its lowest foot touches the ground (small penetration / gap)
        w = O.World()
        for i in range(n):
            w.set_state(S[i])
            zmin = min(w.link_pose(l)[1][2] for l in (6, 10, 14, 18)) - 0.02
            S[i, 2] += -zmin + rng.uniform(-0.002, 0.0005)
    return S


def tick_parity(n_ticks, contact, use_f64, n=256, seed=0):
    rng = np.random.default_rng(seed)
    S = random_states(n, rng, contact).astype(np.float32).astype(np.float64)
    tau = (rng.normal(size=(n, 12)) * 5).astype(np.float32).astype(np.float64)
    mu = rng.uniform(0.5, 1.0, size=n).astype(np.float32).astype(np.float64)
    env = BatchedQuadrupedGymEnv(num_envs=n, enable_springs=True, task_env="JUMPING_IN_PLACE",
                                 observation_space_mode="ARS_BASIC", enable_noise=False, auto_reset=False)
    env.set_state(torch.tensor(S, dtype=torch.float32))
    env._views["mu"][:] = torch.tensor(mu, dtype=torch.float32, device="cuda")
    env.debug_ticks(torch.tensor(tau, dtype=torch.float32), n_ticks, use_f64)
    got = env.get_state().cpu().numpy().astype(np.float64)
    contact_bits = env._views["contact"].cpu().numpy()
    ff = env._views["foot_force"].cpu().numpy().T
    ref = np.zeros_like(S)
    w = O.World()
    refc = np.zeros(n, dtype=int)
    reff = np.zeros((n, 4))
    iters = []
    for i in range(n):
        w.set_params(mu_ground=mu[i])
        w.set_state(S[i])
        for _ in range(n_ticks):
            w.step(tau[i])
        ref[i] = w.get_state()
        iters.append(w.last_iterations)
        for link, nf, dist, pos in w.contacts():
            if link in (5, 9, 13, 17):
                k = (link - 5) // 4
                refc[i] |= 1 << k
                reff[i, k] += nf
    err = np.abs(got - ref)
    names = {"pos": slice(0, 3), "quat": slice(3, 7), "vlin": slice(7, 10), "vang": slice(10, 13), "q": slice(13, 25),
             "qd": slice(25, 37)}
    res = {k: float(err[:, v].max()) for k, v in names.items()}
    res["contact_mismatch"] = int(((contact_bits & 15) != refc).sum())
    res["force_err"] = float(np.abs(ff - reff).max())
    res["force_max"] = float(reff.max())
    res["mean_contacts"] = float(np.mean([bin(c).count("1") for c in refc]))
    res["mean_iters"] = float(np.mean(iters))
    worst = int(err[:, 25:37].max(axis=1).argmax())
    res["worst_env"] = worst
    return res


def rollout_parity(name, max_steps=200):
    g = np.load(os.path.join(GOLDEN, f"rollout_{name}.npz"))
    cfg = json.loads(str(g["cfg"]))
    env = BatchedQuadrupedGymEnv(num_envs=4, enable_noise=False, auto_reset=False, env_randomizer_mode="NO_RANDOMIZER",
                                 solver=dict(mu_ground=float(g["mu"])), **cfg)
    obs = env.reset()
    e0 = np.abs(env.get_state().cpu().numpy()[0] - g["init_state"]).max()
    eo = np.abs(obs.cpu().numpy()[0] - g["init_obs"]).max()
    out = {"init_state_err": float(e0), "init_obs_err": float(eo), "steps": []}
    T = min(len(g["reward"]), max_steps)
    for t in range(T):
        a = torch.tensor(g["actions"][t], dtype=torch.float32, device="cuda").expand(4, -1)
        obs, r, d, info = env.step(a)
        es = np.abs(env.get_state().cpu().numpy()[0] - g["state"][t]).max()
        eo = np.abs(obs.cpu().numpy()[0] - g["obs"][t]).max()
        er = abs(float(r[0]) - float(g["reward"][t]))
        out["steps"].append((t, float(es), float(eo), float(er), bool(d[0]), bool(g["done"][t])))
    return out


if __name__ == "__main__":
    torch.cuda.init()
    print("device", torch.cuda.get_device_name(0))
    for use_f64 in (1, 0):
        for contact in (False, True):
            for nt in (1, 10):
                t = time.time()
                r = tick_parity(nt, contact, use_f64)
                print(f"f64={use_f64} contact={contact} ticks={nt}: " +
                      " ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in r.items()), flush=True)
    for name in ("jip_random", "jip_jump", "jf_cartesian_random", "backflip", "jip_ppo_hp_filter"):
        r = rollout_parity(name)
        print(name, "init", r["init_state_err"], r["init_obs_err"])
        for s in r["steps"]:
            if s[0] % 10 == 0 or s[4] or s[5]:
                print("   t=%d state_err=%.3g obs_err=%.3g rew_err=%.3g done=%s/%s" % s)
    # throughput first look
    for n in (4096, 65536):
        env = BatchedQuadrupedGymEnv(num_envs=n, enable_springs=True, task_env="JUMPING_IN_PLACE",
                                     observation_space_mode="ARS_BASIC", auto_reset=False)
        t = time.time(); env.reset(); torch.cuda.synchronize(); print(f"N={n} reset {time.time()-t:.3f}s")
        a = torch.rand(n, 6, device="cuda") * 2 - 1
        for _ in range(3):
            env.step(a)
        torch.cuda.synchronize()
        t = time.time()
        for _ in range(20):
            a = torch.rand(n, 6, device="cuda") * 2 - 1
            env.step(a)
        torch.cuda.synchronize()
        dt = (time.time() - t) / 20
        print(f"N={n} step {dt*1e3:.3f} ms -> {n/dt/1e6:.2f} M env-steps/s")
