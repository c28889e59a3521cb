import sys, time, torch
sys.path.insert(0, "/root/repo")
import quadruped_springs_b200 as qs
n = 32768
venv = qs.BatchedVecEnv(num_envs=n, enable_springs=True, task_env="BACKFLIP", observation_space_mode="ARS_BACKFLIP", landing_wrapper="LandingWrapperBackflip")
vn = qs.VecNormalizeTorch(venv, training=True, norm_reward=False)
pol = qs.MlpPolicyTorch(venv.env.obs_dim, venv.env.action_dim).cuda()
obs = vn.reset()
def timeit(f, k=200):
    for _ in range(50): f()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.time() - t0) / k * 1e3
a = torch.zeros(n, 6, device="cuda")
state = {"obs": obs}
print("env.step only           ms", timeit(lambda: venv.env.step(a)))
print("policy only             ms", timeit(lambda: pol.predict(state["obs"])))
def full():
    o, r, d, i = vn.step(pol.predict(state["obs"])); state["obs"] = o
print("full loop (training)    ms", timeit(full))
vn.training = False
print("full loop (eval)        ms", timeit(full))
def raw():
    o, r, d, i = venv.step_tensor(pol.predict(state["obs"])); state["obs"] = o
print("policy + step, no norm  ms", timeit(raw))
