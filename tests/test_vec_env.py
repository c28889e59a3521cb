"""SB3-shaped front ends: VecNormalize statistics and the MlpPolicy loader are checked on
the CPU against numpy restatements of stable-baselines3's documented arithmetic; the VecEnv protocol, the terminal
observation and policy-in-the-loop rollouts run on the GPU."""
import io
import json
import zipfile

import numpy as np
import pytest
import torch

from quadruped_springs_b200.vec_env import BatchedVecEnv, MlpPolicyTorch, RunningMeanStdTorch, VecNormalizeTorch


class _NpRunningMeanStd:
    """stable_baselines3/common/running_mean_std.py, restated"""

    def __init__(self, shape=(), epsilon=1e-4):
        self.mean, self.var, self.count = np.zeros(shape), np.ones(shape), epsilon

    def update(self, arr):
        bm, bv, bc = arr.mean(axis=0), arr.var(axis=0), arr.shape[0]
        delta = bm - self.mean
        tot = self.count + bc
        new_mean = self.mean + delta * bc / tot
        m2 = self.var * self.count + bv * bc + np.square(delta) * self.count * bc / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class _FakeVenv:
    """tensor half of BatchedVecEnv fed from a script"""

    def __init__(self, obs, rew, done):
        self.obs, self.rew, self.done, self.t = obs, rew, done, 0
        self.num_envs = obs.shape[1]
        self.observation_space = self.action_space = None
        self.device = torch.device("cpu")
        self.env = type("E", (), {"obs_dim": obs.shape[2]})()

    def reset_tensor(self):
        self.t = 0
        return torch.as_tensor(self.obs[0], dtype=torch.float32)

    def step_tensor(self, actions):
        self.t += 1
        o = torch.as_tensor(self.obs[self.t], dtype=torch.float32)
        return o, torch.as_tensor(self.rew[self.t], dtype=torch.float32), torch.as_tensor(self.done[self.t]), \
            {"terminal_observation": o.clone()}


def test_running_mean_std_matches_sb3_arithmetic():
    rng = np.random.default_rng(0)
    a, b = _NpRunningMeanStd((5,)), RunningMeanStdTorch((5,), "cpu")
    for _ in range(20):
        x = rng.normal(2.0, 3.0, size=(64, 5))
        a.update(x)
        b.update(torch.as_tensor(x))
    np.testing.assert_allclose(b.mean.numpy(), a.mean, rtol=1e-12)
    np.testing.assert_allclose(b.var.numpy(), a.var, rtol=1e-12)
    assert b.count == pytest.approx(a.count)


def test_vecnormalize_matches_sb3_semantics():
    """VecNormalize.step_wait: obs_rms.update(obs); returns = returns * gamma + r; ret_rms.update(returns);
    r = clip(r / sqrt(ret_rms.var + eps)); returns[done] = 0; obs = clip((obs - mean) / sqrt(var + eps))."""
    rng = np.random.default_rng(1)
    T, n, o = 30, 16, 7
    obs = rng.normal(1.0, 4.0, size=(T + 1, n, o)).astype(np.float32)
    rew = rng.normal(0.5, 2.0, size=(T + 1, n)).astype(np.float32)
    done = rng.random((T + 1, n)) < 0.1
    vn = VecNormalizeTorch(_FakeVenv(obs, rew, done), clip_obs=5.0, clip_reward=3.0, gamma=0.97)
    orms, rrms, ret = _NpRunningMeanStd((o,)), _NpRunningMeanStd(()), np.zeros(n)
    got = vn.reset()
    orms.update(obs[0].astype(np.float64))
    np.testing.assert_allclose(got.numpy(), np.clip((obs[0] - orms.mean) / np.sqrt(orms.var + 1e-8), -5, 5), rtol=1e-5, atol=1e-6)
    for t in range(1, T + 1):
        g_obs, g_rew, g_done, _ = vn.step(None)
        orms.update(obs[t].astype(np.float64))
        ret = ret * 0.97 + rew[t]
        rrms.update(ret)
        np.testing.assert_allclose(g_rew.numpy(), np.clip(rew[t] / np.sqrt(rrms.var + 1e-8), -3, 3), rtol=1e-5, atol=1e-6)
        ret[done[t]] = 0
        np.testing.assert_allclose(g_obs.numpy(), np.clip((obs[t] - orms.mean) / np.sqrt(orms.var + 1e-8), -5, 5),
                                   rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(vn.get_original_obs().numpy(), obs[t])
    # evaluation mode of load_model.py:114-116: statistics frozen, rewards untouched
    vn.training, vn.norm_reward = False, False
    mean0 = vn.obs_rms.mean.clone()
    vn.venv.t = 0
    _, r, _, _ = vn.step(None)
    assert torch.equal(vn.obs_rms.mean, mean0) and torch.equal(r, torch.as_tensor(rew[1]))
    # state round trip
    vn2 = VecNormalizeTorch(_FakeVenv(obs, rew, done))
    vn2.load_state_dict(vn.state_dict())
    x = torch.as_tensor(obs[3])
    assert torch.equal(vn2.normalize_obs(x), vn.normalize_obs(x))
    np.testing.assert_allclose(vn.unnormalize_obs(vn.normalize_obs(x * 0.1)).numpy(), obs[3] * 0.1, rtol=1e-4, atol=1e-4)


def _sb3_like_zip(tmp_path, obs_dim=27, act_dim=6, arch=(64, 64), net_arch_json=None):
    """an archive laid out like PPO.save(): `data` (JSON) + `policy.pth` (state dict with SB3's key names)"""
    g = torch.Generator().manual_seed(0)
    sd, last = {}, obs_dim
    for tag in ("policy_net", "value_net"):
        last = obs_dim
        for i, s in enumerate(arch):
            sd[f"mlp_extractor.{tag}.{2 * i}.weight"] = torch.randn(s, last, generator=g) * 0.3
            sd[f"mlp_extractor.{tag}.{2 * i}.bias"] = torch.randn(s, generator=g) * 0.1
            last = s
    sd["action_net.weight"] = torch.randn(act_dim, last, generator=g) * 0.3
    sd["action_net.bias"] = torch.randn(act_dim, generator=g) * 0.1
    sd["value_net.weight"] = torch.randn(1, last, generator=g)
    sd["value_net.bias"] = torch.zeros(1)
    sd["log_std"] = torch.full((act_dim,), -0.5)
    path = tmp_path / "best_model.zip"
    buf = io.BytesIO()
    torch.save(sd, buf)
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps({"policy_kwargs": {"net_arch": net_arch_json} if net_arch_json else {}}))
        z.writestr("policy.pth", buf.getvalue())
        z.writestr("_stable_baselines3_version", "1.6.2")
    return path, sd


@pytest.mark.parametrize("arch,na", [((64, 64), None), ((128, 32), [dict(pi=[128, 32], vf=[128, 32])]),
                                     ((32,), dict(pi=[32], vf=[32]))])
def test_mlp_policy_loads_sb3_archive_and_matches_manual_forward(tmp_path, arch, na):
    path, sd = _sb3_like_zip(tmp_path, arch=arch, net_arch_json=na)
    pol = MlpPolicyTorch.from_sb3_zip(str(path), device="cpu")
    x = torch.randn(33, 27, generator=torch.Generator().manual_seed(1))
    h = v = x
    for i in range(len(arch)):
        h = torch.tanh(h @ sd[f"mlp_extractor.policy_net.{2 * i}.weight"].T + sd[f"mlp_extractor.policy_net.{2 * i}.bias"])
        v = torch.tanh(v @ sd[f"mlp_extractor.value_net.{2 * i}.weight"].T + sd[f"mlp_extractor.value_net.{2 * i}.bias"])
    mean = (h @ sd["action_net.weight"].T + sd["action_net.bias"]).clamp(-1, 1)
    torch.testing.assert_close(pol.predict(x, deterministic=True), mean)
    torch.testing.assert_close(pol.predict_values(x), (v @ sd["value_net.weight"].T + sd["value_net.bias"]).squeeze(-1))
    torch.testing.assert_close(pol.log_std.data, sd["log_std"])
    s = pol.predict(x, deterministic=False, generator=torch.Generator().manual_seed(2))
    assert (s - mean).abs().max() > 0.05 and s.abs().max() <= 1.0


# ------------------------------------------------------------------------------------------------------------- GPU
JIP = dict(enable_springs=True, task_env="JUMPING_IN_PLACE", observation_space_mode="ARS_BASIC")


@pytest.mark.gpu
def test_vecenv_protocol_and_terminal_observation():
    """numpy VecEnv protocol; infos[i]["terminal_observation"] is the observation the same env returns at `done` when it
    is NOT reset inside the step (twin env with auto_reset=False, same seed, same actions), and the obs row is the first
    observation of the next episode."""
    import quadruped_springs_b200 as qs
    n = 512
    venv = BatchedVecEnv(num_envs=n, seed=5, **JIP)
    twin = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=5, auto_reset=False, **JIP)
    obs = venv.reset()
    tobs = twin.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (n, 27) and obs.dtype == np.float32
    np.testing.assert_array_equal(obs, tobs.cpu().numpy())
    assert venv.get_attr("task_env") == ["JUMPING_IN_PLACE"] * n and venv.env_is_wrapped(object) == [False] * n
    assert venv.env_method("are_springs_enabled") == [True] * n
    with pytest.raises(ValueError):      # one handle = all envs: a call on a strict subset cannot be honoured
        venv.env_method("are_springs_enabled", indices=[0, 1])
    rng = np.random.default_rng(0)
    alive = np.ones(n, bool)       # envs still in their first episode: the twin is comparable
    seen = 0
    for t in range(120):
        a = rng.uniform(-1, 1, size=(n, 6)).astype(np.float32)
        obs, rew, done, infos = venv.step(a)
        tobs, trew, tdone, _ = twin.step(torch.as_tensor(a, device="cuda"))
        tobs, trew, tdone = tobs.cpu().numpy(), trew.cpu().numpy(), tdone.cpu().numpy()
        assert rew.dtype == np.float32 and done.dtype == bool and len(infos) == n
        np.testing.assert_array_equal(done[alive], tdone[alive])
        np.testing.assert_array_equal(rew[alive], trew[alive])
        for i in np.flatnonzero(done & alive):
            np.testing.assert_array_equal(infos[i]["terminal_observation"], tobs[i])
            assert infos[i]["TimeLimit.truncated"] is False
            assert np.abs(obs[i] - tobs[i]).max() > 1e-3          # the row already belongs to the next episode
            assert abs(obs[i][25] - 0.3) < 0.1                     # ... a settled robot (height sensor)
            seen += 1
        for i in np.flatnonzero(~done):
            assert "terminal_observation" not in infos[i]
        still = alive & ~done
        np.testing.assert_array_equal(obs[still], tobs[still])
        alive = still
    assert seen > 100


@pytest.mark.gpu
def test_policy_in_the_loop_with_vecnormalize_on_device():
    """BASELINE config 5 shape at test size: BACKFLIP, SB3-shaped MlpPolicy + VecNormalize evaluated on the device each
    step; nothing leaves the GPU, statistics converge to the observation moments."""
    n = 4096
    venv = BatchedVecEnv(num_envs=n, seed=1, enable_springs=True, task_env="BACKFLIP", observation_space_mode="ARS_BACKFLIP")
    vn = VecNormalizeTorch(venv)
    torch.manual_seed(0)
    pol = MlpPolicyTorch(venv.env.obs_dim, venv.env.action_dim).cuda()
    obs = vn.reset()
    raw = [vn.get_original_obs().clone()]
    ep_done = 0
    for t in range(60):
        obs, rew, done, infos = vn.step(pol.predict(obs))
        assert obs.is_cuda and obs.abs().max() <= 10.0 and torch.isfinite(rew).all()
        raw.append(vn.get_original_obs().clone())
        ep_done += int(done.sum())
    allobs = torch.cat(raw).double()
    torch.testing.assert_close(vn.obs_rms.mean, allobs.mean(0), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(vn.obs_rms.var, allobs.var(0, unbiased=False), rtol=1e-3, atol=1e-4)
    assert ep_done > 0 and vn.ret_rms.count > n


@pytest.mark.gpu
def test_evaluation_wrapper_infos():
    """env/wrappers/evaluation_wrapper.py:43-53: feet_forces = sum(F_n) / 4, running max height (restarts at reset) and
    running max jumping distance"""
    import quadruped_springs_b200 as qs
    from quadruped_springs_b200.vec_env import EvaluationWrapper
    n = 256
    env = EvaluationWrapper(qs.BatchedQuadrupedGymEnv(num_envs=n, seed=3, auto_reset=False, enable_noise=False,
                                                      task_env="JUMPING_FORWARD", enable_springs=True,
                                                      observation_space_mode="ARS_BASIC"))
    env.reset()
    a = env.get_last_action().clone()            # the settling action: the robot keeps standing
    obs, r, d, info = env.step(a)
    np.testing.assert_allclose(info["feet_forces"].cpu().numpy() * 4, 12.01301 * 9.8, rtol=2e-2)   # standing
    a = torch.zeros(n, 6, device="cuda")
    hs = []
    for t in range(40):
        a[:, [1, 4]] = 0.9 if t < 12 else -0.7
        a[:, [2, 5]] = -0.9 if t < 12 else 1.0
        obs, r, d, info = env.step(a)
        hs.append(env.robot.GetBasePosition()[:, 2].clone())
    np.testing.assert_allclose(info["max_height"].cpu().numpy(), torch.stack(hs).max(0).values.cpu().numpy(), rtol=1e-6)
    assert (info["max_height"] > 0.4).float().mean() > 0.9 and (info["max_fwd"] >= 0).all()
    flying = env.robot._is_flying()
    assert (info["feet_forces"][flying] == 0).all()
    env.reset()
    assert (env.max_h == 0).all()


def test_policy_loader_orders_layers_numerically(tmp_path):
    """six or more hidden layers put Linear modules at indices 0, 2, ..., 10: '10' must sort after '2'"""
    arch = (24, 20, 16, 12, 10, 8)
    ref = MlpPolicyTorch(9, 3, arch)
    path = tmp_path / "deep.zip"
    buf = io.BytesIO()
    torch.save(ref.state_dict(), buf)
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("policy.pth", buf.getvalue())        # no "data": the architecture is inferred from the weights
    pol = MlpPolicyTorch.from_sb3_zip(str(path), device="cpu")
    obs = torch.randn(5, 9)
    assert torch.equal(pol.predict(obs), ref.predict(obs))


def test_vec_env_rejects_index_subsets():
    class _E:
        num_envs, obs_dim, action_dim, device = 4, 3, 2, torch.device("cpu")
        observation_space = action_space = None
        _auto_reset = True
        foo = 1

        def set_terminal_obs_buffer(self, b):
            pass

        def bar(self):
            return 7

    v = BatchedVecEnv(env=_E())
    assert v.get_attr("foo", [1, 2]) == [1, 1]
    assert v.env_method("bar") == [7] * 4
    v.set_attr("foo", 5)
    with pytest.raises(ValueError):
        v.set_attr("foo", 6, indices=[0])
    with pytest.raises(ValueError):
        v.env_method("bar", indices=[1, 2])


def test_episode_wrappers_refuse_auto_reset_envs():
    from quadruped_springs_b200.demo import DemonstrationRecorder
    from quadruped_springs_b200.vec_env import EvaluationWrapper
    fake = type("E", (), {"_auto_reset": True, "num_envs": 2, "device": torch.device("cpu")})()
    with pytest.raises(ValueError):
        DemonstrationRecorder(fake)
    with pytest.raises(ValueError):
        EvaluationWrapper(fake)
