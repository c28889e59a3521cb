"""The reference's load_model.py (:66-134) without stable-baselines3 / sb3-contrib / gym: args.yml, vecnormalize.pkl and
ARS / PPO policy archives are read directly (quadruped_springs_b200/load_model.py).  The files are fabricated here in the
formats those libraries write (stand-in classes under their module names for the pickle)."""
import io
import json
import os
import pickle
import sys
import types
import zipfile
from collections import OrderedDict

import numpy as np
import pytest
import torch
import yaml

from quadruped_springs_b200 import load_model as lm


def _write_zoo_dir(tmp_path, obs_dim=27):
    src = tmp_path / "logs" / "ppo" / "QuadrupedSpring-v0_3"
    (src / lm.ENV_NAME).mkdir(parents=True)
    args = OrderedDict([("algo", "ppo"), ("env", lm.ENV_NAME), ("seed", 24),
                        ("env_kwargs", {"motor_control_mode": "CARTESIAN_PD", "task_env": "JUMPING_IN_PLACE", "enable_springs": True,
                                        "observation_space_mode": "ARS_BASIC", "action_space_mode": "SYMMETRIC",
                                        "add_noise": True, "enable_env_randomization": False, "aux_seed": 7})])
    with open(src / lm.ENV_NAME / "args.yml", "w") as f:
        yaml.dump(args, f)                       # rl-zoo: `yaml.dump(ordered_args, f)` -> python/object/apply:collections.OrderedDict
    # a pickle that names stable-baselines3 / gym classes, as VecNormalize.save() writes it
    mods = {}
    for name in ("stable_baselines3", "stable_baselines3.common", "stable_baselines3.common.vec_env",
                 "stable_baselines3.common.vec_env.vec_normalize", "stable_baselines3.common.running_mean_std", "gym",
                 "gym.spaces", "gym.spaces.box"):
        mods[name] = types.ModuleType(name)
    ns = {}
    exec("class RunningMeanStd:\n    pass\n", ns)
    ns["RunningMeanStd"].__module__ = "stable_baselines3.common.running_mean_std"
    mods["stable_baselines3.common.running_mean_std"].RunningMeanStd = ns["RunningMeanStd"]
    exec("class VecNormalize:\n    pass\n", ns)
    ns["VecNormalize"].__module__ = "stable_baselines3.common.vec_env.vec_normalize"
    mods["stable_baselines3.common.vec_env.vec_normalize"].VecNormalize = ns["VecNormalize"]
    exec("class Box:\n    pass\n", ns)
    ns["Box"].__module__ = "gym.spaces.box"
    mods["gym.spaces.box"].Box = ns["Box"]
    sys.modules.update(mods)
    try:
        rng = np.random.default_rng(0)
        obs_rms, ret_rms = ns["RunningMeanStd"](), ns["RunningMeanStd"]()
        obs_rms.mean, obs_rms.var, obs_rms.count = rng.normal(size=obs_dim), rng.uniform(0.5, 2, obs_dim), 12345.0
        ret_rms.mean, ret_rms.var, ret_rms.count = np.float64(0.3), np.float64(1.7), 999.0
        box = ns["Box"]()
        box.low, box.high, box.shape, box.dtype = -np.ones(obs_dim, np.float32), np.ones(obs_dim, np.float32), (obs_dim,), np.dtype("float32")
        box._np_random = np.random.RandomState(3)          # gym spaces carry a generator: must load as an inert bag too
        vn = ns["VecNormalize"]()
        vn.obs_rms, vn.ret_rms, vn.clip_obs, vn.clip_reward, vn.gamma, vn.epsilon = obs_rms, ret_rms, 10.0, 10.0, 0.99, 1e-8
        vn.training, vn.norm_obs, vn.norm_reward, vn.observation_space, vn.num_envs = True, True, True, box, 8
        vn.returns = np.zeros(8)
        with open(src / lm.ENV_NAME / "vecnormalize.pkl", "wb") as f:
            pickle.dump(vn, f)
    finally:
        for name in mods:
            sys.modules.pop(name, None)
    return str(src), obs_rms, ret_rms


def test_args_yml_and_env_kwargs(tmp_path):
    src, _, _ = _write_zoo_dir(tmp_path)
    raw = lm.load_env_kwargs(src)
    assert raw["algo"] == "ppo" and raw["env_kwargs"]["motor_control_mode"] == "CARTESIAN_PD"
    kw = lm.get_env_kwargs(src, task="JUMPING_FORWARD")
    # load_model.py:77-85,102-106: task and randomizer overridden, training-only keys removed
    assert kw["task_env"] == "JUMPING_FORWARD" and kw["env_randomizer_mode"] == "GROUND_RANDOMIZER" and kw["render"] is False
    assert not {"add_noise", "enable_env_randomization", "aux_seed"} & set(kw)
    with pytest.raises(RuntimeError):
        lm.load_env_kwargs(str(tmp_path / "nowhere"))
    # nothing but the OrderedDict tag is honoured (the reference's UnsafeLoader would build any object)
    bad = tmp_path / "bad" / lm.ENV_NAME
    bad.mkdir(parents=True)
    (bad / "args.yml").write_text("x: !!python/object/apply:os.system ['echo pwned']\n")
    with pytest.raises(yaml.YAMLError):
        lm.load_env_kwargs(str(tmp_path / "bad"))


def test_vecnormalize_pickle_is_read_without_sb3(tmp_path):
    src, obs_rms, ret_rms = _write_zoo_dir(tmp_path)
    assert "stable_baselines3" not in sys.modules and "gym" not in sys.modules
    st = lm.read_vecnormalize_pkl(os.path.join(src, lm.ENV_NAME, "vecnormalize.pkl"))
    np.testing.assert_array_equal(st["obs_mean"], obs_rms.mean)
    np.testing.assert_array_equal(st["obs_var"], obs_rms.var)
    assert st["obs_count"] == 12345.0 and st["ret_var"] == 1.7 and st["clip_obs"] == 10.0 and st["gamma"] == 0.99
    assert "stable_baselines3" not in sys.modules      # nothing was imported on the way

    class _V:                       # the part of BatchedVecEnv that VecNormalizeTorch touches
        num_envs, observation_space, action_space, device = 4, None, None, torch.device("cpu")
        env = type("E", (), {"obs_dim": 27})()
    vn = lm.load_vecnormalize(os.path.join(src, lm.ENV_NAME, "vecnormalize.pkl"), _V())
    assert vn.training is False and vn.norm_reward is False          # load_model.py:115-116
    obs = torch.randn(4, 27)
    ref = np.clip((obs.numpy().astype(np.float64) - obs_rms.mean) / np.sqrt(obs_rms.var + 1e-8), -10, 10)
    np.testing.assert_allclose(vn.normalize_obs(obs).numpy(), ref, rtol=1e-6, atol=1e-6)
    _V.env = type("E", (), {"obs_dim": 28})()
    with pytest.raises(ValueError):
        lm.load_vecnormalize(os.path.join(src, lm.ENV_NAME, "vecnormalize.pkl"), _V())


def test_a_malicious_pickle_runs_nothing(tmp_path):
    class Evil:
        def __reduce__(self):
            return (os.system, ("touch " + str(tmp_path / "pwned"),))
    p = tmp_path / "evil.pkl"
    p.write_bytes(pickle.dumps({"obs_rms": Evil()}))
    with pytest.raises(Exception):
        lm.read_vecnormalize_pkl(str(p))
    assert not (tmp_path / "pwned").exists()


def _zip_policy(path, sd, data):
    buf = io.BytesIO()
    torch.save(sd, buf)
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("policy.pth", buf.getvalue())
        z.writestr("data", json.dumps(data))


def test_ars_policies_load_and_predict(tmp_path):
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(16, 27, generator=g)
    # ARSLinearPolicy: one bias-free Linear, no squashing; predict clips to the action box
    W = torch.randn(6, 27, generator=g)
    _zip_policy(tmp_path / "lin.zip", {"action_net.0.weight": W},
                {"policy_class": {":type:": "<class 'abc.ABCMeta'>", ":serialized:": "x", "__module__": "sb3_contrib.ars.policies"},
                 "policy_kwargs": {}})
    pol = lm.load_policy(str(tmp_path / "lin.zip"), "ars", device="cpu")
    torch.testing.assert_close(pol.predict(obs), (obs @ W.t()).clamp(-1, 1))
    # ARSPolicy (MlpPolicy): [64, 64] ReLU with bias, Tanh on the output (squash_output defaults to True there)
    sd = {"action_net.0.weight": torch.randn(64, 27, generator=g) * 0.2, "action_net.0.bias": torch.randn(64, generator=g) * 0.1,
          "action_net.2.weight": torch.randn(64, 64, generator=g) * 0.2, "action_net.2.bias": torch.randn(64, generator=g) * 0.1,
          "action_net.4.weight": torch.randn(6, 64, generator=g) * 0.2, "action_net.4.bias": torch.randn(6, generator=g) * 0.1}
    _zip_policy(tmp_path / "mlp.zip", sd, {"policy_class": "ARSPolicy", "policy_kwargs": {}})
    pol = lm.load_policy(str(tmp_path / "mlp.zip"), "ars", device="cpu")
    h = torch.relu(obs @ sd["action_net.0.weight"].t() + sd["action_net.0.bias"])
    h = torch.relu(h @ sd["action_net.2.weight"].t() + sd["action_net.2.bias"])
    torch.testing.assert_close(pol.predict(obs), torch.tanh(h @ sd["action_net.4.weight"].t() + sd["action_net.4.bias"]))
    with pytest.raises(ValueError):
        _zip_policy(tmp_path / "ppo.zip", {"mlp_extractor.policy_net.0.weight": W}, {})
        lm.load_policy(str(tmp_path / "ppo.zip"), "ars", device="cpu")


@pytest.mark.gpu
def test_replay_runs_a_zoo_directory_end_to_end(tmp_path):
    """load_model.py:109-138 on the batched env: args.yml -> env, vecnormalize.pkl -> statistics, best_model.zip -> PPO
    policy, one deterministic episode per env with the go-to-rest wrapper"""
    from quadruped_springs_b200.vec_env import MlpPolicyTorch
    src, _, _ = _write_zoo_dir(tmp_path)
    torch.manual_seed(0)
    ref = MlpPolicyTorch(27, 6)
    buf = io.BytesIO()
    torch.save(ref.state_dict(), buf)
    with zipfile.ZipFile(os.path.join(src, "best_model.zip"), "w") as z:
        z.writestr("policy.pth", buf.getvalue())
        z.writestr("data", json.dumps({"policy_kwargs": {}}))
    ret = lm.replay(src, task="JUMPING_IN_PLACE", algo="ppo", n_envs=64, max_steps=1100)
    assert ret.shape == (64,) and np.isfinite(ret).all()
