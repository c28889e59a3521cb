"""GPU behaviour tests of the batched env at BASELINE sizes: gym contract,
auto-reset, time-limit truncation, determinism, shard invariance, sensor noise,
rollout statistics (pytest -m gpu)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

JIP = dict(enable_springs=True, task_env="JUMPING_IN_PLACE", observation_space_mode="ARS_BASIC")


@pytest.fixture(scope="module")
def qs():
    import quadruped_springs_b200 as m
    return m


def test_gym_contract_shapes_and_dtypes(qs):
    env = qs.BatchedQuadrupedGymEnv(num_envs=4096, **JIP)
    obs = env.reset()
    assert obs.shape == (4096, 27) and obs.dtype == torch.float32 and obs.is_cuda
    assert env.action_space.shape == (6,) and env.observation_space.shape == (27,)
    a = torch.rand(4096, 6, device="cuda") * 2 - 1
    obs, r, d, info = env.step(a)
    assert r.shape == (4096,) and d.dtype == torch.bool and info["TimeLimit.truncated"].dtype == torch.bool
    assert torch.isfinite(obs).all() and torch.isfinite(r).all()
    assert (env.get_sim_time() == 0.01).all() or d.any()
    with pytest.raises(ValueError):
        env.step(torch.zeros(4096, 5, device="cuda"))
    d_obs = env.get_observation_dict()
    assert list(d_obs) == ["Encoder", "JointVelocity", "Pitch", "Height", "Base Linear Velocity z direction"]
    with pytest.raises(ValueError):
        qs.BatchedQuadrupedGymEnv(num_envs=2, motor_control_mode="TORQUE")          # quadruped_gym_env.py:167-168
    with pytest.raises(ValueError):
        qs.BatchedQuadrupedGymEnv(num_envs=2, observation_space_mode="ARS_HEIGHT")  # __init__.py:9 (undefined upstream)


def test_full_size_random_rollout_stays_finite_and_resets(qs):
    n = 65536
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=1, **JIP)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    n_done = 0
    for t in range(60):
        a = torch.rand(n, 6, device="cuda", generator=g) * 2 - 1
        obs, r, d, info = env.step(a)
        n_done += int(d.sum())
        # auto-reset: a finished env starts its next episode inside the same call
        assert (env._views["sim_steps"][d] == 0).all() and (env._views["env_steps"][d] == 0).all()
        assert (env._views["env_steps"][~d] > 0).all()
    assert torch.isfinite(obs).all() and torch.isfinite(r).all()
    S = env.get_state()
    assert torch.isfinite(S).all()
    assert (S[:, 3:7].norm(dim=1) - 1).abs().max() < 1e-4
    assert (S[:, 25:37].abs() <= 30.1 + 1e-4).all()        # maxJointVelocity clamp (quadruped.py:678-683)
    assert n_done > 0                                       # random actions do crash some robots within 60 steps
    out = qs.stats.gather_rollout_stats(env.rollout_stats())
    assert out["num_envs"] == n and out["episodes"] == n_done
    assert 0 < out["mean_length"] <= 60 and out["terminated_fraction"] == 1.0


def test_time_limit_truncation_on_step_1001(qs):
    # holding the settling action keeps the robot standing: the only way out is the 10 s limit,
    # which the reference reaches on control step 1001 (`>` test, quadruped_gym_env.py:245)
    env = qs.BatchedQuadrupedGymEnv(num_envs=32, auto_reset=False, enable_noise=False, **JIP)
    env.reset()
    a = env.get_last_action().clone()
    for t in range(1, 1002):
        obs, r, d, info = env.step(a)
        if t <= 1000:
            assert not d.any(), t
    assert d.all() and info["TimeLimit.truncated"].all()
    # sparse task: reward only at the end (alive bonus, robot_tasks.py:50-52), zero before
    assert (r > 0).all() and (r < 0.2).all()


def test_determinism_and_shard_invariance(qs):
    n = 1024
    acts = [torch.rand(n, 6, device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)) * 2 - 1 for i in range(15)]

    def run(offset, count, sl):
        env = qs.BatchedQuadrupedGymEnv(num_envs=count, seed=7, env_id_offset=offset, **JIP)
        out = [env.reset().clone()]
        for a in acts:
            obs, r, d, _ = env.step(a[sl].contiguous())
            out.append(torch.cat([obs, r[:, None], d[:, None].float()], dim=1).clone())
        return out

    full = run(0, n, slice(0, n))
    again = run(0, n, slice(0, n))
    for x, y in zip(full, again):
        assert torch.equal(x, y)                       # bit-identical replays
    half = run(512, 512, slice(512, n))                # the second shard of a 2-GPU run
    for x, y in zip(full, half):
        assert torch.equal(x[512:], y)                 # RNG streams and results depend on the GLOBAL env id only


def test_sensor_noise_statistics(qs):
    n = 65536
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, auto_reset=False, seed=3, **JIP)
    noisy = env.reset().clone()
    clean = env.get_observation(with_noise=False)
    diff = (noisy - clean).cpu().numpy()
    std = qs.ops.obs_noise_std(enable_springs=True, observation_space_mode="ARS_BASIC")
    np.testing.assert_allclose(diff.std(axis=0), std, rtol=0.03)          # sensor.py:25-32 N(0, sigma) per element
    assert np.abs(diff.mean(axis=0) / std).max() < 0.03
    c = np.corrcoef(diff[:, :6].T)
    assert np.abs(c - np.eye(6)).max() < 0.03                             # independent across elements
    a = env.get_last_action().clone()
    o1 = env.step(a)[0].clone()
    d1 = (o1 - env.get_observation(with_noise=False)).cpu().numpy()
    assert np.abs(np.corrcoef(diff[:, 0], d1[:, 0])[0, 1]) < 0.03         # resampled every control step (:58-60)
    quiet = qs.BatchedQuadrupedGymEnv(num_envs=64, enable_noise=False, auto_reset=False, **JIP)
    o = quiet.reset()
    assert torch.equal(o, quiet.get_observation(with_noise=False))


@pytest.mark.parametrize("late_cap", [None, 2])
def test_step_host_matches_device_step(qs, monkeypatch, late_cap):
    """qs_step_host sends the bulk of the results to the host as soon as k_step_contact is done and the rows of the envs
    the general solver finishes afterwards in a compact side buffer (scattered on the host); with a side buffer of two
    rows the overflow path (everything copied again) runs instead.  Either way: the device-buffer step, bit for bit."""
    if late_cap is not None:
        monkeypatch.setenv("QS_LATE_CAP", str(late_cap))
    n = 2048
    e1 = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=5, **JIP)
    e2 = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=5, **JIP)
    e1.reset(); e2.reset()
    rng = np.random.default_rng(0)
    late = 0
    for _ in range(60):
        a = rng.uniform(-1, 1, size=(n, 6)).astype(np.float32)
        o, r, d, t = e1.step_host(a)
        od, rd, dd, info = e2.step(torch.from_numpy(a).cuda())
        assert np.array_equal(o, od.cpu().numpy()) and np.array_equal(r, rd.cpu().numpy())
        assert np.array_equal(d.astype(bool), dd.cpu().numpy())
        assert np.array_equal(t.astype(bool), info["TimeLimit.truncated"].cpu().numpy())
        late += int(d.sum())
    assert late > 50    # crashes went through the general solver, i.e. through the side buffer


def test_reset_host_matches_device_reset(qs):
    e1 = qs.BatchedQuadrupedGymEnv(num_envs=128, seed=9, auto_reset=False, **JIP)
    e2 = qs.BatchedQuadrupedGymEnv(num_envs=128, seed=9, auto_reset=False, **JIP)
    o1 = e1.reset_host()
    o2 = e2.reset()
    assert np.array_equal(o1, o2.cpu().numpy())


def test_partial_reset_mask(qs):
    env = qs.BatchedQuadrupedGymEnv(num_envs=256, auto_reset=False, enable_noise=False, **JIP)
    env.reset()
    a = torch.rand(256, 6, device="cuda") * 2 - 1
    for _ in range(3):
        env.step(a)
    before = env.get_state().clone()
    mask = torch.zeros(256, dtype=torch.bool, device="cuda")
    mask[::3] = True
    env.reset(mask)
    after = env.get_state()
    assert torch.equal(after[~mask], before[~mask])
    assert (env._views["env_steps"][mask] == 0).all() and (env._views["env_steps"][~mask] == 3).all()
    assert (after[mask, 2] - 0.328).abs().max() < 5e-3       # settled standing height with springs


def test_per_env_gain_override(qs):
    # landing wrappers swap PD gains at run time (landing_wrapper.py:21-33): gains are per-env tensors
    env = qs.BatchedQuadrupedGymEnv(num_envs=8, auto_reset=False, enable_noise=False,
                                    env_randomizer_mode="NO_RANDOMIZER", **JIP)
    env.reset()
    env.robot.set_motor_gains(60.0, 1.5, env_ids=torch.tensor([1, 3], device="cuda"))
    a = torch.full((8, 6), 0.5, device="cuda")
    env.step(a)
    tau = env.robot.GetMotorTorques()
    assert torch.equal(tau[0], tau[2]) and not torch.equal(tau[0], tau[1]) and torch.equal(tau[1], tau[3])


def test_config2_no_springs_4096(qs):
    """BASELINE config 2: PEA off, jumping in place, 4096 envs, fixed-seed random actions"""
    env = qs.BatchedQuadrupedGymEnv(num_envs=4096, enable_springs=False, task_env="JUMPING_IN_PLACE",
                                    observation_space_mode="ARS_BASIC", seed=0)
    obs = env.reset()
    assert (env.robot.GetBasePosition()[:, 2] - 0.309).abs().max() < 5e-3     # settled height without springs
    assert (env.robot.GetSpringTorques() == 0).all()
    g = torch.Generator(device="cuda").manual_seed(0)
    ret = torch.zeros(4096, device="cuda")
    for _ in range(150):
        obs, r, d, info = env.step(torch.rand(4096, 6, device="cuda", generator=g) * 2 - 1)
        ret += r
    assert torch.isfinite(obs).all() and torch.isfinite(ret).all()
    out = qs.stats.gather_rollout_stats(env.rollout_stats())
    assert out["episodes"] > 100 and 0.2 < out["mean_max_height"] < 1.5


def test_config4_cpg_torque_mode_matches_oracle_and_trots(qs):
    """BASELINE config 4: HopfNetwork + Cartesian impedance in TORQUE mode, action_repeat = 1
    (hopf_network.py:176-289): first against the oracle tick by tick, then 4096 robots trotting."""
    from oracle import oracle as O
    n = 4
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, isRLGymInterface=False, action_repeat=1, motor_control_mode="TORQUE",
                                    enable_springs=True, auto_reset=False, enable_noise=False,
                                    env_randomizer_mode="NO_RANDOMIZER", solver=dict(mu_ground=0.8))
    env.reset()
    assert env.action_dim == 12
    ref = O.Env(enable_springs=True, motor_control_mode="TORQUE", isRLGymInterface=False, action_repeat=1)
    ref.reset(mu=0.8)
    np.testing.assert_allclose(env.get_state()[0].cpu().numpy(), ref.world.get_state(), atol=1e-3)  # settle_robot_by_pd
    cpg = qs.HopfNetwork(num_envs=n, gait="TROT", omega_swing=16 * np.pi, omega_stance=4 * np.pi, seed=1)
    X = cpg.X[0].cpu().numpy().copy()
    PHI = cpg.PHI
    for t in range(150):
        xs, zs, tau = cpg.update(env.robot.GetMotorAngles(), env.robot.GetMotorVelocities())
        s = ref.world.get_state()
        X, rx, rz = O.cpg_step(X, PHI, 2, 16 * np.pi, 4 * np.pi, 1, 0.001, 0.04, 0.25, 0.05, 0.01)
        rtau = O.cpg_torque(rx, rz, s[13:25], s[25:37], 0.0838, [150, 70, 70], [2, 0.5, 0.5], 2500.0, 40.0)
        env.step(tau)
        ref.step(rtau)
        if t < 40:   # the impedance gains (2500 N/m) amplify fp32 differences quickly; hold the first 40 ms tightly
            np.testing.assert_allclose(env.get_state()[0].cpu().numpy()[:25], ref.world.get_state()[:25], atol=2e-3)
    assert abs(float(env.get_state()[0, 2]) - ref.world.get_state()[2]) < 0.02
    # at size: robots keep walking forward and stay up
    sys_path = __import__("sys").path
    sys_path.insert(0, __import__("os").path.join(__import__("conftest").ROOT, "examples"))
    import cpg_trot
    env2, _ = cpg_trot.main(4096, 1500)
    pos = env2.robot.GetBasePosition()
    assert torch.isfinite(pos).all()
    assert float(pos[:, 2].mean()) > 0.2 and float((pos[:, 2] > 0.15).float().mean()) > 0.95
    assert float(pos[:, 0].mean()) > 0.05     # moved forward in 1.5 s of trotting


def test_config5_policy_in_the_loop(qs):
    sys_path = __import__("sys").path
    sys_path.insert(0, __import__("os").path.join(__import__("conftest").ROOT, "examples"))
    import policy_rollout
    policy_rollout.main(2048, 40)


@pytest.mark.parametrize("mode,variants", [("GROUND_RANDOMIZER", ((100, 100), (2500, 2500), (7, 13))),
                                           ("TEST_RANDOMIZER", ((37, 37),))])   # masses + springs ride the conveyor too
def test_settle_conveyor_is_invisible(qs, monkeypatch, mode, variants):
    """Episodes are settled ahead of time in slices of ticks (csrc/qs_step_kernels.cuh, settle
    conveyor).  However the 2500 ticks are cut -- tiny slices, whole settles, or so slowly that
    envs run out of settled slots and take the urgent path -- every env sees bit-identical
    episodes: the start state depends on (seed, global env id, episode number) only."""
    import ctypes as C
    from quadruped_springs_b200 import _lib
    n, steps = 1024, 520
    cfg = dict(enable_springs=True, task_env="JUMPING_FORWARD", motor_control_mode="CARTESIAN_PD",
               observation_space_mode="ARS_BASIC", env_randomizer_mode=mode)   # random actions end an episode every ~85 steps
    g = torch.Generator(device="cuda").manual_seed(11)
    acts = [torch.rand(n, 6, device="cuda", generator=g) * 2 - 1 for _ in range(steps)]

    def run(slice_min, slice_max):
        if slice_min is None:
            monkeypatch.delenv("QS_SETTLE_SLICE_MIN", raising=False)
            monkeypatch.delenv("QS_SETTLE_SLICE_MAX", raising=False)
        else:
            monkeypatch.setenv("QS_SETTLE_SLICE_MIN", str(slice_min))
            monkeypatch.setenv("QS_SETTLE_SLICE_MAX", str(slice_max))
        env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=5, **cfg)
        out, cnt, urgent, ndone = [env.reset().clone()], (C.c_int32 * 4)(), 0, 0
        for a in acts:
            obs, r, d, _ = env.step(a)
            out.append(torch.cat([obs, r[:, None], d[:, None].float()], dim=1).clone())
            _lib.check(env._L.qs_debug_counters(env._h, cnt, None))
            urgent += cnt[1]
            ndone += int(d.sum())
        env.close()
        return out, urgent, ndone

    ref, urgent_ref, ndone = run(None, None)
    assert ndone > 5 * n and urgent_ref <= ndone // 100   # more episodes per env than ring slots, settled in time
    for lo, hi in variants:
        out, urgent, _ = run(lo, hi)
        for x, y in zip(ref, out):
            assert torch.equal(x, y)
    out, urgent, _ = run(1, 1)                         # 1 tick per step: the ring of 4 runs dry
    assert urgent > 0
    for x, y in zip(ref, out):
        assert torch.equal(x, y)


def test_step_host_fetches_urgent_first_observations(qs, monkeypatch):
    """qs_step_host copies the results to the host while the settle slice is still running; an episode that
    had to be settled on the spot writes its first observation after that copy, so the call repeats the
    observation copy.  Starve the conveyor (1 tick per step) and compare with the device-buffer step."""
    monkeypatch.setenv("QS_SETTLE_SLICE_MIN", "1")
    monkeypatch.setenv("QS_SETTLE_SLICE_MAX", "1")
    n, steps = 256, 460
    cfg = dict(enable_springs=True, task_env="JUMPING_FORWARD", motor_control_mode="CARTESIAN_PD",
               observation_space_mode="ARS_BASIC")
    a_dev = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=3, **cfg)
    a_host = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=3, **cfg)
    a_dev.reset()
    a_host.reset()
    g = torch.Generator(device="cuda").manual_seed(4)
    out = (np.empty((n, a_host.obs_dim), np.float32), np.empty(n, np.float32), np.empty(n, np.uint8), np.empty(n, np.uint8))
    import ctypes as C
    from quadruped_springs_b200 import _lib
    cnt, urgent = (C.c_int32 * 4)(), 0
    for t in range(steps):
        a = torch.rand(n, 6, device="cuda", generator=g) * 2 - 1
        obs, r, d, info = a_dev.step(a)
        a_host.step_host(a.cpu().numpy(), out)
        _lib.check(a_host._L.qs_debug_counters(a_host._h, cnt, None))
        urgent += cnt[1]
        assert np.array_equal(out[0], obs.cpu().numpy()), t
        assert np.array_equal(out[1], r.cpu().numpy()) and np.array_equal(out[2].astype(bool), d.cpu().numpy()), t
    assert urgent > 0


def test_everything_on_is_shard_invariant_and_consistent(qs):
    """all the widened pieces at once -- curriculum randomizers (ground, masses + payload, springs), LandingWrapper2,
    GoToRestWrapper, action filter, sensor noise, auto-reset with the settle conveyor -- for 400 steps: finite, legal
    controller transitions, controller gains where they belong, and the second shard of a two-GPU run (global env ids
    2048..4095) reproduces the second half of the full run bit for bit, resets included."""
    n, steps = 4096, 400
    cfg = dict(enable_springs=True, task_env="JUMPING_FORWARD", motor_control_mode="CARTESIAN_PD", observation_space_mode="ARS_BASIC",
               env_randomizer_mode="TEST_RANDOMIZER_CURRICULUM", curriculum_level=0.5, landing_wrapper="LandingWrapper2",
               go_to_rest_wrapper=True, enable_action_filter=True, seed=13)
    g = torch.Generator(device="cuda").manual_seed(3)
    # mostly hops (the scripted controllers need take-offs), some random steps
    acts = []
    for t in range(steps):
        ph = t % 70
        z, x = (1.0, 0.3) if ph < 25 else ((-1.0, -0.5) if ph < 37 else (-0.1, 0.0))
        a = torch.tensor([x, 0, z, x, 0, z], device="cuda").expand(n, -1) + 0.15 * torch.randn(n, 6, device="cuda", generator=g)
        acts.append(a.clamp(-1, 1).contiguous())

    def run(offset, count, sl, check):
        env = qs.BatchedQuadrupedGymEnv(num_envs=count, env_id_offset=offset, **cfg)
        out = [env.reset().clone()]
        prev_land = torch.zeros(count, dtype=torch.int32, device="cuda")
        seen_rest = seen_land = episodes = 0
        for a in acts:
            obs, r, d, info = env.step(a[sl].contiguous())
            out.append(torch.cat([obs, r[:, None], d[:, None].float()], dim=1).clone())
            if check:
                land, rest = info["landing_mode"].clone(), info["rest_active"].clone()
                assert torch.isfinite(obs).all() and torch.isfinite(r).all()
                ok = (land == prev_land) | ((prev_land == 0) & (land == 1)) | ((prev_land == 1) & (land == 2)) | \
                     ((prev_land == 2) & (land == 3)) | (d & (land == 0))
                assert ok.all()
                kp = env._views["kp"][0]
                assert (kp[rest != 0] == 60).all() and (kp[(rest == 0)] == 75).all()     # LandingWrapper2 keeps the default gains
                assert (rest[d] == 0).all()                                                # a new episode starts unscripted
                seen_rest += int((rest != 0).sum()); seen_land += int((land == 2).sum()); episodes += int(d.sum())
                prev_land = land
        if check:
            assert seen_rest > 0 and seen_land > 0 and episodes > n
            md = env._views["mass_draw"]
            assert float(md[4].max()) > 1.5 and float(md[4].max()) < 2.5                 # curriculum level 0.5: block up to 2.5 kg
        return out

    full = run(0, n, slice(0, n), True)
    half = run(n // 2, n // 2, slice(n // 2, n), False)
    for x, y in zip(full, half):
        assert torch.equal(x[n // 2:], y)


def test_nan_guard_cuts_the_env_off_and_spares_the_others(qs):
    """a state that stopped being finite ends the episode (done, not truncated, reward 0, finite
    observation); with auto_reset the env starts a fresh episode like after any other done.  Nothing leaks into the other
    envs, the rewards or the rollout statistics."""
    n = 256
    for auto in (False, True):
        env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=5, auto_reset=auto, **JIP)
        twin = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=5, auto_reset=auto, **JIP)
        env.reset(); twin.reset()
        a = torch.zeros(n, 6, device="cuda")
        env.step(a); twin.step(a)
        S = env.get_state()
        bad = [3, 77, 200]
        S[3, 2] = float("nan")            # base height
        S[77, 15] = float("inf")          # a joint angle (an Inf / NaN joint RATE is healed by the +-30.1 velocity clamp)
        S[200, 4] = float("nan")          # a quaternion component
        contact = env._views["contact"].clone(); ff = env._views["foot_force"].clone()
        env.set_state(S)                  # (set_state clears the contact history: put it back for the healthy envs)
        env._views["contact"].copy_(contact); env._views["foot_force"].copy_(ff)
        obs, r, d, info = env.step(a)
        o2, r2, d2, _ = twin.step(a)
        assert d[bad].all() and not info["TimeLimit.truncated"][bad].any()
        assert (r[bad] == 0).all()
        assert torch.isfinite(obs).all() and torch.isfinite(r).all() and torch.isfinite(env.get_state()).all()
        good = torch.ones(n, dtype=torch.bool, device="cuda"); good[bad] = False
        assert torch.equal(obs[good], o2[good]) and torch.equal(r[good], r2[good]) and torch.equal(d[good], d2[good])
        out = qs.stats.gather_rollout_stats(env.rollout_stats())
        assert out["nonfinite"] == 3 and np.isfinite(list(out.values())).all()
        if auto:
            assert (env._views["env_steps"][bad] == 0).all()       # a new episode has started
            for _ in range(3):
                obs, r, d, info = env.step(a)
            assert torch.isfinite(obs).all()


def test_partial_host_reset_keeps_the_other_rows(qs):
    """reset_host(mask) after step_host must return the last observation for the envs that are not reset"""
    n = 512
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=2, auto_reset=False, **JIP)
    env.reset_host()
    rng = np.random.default_rng(0)
    for _ in range(3):
        obs, r, d, t = env.step_host(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
    last = obs.copy()
    mask = np.zeros(n, np.uint8); mask[::7] = 1
    out = env.reset_host(mask)
    keep = mask == 0
    assert np.array_equal(out[keep], last[keep])
    assert not np.array_equal(out[~keep], last[~keep])
    assert (env._views["env_steps"].cpu().numpy()[~keep] == 0).all() and (env._views["env_steps"].cpu().numpy()[keep] == 3).all()
    # and the device path after a host step: the rows of the tensor it returns are the same
    obs, r, d, t = env.step_host(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
    last = obs.copy()
    dev = env.reset(mask=torch.as_tensor(mask, device="cuda"))
    assert np.array_equal(dev.cpu().numpy()[keep], last[keep])


def test_cpg_drive_equals_the_python_loop(qs):
    """HopfNetwork.drive (qs_cpg_steps: the reference's CPG loop, hopf_network.py:241-289, inside the library) gives bit
    for bit the states of the caller's own loop over cpg.update + env.step"""
    n = 512
    kw = dict(num_envs=n, isRLGymInterface=False, time_step=0.001, action_repeat=1, motor_control_mode="TORQUE",
              enable_springs=True, auto_reset=False, seed=4)
    a, b = qs.BatchedQuadrupedGymEnv(**kw), qs.BatchedQuadrupedGymEnv(**kw)
    a.reset(); b.reset()
    ca = qs.HopfNetwork(num_envs=n, gait="BOUND", omega_swing=16 * np.pi, omega_stance=4 * np.pi, time_step=0.001, seed=1)
    cb = qs.HopfNetwork(num_envs=n, gait="BOUND", omega_swing=16 * np.pi, omega_stance=4 * np.pi, time_step=0.001, seed=1)
    for _ in range(60):
        _, _, tau = ca.update(a.robot.GetMotorAngles(), a.robot.GetMotorVelocities())
        oa, ra, da, _ = a.step(tau)
    ob, rb, db, _ = cb.drive(b, 60)
    assert torch.equal(a.get_state(), b.get_state()) and torch.equal(ca.X, cb.X)
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    with pytest.raises(ValueError if False else Exception):
        cb.drive(qs.BatchedQuadrupedGymEnv(num_envs=4, **JIP), 1)      # not a TORQUE-mode env


def test_graph_replay_equals_direct_launches(qs, monkeypatch):
    """qs_step replays a CUDA graph of the step (side stream, programmatic dependent launch and all); with QS_GRAPH=0 the
    same kernels are launched one by one.  Same numbers, bit for bit, with auto-reset and sensor noise on -- also when the
    caller hands over a new action tensor every step and when the output buffers change (a second graph key)."""
    n = 2048
    kw = dict(num_envs=n, seed=11, auto_reset=True, enable_springs=True, task_env="JUMPING_FORWARD",
              motor_control_mode="CARTESIAN_PD", observation_space_mode="ARS_BASIC")
    a = qs.BatchedQuadrupedGymEnv(**kw)
    monkeypatch.setenv("QS_GRAPH", "0")
    b = qs.BatchedQuadrupedGymEnv(**kw)
    monkeypatch.delenv("QS_GRAPH")
    assert a._L.qs_timing_window(a._h) == 16 and b._L.qs_timing_window(b._h) == 512
    a.reset(); b.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(70):
        act = torch.rand(n, 6, device="cuda", generator=g) * 2 - 1          # a fresh tensor every step
        if t == 40:                                                          # new output buffers: a second graph key
            a._obs = torch.zeros_like(a._obs); a._reward = torch.zeros_like(a._reward)
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act.clone())
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), t
        assert torch.equal(ia["TimeLimit.truncated"], ib["TimeLimit.truncated"])
    assert torch.equal(a.get_state(), b.get_state())
    assert int(da.sum()) >= 0 and float(a.rollout_stats()[1]) > 0           # episodes did end and restart on the way
