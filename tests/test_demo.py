"""Demonstration recording / reference-state initialisation: row layout of
get_demonstration_wrapper.py:35-70 and the reset-from-a-desired-state path of quadruped_gym_env.py:288-289."""
import numpy as np
import pytest
import torch

from quadruped_springs_b200 import demo as D


def _reference_read_demo(demo, action_dim=6, num_joints=12):
    """get_demonstration_wrapper.py:59-70, restated"""
    ret, first = [], 0
    for last in np.cumsum(np.array([action_dim, num_joints, num_joints, 3, 4, 3, 3, 1])):
        ret.append(demo[first:last])
        first = last
    return ret


@pytest.mark.parametrize("A", [4, 6, 12])
def test_read_demo_layout_matches_reference(A):
    row = np.arange(A + 38, dtype=np.float64)
    got, ref = D.read_demo(row, A), _reference_read_demo(row, A)
    assert len(got) == len(ref) == len(D.DEMO_FIELDS) == 8
    for g, r in zip(got, ref):
        np.testing.assert_array_equal(g, r)
    s = D.demo_rows_to_states(row[None], A)[0]
    _, q, qd, pos, quat, lin, ang, _ = ref
    np.testing.assert_array_equal(s, np.concatenate([pos, quat, lin, ang, q, qd]))
    st = D.demo_rows_to_states(torch.as_tensor(row)[None], A)[0]
    np.testing.assert_array_equal(st.numpy(), s)


JIP = dict(enable_springs=True, task_env="JUMPING_IN_PLACE", observation_space_mode="ARS_BASIC")


def _jump_actions(T, n, device):
    a = torch.zeros(T, n, 6, device=device)
    a[:12, :, [1, 4]] = 0.9; a[:12, :, [2, 5]] = -0.9          # crouch
    a[12:20, :, [1, 4]] = -0.7; a[12:20, :, [2, 5]] = 1.0     # push
    return a


@pytest.mark.gpu
def test_recorder_rows_and_reference_state_initialisation(tmp_path):
    import quadruped_springs_b200 as qs
    n, T = 64, 60
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=2, enable_noise=False, auto_reset=False, enable_action_filter=True, **JIP)
    rec = D.DemonstrationRecorder(env, path=str(tmp_path))
    rec.reset()
    acts = _jump_actions(T, n, env.device)
    states = []
    for t in range(T):
        obs, r, d, info = rec.step(acts[t])
        states.append(env.get_state().clone())
        if t == 5:   # a row is exactly what the accessors return, in the reference's order
            row = rec.rows[-1][3].cpu().numpy()
            a, q, qd, pos, quat, lin, ang, flag = D.read_demo(row, 6)
            np.testing.assert_array_equal(a, env.get_last_filtered_action()[3].cpu().numpy())
            np.testing.assert_array_equal(q, env.robot.GetMotorAngles()[3].cpu().numpy())
            np.testing.assert_array_equal(qd, env.robot.GetMotorVelocities()[3].cpu().numpy())
            np.testing.assert_array_equal(pos, env.robot.GetBasePosition()[3].cpu().numpy())
            np.testing.assert_array_equal(quat, env.robot.GetBaseOrientation()[3].cpu().numpy())
            np.testing.assert_array_equal(lin, env.robot.GetBaseLinearVelocity()[3].cpu().numpy())
            np.testing.assert_array_equal(ang, env.robot.GetBaseAngularVelocity()[3].cpu().numpy())
            assert flag[0] == 0.0
            np.testing.assert_array_equal(D.demo_rows_to_states(row[None], 6)[0], states[-1][3].cpu().numpy())
    demo = rec.save_demo(0)
    assert demo.shape == (T - 1, 6 + 38)                          # save_demo drops the last row (:29-31)
    np.testing.assert_array_equal(np.load(tmp_path / "demo_list.npy"), demo)
    flag = demo[:, -1]
    assert flag[0] == 0 and flag[-1] == 1 and (np.diff(flag) >= 0).all()     # latches once, after the apex
    k = int(np.argmax(flag))
    assert demo[k, 6 + 24 + 2] > 0.33 and demo[k, 6 + 24 + 7 + 2] <= 0.0      # airborne, moving down
    # the filter's output, not the raw action, is what is recorded (:37)
    assert np.abs(demo[:12, 1] - 0.9).max() > 0.05

    # ---- reference-state initialisation: episodes start on demonstration rows, unsettled
    env2 = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=2, enable_noise=False, auto_reset=False, **JIP)
    env2.reset()
    rsi = D.ReferenceStateInitialization(env2, demo, seed=0)
    obs = rsi.reset()
    els = rsi.random_el.cpu().numpy()
    assert els.min() >= 0 and els.max() < len(demo) - 5 and len(np.unique(els)) > 10
    S = env2.get_state().cpu().numpy()
    np.testing.assert_array_equal(S, D.demo_rows_to_states(demo[els], 6))
    assert (env2._views["sim_steps"] == 0).all() and (env2._views["env_steps"] == 0).all()
    assert (env2.get_last_action() == 0).all()                    # no settle, no settling action (:284,288-289)
    assert ((env2._views["contact"] & 15) == 0).all()             # no contact points before the first stepSimulation
    np.testing.assert_array_equal(env2._views["task"][6].cpu().numpy(), S[:, 2])   # task._reset took ITS height
    np.testing.assert_allclose(obs.cpu().numpy(), env2.get_observation(with_noise=False).cpu().numpy(), atol=0)
    np.testing.assert_allclose(obs[:, 25].cpu().numpy(), S[:, 2], atol=1e-6)       # the height sensor reads the demo state
    # stepping from there is the same as stepping the recorded run from that state (same kernels, same inputs)
    env3 = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=2, enable_noise=False, auto_reset=False, **JIP)
    env3.reset()
    env3.set_state(env2.get_state())
    env3._views["contact"][:] = 0
    env3._views["foot_force"][:] = 0
    env3._views["mu"][:] = env2._views["mu"]                      # env2 is in its second episode: another friction draw
    a = torch.zeros(n, 6, device=env.device)
    o2, r2, d2, _ = env2.step(a)
    o3, r3, d3, _ = env3.step(a)
    torch.testing.assert_close(env2.get_state(), env3.get_state(), rtol=0, atol=0)
    # a masked RSI reset leaves the other envs alone; every sixth draw comes from the first fifth of the demo (:34-43)
    before = env2.get_state().clone()
    mask = torch.zeros(n, dtype=torch.bool, device=env.device); mask[::2] = True
    rsi.reset(mask)
    after = env2.get_state()
    assert torch.equal(after[1::2], before[1::2]) and (env2._views["env_steps"][::2] == 0).all()
    assert (env2._views["env_steps"][1::2] == 1).all()
    r = D.ReferenceStateInitialization(env2, demo, seed=1)
    draws = [r.compute_random_el() for _ in range(60)]
    assert all(d < len(demo) // 5 for d in draws[5::6]) and max(draws) >= len(demo) // 5


@pytest.mark.gpu
def test_reset_to_state_with_auto_reset_keeps_the_conveyor_consistent():
    """reset_to_state skips the env's prefetched episode; later automatic resets still start settled episodes and the
    run stays identical to one whose envs were never touched (the other envs) -- the ring bookkeeping is intact."""
    import quadruped_springs_b200 as qs
    n = 1024
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=6, **JIP)
    twin = qs.BatchedQuadrupedGymEnv(num_envs=n, seed=6, **JIP)
    env.reset(); twin.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    mask = torch.zeros(n, dtype=torch.bool, device="cuda"); mask[:100] = True
    S = env.get_state().clone()
    S[:, 2] += 0.2                                             # dropped from 20 cm above the settled pose
    env.reset_to_state(S, mask)
    assert (env.get_state()[:100, 2] > 0.45).all()
    for t in range(150):
        a = torch.rand(n, 6, device="cuda", generator=g) * 2 - 1
        o, r, d, _ = env.step(a)
        ot, rt, dt, _ = twin.step(a)
        assert torch.isfinite(o).all()
        assert torch.equal(o[100:], ot[100:]) and torch.equal(d[100:], dt[100:])      # untouched envs: bit-identical
    # the touched envs went through automatic resets onto settled robots again
    assert (env._views["env_steps"][:100] < 150).any()
    z = env.get_state()[:100, 2]
    assert ((z > 0.05) & (z < 1.5)).all()
