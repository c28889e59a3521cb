"""The CPU oracle against the fixtures produced by the reference's own code
(oracle/gen_golden.py).  Runs without a GPU."""
import json

import numpy as np
import pytest

from conftest import ROLLOUTS, load_golden
from oracle import oracle as O

TOL = dict(rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("tag", ["s1", "s0"])
def test_pd_and_spring_torque(analytic, tag):
    g = analytic
    kp, kd, tm = g[f"{tag}_cfg_MOTOR_KP"], g[f"{tag}_cfg_MOTOR_KD"], g[f"{tag}_cfg_RL_TORQUE_LIMITS"]
    q, qd, cmd = g[f"{tag}_q"], g[f"{tag}_qd"], g[f"{tag}_cmd"]
    for i in range(len(q)):
        np.testing.assert_allclose(O.pd_torque(kp, kd, tm, cmd[i], q[i], qd[i]), g[f"{tag}_tau_pd"][i], **TOL)
        np.testing.assert_allclose(O.pd_torque(kp, kd, tm, g[f"{tag}_tcmd"][i], q[i], qd[i], torque_mode=True),
                                   g[f"{tag}_tau_torque"][i], **TOL)
        if tag == "s1":
            np.testing.assert_allclose(
                O.spring_torque(g["s1_cfg_SPRINGS_STIFFNESS"], g["s1_cfg_SPRINGS_DAMPING"],
                                g["s1_cfg_SPRINGS_REST_ANGLE"], q[i], qd[i]), g["s1_tau_spring"][i], **TOL)


def test_known_answers_survey_appendix_c(analytic):
    # PD (-7.9,-8,-8)x4 and the PEA vector at q = INIT+0.1, qd = 0.5
    g = analytic
    np.testing.assert_allclose(g["s1_tau_pd"][1], [-7.9, -8.0, -8.0] * 4, atol=1e-12)
    np.testing.assert_allclose(g["s1_tau_spring"][1],
                               [-0, -2.15, 5.85, -2.15, -2.15, 5.85, -0, -2.15, 5.85, -2.15, -2.15, 5.85], atol=1e-12)
    J, pos = O.fk_jacobian([0, np.pi / 4, -np.pi / 2], 0)
    np.testing.assert_allclose(pos, [0, -0.0847, -0.30122749], atol=1e-8)
    np.testing.assert_allclose(J, [[0, -0.30122749, -0.15061374], [0.30122749, 0, 0], [-0.0847, 0, -0.15061374]], atol=1e-8)


@pytest.mark.parametrize("tag", ["s1", "s0"])
def test_fk_jacobian_ik(analytic, tag):
    g = analytic
    q, qd = g[f"{tag}_q"], g[f"{tag}_qd"]
    for i in range(len(q)):
        for leg in range(4):
            J, pos = O.fk_jacobian(q[i, 3 * leg:3 * leg + 3], leg)
            np.testing.assert_allclose(J, g[f"{tag}_fk_J"][i, leg], **TOL)
            np.testing.assert_allclose(pos, g[f"{tag}_fk_pos"][i, leg], **TOL)
            np.testing.assert_allclose(J @ qd[i, 3 * leg:3 * leg + 3], g[f"{tag}_foot_vel"][i, leg], **TOL)
            np.testing.assert_allclose(O.ik(g[f"{tag}_ik_xyz"][i, leg], leg), g[f"{tag}_ik_q"][i, leg], **TOL)
            np.testing.assert_allclose(O.ik(pos, leg), g[f"{tag}_ik_of_fk"][i, leg], rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("tag", ["s1", "s0"])
@pytest.mark.parametrize("ctrl", ["PD", "CARTESIAN_PD"])
@pytest.mark.parametrize("am", ["DEFAULT", "SYMMETRIC", "SYMMETRIC_NO_HIP"])
def test_action_to_command(analytic, tag, ctrl, am):
    g = analytic
    a, cmd = g[f"{tag}_{ctrl}_{am}_a"], g[f"{tag}_{ctrl}_{am}_cmd"]
    for i in range(len(a)):
        got = O.action_to_command(a[i], enable_springs=(tag == "s1"), control=ctrl, action_mode=am)
        np.testing.assert_allclose(got, cmd[i], rtol=1e-9, atol=1e-9)


def test_backflip_limits(analytic):
    g = analytic
    for i in range(len(g["backflip_PD_SYMMETRIC_a"])):
        got = O.action_to_command(g["backflip_PD_SYMMETRIC_a"][i], True, "PD", "SYMMETRIC", task="BACKFLIP")
        np.testing.assert_allclose(got, g["backflip_PD_SYMMETRIC_cmd"][i], **TOL)


def test_orientation(analytic):
    g = analytic
    for i in range(len(g["orient_quat"])):
        qt = g["orient_quat"][i]
        np.testing.assert_allclose(O.rpy_from_quat(qt), g["orient_rpy"][i], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(O.backflip_pitch(qt, 0), g["orient_bf_pitch0"][i], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(O.backflip_pitch(qt, 1), g["orient_bf_pitch1"][i], rtol=1e-9, atol=1e-9)


def test_butterworth_coefficients(analytic):
    # SURVEY.md App. C
    np.testing.assert_allclose(analytic["filter_b"], [0.00782021, 0.01564042, 0.00782021], atol=1e-8)
    np.testing.assert_allclose(analytic["filter_a"], [1, -1.73472577, 0.7660066], atol=1e-8)


def test_cpg(hopf):
    for gait in ("TROT", "BOUND", "WALK", "PACE"):
        p = hopf[f"{gait}_params"]
        X = hopf[f"{gait}_X0"].copy()
        for t in range(len(hopf[f"{gait}_X"])):
            X, xs, zs = O.cpg_step(X, hopf[f"{gait}_PHI"], *p)
            np.testing.assert_allclose(X, hopf[f"{gait}_X"][t], rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(xs, hopf[f"{gait}_xs"][t], rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(zs, hopf[f"{gait}_zs"][t], rtol=1e-9, atol=1e-11)
    for i in range(len(hopf["law_q"])):
        tau = O.cpg_torque(hopf["law_xs"][i], hopf["law_zs"][i], hopf["law_q"][i], hopf["law_dq"][i], 0.0838,
                           [150, 70, 70], [2, 0.5, 0.5], 2500.0, 40.0)
        np.testing.assert_allclose(tau, hopf["law_tau"][i], rtol=1e-9, atol=1e-8)


@pytest.mark.parametrize("name", ROLLOUTS)
def test_rollout_matches_reference_env(name):
    """qso_env_* (C) against the reference's QuadrupedGymEnv run over the same
    oracle world: pins control flow, tasks, sensors and bookkeeping."""
    g = load_golden(f"rollout_{name}.npz")
    cfg = json.loads(str(g["cfg"]))
    env = O.Env(enable_springs=cfg["enable_springs"], motor_control_mode=cfg["motor_control_mode"],
                action_space_mode=cfg["action_space_mode"], task_env=cfg["task_env"],
                observation_space_mode=cfg["observation_space_mode"],
                enable_action_filter=cfg.get("enable_action_filter", False))
    if "springs" in g.files:   # SPRING_RANDOMIZER: the episode's draw, used by the settle too (env_randomizer.py:101-122)
        env.set_springs(g["springs"][:3], g["springs"][3:6], g["springs"][6:])
    if "masses" in g.files:    # MASS_RANDOMIZER: link masses and payload block of the episode (env_randomizer.py:56-84)
        m = g["masses"]
        for leg in range(4):
            for j in range(3):
                env.world.set_mass(2 + 4 * leg + j, m[j])
        env.world.set_mass(0, m[3])
        env.world.set_payload(m[4], m[5:8])
        assert 0 < m[4] < 1 and abs(m[3] + m[4] + 4 * m[:3].sum() + 0.24 - 12.01301) < 1e-9   # total mass is kept (:56-60)
    if "demo" in g.files:      # *_DEMO tasks: the demonstration the reference task loaded (task_base.py:169-176)
        env.set_demo(g["demo"][:, :env.action_dim])
    if "rsi" in name:          # ReferenceStateInitializationWrapper: start on a demonstration row, unsettled (:24-32)
        from quadruped_springs_b200.demo import demo_rows_to_states
        el = int(g["demo_start"])
        env.set_demo_counter(el)
        obs = env.reset_to_state(demo_rows_to_states(g["demo"][el], env.action_dim), mu=float(g["mu"]))
        assert np.abs(g["init_last_action"]).max() == 0 and env.demo_counter() == el
    else:
        obs = env.reset(mu=float(g["mu"]))
    np.testing.assert_allclose(env.world.get_state(), g["init_state"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(obs, g["init_obs"], rtol=1e-9, atol=1e-10)
    for t in range(len(g["reward"])):
        obs, r, d, tr = env.step(g["actions"][t])
        np.testing.assert_allclose(env.world.get_state(), g["state"][t], rtol=1e-8, atol=1e-9, err_msg=f"step {t}")
        np.testing.assert_allclose(obs, g["obs"][t], rtol=1e-8, atol=1e-9, err_msg=f"obs step {t}")
        np.testing.assert_allclose(r, g["reward"][t], rtol=1e-8, atol=1e-10, err_msg=f"reward step {t}")
        assert d == bool(g["done"][t]) and tr == bool(g["truncated"][t]), t
        ts = env.task_state()
        np.testing.assert_allclose(ts[23:27], g["foot_force"][t], rtol=1e-8, atol=1e-8)
        np.testing.assert_array_equal(ts[19:23], g["foot_contact"][t])
        np.testing.assert_allclose(env.torques()[0], g["tau"][t], rtol=1e-8, atol=1e-9)
        if cfg["task_env"] != "NO_TASK":
            gt = g["task"][t]
            # switched, in_air, t_takeoff, init_h, max_flight, max_fwd, max_pitch, rel_max_h, max_dx, max_h
            np.testing.assert_allclose([ts[0], ts[1], ts[2], ts[6], ts[8], ts[9], ts[10], ts[11], ts[12], ts[13]],
                                       gt[:10], rtol=1e-8, atol=1e-10, err_msg=f"task step {t}")
            if cfg["task_env"].startswith("CONTINUOUS"):
                # is_jumping, cumulative_fwd, cumulative_flight_time | jump_counter, good_jumps, first_jump, max_jump_h, end_jump
                fwd, perf, cnt = env.jump_arrays()
                np.testing.assert_allclose(ts[29:32], gt[13:16], rtol=1e-8, atol=1e-10, err_msg=f"continuous step {t}")
                if cfg["task_env"] in ("CONTINUOUS_JUMPING_FORWARD3", "CONTINUOUS_JUMPING_FORWARD_PPO"):
                    np.testing.assert_allclose(cnt, gt[16:21], rtol=1e-8, atol=1e-10, err_msg=f"jump counters step {t}")
    if "demo" in g.files:
        assert env.demo_counter() == int(g["demo_counter_end"])
    if "jumps" in g.files:
        fwd, perf, _ = env.jump_arrays()
        np.testing.assert_allclose(fwd, g["jumps"][0], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(perf, g["jumps"][1], rtol=1e-8, atol=1e-10)


LANDINGS = ["w1_jip_pd", "w1_jf_cartesian", "w2_jip_pd_nosprings", "w2_jf_cartesian", "w3_continuous", "w4_backflip",
            "w5_backflip2", "w5_backflip2_late", "rest_w2_jip_pd", "rest_w2_jf_cartesian", "rest_only_jip_pd_full"]


@pytest.mark.parametrize("name", LANDINGS)
def test_landing_controller_matches_reference_wrapper(name):
    """The reference's LandingWrapper / LandingWrapper2 loop env.step inside one wrapper step
    (landing_wrapper.py:38-66, landing_wrapper_2.py:39-72).  The oracle runs the same control flow
    as a mode machine, one qso_env_step per inner step: fed with the policy's action at every
    control step, it must apply the action the wrapper applied (hold, then landing action, with the
    landing gains) and reproduce every inner state, reward and done."""
    g = load_golden(f"landing_{name}.npz")
    cfg = json.loads(str(g["cfg"]))
    env = O.Env(enable_springs=cfg["enable_springs"], motor_control_mode=cfg["motor_control_mode"],
                action_space_mode=cfg["action_space_mode"], task_env=cfg["task_env"],
                observation_space_mode=cfg["observation_space_mode"], landing_mode=int(g["landing_mode"]),
                rest_mode=int(g["rest_mode"]) if "rest_mode" in g.files else 0)
    obs = env.reset(mu=float(g["mu"]))
    np.testing.assert_allclose(env.world.get_state(), g["init_state"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(obs, g["init_obs"], rtol=1e-9, atol=1e-10)
    modes, rest = [], 0
    for t in range(len(g["reward"])):
        obs, r, d, tr = env.step(g["policy_action"][t])      # ignored while the controller is scripted
        rest += env.rest_state()[0]
        np.testing.assert_allclose(env.last_action(), g["applied_action"][t], rtol=1e-9, atol=1e-12, err_msg=f"action {t}")
        np.testing.assert_allclose(env.world.get_state(), g["state"][t], rtol=1e-8, atol=1e-9, err_msg=f"step {t}")
        np.testing.assert_allclose(obs, g["obs"][t], rtol=1e-8, atol=1e-9, err_msg=f"obs {t}")
        np.testing.assert_allclose(r, g["reward"][t], rtol=1e-8, atol=1e-10, err_msg=f"reward {t}")
        assert d == bool(g["done"][t]) and tr == bool(g["truncated"][t]), t
        np.testing.assert_allclose(env.torques()[0], g["tau"][t], rtol=1e-8, atol=1e-9)
        modes.append(env.landing_state()[0])
    # the wrapper returns the reward / done of its LAST inner step
    last = np.flatnonzero(np.diff(np.append(g["wrapper_step"], -1)) != 0)
    np.testing.assert_allclose(g["reward"][last], g["wrapper_out"][:, 0], rtol=0, atol=0)
    lm = int(g["landing_mode"])
    if "rest_mode" in g.files and int(g["rest_mode"]):
        assert rest > 5                                      # GoToRestWrapper took over (go_to_rest_wrapper.py:58-81)
        np.testing.assert_allclose(g["kp"][-1], 60.0)        # and ran on its own gains until the episode ended
    if lm == 0:
        return
    assert (4 if lm >= 4 else 1) in modes and 2 in modes     # take-off phase and landing both happened
    if lm == 2:
        assert 3 in modes                                    # LandingWrapper2 hands control back to the policy
    if lm == 3:                                              # the continuous variant re-arms after every jump
        assert any(a == 2 and b == 0 for a, b in zip(modes, modes[1:]))


def test_self_collision_contact_info_matches_reference():
    """tests/golden/self_contact_info.npz: the reference's own GetContactInfo (quadruped.py:224-258) over getContactPoints
    rows with bodyA == bodyB, 96 airborne states (half with crossed legs), one stepSimulation each.  The oracle env's
    contact bookkeeping must count the same invalid contacts; the states without a hit must have none."""
    from oracle import oracle as O
    g = load_golden("self_contact_info.npz")
    w = O.World()
    for s, info in zip(g["state"], g["info"]):
        w.set_state(s)
        w.step()
        calf_pairs = [(a, b) for a, b, d in w.self_contacts() if a in (4, 8, 12, 16) or b in (4, 8, 12, 16)]
        assert len(calf_pairs) == int(info[1]) and len(w.contacts()) == 0
    # parent-child pairs are never reported, and the flag turns the detection off
    assert all((b - a) not in (1, -1) or a // 4 != b // 4 for s in g["state"][:8] for a, b, _ in (w.set_state(s), w.self_contacts(detect=True))[1])
    w2 = O.World(self_collision=0)
    w2.set_state(g["state"][0]); w2.step()
    assert w2.self_contacts() == []
