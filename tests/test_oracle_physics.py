"""Physics-level checks of the CPU oracle (no GPU): model tables against the
reference URDF fixture, mass matrix and gravity against an independent numpy
model, conservation laws, contact statics."""
import numpy as np
import pytest

from oracle import oracle as O
from urdf_numpy_model import LINK_ORDER, Model, bullet_inertia_diag


def random_state(rng, height=3.0, vel=1.0):
    s = np.zeros(37)
    s[0:3] = [rng.normal() * 0.1, rng.normal() * 0.1, height]
    q = rng.normal(size=4)
    s[3:7] = q / np.linalg.norm(q)
    s[7:13] = rng.normal(size=6) * vel
    s[13:25] = np.array([0, np.pi / 4, -np.pi / 2] * 4) + rng.normal(size=12) * 0.3
    s[25:37] = rng.normal(size=12) * 3 * vel
    return s


def test_model_tables_match_urdf():
    m = Model()
    w = O.World()
    for i, name in enumerate(LINK_ORDER):
        mass, I, com = w.dynamics(i - 1)
        assert mass == pytest.approx(m.links[name]["mass"], rel=1e-12)
        np.testing.assert_allclose(com, m.links[name]["com"], atol=1e-12)
        np.testing.assert_allclose(I, bullet_inertia_diag(m.links[name]), rtol=1e-9, atol=1e-15)
    # SURVEY.md App. B.2 expected diagonals
    np.testing.assert_allclose(w.dynamics(0)[1], [0.009427, 0.067012, 0.065167], rtol=2e-4)
    np.testing.assert_allclose(w.dynamics(2)[1], [0.00049566, 0.00083370, 0.00049566], rtol=2e-4)
    np.testing.assert_allclose(w.dynamics(3)[1], [0.0035243, 0.0035669, 0.00013465], rtol=2e-4)
    np.testing.assert_allclose(w.dynamics(4)[1], [0.00049807, 0.00049807, 5.59e-6], rtol=2e-3)
    np.testing.assert_allclose(w.dynamics(5)[1], [9.6e-6] * 3, rtol=1e-12)
    assert m.total_mass() == pytest.approx(12.01301, abs=1e-9)


def test_link_poses_match_independent_fk():
    m = Model()
    w = O.World()
    rng = np.random.default_rng(3)
    for _ in range(5):
        s = random_state(rng)
        w.set_state(s)
        poses, _ = m.fk(s[0:3], s[3:7], s[13:25])
        for i, name in enumerate(LINK_ORDER):
            R, p = w.link_pose(i)
            np.testing.assert_allclose(R, poses[name][0], atol=1e-12)
            np.testing.assert_allclose(p, poses[name][1], atol=1e-12)


def test_mass_matrix_and_gravity_match_independent_model():
    m = Model()
    w = O.World(enable_limits=0)
    rng = np.random.default_rng(4)
    for _ in range(5):
        s = random_state(rng, vel=0.0)
        w.set_state(s)
        M = w.mass_matrix()
        Mi = m.mass_matrix(s[0:3], s[3:7], s[13:25])
        np.testing.assert_allclose(M, Mi, rtol=1e-9, atol=1e-12)
        # zero velocity, zero torque: M nudot = -dV/dq (joint part by central differences)
        acc = w.accel()
        lhs = M @ acc
        eps = 1e-6
        for j in range(12):
            qp, qm = s[13:25].copy(), s[13:25].copy()
            qp[j] += eps; qm[j] -= eps
            dV = (m.potential(s[0:3], s[3:7], qp) - m.potential(s[0:3], s[3:7], qm)) / (2 * eps)
            assert lhs[6 + j] == pytest.approx(-dV, abs=1e-6)
        # base linear part: total weight expressed in base coordinates
        from urdf_numpy_model import quat_R
        np.testing.assert_allclose(lhs[3:6], quat_R(s[3:7]).T @ [0, 0, -9.8 * m.total_mass()], atol=1e-9)


def test_free_flight_conservation():
    w = O.World(enable_limits=0)
    rng = np.random.default_rng(0)
    s = random_state(rng, height=50.0)
    drift = []
    for dt in (1e-3, 2.5e-4):
        w.set_params(dt=dt)
        w.set_state(s)
        E0, P0, L0 = w.energy()
        for _ in range(int(round(0.5 / dt))):
            w.step()
        E1, P1, L1 = w.energy()
        assert len(w.contacts()) == 0
        np.testing.assert_allclose(P1[:2], P0[:2], atol=2e-2 * dt / 1e-3)
        assert P1[2] - P0[2] == pytest.approx(-9.8 * 12.01301 * 0.5, rel=1e-4)
        drift.append(abs(E1 - E0))
    assert drift[0] < 1.0            # symplectic-Euler drift stays small ...
    assert drift[1] < 0.4 * drift[0]  # ... and shrinks with dt (first order)


def test_power_balance_with_torques():
    # dE/dt = tau . qd for the unconstrained dynamics
    w = O.World(enable_limits=0, dt=1e-5, max_coord_vel=1e9)
    rng = np.random.default_rng(7)
    s = random_state(rng, height=20.0)
    tau = rng.normal(size=12) * 5
    w.set_state(s)
    E0 = w.energy()[0]
    work = 0.0
    for _ in range(2000):
        qd0 = w.get_state()[25:37]
        w.step(tau)
        qd1 = w.get_state()[25:37]
        work += 1e-5 * tau @ (0.5 * (qd0 + qd1))
    E1 = w.energy()[0]
    assert E1 - E0 == pytest.approx(work, rel=2e-3, abs=1e-3)


def test_standing_contact_statics_and_settle_depends_weakly_on_mu():
    states = []
    for mu in (0.5, 0.999):
        e = O.Env(enable_springs=True, task_env="JUMPING_IN_PLACE", observation_space_mode="ARS_BASIC")
        e.reset(mu=mu)
        forces = [c[1] for c in e.world.contacts() if c[0] in (5, 9, 13, 17)]
        assert len(forces) == 4
        assert sum(forces) == pytest.approx(12.01301 * 9.8, rel=2e-3)
        states.append(e.world.get_state())
    # friction saturates briefly while the legs splay on touch-down, so the settled state
    # depends (weakly) on mu: a cached settle would NOT be exact -> the product settles per env
    assert 0 < np.abs(states[0] - states[1]).max() < 1e-2


def test_joint_limit_row_stops_the_joint():
    w = O.World()
    s = w.get_state()
    s[2] = 5.0
    s[13 + 2] = -0.85   # FR calf close to its upper limit -0.8378
    s[25 + 2] = 5.0     # moving into it
    w.set_state(s)
    for _ in range(40):
        w.step()
    st = w.get_state()
    assert st[13 + 2] < -0.8378 + 2e-3
    assert abs(st[25 + 2]) < 0.5


def test_payload_block_adds_a_welded_rigid_body_to_the_trunk():
    """Quadruped._add_base_mass_offset (quadruped.py:778-819) restated as a welded block: the generalized mass matrix
    gains exactly the block's mass, first moment and inertia about the base origin (0.1 m cube)."""
    w0, w1 = O.World(), O.World()
    mp, p = 0.73, np.array([0.08, 0.0, -0.06])
    w1.set_payload(mp, p)
    s = w0.get_state()
    s[13:25] += np.random.default_rng(0).normal(size=12) * 0.2
    w0.set_state(s); w1.set_state(s)
    dM = w1.mass_matrix() - w0.mass_matrix()
    px = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
    np.testing.assert_allclose(dM[3:6, 3:6], mp * np.eye(3), atol=1e-12)
    np.testing.assert_allclose(dM[0:3, 0:3], mp * 0.01 / 6 * np.eye(3) + mp * (p @ p * np.eye(3) - np.outer(p, p)), atol=1e-12)
    np.testing.assert_allclose(dM[0:3, 3:6], mp * px, atol=1e-12)
    np.testing.assert_allclose(dM[6:, :], 0, atol=1e-12)     # the legs do not see it
    # free fall is unchanged, and the robot still falls at g
    for w in (w0, w1):
        s2 = s.copy(); s2[2] = 1.0
        w.set_state(s2); w.step()
        assert w.get_state()[9] == pytest.approx(-9.8e-3, rel=1e-6)
