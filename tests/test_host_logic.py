"""CPU-side checks of the product's host layer: the C-ABI library builds, loads
and exports every symbol of include/qs_b200.h; registries, observation spaces
and noise tables match the reference fixtures; sharding / statistics helpers.
No compute call is made (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from quadruped_springs_b200 import _lib, configs, ops, stats
from quadruped_springs_b200.env import (ActionInterfaceCollection, MotorInterfaceCollection, SensorCollection,
                                        TaskCollection)


def test_library_builds_loads_and_exports_every_header_symbol():
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "qs_b200.h")).read()
    declared = set(re.findall(r"\b(qs_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/qs_b200.h but not exported"
    assert set(_lib.EXPORTS) == declared


def test_config_struct_layout_matches_c():
    L = _lib.lib()
    cfg = _lib.QsConfig()
    L.qs_default_config(C.byref(cfg))
    assert cfg.action_repeat == 10 and cfg.settling_steps == 2500 and cfg.is_rl_interface == 1
    assert cfg.time_step == 0.001 and cfg.max_episode_time == 10.0
    assert cfg.gravity_z == pytest.approx(-9.8) and cfg.max_coord_vel == pytest.approx(30.1)
    assert cfg.breaking_threshold == pytest.approx(0.02) and cfg.warmstart == pytest.approx(0.1)
    # the fields after the solver block (wrappers, randomizers): a layout drift would scramble these
    assert cfg.landing_mode == 0 and cfg.spring_randomizer == 0 and cfg.rest_mode == 0 and cfg.mass_randomizer == 0
    assert cfg.rand_leg_mass_err == pytest.approx(0.1) and cfg.rand_payload_max == pytest.approx(1.0)
    assert list(cfg.rand_payload_pos) == pytest.approx([0.1, 0.0, 0.1]) and cfg.rand_spring_err == pytest.approx(0.1)


def test_integration_md_binding_matches_the_header_struct():
    """the ctypes QsConfig shown to reference maintainers in INTEGRATION.md has the fields of include/qs_b200.h, in order"""
    header = open(os.path.join(ROOT, "include", "qs_b200.h")).read()
    body = header[header.index("typedef struct qs_config {"):header.index("} qs_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        m = re.match(r"\s*(?:int32_t|uint64_t|int64_t|double|float)\s+(.*)", decl, flags=re.S)
        if m:
            fields += [re.match(r"[a-z_0-9]+", x.strip()).group(0) for x in m.group(1).split(",")]
    assert fields == [f[0] for f in _lib.QsConfig._fields_]
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index("class QsConfig(C.Structure)"):doc.index("CONTROL = {")]
    names = re.findall(r'"([a-z_0-9]+)"', block)
    assert names == fields


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    cfg = _lib.QsConfig()
    L.qs_default_config(C.byref(cfg))
    h = C.c_void_p()
    assert L.qs_create(C.byref(cfg), 8, 0, C.byref(h)) == -2  # QS_ERR_CUDA, never a CPU fallback
    assert b"cuda" in L.qs_last_error().lower()
    from quadruped_springs_b200 import BatchedQuadrupedGymEnv
    with pytest.raises(RuntimeError):
        BatchedQuadrupedGymEnv(num_envs=2)


def test_invalid_configs_are_rejected():
    L = _lib.lib()
    cfg = _lib.QsConfig()
    L.qs_default_config(C.byref(cfg))
    cfg.control_mode = 2  # TORQUE with the RL interface (quadruped_gym_env.py:167-168)
    out = (C.c_float * 32)()
    assert L.qs_obs_noise_std(C.byref(cfg), out) == -1
    assert b"TORQUE" in L.qs_last_error()
    for reg, bad in ((TaskCollection(), "LR_COURSE_TASK"), (SensorCollection(), "ARS_HEIGHT"),
                     (MotorInterfaceCollection(), "PID"), (ActionInterfaceCollection(), "FULL")):
        with pytest.raises(ValueError):
            reg.get_el(bad)


def test_registry_keys_follow_the_reference():
    assert MotorInterfaceCollection().keys() == ["PD", "CARTESIAN_PD", "TORQUE"]
    assert ActionInterfaceCollection().keys() == ["DEFAULT", "SYMMETRIC", "SYMMETRIC_NO_HIP"]
    for k in ("JUMPING_IN_PLACE", "JUMPING_FORWARD", "BACKFLIP", "JUMPING_IN_PLACE_PPO", "JUMPING_FORWARD_PPO",
              "BACKFLIP_PPO", "JUMPING_IN_PLACE_PPO_HP", "JUMPING_FORWARD_PPO_HP", "NO_TASK"):
        TaskCollection().get_el(k)


@pytest.mark.parametrize("springs", [True, False])
def test_observation_spaces_and_noise_match_reference(obs_spaces, springs):
    tag = "s1" if springs else "s0"
    cfg = configs.go1_config(springs)
    for mode in ("ENCODER", "ENCODER_2", "CARTESIAN_NO_IMU", "ARS_BASIC", "ARS_SENSOR", "LANDING_SENSOR", "PPO_BASIC",
                 "PPO_BASIC_X", "PPO_BASIC_CONTACT", "ARS_BACKFLIP", "PPO_BACKFLIP"):
        layout, hi, lo = configs.observation_layout(mode, cfg)
        # float32 Box bounds of the reference (quadruped_gym_env.py:160-164)
        np.testing.assert_allclose((hi + 0.01).astype(np.float32), obs_spaces[f"{tag}_{mode}_high"], rtol=1e-6)
        np.testing.assert_allclose((lo - 0.01).astype(np.float32), obs_spaces[f"{tag}_{mode}_low"], rtol=1e-6)
        got = ops.obs_noise_std(enable_springs=springs, observation_space_mode=mode)
        np.testing.assert_allclose(got, obs_spaces[f"{tag}_{mode}_noise_std"], rtol=1e-6, atol=1e-9)
        assert layout[-1][2] == len(hi) == len(got)


def test_python_config_constants_match_reference(analytic):
    for springs, tag in ((True, "s1"), (False, "s0")):
        cfg = configs.go1_config(springs)
        for name in ("INIT_MOTOR_ANGLES", "RL_UPPER_ANGLE_JOINT", "RL_LOWER_ANGLE_JOINT", "RL_UPPER_CARTESIAN_POS",
                     "RL_LOWER_CARTESIAN_POS", "RL_TORQUE_LIMITS", "MOTOR_KP", "MOTOR_KD", "NOMINAL_FOOT_POS_LEG_FRAME",
                     "IS_FALLEN_HEIGHT", "JOINT_ANGLES_NOISE", "JOINT_VELOCITIES_NOISE", "HEIGHT_NOISE", "PITCH_NOISE",
                     "VEL_LIN_NOISE", "VEL_ANG_NOISE", "PITCH_RATE_NOISE", "FEET_POS_NOISE", "FEET_VEL_NOISE"):
            np.testing.assert_allclose(np.asarray(getattr(cfg, name), dtype=np.float64), analytic[f"{tag}_cfg_{name}"],
                                       rtol=1e-12, atol=1e-15, err_msg=name)
    cfg = configs.go1_config(True)
    for name in ("SPRINGS_STIFFNESS", "SPRINGS_DAMPING", "SPRINGS_REST_ANGLE"):
        np.testing.assert_allclose(getattr(cfg, name), analytic[f"s1_cfg_{name}"], rtol=1e-12)


def test_shard_ranges_cover_all_envs():
    for total, world in ((65536, 8), (262144, 8), (1000, 3), (7, 8)):
        seen = []
        for r in range(world):
            s, n = stats.shard_range(total, r, world)
            seen += list(range(s, s + n))
        assert seen == list(range(total))


def test_stats_combine():
    v = torch.zeros(2, 16)
    v[0, :12] = torch.tensor([100, 4, 2.0, 0.9, 1.0, 0.8, 0.5, 0.6, 0.4, 3.0, 400, 2])
    v[1, :12] = torch.tensor([100, 6, 3.0, 1.1, 2.0, 1.2, 0.7, 0.9, 0.6, 5.0, 600, 3])
    out = stats.combine(v)
    assert out["num_envs"] == 200 and out["episodes"] == 10
    assert out["max_max_height"] == pytest.approx(1.1) and out["max_max_fwd"] == pytest.approx(0.7)
    assert out["mean_max_height"] == pytest.approx(0.5) and out["mean_length"] == pytest.approx(100.0)
    assert out["terminated_fraction"] == pytest.approx(0.5)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, n = stats.shard_range(1000, rank, world)
    local = torch.zeros(16)
    local[0] = n
    local[1] = rank + 1          # episodes
    local[2] = 0.5 * (rank + 1)  # sum max height
    local[3] = 0.3 + rank        # max max height
    out = stats.gather_rollout_stats(local)
    q.put((rank, out))
    dist.destroy_process_group()


def test_rollout_stats_all_gather_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert res[r]["num_envs"] == 1000 and res[r]["episodes"] == 3
        assert res[r]["max_max_height"] == pytest.approx(1.3)
        assert res[r]["mean_max_height"] == pytest.approx(1.5 / 3)
