#!/usr/bin/env python
"""bench.py -- control env-steps/s of the batched Go1+PEA step path on B200.

    python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one control step (action_repeat = 10 physics ticks + task/obs
epilogue, finished envs re-settled in place) of every env of the job.
Workload = BASELINE.json configs[2]: Go1 + PEA, JUMPING_FORWARD, CARTESIAN_PD
(IK -> joint PD), SYMMETRIC actions, ARS_BASIC observations, 65536 envs per
GPU, uniform random actions.  Envs are independent, so N GPUs run N shards of
65536 envs with no data-path collective ("weak" scaling); NCCL is used for the
timing barrier / max-over-ranks and for the all-gather of rollout statistics.

The reference arm times the CPU restatement of the reference path (the oracle,
kind "port": the reference itself needs pybullet, which is neither vendored nor
installable here) on all host cores, one process per core.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(enable_springs=True, task_env="JUMPING_FORWARD", motor_control_mode="CARTESIAN_PD",
                action_space_mode="SYMMETRIC", observation_space_mode="ARS_BASIC")
WORKLOAD_NAME = "go1_pea_jumping_forward_cartesian_pd_65536env_per_gpu"
ENVS_PER_GPU = 65536
METRIC = "control_env_steps_per_sec"
UNIT = "env-steps/s"
SM_COUNT, FP32_LANES = 148, 128


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(args):
    seed, budget_s, max_steps = args
    import numpy as np
    from oracle import oracle as O
    env = O.Env(enable_limits=1, body_contact_response=1, **WORKLOAD)
    rng = np.random.default_rng(seed)
    t_reset = time.perf_counter()
    env.reset(mu=0.5 + 0.5 * rng.random())
    t_reset = time.perf_counter() - t_reset
    steps = resets = 0
    t0 = time.perf_counter()
    while steps < max_steps and time.perf_counter() - t0 < budget_s:
        o, r, d, tr = env.step(rng.uniform(-1, 1, env.action_dim))
        steps += 1
        if d:
            env.reset(mu=0.5 + 0.5 * rng.random())
            resets += 1
    return steps, time.perf_counter() - t0, resets, t_reset


def cpu_sample(budget_s, max_steps=10**9, procs=None, seed0=0):
    """P processes (one per host core) each stepping the oracle env; returns aggregate env-steps/s"""
    from oracle import oracle as O
    O.build()
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(seed0 + i, budget_s, max_steps) for i in range(procs)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return dict(value=steps / wall, steps=steps, wall_s=wall, cores=procs, resets=sum(r[2] for r in res),
                reset_s=statistics.mean(r[3] for r in res))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step_budget = min(8.0, 150.0 / max(args.steps + args.warmup, 1))
    for _ in range(args.warmup):
        cpu_sample(min(per_step_budget, 2.0))
    vals, cores, tot_steps, tot_wall = [], 0, 0, 0.0
    for i in range(args.steps):
        s = cpu_sample(per_step_budget, seed0=1000 * (i + 1))
        vals.append(s["value"]); cores = s["cores"]; tot_steps += s["steps"]; tot_wall += s["wall_s"]
    value = tot_steps / tot_wall
    sample = (f"{cores} processes x {per_step_budget:.1f} s per step of the oracle env (C port, fp64) on the same "
              f"workload, one env per process, re-reset (2500 settle ticks) on done")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_wall / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, **WORKLOAD, "envs": cores, "host": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples), "reasons": reasons}


# ----------------------------------------------------------------------------- ours
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import quadruped_springs_b200 as qs
    from quadruped_springs_b200 import _lib, stats

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs_per_gpu
    L = _lib.lib()
    env = qs.BatchedQuadrupedGymEnv(num_envs=n, device=dev, seed=args.seed, env_id_offset=rank * n, auto_reset=True,
                                    enable_noise=True, **WORKLOAD)
    A, O = env.action_dim, env.obs_dim
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def fresh_action():
        # fixed-seed uniform(-1, 1) actions, a new draw for every env and step, generated on the device
        return (torch.rand(n, A, device=dev, generator=gen) * 2 - 1).contiguous()
    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident timing ("value")
    for i in range(args.warmup):
        env.step(fresh_action())
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    work0, swork0 = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
    _lib.check(L.qs_work_counters(env._h, work0, None))
    _lib.check(L.qs_settle_work_counters(env._h, swork0, None))
    launches0 = L.qs_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(float(i))          # L2 flush between timed iterations (outside the event pair)
        a = fresh_action()             # inputs resident in HBM before the timed region of the step starts
        ev[i][0].record()
        env.step(a)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = L.qs_launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    kms = C.c_float()
    _lib.check(L.qs_step_kernel_time(env._h, min(args.steps, 512), C.byref(kms)))
    k_step_ms = kms.value / min(args.steps, 512)
    _lib.check(L.qs_settle_kernel_time(env._h, min(args.steps, 512), C.byref(kms)))
    k_settle_ms = kms.value / min(args.steps, 512)
    work1, swork1 = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
    _lib.check(L.qs_work_counters(env._h, work1, None))
    _lib.check(L.qs_settle_work_counters(env._h, swork1, None))
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    ms_per_step = dev_ms_max / args.steps
    value = world * n * args.steps / (dev_ms_max * 1e-3)

    # ---------------- end-to-end through the host-buffer C-ABI call ("e2e")
    n_act = 16
    h_act = [torch.empty(n, A, dtype=torch.float32).pin_memory() for _ in range(n_act)]
    for i in range(n_act):
        h_act[i].copy_(fresh_action())
    h_out = (torch.empty(n, O, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory(),
             torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory())
    np_act = [a.numpy() for a in h_act]
    np_out = tuple(x.numpy() for x in h_out)
    e2e_steps = max(min(args.steps, 200), 5)   # long enough to contain the periodic settle batches
    for i in range(2):
        env.step_host(np_act[i % n_act], np_out)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        env.step_host(np_act[i % n_act], np_out)   # H2D actions, step(+resets), D2H obs/reward/done/truncated, sync
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t.item())

    # ---------------- rollout statistics: the only exchange of the path (one all-gather of 16 floats per shard)
    rollout = stats.gather_rollout_stats(env.rollout_stats())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline, FP32 pipe: the settle slice (dominant by time) and the step kernels
    fm = json.load(open(os.path.join(ROOT, "quadruped_springs_b200", "flop_model.json")))

    def tick_flops(w0, w1):
        ticks, cticks, csweeps = (int(w1[i] - w0[i]) for i in range(3))
        f = (ticks * fm["W0_flight_tick"] + cticks * (fm["W_per_contact"] + fm["W_any_contact"] / 4.0)
             + csweeps * fm["W_per_contact_sweep"])
        return f, ticks, cticks, csweeps

    step_flops, ticks, cticks, csweeps = tick_flops(work0, work1)
    step_flops += n * args.steps * fm["epilogue_per_control_step_estimate"]
    settle_flops, sticks, scticks, scsweeps = tick_flops(swork0, swork1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_mhz_max = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = SM_COUNT * FP32_LANES * 2 * sm_mhz_max * 1e6 / 1e12     # TFLOP/s, non-tensor FP32 FMA
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    state_bytes = n * 4 * (2 * (37 + 12 + 12 + 4 + 1 + 29 + 12 + 3 + 3) + 24 + 9 + 1 + A + O + 2)  # per step, algorithmic

    def fp32_line(kernel, flops, ms, extra):
        ach = flops / args.steps / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        return {"bound": "fp32", "kernel": kernel, "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach / fp32_peak, "kernel_ms": ms, "kernel_share_of_step": ms / ms_per_step, **extra}

    settle_line = fp32_line("k_settle_slice", settle_flops, k_settle_ms, {
        "what": "reset()'s 2500-tick settle of the next episodes: two slices of ticks per control step on a second stream "
                "(timed there with CUDA events; the early one runs next to k_step_contact, the late one next to "
                "k_step_slow); achieved = flops of both launches / their summed duration",
        "launches_per_step": 2,
        "settle_ticks_per_step": sticks / args.steps, "algorithmic_flops_per_settle_tick": settle_flops / max(sticks, 1),
        "mean_foot_contacts_per_tick": scticks / max(sticks, 1), "mean_pgs_sweeps_per_contact_tick": scsweeps / max(scticks, 1)})
    step_line = fp32_line("k_step + k_step_contact", step_flops, k_step_ms, {
        "what": "the action_repeat physics ticks of step(): flight variant, then the envs with foot contacts",
        "algorithmic_flops_per_env_step": step_flops / args.steps / n,
        "mean_foot_contacts_per_tick": cticks / max(ticks, 1), "mean_pgs_sweeps_per_contact_tick": csweeps / max(cticks, 1),
        "hbm": {"algorithmic_bytes_per_launch": state_bytes, "achieved_GBps": state_bytes / (k_step_ms * 1e-3) / 1e9,
                "peak_GBps": hbm_peak, "frac": state_bytes / (k_step_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"}})
    dominant, other = (settle_line, step_line) if k_settle_ms >= k_step_ms else (step_line, settle_line)
    roofline = dict(dominant)
    roofline["traffic"] = None
    roofline["peak_source"] = (f"{SM_COUNT} SMs x {FP32_LANES} FP32 lanes x 2 x {sm_mhz_max:.0f} MHz (clocks.max.sm from "
                               f"{'MEASURED_PEAKS.json' if peaks else 'nominal'}); no tensor cores: per-env matrices <= 6x6")
    roofline["other_kernels"] = [other]
    ncu_path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(ncu_path):
        try:
            roofline["traffic"] = json.load(open(ncu_path)).get(dominant["kernel"], {}).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---------------- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        s = cpu_sample(args.cpu_seconds)
        cpu = {"value": s["value"], "unit": UNIT, "cores": s["cores"], "kind": "port",
               "sample": f"{s['cores']} processes x {args.cpu_seconds:.0f} s of the oracle env (C port of the reference "
                         f"path, fp64) on the same workload; {s['steps']} env-steps, {s['resets']} resets, "
                         f"one reset = {s['reset_s']*1e3:.0f} ms"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, **WORKLOAD, "envs_per_gpu": n, "envs_total": world * n, "action_repeat": 10,
                   "actions": "fixed-seed uniform(-1,1), fresh draw per env and step, generated on device before each timed step", "auto_reset": True,
                   "sensor_noise": True, "l2": "flushed between timed iterations (192 MiB fill outside the event pair)",
                   "parallelism": f"env-sharded x{world}, no per-step collective"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * A * 4, "d2h_bytes_per_step": n * (O * 4 + 4 + 2),
                "steps": e2e_steps, "path": "qs_step_host: pinned host actions -> H2D -> step kernels + settle slice -> D2H obs, "
                                            "reward, done, truncated -> stream sync"},
        "roofline": roofline, "cpu_baseline": cpu,
        "wall_s_timed_region": t_wall,
        "rollout_stats": {k: rollout[k] for k in ("episodes", "mean_length", "mean_max_height", "mean_max_fwd",
                                                  "mean_return", "terminated_fraction")},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=100)  # past the start-up transient (all envs begin in phase)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
