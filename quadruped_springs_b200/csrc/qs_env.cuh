// qs_env.cuh -- per-env control-step logic around the physics tick: motor +
// PEA torques per substep, jumping-task bookkeeping, rewards, terminations and
// observation assembly (fp32).  Restates, per env and in registers, what the
// reference does in Python after its ten stepSimulation calls
// (quadruped_gym_env.py:239-256).
#pragma once
#include "qs_physics.cuh"

namespace qs {

// ---------------------------------------------------------------- Philox4x32-10
QS_DEVONLY void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
QS_DEVONLY float u01(uint32_t x) { return (float(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)
// four N(0,1) samples for (stream, step, block)
QS_DEVONLY void normal4(uint64_t seed, uint64_t gid, uint32_t epoch, uint32_t step, uint32_t blk, float* n) {
  uint32_t r[4];
  philox4x32(uint32_t(gid), uint32_t(gid >> 32) ^ (blk << 16), step, epoch, uint32_t(seed), uint32_t(seed >> 32), r);
  float s, c;
  float m = sqrtf(-2.0f * __logf(u01(r[0])));
  sincospif(2.0f * u01(r[1]), &s, &c);
  n[0] = m * c; n[1] = m * s;
  m = sqrtf(-2.0f * __logf(u01(r[2])));
  sincospif(2.0f * u01(r[3]), &s, &c);
  n[2] = m * c; n[3] = m * s;
}
QS_DEVONLY float uniform1(uint64_t seed, uint64_t gid, uint32_t epoch, uint32_t blk) {
  uint32_t r[4];
  philox4x32(uint32_t(gid), uint32_t(gid >> 32) ^ (blk << 16), 0xFFFFFFFFu, epoch, uint32_t(seed), uint32_t(seed >> 32), r);
  return float(r[0] >> 8) * (1.0f / 16777216.0f);  // [0,1) like np.random.random
}

// ---------------------------------------------------------------- substeps
struct EnvCfg {
  int enable_springs, control_mode, action_mode, task, obs_mode, action_repeat, is_rl, enable_filter;
  int enable_noise, obs_dim, action_dim, settling_steps, ground_randomizer, auto_reset, landing_mode, spring_randomizer, rest_mode, mass_randomizer;
  float max_episode_time, mu_ground, leg_mass_err, payload_max, payload_pos[3], spring_err;
  const float* demo;  // [demo_len][action_dim] demonstration actions of the *_DEMO tasks (device)
  int demo_len;
  uint64_t seed;
  int64_t gid0;
};

// PD + PEA torque of one tick (quadruped.py:288-320, quadruped_motor.py:45-104)
__device__ __forceinline__ void tick_torques(const EnvState<float>& st, const float* cmd, bool torque_mode, int env,
                                             const DeviceView& D, const EnvCfg& C, const RobotConst& RC, const float* sk,
                                             const float* sb, const float* sr, float* tau, float* tau_m, float* tau_s,
                                             bool custom_gains) {
  const int n = D.n;
  if (custom_gains) {  // gains swapped at run time for this env (landing wrappers, landing_wrapper.py:21-33)
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const float kp = D.kp[i * n + env], kd = D.kd[i * n + env];
      tau_m[i] = pd_torque1(kp, kd, RC.tau_max[i], cmd[i], st.q[i], st.qd[i], torque_mode);
    }
  } else {             // config gains straight from the constant bank: no global loads on the per-tick path
#pragma unroll
    for (int i = 0; i < 12; i++) tau_m[i] = pd_torque1(RC.kp[i], RC.kd[i], RC.tau_max[i], cmd[i], st.q[i], st.qd[i], torque_mode);
  }
#pragma unroll
  for (int i = 0; i < 12; i++) tau[i] = tau_m[i];
  if (C.enable_springs) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      spring_torque_leg(k, sk, sb, sr, st.q + 3 * k, st.qd + 3 * k, tau_s + 3 * k);
#pragma unroll
      for (int j = 0; j < 3; j++) tau[3 * k + j] += tau_s[3 * k + j];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 12; i++) tau_s[i] = 0.f;
  }
}

QS_DEVONLY void load_springs(const DeviceView& D, int env, float* sk, float* sb, float* sr) {
  const int n = D.n;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    sk[j] = D.spring[(0 + j) * n + env];
    sb[j] = D.spring[(3 + j) * n + env];
    sr[j] = D.spring[(6 + j) * n + env];
  }
}

// ApplyAction + stepSimulation for ticks [t0, n_ticks) (quadruped_gym_env.py:207-219).
// cmd = desired joint angles (PD) or torques (TORQUE).  Returns n_ticks when all ticks ran,
// or the index of the tick that needs another solver (state untouched by that tick; *why =
// TICK_NEEDS_GENERAL / TICK_NEEDS_CONTACT).  kContacts = false is the flight variant: it
// hands the env over as soon as a foot comes within its contact threshold.
// block size of the step / reset / settle kernels: a compile-time constant so that the per-leg scratch is addressed
// with immediate offsets (qs_physics.cuh Scratch)
constexpr int QS_BLOCK = 128;
using StepScratch = Scratch<float, QS_BLOCK>;

template <bool kContacts, bool kEM>
__device__ __forceinline__ int run_ticks(EnvState<float>& st, ContactState<float>& cs, const float* cmd,
                                         bool torque_mode, int t0, int n_ticks, int env, const DeviceView& D,
                                         const EnvCfg& C, const RobotConst& RC, const ModelConstT<float>& M,
                                         const SolverConst& SC, float* tau_m /*12 out*/, float* tau_s /*12 out*/,
                                         bool detect_invalid_last, const StepScratch& scr, int* why, bool skip = false) {
  const float mu = D.mu[env];
  const bool custom = D.custom_gains[env] != 0;
  float sk[3], sb[3], sr[3];
  load_springs(D, env, sk, sb, sr);
  // skip: this thread only keeps the block's barriers matched (its env is handed over at t0)
  int bail = skip ? t0 : n_ticks;
  *why = skip ? TICK_NEEDS_CONTACT : TICK_DONE;
  for (int t = t0; t < n_ticks; t++) {
    // keep the warps of the block in lockstep: the tick is several times larger than the
    // instruction cache, so warps that run it together share every fetched line
    if (!__syncthreads_or(bail == n_ticks)) break;  // every env of the block was handed over
    if (bail == n_ticks) {
      float tau[12];
      tick_torques(st, cmd, torque_mode, env, D, C, RC, sk, sb, sr, tau, tau_m, tau_s, custom);
      const int r = physics_tick<float, kContacts, QS_BLOCK, kEM>(st, tau, mu, cs, M, SC, detect_invalid_last && (t == n_ticks - 1), scr,
                                                             EnvModelRef{D.model, D.n, env});
      if (r != TICK_DONE) { bail = t; *why = r; }
    }
  }
  return bail;
}

// same loop on the general solver (joint limits, every collision shape); rare path
template <bool kEM>
__device__ __noinline__ void run_ticks_general(EnvState<float>& st, ContactState<float>& cs, const float* cmd,
                                               bool torque_mode, int t0, int n_ticks, int env, const DeviceView& D,
                                               const EnvCfg& C, const RobotConst& RC, const ModelConstT<float>& M,
                                               const SolverConst& SC, float* tau_m, float* tau_s) {
  const float mu = D.mu[env];
  const bool custom = D.custom_gains[env] != 0;
  float sk[3], sb[3], sr[3];
  load_springs(D, env, sk, sb, sr);
  for (int t = t0; t < n_ticks; t++) {
    float tau[12];
    tick_torques(st, cmd, torque_mode, env, D, C, RC, sk, sb, sr, tau, tau_m, tau_s, custom);
    physics_tick_general<float, kEM>(st, tau, mu, cs, M, SC, EnvModelRef{D.model, D.n, env});
  }
}

// ---------------------------------------------------------------- task logic
QS_DEV bool is_jump_task(int task) { return task != QS_TASK_NO_TASK; }
// 0 = TaskJumping, 1 = TaskContinuousJumping, 2 = TaskContinuousJumping2 (task_base.py:222,280)
QS_DEV int task_family(int task) {
  if (task == QS_TASK_CONTINUOUS_JUMPING_FORWARD || task == QS_TASK_CONTINUOUS_JUMPING_FORWARD2) return 1;
  if (task == QS_TASK_CONTINUOUS_JUMPING_FORWARD3 || task == QS_TASK_CONTINUOUS_JUMPING_FORWARD_PPO) return 2;
  return 0;
}
// rows of DeviceView::task this task reads and writes
QS_DEV bool is_demo_task(int task) { return task >= QS_TASK_JUMPING_IN_PLACE_DEMO && task <= QS_TASK_BACKFLIP_DEMO; }
QS_DEV int task_slots(int task) { return task_family(task) ? int(TS_END) : (is_demo_task(task) ? int(TS_END_DEMO) : int(TS_END_BASIC)); }

QS_DEVONLY float jumping_distance(const float* ts, const float* pos) {  // task_base.py:109-116
  float s, c;
  sincosf(ts[TS_TAKEOFF_YAW], &s, &c);
  const float dx = pos[0] - ts[TS_TAKEOFF_X], dy = pos[1] - ts[TS_TAKEOFF_Y];
  return fmaxf(c * dx - s * dy, 0.f);
}

// TaskJumping._on_step (task_base.py:61-107) + subclass overrides
QS_DEVONLY void task_on_step(float* ts, const EnvState<float>& st, const ContactState<float>& cs, const float* tau_m,
                         const float* rpy, const float* Rb, float sim_time, int task) {
  if (!is_jump_task(task)) return;
  const bool flying = (cs.mask & 15) == 0;
  if (ts[TS_SWITCHED] == 0.f && flying && st.vlin[2] / 9.81f > 0.06f) ts[TS_SWITCHED] = 1.f;  // :152-160
  const float z = st.pos[2];
  ts[TS_REL_MAX_H] = fmaxf(ts[TS_REL_MAX_H], fmaxf(z - ts[TS_INIT_HEIGHT], 0.f));
  ts[TS_MAX_H] = fmaxf(ts[TS_MAX_H], fabsf(z));
  ts[TS_MAX_DX] = fmaxf(ts[TS_MAX_DX], fabsf(st.pos[0]));
  ts[TS_MAX_PITCH] = fmaxf(ts[TS_MAX_PITCH], fabsf(rpy[1]));
  const int fam = task_family(task);
  if (fam == 0) {
    if (flying) {
      if (ts[TS_IN_AIR] == 0.f) {
        ts[TS_IN_AIR] = 1.f;
        ts[TS_T_TAKEOFF] = sim_time;
        ts[TS_TAKEOFF_X] = st.pos[0]; ts[TS_TAKEOFF_Y] = st.pos[1]; ts[TS_TAKEOFF_Z] = st.pos[2];
        ts[TS_TAKEOFF_YAW] = rpy[2];
      } else {
        ts[TS_MAX_FWD] = fmaxf(ts[TS_MAX_FWD], jumping_distance(ts, st.pos));
      }
    } else {
      if (ts[TS_IN_AIR] != 0.f) {
        ts[TS_MAX_FLIGHT] = fmaxf(ts[TS_MAX_FLIGHT], sim_time - ts[TS_T_TAKEOFF]);
        ts[TS_MAX_FWD] = fmaxf(ts[TS_MAX_FWD], jumping_distance(ts, st.pos));
        ts[TS_IN_AIR] = 0.f;
      } else {
        ts[TS_MAX_FWD] = 0.f;  // task_base.py:106-107
      }
    }
  } else {
    // TaskContinuousJumping(2)._compute_jumping_info (task_base.py:243-262, 319-338)
    const bool cj2 = fam == 2;
    ts[TS_END_JUMP] = 0.f;
    if (flying) {
      if (ts[TS_IN_AIR] == 0.f) {
        ts[TS_IN_AIR] = 1.f;
        ts[TS_T_TAKEOFF] = sim_time;
        ts[TS_TAKEOFF_X] = st.pos[0]; ts[TS_TAKEOFF_Y] = st.pos[1]; ts[TS_TAKEOFF_Z] = st.pos[2];
        ts[TS_TAKEOFF_YAW] = rpy[2];
        ts[TS_IS_JUMPING] = st.vlin[2] / 9.81f > 0.06f ? 1.f : 0.f;  // detect_jumping, :236-241
        ts[TS_MAX_JUMP_H] = 0.f;  // set to z, then zeroed by restart_jump_performance_variables (:326-330)
      } else if (cj2) {
        ts[TS_MAX_JUMP_H] = fmaxf(ts[TS_MAX_JUMP_H], st.pos[2]);
      }
    } else if (ts[TS_IN_AIR] != 0.f) {
      ts[TS_MAX_FLIGHT] = fmaxf(ts[TS_MAX_FLIGHT], sim_time - ts[TS_T_TAKEOFF]);
      const float d = jumping_distance(ts, st.pos);
      if (!cj2) {  // update_end_jump :264-266
        const float time_limit = task == QS_TASK_CONTINUOUS_JUMPING_FORWARD ? 0.15f : 0.35f;  // robot_tasks.py:105-106,139-140
        ts[TS_MAX_FWD] = fmaxf(ts[TS_MAX_FWD], d);
        ts[TS_CUM_FWD] += fminf(ts[TS_MAX_FWD], 0.5f);
        ts[TS_CUM_FLIGHT] += fminf(ts[TS_MAX_FLIGHT], time_limit);
      } else if (ts[TS_FIRST_JUMP] == 0.f) {  // update_end_jump :340-353
        const bool t3 = task == QS_TASK_CONTINUOUS_JUMPING_FORWARD3;  // robot_tasks.py:172-177 / 553-561
        const float jump_limit = 0.6f, height_limit = t3 ? 0.45f : 0.5f, bound = t3 ? 0.7f : 0.85f;
        const float fwd = fminf(d, jump_limit);
        const float perf = 0.7f * fwd / jump_limit + 0.3f * fminf(ts[TS_MAX_JUMP_H], height_limit) / height_limit;
        ts[TS_JUMP_COUNT] += 1.f;
        if (perf >= bound) ts[TS_GOOD_JUMPS] += 1.f;
        ts[TS_SUM_FWD] += fwd;
        ts[TS_SUM_FLOG] += fwd > 0.f ? fwd * log2f(fwd) : 0.f;
        ts[TS_SUM_PERF] += perf;
        ts[TS_MAX_PERF] = ts[TS_JUMP_COUNT] == 1.f ? perf : fmaxf(ts[TS_MAX_PERF], perf);
        ts[TS_LAST_PERF] = perf;
        ts[TS_END_JUMP] = 1.f;
      } else {
        ts[TS_FIRST_JUMP] = 0.f;
      }
      ts[TS_IN_AIR] = 0.f;
      ts[TS_IS_JUMPING] = 0.f;
    }
  }
  if (task == QS_TASK_BACKFLIP)  // robot_tasks.py:527-530
    ts[TS_MAX_PITCH_BF] = fmaxf(ts[TS_MAX_PITCH_BF], backflip_pitch(Rb, ts[TS_SWITCHED] != 0.f));
  if (task == QS_TASK_BACKFLIP_PPO)  // robot_tasks.py:745-747
    ts[TS_MAX_PITCH] = fmaxf(ts[TS_MAX_PITCH], backflip_pitch(Rb, ts[TS_SWITCHED] != 0.f));
  if (task == QS_TASK_JUMPING_FORWARD_PPO || task == QS_TASK_JUMPING_FORWARD_PPO_HP) {  // :420-426
    ts[TS_OLD_FWD] = ts[TS_ACTUAL_FWD];
    ts[TS_ACTUAL_FWD] = ts[TS_MAX_FWD];
  }
}

QS_DEVONLY bool task_terminated(const float* ts, const EnvState<float>& st, const ContactState<float>& cs,
                            const float* Rb, float fallen_height, int task) {
  if (!is_jump_task(task)) return false;
  const bool fallen_ground = st.pos[2] < fallen_height;  // task_base.py:123-124
  const bool fallen_orient = Rb[8] < 0.85f;              // task_base.py:126-130
  if (task == QS_TASK_BACKFLIP || task == QS_TASK_BACKFLIP_DEMO) return fallen_ground || cs.invalid > 0;  // robot_tasks.py:532-533,239-241
  return (fallen_orient && fallen_ground) || cs.invalid > 0;            // task_base.py:146-147
}

// per-step reward (robot_tasks.py:334-344, 461-471, 783-800); 0 for the sparse tasks.
// old_tau = torque of the previous control step, tau_m = of this one.
QS_DEVONLY float task_reward(const float* ts, const EnvState<float>& st, const float* foot_force, const float* old_tau,
                         const float* tau_m, const float* rpy, const float* Rb, int task) {
  float max_h_task, k_h;
  switch (task) {
    case QS_TASK_JUMPING_IN_PLACE_PPO: max_h_task = 1.0f; k_h = 0.023f; break;
    case QS_TASK_JUMPING_IN_PLACE_PPO_HP: max_h_task = 1.25f; k_h = 0.023f; break;
    case QS_TASK_JUMPING_FORWARD_PPO: max_h_task = 0.9f; k_h = 0.026f; break;
    case QS_TASK_JUMPING_FORWARD_PPO_HP: max_h_task = 1.1f; k_h = 0.026f; break;
    case QS_TASK_BACKFLIP_PPO: max_h_task = 0.7f; k_h = 0.026f; break;
    default: return 0.f;
  }
  const float z = st.pos[2];
  const float h_clip = (z < 0.29f || z > max_h_task) ? 0.f : z;
  const float F = foot_force[0] + foot_force[1] + foot_force[2] + foot_force[3];
  const float over = F > 800.f ? F : 0.f;
  float dn = 0.f;
#pragma unroll
  for (int i = 0; i < 12; i++) { const float d = old_tau[i] - tau_m[i]; dn += d * d; }
  const float rew_h = k_h * h_clip;
  const float rew_smooth = 0.015f * expf(-0.1f * sqrtf(dn));
  const float rew_contact = -3e-4f * over;
  const float rew_pitch = 0.014f * expf(-26.f * fabsf(rpy[1]));
  if (task == QS_TASK_JUMPING_IN_PLACE_PPO || task == QS_TASK_JUMPING_IN_PLACE_PPO_HP) {
    const float rew_pos = 0.013f * expf(-40.f * fabsf(st.pos[0]));
    return 0.05f * rew_pos + 0.5f * rew_contact + 0.2f * rew_smooth + 0.45f * rew_h + 0.3f * rew_pitch;
  }
  if (task == QS_TASK_JUMPING_FORWARD_PPO || task == QS_TASK_JUMPING_FORWARD_PPO_HP) {
    const float max_fwd = task == QS_TASK_JUMPING_FORWARD_PPO ? 1.3f : 1.4f;
    const float fwd = (ts[TS_ACTUAL_FWD] > max_fwd || ts[TS_ACTUAL_FWD] == ts[TS_OLD_FWD]) ? 0.f : ts[TS_ACTUAL_FWD];
    return 0.4f * rew_contact + 0.2f * rew_smooth + 0.25f * rew_h + 0.3f * rew_pitch + 0.4f * (0.038f * fwd);
  }
  const float pbf = z > 0.5f ? backflip_pitch(Rb, ts[TS_SWITCHED] != 0.f) : 0.f;
  return 0.4f * rew_contact + 0.2f * rew_smooth + 0.25f * rew_h + 0.3f * (0.014f * pbf);
}

// end-of-episode bonus / malus (robot_tasks.py:31-57, 70-99, 535-550, 349-358, 476-485, 802-809)
QS_DEVONLY float task_reward_end(const float* ts, bool term, int task, float sim_time, float max_ep_time) {
  float r = 0.f;
  switch (task) {
    case QS_TASK_JUMPING_IN_PLACE: {
      const float h = ts[TS_REL_MAX_H] > 0.9f ? 1.f : ts[TS_REL_MAX_H] / 0.9f;
      r += 0.7f * h;
      r += h * 0.3f * expf(-ts[TS_MAX_PITCH] * ts[TS_MAX_PITCH] / (0.15f * 0.15f));
      r += h * 0.05f * expf(-ts[TS_MAX_DX] * ts[TS_MAX_DX] / 0.05f);
      if (!term) r += 0.1f * h; else r -= 0.08f * (1.f + 0.8f * h);
      return r;
    }
    case QS_TASK_JUMPING_FORWARD: {
      const float h = ts[TS_REL_MAX_H] > 0.3f ? 1.f : ts[TS_REL_MAX_H] / 0.3f;
      const float d = ts[TS_MAX_FWD] > 1.3f ? 1.f : ts[TS_MAX_FWD] / 1.3f;
      const float bm = (h + d) / 2.f;
      r += 0.25f * h;
      r += 0.5f * d * h;
      r += h * 0.25f * expf(-ts[TS_MAX_PITCH] * ts[TS_MAX_PITCH] / (0.15f * 0.15f));
      if (!term) r += 0.1f * bm; else r -= 0.08f * (1.f + 1.2f * bm);
      return r;
    }
    case QS_TASK_BACKFLIP: {
      const float h = fminf(fmaxf(ts[TS_MAX_H] - 0.3f, 0.f), 0.4f) / 0.4f;
      const float p = ts[TS_MAX_PITCH_BF] / float(2 * QS_PI);
      r += p * 0.4f; r += h * 0.4f; r += h * p;
      if (ts[TS_SWITCHED] != 0.f && !term) r += 0.2f;
      return r;
    }
    case QS_TASK_JUMPING_IN_PLACE_PPO:
    case QS_TASK_JUMPING_IN_PLACE_PPO_HP: return term ? -0.25f * ts[TS_MAX_H] : 0.f;
    case QS_TASK_JUMPING_FORWARD_PPO:
    case QS_TASK_JUMPING_FORWARD_PPO_HP: return term ? 0.f : 0.05f * (ts[TS_MAX_FWD] + ts[TS_MAX_H]) / 2.f;
    case QS_TASK_BACKFLIP_PPO: return term ? 0.f : 0.2f * (0.7f * ts[TS_MAX_PITCH] / 5.f + 0.3f * ts[TS_MAX_H]) / 2.f;
    case QS_TASK_CONTINUOUS_JUMPING_FORWARD: {  // robot_tasks.py:112-131
      const float a = ts[TS_CUM_FLIGHT] / 0.15f, b = ts[TS_CUM_FWD] / 0.5f;
      r += 0.25f * a; r += 0.5f * b;
      r += a * 0.25f * expf(-ts[TS_MAX_PITCH] * ts[TS_MAX_PITCH] / (0.15f * 0.15f));
      if (!term) r += 0.1f * (a + b) / 2.f;
      return r;
    }
    case QS_TASK_CONTINUOUS_JUMPING_FORWARD2: {  // robot_tasks.py:146-166
      const float a = fminf(ts[TS_MAX_FLIGHT], 0.35f) / 0.35f, b = fminf(ts[TS_MAX_FWD], 0.5f) / 0.5f, bm = (a + b) / 2.f;
      r += 0.25f * a; r += 0.5f * b;
      r += b * 0.15f * expf(-(ts[TS_MAX_PITCH] * ts[TS_MAX_PITCH] / (0.15f * 0.15f)));
      r += 0.4f * (sim_time / max_ep_time) * bm;
      if (!term) r += 0.2f * bm;
      return r;
    }
    case QS_TASK_CONTINUOUS_JUMPING_FORWARD3:      // robot_tasks.py:183-212
    case QS_TASK_CONTINUOUS_JUMPING_FORWARD_PPO: {  // robot_tasks.py:687-698
      const float n = ts[TS_JUMP_COUNT], size = fmaxf(n, 3.f);  // arrays are zero-padded to >= 3 entries
      const float avg = ts[TS_SUM_PERF] / size;
      float entropy = 0.f;  // get_entropy_fwd, task_base.py:376-383: -sum p log2 p / log2(size), p = f / S
      if (n > 0.f && ts[TS_SUM_FWD] >= 0.05f)
        entropy = (log2f(ts[TS_SUM_FWD]) - ts[TS_SUM_FLOG] / ts[TS_SUM_FWD]) / log2f(size);
      const float rew_entropy = expf((entropy - 1.f) / 0.3f);
      if (task == QS_TASK_CONTINUOUS_JUMPING_FORWARD_PPO) return avg * rew_entropy - (term ? 1.f : 0.f);
      const float mx = n < 3.f ? fmaxf(n > 0.f ? ts[TS_MAX_PERF] : 0.f, 0.f) : ts[TS_MAX_PERF];
      float rew_avg = avg * 0.15f * expf(-ts[TS_MAX_PITCH] * ts[TS_MAX_PITCH] / (0.15f * 0.15f));
      rew_avg += avg * 0.4f * (sim_time / max_ep_time);
      rew_avg += avg * rew_entropy * 0.2f;
      rew_avg += avg * 0.25f;
      r = 0.8f * rew_avg + 0.2f * mx;
      r += 0.1f * ts[TS_GOOD_JUMPS];
      if (!term) r += 0.2f * avg;
      return r;
    }
    default: return 0.f;
  }
}

// TaskJumping._reset (task_base.py:40-59): reset_params + one _on_step
QS_DEVONLY void task_reset(float* ts, const EnvState<float>& st, const ContactState<float>& cs, const float* tau_m,
                       const float* rpy, const float* Rb, float sim_time, int task) {
  if (!is_jump_task(task)) return;
  const float keep_bf = ts[TS_MAX_PITCH_BF];  // BackFlip.max_pitch lives in __init__ only (robot_tasks.py:524)
#pragma unroll
  for (int i = 0; i < TS_END; i++) ts[i] = 0.f;
  ts[TS_MAX_PITCH_BF] = keep_bf;
  ts[TS_FIRST_JUMP] = 1.f;
  ts[TS_T_TAKEOFF] = sim_time;
  ts[TS_TAKEOFF_X] = st.pos[0]; ts[TS_TAKEOFF_Y] = st.pos[1]; ts[TS_TAKEOFF_Z] = st.pos[2];
  ts[TS_INIT_HEIGHT] = st.pos[2];
  ts[TS_TAKEOFF_YAW] = rpy[2];
#pragma unroll
  for (int i = 0; i < 12; i++) ts[TS_OLD_TAU0 + i] = tau_m[i];
  task_on_step(ts, st, cs, tau_m, rpy, Rb, sim_time, task);
}

// ---------------------------------------------------------------- sensors
// SensorList.get_obs / get_noisy_obs (sensor.py:101-111) for the modes of
// sensor_collection.py:18-105; obs is written row-major [N, O].
QS_DEVONLY void observe(const EnvState<float>& st, const ContactState<float>& cs, const float* ts, const float* rpy,
                    const float* Rb, int obs_mode, int task, float* o /*QS_MAX_OBS regs*/) {
  const float* q = st.q;
  const float* qd = st.qd;
  float wl[3];
  m3t_v(Rb, st.vang, wl);  // quadruped.py:141-170
  const float landing = ts[TS_SWITCHED];
  const float jumping = task_family(task) ? ts[TS_IS_JUMPING] : 0.f;  // the attribute exists on those tasks only
  int n = 0;
#define PUT(x) o[n++] = (x)
#define PUT12(p) _Pragma("unroll") for (int _i = 0; _i < 12; _i++) o[n++] = (p)[_i]
  switch (obs_mode) {
    case QS_OBS_ENCODER: PUT12(q); PUT12(qd); break;
    case QS_OBS_ENCODER_2:
      PUT(st.vlin[0]); PUT(st.vlin[1]); PUT(st.vlin[2]); PUT(st.vang[0]); PUT(st.vang[1]); PUT(st.vang[2]);
      PUT12(q); PUT12(qd); break;
    case QS_OBS_CARTESIAN_NO_IMU: {
      float fp[12], fv[12];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float J[9];
        fk_jacobian(q + 3 * k, k, fp + 3 * k, J);
#pragma unroll
        for (int a = 0; a < 3; a++)
          fv[3 * k + a] = J[3 * a] * qd[3 * k] + J[3 * a + 1] * qd[3 * k + 1] + J[3 * a + 2] * qd[3 * k + 2];
      }
      PUT12(fp); PUT12(fv);
      break;
    }
    case QS_OBS_ARS_BASIC: PUT12(q); PUT12(qd); PUT(rpy[1]); PUT(st.pos[2]); PUT(st.vlin[2]); break;
    case QS_OBS_ARS_SENSOR: PUT12(q); PUT12(qd); PUT(rpy[1]); PUT(wl[1]); PUT(st.pos[2]); PUT(st.vlin[2]); break;
    case QS_OBS_LANDING_SENSOR:
      PUT12(q); PUT12(qd); PUT(rpy[1]); PUT(wl[1]); PUT(st.pos[2]); PUT(st.vlin[2]); PUT(landing); break;
    case QS_OBS_PPO_BASIC: PUT12(q); PUT12(qd); PUT(rpy[1]); PUT(st.pos[2]); PUT(st.vlin[2]); PUT(landing); break;
    case QS_OBS_PPO_BASIC_X:
      PUT12(q); PUT12(qd); PUT(rpy[1]); PUT(st.pos[2]); PUT(st.vlin[2]); PUT(st.vlin[0]); PUT(landing); break;
    case QS_OBS_PPO_BASIC_CONTACT:
      PUT12(q); PUT12(qd); PUT(rpy[1]); PUT(st.pos[2]); PUT(st.vlin[2]); PUT(landing);
#pragma unroll
      for (int k = 0; k < 4; k++) PUT((cs.mask >> k) & 1 ? 1.f : 0.f);
      break;
    case QS_OBS_ARS_BACKFLIP:
      PUT12(q); PUT12(qd); PUT(st.pos[2]); PUT(st.vlin[2]); PUT(backflip_pitch(Rb, landing != 0.f)); break;
    case QS_OBS_PPO_BACKFLIP:
      PUT12(q); PUT12(qd); PUT(st.pos[2]); PUT(st.vlin[2]); PUT(backflip_pitch(Rb, landing != 0.f)); PUT(landing); break;
    default:  // QS_OBS_PPO_CONTINUOUS_JUMPING_FORWARD: is_jumping only exists on the continuous tasks
      PUT12(q); PUT12(qd); PUT(st.pos[2]); PUT(st.vlin[2]); PUT(rpy[1]); PUT(landing); PUT(jumping); break;
  }
#undef PUT
#undef PUT12
}

// add N(0, sigma) per element (sensor.py:25-32,46-52) and store the row
QS_DEVONLY void store_obs(float* obs_row, float* o, const EnvCfg& C, const RobotConst& RC, uint64_t gid, uint32_t epoch,
                      uint32_t step, bool with_noise) {
  if (with_noise) {
#pragma unroll
    for (int b = 0; b < QS_MAX_OBS / 4; b++) {
      if (4 * b < C.obs_dim) {
        float nz[4];
        normal4(C.seed, gid, epoch, step, b, nz);
#pragma unroll
        for (int j = 0; j < 4; j++) o[4 * b + j] += RC.obs_noise[4 * b + j] * nz[j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < QS_MAX_OBS; i++)
    if (i < C.obs_dim) obs_row[i] = o[i];
}

// ---------------------------------------------------------------- SoA load/store
QS_DEVONLY void load_state(const DeviceView& D, int env, EnvState<float>& st, ContactState<float>& cs, float dt) {
  const int n = D.n;
  const float* s = D.state + env;
#pragma unroll
  for (int i = 0; i < 3; i++) st.pos[i] = s[i * n];
#pragma unroll
  for (int i = 0; i < 4; i++) st.quat[i] = s[(3 + i) * n];
#pragma unroll
  for (int i = 0; i < 3; i++) st.vlin[i] = s[(7 + i) * n];
#pragma unroll
  for (int i = 0; i < 3; i++) st.vang[i] = s[(10 + i) * n];
#pragma unroll
  for (int i = 0; i < 12; i++) st.q[i] = s[(13 + i) * n];
#pragma unroll
  for (int i = 0; i < 12; i++) st.qd[i] = s[(25 + i) * n];
  const int c = D.contact[env];
  cs.mask = c & 15;
  cs.invalid = c >> 8;
  cs.work_contacts = 0;
  cs.work_row_iters = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) cs.lam_n[k] = D.foot_force[k * n + env] * dt;
}
QS_DEVONLY void store_state(const DeviceView& D, int env, const EnvState<float>& st, const ContactState<float>& cs, float dt) {
  const int n = D.n;
  float* s = D.state + env;
#pragma unroll
  for (int i = 0; i < 3; i++) s[i * n] = st.pos[i];
#pragma unroll
  for (int i = 0; i < 4; i++) s[(3 + i) * n] = st.quat[i];
#pragma unroll
  for (int i = 0; i < 3; i++) s[(7 + i) * n] = st.vlin[i];
#pragma unroll
  for (int i = 0; i < 3; i++) s[(10 + i) * n] = st.vang[i];
#pragma unroll
  for (int i = 0; i < 12; i++) s[(13 + i) * n] = st.q[i];
#pragma unroll
  for (int i = 0; i < 12; i++) s[(25 + i) * n] = st.qd[i];
  D.contact[env] = (cs.mask & 15) | (cs.invalid << 8);
#pragma unroll
  for (int k = 0; k < 4; k++) D.foot_force[k * n + env] = cs.lam_n[k] / dt;
}

}  // namespace qs
