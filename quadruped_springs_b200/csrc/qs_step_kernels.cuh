// qs_step_kernels.cuh -- the step / reset kernel family (K1, K2).
//
//   k_step       one thread per env: action map, action_repeat fast ticks, epilogue.
//                An env whose tick needs the general solver (joint at its limit, body
//                shape on the ground) is parked untouched on the slow list.
//   k_step_slow  resumes parked envs with the general tick and runs the same epilogue
//                (a handful of envs per step: launched for N, exits on `count`).
//   k_reset      in-place reset + 2500-tick settle (exact QuadrupedGymEnv.reset), for a
//                mask / list / all envs.
//   k_refill     pre-settles the NEXT episode of envs whose spare slot is empty.
//
// Auto-reset without stalls: the state an env starts its next episode from is a pure
// function of (seed, global env id, episode number): mu ~ U[0.5,1) from Philox, then the
// reference's 2500 settle ticks.  Each env owns a ring of QS_SLOTS spare slots holding its
// next episodes' settled states; a finished env copies its slot inside k_step and queues an
// (env, episode) refill entry.  k_refill is launched after every step but returns at once
// unless a machine-filling batch of entries is pending, so the 2500-tick settles always run
// as dense, full-occupancy launches.  If a slot is not ready the env falls back to k_reset in
// place -- same numbers either way, so results never depend on scheduling or sharding.
#pragma once

// settling / reset-time command: _convert_reference_to_command(get_init_pose())
// (interface_base.py:68-72,182-200); also returns the settling action.
__device__ __forceinline__ void settle_command(const EnvCfg& C, const RobotConst& RC, float* cmd, float* act12) {
  if (C.is_rl) {
    const bool cart = C.control_mode == QS_CTRL_CARTESIAN_PD;
    const int sidx = cart ? 1 : 0;
    float a12[12];
#pragma unroll
    for (int i = 0; i < 12; i++)
      a12[i] = cart ? command_to_action1(RC.nominal_foot[i], RC.cart_lo[i], RC.cart_hi[i])
                    : command_to_action1(RC.init_angles[i], RC.ang_lo[i], RC.ang_hi[i]);
    // _convert_to_actual_action_space (action_interface.py:17-18,41-44,67-74)
#pragma unroll
    for (int i = 0; i < 12; i++) act12[i] = 0.f;
    if (C.action_mode == QS_ACT_DEFAULT) {
#pragma unroll
      for (int i = 0; i < 12; i++) act12[i] = a12[i];
    } else if (C.action_mode == QS_ACT_SYMMETRIC) {
#pragma unroll
      for (int j = 0; j < 3; j++) { act12[j] = a12[j]; act12[3 + j] = a12[6 + j]; }
    } else {
      if (sidx == 0) { act12[0] = a12[1]; act12[1] = a12[2]; act12[2] = a12[7]; act12[3] = a12[8]; }
      else { act12[0] = a12[0]; act12[1] = a12[2]; act12[2] = a12[6]; act12[3] = a12[8]; }
    }
    float b12[12];
    expand_action(C.action_mode, sidx, act12, b12);
    action12_to_command(RC, C.control_mode, b12, cmd);
    // settling_action = _transform_motor_command_to_action(settling_command)
    // (interface_base.py:196-200).  In CARTESIAN_PD mode the command holds JOINT
    // ANGLES (it went through IK) yet is scaled with the CARTESIAN limits, so the
    // reference stores (0, 1, -1) per leg as _last_action: reproduced as is.
#pragma unroll
    for (int i = 0; i < 12; i++)
      a12[i] = cart ? command_to_action1(cmd[i], RC.cart_lo[i], RC.cart_hi[i])
                    : command_to_action1(cmd[i], RC.ang_lo[i], RC.ang_hi[i]);
#pragma unroll
    for (int i = 0; i < 12; i++) act12[i] = 0.f;
    if (C.action_mode == QS_ACT_DEFAULT) {
#pragma unroll
      for (int i = 0; i < 12; i++) act12[i] = a12[i];
    } else if (C.action_mode == QS_ACT_SYMMETRIC) {
#pragma unroll
      for (int j = 0; j < 3; j++) { act12[j] = a12[j]; act12[3 + j] = a12[6 + j]; }
    } else {
      if (sidx == 0) { act12[0] = a12[1]; act12[1] = a12[2]; act12[2] = a12[7]; act12[3] = a12[8]; }
      else { act12[0] = a12[0]; act12[1] = a12[2]; act12[2] = a12[6]; act12[3] = a12[8]; }
    }
  } else {
    // settle_robot_by_pd (control_interface/utils.py:22-30): PD limits, DEFAULT space
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const float a = command_to_action1(RC.init_angles[i], RC.ang_lo[i], RC.ang_hi[i]);
      cmd[i] = clampt(RC.ang_lo[i] + 0.5f * (a + 1.f) * (RC.ang_hi[i] - RC.ang_lo[i]), RC.ang_lo[i], RC.ang_hi[i]);
      act12[i] = 0.f;
    }
  }
}

__device__ __forceinline__ void finish_episode_stats(const DeviceView& D, int env, const float* ts, float ep_return,
                                                     int ep_len, bool terminated, int task) {
  const int n = D.n;
  float* s = D.stats + env;
  s[0 * n] += 1.f;
  s[1 * n] += ts[TS_MAX_H];
  s[2 * n] = fmaxf(s[2 * n], ts[TS_MAX_H]);
  s[3 * n] += ts[TS_REL_MAX_H];
  s[4 * n] += ts[TS_MAX_FWD];
  s[5 * n] = fmaxf(s[5 * n], ts[TS_MAX_FWD]);
  s[6 * n] += ts[TS_MAX_FLIGHT];
  const float flip = (task == QS_TASK_BACKFLIP ? ts[TS_MAX_PITCH_BF] : ts[TS_MAX_PITCH]) / float(2 * QS_PI);
  s[7 * n] += flip;
  s[8 * n] += ep_return;
  s[9 * n] += float(ep_len);
  s[10 * n] += terminated ? 1.f : 0.f;
}

struct StepIO {
  const float* actions;
  float* obs;
  float* reward;
  uint8_t* done;
  uint8_t* truncated;
  int* slow_list;    // envs parked for the general solver, slow_list[n] = count
  int* reset_list;   // envs that must be reset in place,   reset_list[n] = count
  int* refill_list;  // (env, episode) pairs to pre-settle, refill_list[2 * refill_cap] = count
  int refill_cap;
};

constexpr int QS_SLOTS = 2;
constexpr uint32_t QS_URGENT = 0x80000000u;  // refill entry flag: settle AND start the episode now (no slot was ready)

__device__ __forceinline__ void refill_push(int* list, int cap, int env, uint32_t epoch) {
  const int i = atomicAdd(list + 2 * cap, 1);
  // an overflowing entry is only a missed prefetch: the env falls back to k_reset
  if (i < cap) { list[2 * i] = env; list[2 * i + 1] = int(epoch); }
}

// ---- spare slot of an env: the settled state its next episode starts from
// rows: state 37 | tau_motor 12 | tau_spring 12 | foot_force 4 | mu 1   (floats), contact, epoch (ints)
constexpr int SLOT_ROWS = 37 + 12 + 12 + 4 + 1;

__device__ __forceinline__ void slot_store(const DeviceView& D, int env, const EnvState<float>& st,
                                           const ContactState<float>& cs, const float* tau_m, const float* tau_s,
                                           float mu, uint32_t epoch, float dt) {
  const int n = D.n;
  const int sl = int(epoch % QS_SLOTS);
  float* s = D.slot + size_t(sl) * SLOT_ROWS * n + env;
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
#pragma unroll
  for (int i = 0; i < 12; i++) { s[(37 + i) * n] = tau_m[i]; s[(49 + i) * n] = tau_s[i]; }
#pragma unroll
  for (int k = 0; k < 4; k++) s[(61 + k) * n] = cs.lam_n[k] / dt;
  s[65 * n] = mu;
  D.slot_contact[sl * n + env] = (cs.mask & 15) | (cs.invalid << 8);
  D.slot_epoch[sl * n + env] = epoch;
}

__device__ __forceinline__ void slot_load(const DeviceView& D, int env, uint32_t epoch, EnvState<float>& st,
                                          ContactState<float>& cs, float* tau_m, float* tau_s, float* mu, float dt) {
  const int n = D.n;
  const int sl = int(epoch % QS_SLOTS);
  const float* s = D.slot + size_t(sl) * SLOT_ROWS * n + env;
#pragma unroll
  for (int i = 0; i < 3; i++) st.pos[i] = __ldcg(s + i * n);
#pragma unroll
  for (int i = 0; i < 4; i++) st.quat[i] = __ldcg(s + (3 + i) * n);
#pragma unroll
  for (int i = 0; i < 3; i++) st.vlin[i] = __ldcg(s + (7 + i) * n);
#pragma unroll
  for (int i = 0; i < 3; i++) st.vang[i] = __ldcg(s + (10 + i) * n);
#pragma unroll
  for (int i = 0; i < 12; i++) st.q[i] = __ldcg(s + (13 + i) * n);
#pragma unroll
  for (int i = 0; i < 12; i++) st.qd[i] = __ldcg(s + (25 + i) * n);
#pragma unroll
  for (int i = 0; i < 12; i++) { tau_m[i] = __ldcg(s + (37 + i) * n); tau_s[i] = __ldcg(s + (49 + i) * n); }
#pragma unroll
  for (int k = 0; k < 4; k++) cs.lam_n[k] = __ldcg(s + (61 + k) * n) * dt;
  *mu = __ldcg(s + 65 * n);
  const int c = __ldcg(D.slot_contact + sl * n + env);
  cs.mask = c & 15;
  cs.invalid = c >> 8;
  cs.work_contacts = 0;
  cs.work_row_iters = 0;
}

// A fresh robot settled for episode `epoch` of global env `gid`
// (quadruped_gym_env.py:278-289,323-327; env_randomizer.py:287-289).
__device__ __forceinline__ void settle_fresh(const KernelArgs& A, int env, uint64_t gid, uint32_t epoch,
                                             EnvState<float>& st, ContactState<float>& cs, float* tau_m, float* tau_s,
                                             float* mu_out, const Scratch<float>& scr, bool need) {
  // `need` = this thread really settles; the others only keep the block's barriers matched.
  // Must be called by every thread of the block.
  if (!__syncthreads_or(need)) return;
  const EnvCfg& C = A.C;
  const float mu = C.ground_randomizer ? 0.5f + 0.5f * uniform1(C.seed, gid, epoch, 100) : C.mu_ground;
  if (need) *mu_out = mu;
  EnvState<float> st_keep = st;  // a thread that does not need the settle gets its state back untouched
  ContactState<float> cs_keep = cs;
  st.pos[0] = 0.f; st.pos[1] = 0.f; st.pos[2] = 0.32f;  // INIT_POSITION, configs:23
  st.quat[0] = st.quat[1] = st.quat[2] = 0.f; st.quat[3] = 1.f;
#pragma unroll
  for (int i = 0; i < 3; i++) { st.vlin[i] = 0.f; st.vang[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < 12; i++) { st.q[i] = A.RC.init_angles[i]; st.qd[i] = 0.f; tau_m[i] = 0.f; tau_s[i] = 0.f; }
  cs.mask = 0; cs.invalid = 0; cs.work_contacts = 0; cs.work_row_iters = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) cs.lam_n[k] = 0.f;
  float cmd[12], act12[12];
  settle_command(C, A.RC, cmd, act12);
  // a fresh Quadruped has the default gains / springs (quadruped_gym_env.py:299-319); the settle
  // runs on the fast tick only: the standing pose is far from every joint limit and body contact
  SolverConst SCs = A.SC;
  SCs.enable_limits = 0;
  SCs.body_response = 0;
  const int nsettle = C.is_rl ? C.settling_steps : 1500;
  float sk[3], sb[3], sr[3];
#pragma unroll
  for (int j = 0; j < 3; j++) { sk[j] = A.RC.spring_k[j]; sb[j] = A.RC.spring_b[j]; sr[j] = A.RC.spring_rest[j]; }
  for (int t = 0; t < nsettle; t++) {
    __syncthreads();  // lockstep across the block (see run_ticks)
    float tau[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
      tau_m[i] = pd_torque1(A.RC.kp[i], A.RC.kd[i], A.RC.tau_max[i], cmd[i], st.q[i], st.qd[i], false);
      tau[i] = tau_m[i];
    }
    if (C.enable_springs) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        spring_torque_leg(k, sk, sb, sr, st.q + 3 * k, st.qd + 3 * k, tau_s + 3 * k);
#pragma unroll
        for (int j = 0; j < 3; j++) tau[3 * k + j] += tau_s[3 * k + j];
      }
    }
    if (need) physics_tick(st, tau, mu, cs, A.M, SCs, t == nsettle - 1, scr);
  }
  if (!need) { st = st_keep; cs = cs_keep; }
}

// Everything QuadrupedGymEnv.reset does after the settle (quadruped_gym_env.py:282-297),
// from a settled state: counters, task._reset, sensors, filter history; writes the env back.
__device__ __forceinline__ void begin_episode(const KernelArgs& A, int env, uint32_t epoch, float mu,
                                              const EnvState<float>& st, const ContactState<float>& cs,
                                              const float* tau_m, const float* tau_s, float* obs) {
  const DeviceView& D = A.D;
  const EnvCfg& C = A.C;
  const int n = D.n;
  const float dt = A.SC.dt;
  D.reset_count[env] = epoch;
  D.mu[env] = mu;
  D.custom_gains[env] = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) { D.kp[i * n + env] = A.RC.kp[i]; D.kd[i * n + env] = A.RC.kd[i]; }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    D.spring[(0 + j) * n + env] = A.RC.spring_k[j];
    D.spring[(3 + j) * n + env] = A.RC.spring_b[j];
    D.spring[(6 + j) * n + env] = A.RC.spring_rest[j];
  }
  float cmd[12], act12[12];
  settle_command(C, A.RC, cmd, act12);
  float Rb[9], rpy[3];
  quat_to_R(st.quat, Rb);
  rpy_from_quat(st.quat, rpy);
  float ts[QS_TASK_DIM];
#pragma unroll
  for (int i = 0; i < TS_END; i++) ts[i] = i < task_slots(C.task) ? D.task[i * n + env] : 0.f;
  task_reset(ts, st, cs, tau_m, rpy, Rb, 0.f, C.task);
#pragma unroll
  for (int i = 0; i < TS_END; i++) if (i < task_slots(C.task)) D.task[i * n + env] = ts[i];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    D.last_action[i * n + env] = act12[i];
    D.tau_motor[i * n + env] = tau_m[i];
    D.tau_spring[i * n + env] = tau_s[i];
    // action_filter.py:123-127 init_history(last_action)
    D.filt[(0 * 12 + i) * n + env] = act12[i]; D.filt[(1 * 12 + i) * n + env] = act12[i];
    D.filt[(2 * 12 + i) * n + env] = act12[i]; D.filt[(3 * 12 + i) * n + env] = act12[i];
  }
  D.sim_steps[env] = 0;
  D.env_steps[env] = 0;
  D.ep_return[env] = 0.f;
  store_state(D, env, st, cs, dt);
  if (obs) {
    float o[QS_MAX_OBS];
#pragma unroll
    for (int i = 0; i < QS_MAX_OBS; i++) o[i] = 0.f;
    observe(st, cs, ts, rpy, Rb, C.obs_mode, C.task, o);
    store_obs(obs + size_t(env) * C.obs_dim, o, C, A.RC, uint64_t(C.gid0 + env), epoch, 0u, C.enable_noise);
  }
}

// Epilogue of a control step (quadruped_gym_env.py:239-256) + write-back + auto-reset.
__device__ __forceinline__ void finish_step(const KernelArgs& A, const StepIO& io, int env, EnvState<float>& st,
                                            ContactState<float>& cs, const float* tau_m, const float* tau_s) {
  const DeviceView& D = A.D;
  const EnvCfg& C = A.C;
  const int n = D.n;
  const float dt = A.SC.dt;
  const int sim_steps = D.sim_steps[env] + C.action_repeat;
  const int env_steps = D.env_steps[env] + 1;
  float Rb[9], rpy[3];
  quat_to_R(st.quat, Rb);
  rpy_from_quat(st.quat, rpy);
  const float sim_time = float(double(sim_steps) * A.time_step_d);
  float ts[QS_TASK_DIM];
#pragma unroll
  for (int i = 0; i < TS_END; i++) ts[i] = i < task_slots(C.task) ? D.task[i * n + env] : 0.f;
  float foot_force[4];
#pragma unroll
  for (int k = 0; k < 4; k++) foot_force[k] = (cs.mask >> k) & 1 ? cs.lam_n[k] / dt : 0.f;
  task_on_step(ts, st, cs, tau_m, rpy, Rb, sim_time, C.task);
  float r = task_reward(ts, st, foot_force, ts + TS_OLD_TAU0, tau_m, rpy, Rb, C.task);
  const bool term = task_terminated(ts, st, cs, Rb, A.RC.fallen_height, C.task);
  const bool dn = term || (double(sim_steps) * A.time_step_d > A.max_time_d);
  if (dn) r += task_reward_end(ts, term, C.task, sim_time, C.max_episode_time);
#pragma unroll
  for (int i = 0; i < 12; i++) ts[TS_OLD_TAU0 + i] = tau_m[i];
  const float ep_ret = D.ep_return[env] + r;
  io.reward[env] = r;
  io.done[env] = dn;
  io.truncated[env] = dn && !term;
  D.work[0 * n + env] += uint32_t(C.action_repeat);
  D.work[1 * n + env] += uint32_t(cs.work_contacts);
  D.work[2 * n + env] += uint32_t(cs.work_row_iters);
  const uint64_t gid = uint64_t(C.gid0 + env);
  if (dn) finish_episode_stats(D, env, ts, ep_ret, env_steps, term, C.task);
  if (dn && C.auto_reset) {
    // the finished env starts its next episode inside the same call; its obs row becomes the
    // first observation of that episode (SB3 VecEnv convention)
#pragma unroll
    for (int i = 0; i < TS_END; i++) if (i < task_slots(C.task)) D.task[i * n + env] = ts[i];  // BackFlip.max_pitch survives resets
    const uint32_t epoch = D.reset_count[env] + 1;
    uint32_t* tag = D.slot_epoch + int(epoch % QS_SLOTS) * n + env;
    if (*tag == epoch) {
      float tm[12], tsp[12], mu;
      slot_load(D, env, epoch, st, cs, tm, tsp, &mu, dt);
      *tag = 0;
      refill_push(io.refill_list, io.refill_cap, env, epoch + QS_SLOTS);  // the freed slot's next tenant
      begin_episode(A, env, epoch, mu, st, cs, tm, tsp, io.obs);
    } else {
      // no spare slot ready: an urgent refill entry makes k_refill (launched right after this kernel)
      // settle this episode now and start it; same numbers as the prefetched path
      store_state(D, env, st, cs, dt);
      refill_push(io.refill_list, io.refill_cap, env, epoch | QS_URGENT);
#pragma unroll
      for (int d = 1; d <= QS_SLOTS; d++)  // and its (empty or stale) ring is rebuilt in the same launch
        if (D.slot_epoch[int((epoch + d) % QS_SLOTS) * n + env] != epoch + d)
          refill_push(io.refill_list, io.refill_cap, env, epoch + d);
      atomicAdd(io.reset_list + n, 1);  // urgent counter (forces the refill to run)
    }
    return;
  }
  // ---- sensors (:253-254) and write-back
  float o[QS_MAX_OBS];
#pragma unroll
  for (int i = 0; i < QS_MAX_OBS; i++) o[i] = 0.f;
  observe(st, cs, ts, rpy, Rb, C.obs_mode, C.task, o);
  store_obs(io.obs + size_t(env) * C.obs_dim, o, C, A.RC, gid, D.reset_count[env], uint32_t(env_steps), C.enable_noise);
  store_state(D, env, st, cs, dt);
#pragma unroll
  for (int i = 0; i < 12; i++) { D.tau_motor[i * n + env] = tau_m[i]; D.tau_spring[i * n + env] = tau_s[i]; }
#pragma unroll
  for (int i = 0; i < TS_END; i++) if (i < task_slots(C.task)) D.task[i * n + env] = ts[i];
  D.sim_steps[env] = sim_steps;
  D.env_steps[env] = env_steps;
  D.ep_return[env] = ep_ret;
}

// -------------------------------------------------------------------- K1: step
__global__ void __launch_bounds__(256, 1)
k_step(const __grid_constant__ KernelArgs A, const StepIO io) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const DeviceView& D = A.D;
  const EnvCfg& C = A.C;
  const int n = D.n;
  // threads past the end shadow the last env (block-wide barriers inside run_ticks need every
  // thread); they compute the same thing and write nothing
  const bool live = tid < n;
  const int env = live ? tid : n - 1;
  const float dt = A.SC.dt;
  EnvState<float> st;
  ContactState<float> cs;
  load_state(D, env, st, cs, dt);

  // ---- action (quadruped_gym_env.py:229-234)
  float act[12];
#pragma unroll
  for (int i = 0; i < 12; i++) act[i] = i < C.action_dim ? io.actions[size_t(env) * C.action_dim + i] : 0.f;
  if (live) {
#pragma unroll
    for (int i = 0; i < 12; i++) D.last_action[i * n + env] = act[i];
  }
  if (C.enable_filter) {  // utils/action_filter.py:110-121
#pragma unroll
    for (int i = 0; i < 12; i++) {
      if (i < C.action_dim) {
        float* f = D.filt + env;
        const float x0 = f[(0 * 12 + i) * n], x1 = f[(1 * 12 + i) * n];
        const float y0 = f[(2 * 12 + i) * n], y1 = f[(3 * 12 + i) * n];
        const float y = act[i] * A.RC.filt_b[0] + (x0 * A.RC.filt_b[1] + x1 * A.RC.filt_b[2]) -
                        (y0 * A.RC.filt_a[1] + y1 * A.RC.filt_a[2]);
        if (live) {
          f[(1 * 12 + i) * n] = x0; f[(0 * 12 + i) * n] = act[i];
          f[(3 * 12 + i) * n] = y0; f[(2 * 12 + i) * n] = y;
        }
        act[i] = y;
      }
    }
  }
  float cmd[12];
  bool torque_mode = false;
  if (C.is_rl) {
    // _interpolate_actions (:187-205) is a no-op in the reference: step() overwrites
    // _last_action with the current action before the substeps (:229-234).
    float a12[12];
    expand_action(C.action_mode, C.control_mode == QS_CTRL_CARTESIAN_PD ? 1 : 0, act, a12);
    action12_to_command(A.RC, C.control_mode, a12, cmd);
  } else {
#pragma unroll
    for (int i = 0; i < 12; i++) cmd[i] = act[i];
    torque_mode = C.control_mode == QS_CTRL_TORQUE;
  }

  // ---- action_repeat substeps (:236-237)
  float tau_m[12], tau_s[12];
  extern __shared__ float qs_smem[];
  const Scratch<float> scr{qs_smem + threadIdx.x, int(blockDim.x)};
  const int t_done = run_ticks(st, cs, cmd, torque_mode, 0, C.action_repeat, env, D, C, A.RC, A.M, A.SC, tau_m, tau_s, true, scr);
  if (!live) return;
  if (t_done < C.action_repeat) {
    // park the env (state as of the start of tick t_done) for the general solver
    store_state(D, env, st, cs, dt);
    D.work[1 * n + env] += uint32_t(cs.work_contacts);
    D.work[2 * n + env] += uint32_t(cs.work_row_iters);
#pragma unroll
    for (int i = 0; i < 12; i++) D.cmd[i * n + env] = cmd[i];
    D.resume_tick[env] = t_done;
    io.slow_list[atomicAdd(io.slow_list + n, 1)] = env;
    return;
  }
  finish_step(A, io, env, st, cs, tau_m, tau_s);
}

// -------------------------------------------------------------------- K1b: general-solver continuation
__global__ void __launch_bounds__(64)
k_step_slow(const __grid_constant__ KernelArgs A, const StepIO io) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const DeviceView& D = A.D;
  const int n = D.n;
  if (tid >= io.slow_list[n]) return;
  const int env = io.slow_list[tid];
  EnvState<float> st;
  ContactState<float> cs;
  load_state(D, env, st, cs, A.SC.dt);
  float cmd[12], tau_m[12], tau_s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) cmd[i] = D.cmd[i * n + env];
  const bool torque_mode = !A.C.is_rl && A.C.control_mode == QS_CTRL_TORQUE;
  run_ticks_general(st, cs, cmd, torque_mode, D.resume_tick[env], A.C.action_repeat, env, D, A.C, A.RC, A.M, A.SC,
                    tau_m, tau_s);
  finish_step(A, io, env, st, cs, tau_m, tau_s);
}

// -------------------------------------------------------------------- K2: reset + settle
// list == nullptr: thread i resets env i (all envs).  Otherwise thread i resets env list[i]
// for i < list[n] (dense warps whatever the done pattern).
__global__ void __launch_bounds__(256, 1)
k_reset(const __grid_constant__ KernelArgs A, const int* __restrict__ list, int* __restrict__ refill_list,
        int refill_cap, float* __restrict__ obs) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const DeviceView& D = A.D;
  const int n = D.n;
  const int count = list ? list[n] : n;
  if (count == 0) return;                       // uniform over the grid
  if (blockIdx.x * blockDim.x >= count) return;  // uniform over the block
  const bool live = tid < count;                 // stragglers shadow the last entry and write nothing
  const int idx = live ? tid : count - 1;
  const int env = list ? list[idx] : idx;
  const uint64_t gid = uint64_t(A.C.gid0 + env);
  const uint32_t epoch = D.reset_count[env] + 1;
  EnvState<float> st;
  ContactState<float> cs;
  float tau_m[12], tau_s[12], mu;
  uint32_t* tag = D.slot_epoch + int(epoch % QS_SLOTS) * n + env;
  const bool have = *tag == epoch;
  extern __shared__ float qs_smem[];
  const Scratch<float> scr{qs_smem + threadIdx.x, int(blockDim.x)};
  float tm2[12], ts2[12];
  settle_fresh(A, env, gid, epoch, st, cs, tm2, ts2, &mu, scr, !have);
  if (have) {
    slot_load(D, env, epoch, st, cs, tau_m, tau_s, &mu, A.SC.dt);
  } else {
#pragma unroll
    for (int i = 0; i < 12; i++) { tau_m[i] = tm2[i]; tau_s[i] = ts2[i]; }
  }
  if (!live) return;
  if (A.C.auto_reset) {
    if (have) *tag = 0;
    // every future episode within the ring must be present or queued (duplicates are skipped by k_refill)
#pragma unroll
    for (int d = 1; d <= QS_SLOTS; d++) {
      const uint32_t e = epoch + d;
      if (D.slot_epoch[int(e % QS_SLOTS) * n + env] != e) refill_push(refill_list, refill_cap, env, e);
    }
  }
  begin_episode(A, env, epoch, mu, st, cs, tau_m, tau_s, obs);
}

// pre-settle queued (env, episode) pairs into the envs' spare slots.  Launched after every
// step: returns immediately unless at least `threshold` entries are pending or an entry is
// urgent (an env finished without a ready slot: its episode is settled and started here).
__global__ void __launch_bounds__(256, 1)
k_refill(const __grid_constant__ KernelArgs A, const int* __restrict__ list, int cap, int threshold,
         const int* __restrict__ urgent_count, float* __restrict__ obs) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const DeviceView& D = A.D;
  const int n = D.n;
  const int count = min(list[2 * cap], cap);
  if ((count < threshold && *urgent_count == 0) || count == 0) return;  // uniform over the grid
  if (blockIdx.x * blockDim.x >= count) return;                          // uniform over the block
  const bool live = tid < count;
  const int i = live ? tid : count - 1;
  const int env = list[2 * i];
  const uint32_t raw = uint32_t(list[2 * i + 1]);
  const bool urgent = (raw & QS_URGENT) != 0;
  const uint32_t epoch = raw & ~QS_URGENT;
  // skip duplicates and entries that became stale (the env moved past that episode)
  const bool need = urgent ? (D.reset_count[env] + 1 == epoch)
                           : (D.slot_epoch[int(epoch % QS_SLOTS) * n + env] != epoch && epoch > D.reset_count[env]);
  EnvState<float> st;
  ContactState<float> cs;
  float tau_m[12], tau_s[12], mu = 0.f;
  extern __shared__ float qs_smem[];
  const Scratch<float> scr{qs_smem + threadIdx.x, int(blockDim.x)};
  settle_fresh(A, env, uint64_t(A.C.gid0 + env), epoch, st, cs, tau_m, tau_s, &mu, scr, need);
  if (!live || !need) return;
  if (urgent) {
    begin_episode(A, env, epoch, mu, st, cs, tau_m, tau_s, obs);
  } else {
    slot_store(D, env, st, cs, tau_m, tau_s, mu, epoch, A.SC.dt);
  }
}
// clears the queue after a refill that ran (same trigger condition)
__global__ void k_refill_done(int* __restrict__ list, int cap, int threshold, int* __restrict__ urgent_count) {
  const int count = min(list[2 * cap], cap);
  if ((count >= threshold || *urgent_count > 0) && count > 0) {
    list[2 * cap] = 0;
    *urgent_count = 0;
  }
}

__global__ void k_compact(const uint8_t* __restrict__ mask, int n, int* __restrict__ list) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && mask[i]) list[atomicAdd(list + n, 1)] = i;
}
