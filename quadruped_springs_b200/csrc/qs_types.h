// qs_types.h -- POD types shared by host and device code of the product.
#pragma once
#include <stdint.h>

#include "../../include/qs_b200.h"

namespace qs {

// Robot-level constants of go1/configs_go1_{with,without}_springs.py (float).
struct RobotConst {
  float init_angles[12];           // configs:31-36
  float ang_lo[12], ang_hi[12];    // RL_{LOWER,UPPER}_ANGLE_JOINT, configs_with:84-87 / without:80-83
  float cart_lo[12], cart_hi[12];  // RL_{LOWER,UPPER}_CARTESIAN_POS, configs_with:90-96 / without:86-92
  float nominal_foot[12];          // NOMINAL_FOOT_POS_LEG_FRAME, configs:69-71
  float tau_max[12];               // RL_TORQUE_LIMITS, configs:100-101
  float kp[12], kd[12];            // MOTOR_KP/KD, configs_with:106-107 / without:108-109
  float spring_k[3], spring_b[3], spring_rest[3];  // configs_with:150-160
  float fallen_height;             // IS_FALLEN_HEIGHT, configs:24
  float obs_noise[QS_MAX_OBS];     // per-element sensor noise std of the selected obs mode
  float filt_b[3], filt_a[3];      // Butterworth(2, 3 Hz) coefficients, action_filter.py:191-213
  float landing_action[12];        // env.get_landing_action() in the configured action space, quadruped_gym_env.py:375-379
  float env_dt;                    // action_repeat * time_step
  float settle_cmd[12];            // the settling command (settle_command(), qs_step_kernels.cuh), evaluated once on the device at qs_create:
                                   // an operand from the constant bank in the settle kernels instead of twelve registers
};

// Merged 13-body dynamics model of go1.urdf: fixed links folded into their
// parents (base+trunk+imu -> body 0, calf+foot -> calf).  Inertia tensors follow
// Bullet's collision-geometry rule.  Sym3 order: xx xy xz yy yz zz.
template <typename T> struct ModelConstT {
  T trunk_m, trunk_h[3], trunk_I[6];  // about the base origin, base axes
  T hip_pos[4][3];                    // go1.urdf:112-118 (+ mirrored copies)
  T thigh_off_y[4];                   // go1.urdf:164-170: (0, -+0.08, 0)
  T link_len;                         // go1.urdf:191-197,218-222: 0.213
  T body_m[4][3];                     // hip, thigh, calf(+foot)
  T body_com[4][3][3];                // in the link frame
  T body_Ic[4][3][6];                 // about the body's own COM, link axes
  T foot_radius, foot_thresh;         // go1.urdf:230-236, relative breaking threshold
  T trunk_half[3], trunk_thresh;      // collision shapes used for invalid-contact detection
  T imu_pos[3], imu_half, imu_thresh;
  T hip_r, hip_hl, hip_thresh;
  T thigh_half[3], thigh_c[3], thigh_thresh;
  T calf_half[3], calf_c[3], calf_thresh;
  T joint_lo[3], joint_hi[3];         // go1.urdf limits per hip/thigh/calf
};

struct SolverConst {
  float dt, gravity_z, contact_erp, limit_erp, linear_slop, warmstart, residual_threshold;
  float max_coord_vel, mu_link;
  int32_t num_iterations, enable_limits, body_response, self_collision;
};

// Device view of the handle-owned SoA state (all arrays [dim][N]).
struct DeviceView {
  int32_t n;
  float* state;
  float* tau_motor;
  float* tau_spring;
  float* kp;
  float* kd;
  float* spring;
  float* mu;
  float* foot_force;  // [4][N]
  int32_t* contact;   // [N]
  float* task;        // [QS_TASK_DIM][N]
  float* last_action; // [12][N]
  float* filt;        // [4*12][N] xh0 xh1 yh0 yh1
  int32_t* sim_steps;
  int32_t* env_steps;
  float* ep_return;
  float* stats;       // [QS_STATS_DIM][N] finished-episode accumulators
  uint32_t* reset_count; // [N] number of resets (RNG stream separation)
  uint32_t* work;        // [3][N] k_step work counters: ticks, contact-ticks, contact-sweeps
  float* cmd;            // [12][N] motor command of the current control step (slow-path hand-over)
  int32_t* resume_tick;  // [N] tick at which the fast kernel handed the env to the general solver
  int32_t* land_mode;    // [N] landing controller: 0 policy, 1 take-off hold, 2 landing, 3 spent
  float* land_timer;     // [2][N] timer time, timer end (utils/timer.py)
  int32_t* rest_active;  // [N] go-to-rest controller engaged (go_to_rest_wrapper.py:58-81)
  float* rest;           // [14][N] h_actual, sim step of activation, start action[12]
  float* model;          // [EM_ROWS][N] mass randomizer: per-env mass properties of the current episode (qs_physics.cuh EnvModelRef)
  float* mass_draw;      // [8][N] the draws they come from: hip, thigh, calf, trunk mass, block mass, block position
  uint8_t* custom_gains; // [N] non-zero: read kp/kd of this env from the arrays instead of the config constants
  float* slot;           // [slots][66][N] settled states of the next episodes (see qs_step_kernels.cuh)
  int32_t* slot_contact; // [slots][N]
  uint32_t* slot_epoch;  // [slots][N] episode number the slot was settled for (0 = empty)
};

// task state slots (rows of DeviceView::task)
enum TaskSlot {
  TS_SWITCHED = 0, TS_IN_AIR, TS_T_TAKEOFF, TS_TAKEOFF_X, TS_TAKEOFF_Y, TS_TAKEOFF_Z, TS_INIT_HEIGHT,
  TS_TAKEOFF_YAW, TS_MAX_FLIGHT, TS_MAX_FWD, TS_MAX_PITCH, TS_REL_MAX_H, TS_MAX_DX, TS_MAX_H,
  TS_MAX_PITCH_BF, TS_OLD_FWD, TS_ACTUAL_FWD, TS_OLD_TAU0 /* ..+11 */, TS_END_BASIC = TS_OLD_TAU0 + 12,
  // continuous-jumping tasks only (task_base.py:222-400).  The reference keeps per-jump arrays; the
  // end-of-episode reward needs only their count, sum, max and the entropy sums S = sum f, Q = sum f log2 f.
  TS_IS_JUMPING = TS_END_BASIC, TS_CUM_FWD, TS_CUM_FLIGHT, TS_FIRST_JUMP, TS_JUMP_COUNT, TS_GOOD_JUMPS, TS_MAX_JUMP_H,
  TS_SUM_FWD, TS_SUM_FLOG, TS_SUM_PERF, TS_MAX_PERF, TS_LAST_PERF, TS_END_JUMP, TS_END,
  // imitation tasks only (task_base.py:169-220): position in the demonstration, rows left at reset (they share the
  // continuous-jumping rows: a task is one or the other)
  TS_DEMO_COUNTER = TS_END_BASIC, TS_DELTA_DEMO, TS_END_DEMO,
  // TaskJumpingDemo2 (task_base.py:402-452) is an imitation task ON TOP of the continuous-jumping bookkeeping: its two rows
  // come after those
  TS_DEMO2_COUNTER = TS_END, TS_DEMO2_DELTA, TS_END_ALL
};
static_assert(TS_DEMO_COUNTER == QS_TS_DEMO_COUNTER, "header constant out of date");
static_assert(TS_DEMO2_COUNTER == QS_TS_DEMO2_COUNTER, "header constant out of date");
static_assert(TS_END_ALL <= QS_TASK_DIM, "task state too large");

}  // namespace qs
