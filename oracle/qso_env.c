/*
 * qso_env.c -- CPU ORACLE (test infrastructure, not the product).
 *
 * Restatement of the reference's own numpy logic around pybullet for ONE env:
 * action mapping, PD / PEA torques, leg FK / Jacobian / IK, jumping tasks,
 * sensors, reset + settle, and QuadrupedGymEnv.step.  Every function cites the
 * reference file:line it follows (paths relative to
 * /root/reference/quadruped_spring/).  Pinned by the tests/golden npz fixtures, which
 * oracle/gen_golden.py produces by importing the reference's modules.
 */
#include "qso.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846
#define QSO_MAX_JUMPS 512 /* a jump needs >= 2 control steps and an episode has <= 1000 */

/* ------------------------------------------------------------ constants */
/* go1/configs_go1_with_springs.py:56-58 */
static const double L1 = 0.0847, L2 = 0.213, L3 = 0.213;
static const double SIDE[4] = {-1, 1, -1, 1}; /* quadruped.py:360-362 */

typedef struct {
  double init_angles[12];    /* configs:31-36 */
  double ang_lo[12], ang_hi[12]; /* configs_with:84-87 / configs_without:80-83 */
  double cart_lo[12], cart_hi[12]; /* configs_with:69-96 / configs_without:86-92 */
  double nominal_foot[12];
  double tau_max[12];        /* configs:100-101 */
  double kp[12], kd[12];     /* configs_with:106-107 / configs_without:108-109 */
  double spring_k[3], spring_b[3], spring_rest[3]; /* configs_with:150-160 */
  double fallen_height;      /* configs:24 */
} RobotCfg;

static void robot_cfg(int springs, int task, RobotCfg* c) {
  for (int k = 0; k < 4; k++) {
    c->init_angles[3 * k] = 0; c->init_angles[3 * k + 1] = PI / 4; c->init_angles[3 * k + 2] = -PI / 2;
    c->ang_hi[3 * k] = 0.2; c->ang_hi[3 * k + 1] = PI / 4 + 0.5; c->ang_hi[3 * k + 2] = -0.95;
    c->ang_lo[3 * k] = -0.2; c->ang_lo[3 * k + 1] = PI / 4 - 0.5; c->ang_lo[3 * k + 2] = springs ? -2.5 : -2.12;
    c->nominal_foot[3 * k] = 0; c->nominal_foot[3 * k + 1] = SIDE[k] * L1; c->nominal_foot[3 * k + 2] = -0.32;
    c->cart_hi[3 * k] = c->nominal_foot[3 * k] + 0.2;
    c->cart_hi[3 * k + 1] = c->nominal_foot[3 * k + 1] + 0.05;
    c->cart_hi[3 * k + 2] = c->nominal_foot[3 * k + 2] + (springs ? 0.18 : 0.11);
    c->cart_lo[3 * k] = c->nominal_foot[3 * k] - 0.2;
    c->cart_lo[3 * k + 1] = c->nominal_foot[3 * k + 1] - 0.05;
    c->cart_lo[3 * k + 2] = c->nominal_foot[3 * k + 2] - 0.07;
    c->tau_max[3 * k] = 23.7; c->tau_max[3 * k + 1] = 23.7; c->tau_max[3 * k + 2] = 33.55;
    if (springs) {
      c->kp[3 * k] = c->kp[3 * k + 1] = c->kp[3 * k + 2] = 75.0;
    } else {
      c->kp[3 * k] = 55; c->kp[3 * k + 1] = 60; c->kp[3 * k + 2] = 60;
    }
    c->kd[3 * k] = 0.8; c->kd[3 * k + 1] = 1.0; c->kd[3 * k + 2] = 1.0;
  }
  /* motor_interface.py:17-22: BACKFLIP raises two thigh upper limits */
  if (task == QSO_TASK_BACKFLIP) { c->ang_hi[7] = PI / 2; c->ang_hi[10] = PI / 2; }
  c->spring_k[0] = 20; c->spring_k[1] = 20; c->spring_k[2] = 30;
  c->spring_b[0] = c->spring_b[1] = c->spring_b[2] = 0.3;
  c->spring_rest[0] = 0; c->spring_rest[1] = PI / 4; c->spring_rest[2] = -PI / 2 + 0.3;
  c->fallen_height = springs ? 0.10 : 0.12;
}

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ---------------------------------------------------- kinematics (a9,a16) */
/* quadruped.py:348-392 */
void qso_fk_jacobian(const double* q, int leg, double* pos, double* J) {
  double sg = SIDE[leg];
  double s1 = sin(q[0]), s2 = sin(q[1]), s3 = sin(q[2]);
  double c1 = cos(q[0]), c2 = cos(q[1]), c3 = cos(q[2]);
  double c23 = c2 * c3 - s2 * s3, s23 = s2 * c3 + c2 * s3;
  if (J) {
    J[0] = 0;
    J[3] = -sg * L1 * s1 + L2 * c2 * c1 + L3 * c23 * c1;
    J[6] = sg * L1 * c1 + L2 * c2 * s1 + L3 * c23 * s1;
    J[1] = -L3 * c23 - L2 * c2;
    J[4] = -L2 * s2 * s1 - L3 * s23 * s1;
    J[7] = L2 * s2 * c1 + L3 * s23 * c1;
    J[2] = -L3 * c23;
    J[5] = -L3 * s23 * s1;
    J[8] = L3 * s23 * c1;
  }
  pos[0] = -L3 * s23 - L2 * s2;
  pos[1] = L1 * sg * c1 + L3 * (s1 * c23) + L2 * c2 * s1;
  pos[2] = L1 * sg * s1 - L3 * (c1 * c23) - L2 * c1 * c2;
}

/* quadruped.py:399-438 */
void qso_ik(const double* xyz, int leg, double* q) {
  double x = xyz[0], y = xyz[1], z = xyz[2];
  double D = (y * y + z * z - L1 * L1 + x * x - L2 * L2 - L3 * L3) / (2 * L3 * L2);
  D = clampd(D, -1.0, 1.0);
  double sg = SIDE[leg];
  double wrist = atan2(-sqrt(1 - D * D), D);
  double sc = y * y + z * z - L1 * L1;
  if (sc < 0.0) sc = 0.0;
  double shoulder = -atan2(z, y) - atan2(sqrt(sc), sg * L1);
  double elbow = atan2(-x, sqrt(sc)) - atan2(L3 * sin(wrist), L2 + L3 * cos(wrist));
  q[0] = -shoulder; q[1] = elbow; q[2] = wrist;
}

/* ------------------------------------------------ action -> command (a5-a8) */
/* action_interface.py:14-15,29-39,58-65 */
static void expand_action(int action_mode, int symm_idx, const double* a, double* a12) {
  if (action_mode == QSO_ACT_DEFAULT) {
    memcpy(a12, a, 12 * sizeof(double));
  } else if (action_mode == QSO_ACT_SYMMETRIC) {
    for (int j = 0; j < 3; j++) {
      a12[j] = a[j];
      a12[3 + j] = (j == symm_idx) ? -a[j] : a[j];
      a12[6 + j] = a[3 + j];
      a12[9 + j] = (j == symm_idx) ? -a[3 + j] : a[3 + j];
    }
  } else {
    double fr[3], rr[3];
    int s = 0;
    for (int j = 0; j < 3; j++) {
      if (j == symm_idx) { fr[j] = 0; rr[j] = 0; }
      else { fr[j] = a[s]; rr[j] = a[2 + s]; s++; }
    }
    for (int j = 0; j < 3; j++) { a12[j] = fr[j]; a12[3 + j] = fr[j]; a12[6 + j] = rr[j]; a12[9 + j] = rr[j]; }
  }
}

/* interface_base.py:84-90 */
static void scale_action(const double* lo, const double* hi, const double* a, double* out) {
  for (int i = 0; i < 12; i++) {
    double x = clampd(a[i], -1, 1);
    double c = lo[i] + 0.5 * (x + 1) * (hi[i] - lo[i]);
    out[i] = clampd(c, lo[i], hi[i]);
  }
}
/* interface_base.py:92-100 */
static void unscale_command(const double* lo, const double* hi, const double* c, double* out) {
  for (int i = 0; i < 12; i++) {
    double x = clampd(c[i], lo[i], hi[i]);
    out[i] = clampd(-1 + 2 * (x - lo[i]) / (hi[i] - lo[i]), -1, 1);
  }
}

static void command_from_a12(const RobotCfg* c, int control_mode, const double* a12, double* cmd) {
  if (control_mode == QSO_CTRL_PD) {
    scale_action(c->ang_lo, c->ang_hi, a12, cmd); /* motor_interface.py:35-37 */
  } else if (control_mode == QSO_CTRL_CARTESIAN_PD) {
    double foot[12];
    scale_action(c->cart_lo, c->cart_hi, a12, foot); /* motor_interface.py:70-80 */
    for (int k = 0; k < 4; k++) qso_ik(foot + 3 * k, k, cmd + 3 * k);
  } else {
    memcpy(cmd, a12, 12 * sizeof(double));
  }
}

void qso_action_to_command(int springs, int control_mode, int action_mode, int task,
                           const double* action, double* cmd12) {
  RobotCfg c;
  robot_cfg(springs, task, &c);
  double a12[12];
  expand_action(action_mode, control_mode == QSO_CTRL_CARTESIAN_PD ? 1 : 0, action, a12);
  command_from_a12(&c, control_mode, a12, cmd12);
}

/* ------------------------------------------------------- torques (a11,a12) */
/* quadruped_motor.py:45-99 */
void qso_pd_torque(const double* kp, const double* kd, const double* tmax, const double* cmd,
                   const double* q, const double* qd, int torque_mode, double* tau) {
  for (int i = 0; i < 12; i++) {
    double t = torque_mode ? cmd[i] : (-1 * (kp[i] * (q[i] - cmd[i])) - kd[i] * (qd[i] - 0.0));
    tau[i] = clampd(t, -tmax[i], tmax[i]);
  }
}
/* quadruped_motor.py:101-104, springs.py:28-74 */
void qso_spring_torque(const double* k3, const double* b3, const double* rest3, const double* q,
                       const double* qd, double* tau) {
  for (int leg = 0; leg < 4; leg++) {
    double k[3] = {k3[0], k3[1], k3[2]}, b[3] = {b3[0], b3[1], b3[2]};
    const double* ql = q + 3 * leg;
    int right = (leg == 0 || leg == 2);
    int hip_cond = right ? (ql[0] > rest3[0]) : (ql[0] < rest3[0]);
    if (hip_cond) { k[0] = 0; b[0] = 0; }
    if (ql[1] < rest3[1]) { k[1] = 0; b[1] = 0; }
    if (ql[2] > rest3[2]) { k[2] = 0; b[2] = 0; }
    for (int j = 0; j < 3; j++) tau[3 * leg + j] = -k[j] * (ql[j] - rest3[j]) - b[j] * qd[3 * leg + j];
  }
}

/* ------------------------------------------------------ orientation (a14) */
/* pybullet getEulerFromQuaternion, as used by quadruped.py:131-139 */
void qso_rpy_from_quat(const double* q, double* rpy) {
  double sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], sqw = q[3] * q[3];
  double sarg = -2 * (q[0] * q[2] - q[3] * q[1]);
  if (sarg <= -0.99999) { rpy[0] = 0; rpy[1] = -0.5 * PI; rpy[2] = 2 * atan2(q[0], -q[1]); }
  else if (sarg >= 0.99999) { rpy[0] = 0; rpy[1] = 0.5 * PI; rpy[2] = 2 * atan2(-q[0], q[1]); }
  else {
    rpy[0] = atan2(2 * (q[1] * q[2] + q[3] * q[0]), sqw - sqx - sqy + sqz);
    rpy[1] = asin(sarg);
    rpy[2] = atan2(2 * (q[0] * q[1] + q[3] * q[2]), sqw + sqx - sqy - sqz);
  }
}
static void quat_R(const double* q, double* R) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double s = 2.0 / (x * x + y * y + z * z + w * w);
  R[0] = 1 - s * (y * y + z * z); R[1] = s * (x * y - w * z); R[2] = s * (x * z + w * y);
  R[3] = s * (x * y + w * z); R[4] = 1 - s * (x * x + z * z); R[5] = s * (y * z - w * x);
  R[6] = s * (x * z - w * y); R[7] = s * (y * z + w * x); R[8] = 1 - s * (x * x + y * y);
}
/* robot_sensors.py:333-340: -Rotation.as_euler("yxz")[0], +2pi when negative
 * and the controller has switched */
double qso_backflip_pitch(const double* q, int switched) {
  double R[9];
  quat_R(q, R);
  double pitch = -atan2(-R[6], R[8]);
  if (pitch < 0 && switched) pitch = 2 * PI + pitch;
  return pitch;
}

/* ------------------------------------------------------------- CPG (a22) */
/* hopf_network.py:117-173 */
void qso_cpg_step(double* X, const double* PHI, double mu, double om_sw, double om_st, double coupling,
                  double dt, double dstep, double height, double gc, double gp, double* xs, double* zs) {
  double r0[4], th0[4], rd[4], thd[4];
  memcpy(r0, X, sizeof r0);
  memcpy(th0, X + 4, sizeof th0);
  const double alpha = 50;
  for (int i = 0; i < 4; i++) {
    rd[i] = alpha * (mu - r0[i] * r0[i]) * r0[i];
    thd[i] = sin(th0[i]) > 0 ? om_sw : om_st;
    for (int j = 0; j < 4; j++)
      if (j != i) thd[i] += r0[j] * coupling * sin(th0[j] - th0[i] - PHI[4 * i + j]);
  }
  for (int i = 0; i < 4; i++) {
    X[i] = r0[i] + dt * rd[i];
    double th = th0[i] + dt * thd[i];
    th = fmod(th, 2 * PI);
    if (th < 0) th += 2 * PI; /* numpy % is a floored modulo */
    X[4 + i] = th;
  }
  for (int i = 0; i < 4; i++) {
    double r = X[i], th = X[4 + i];
    xs[i] = -dstep * r * cos(th);
    zs[i] = sin(th) > 0 ? -height + gc * sin(th) : -height + gp * sin(th);
  }
}
/* hopf_network.py:241-289 */
void qso_cpg_torque(const double* xs, const double* zs, const double* q, const double* qd, double foot_y,
                    const double* kp, const double* kd, double kpc, double kdc, int add_cart, double* tau) {
  for (int i = 0; i < 4; i++) {
    double xyz_d[3] = {xs[i], SIDE[i] * foot_y, zs[i]}, qdes[3], t[3];
    qso_ik(xyz_d, i, qdes);
    for (int j = 0; j < 3; j++) t[j] = -kp[j] * (q[3 * i + j] - qdes[j]) - kd[j] * qd[3 * i + j];
    if (add_cart) {
      double J[9], p[3], dx[3], F[3];
      qso_fk_jacobian(q + 3 * i, i, p, J);
      for (int a = 0; a < 3; a++)
        dx[a] = J[3 * a] * qd[3 * i] + J[3 * a + 1] * qd[3 * i + 1] + J[3 * a + 2] * qd[3 * i + 2];
      for (int a = 0; a < 3; a++) F[a] = -kpc * (p[a] - xyz_d[a]) - kdc * dx[a];
      for (int j = 0; j < 3; j++) t[j] += J[j] * F[0] + J[3 + j] * F[1] + J[6 + j] * F[2];
    }
    for (int j = 0; j < 3; j++) tau[3 * i + j] = t[j];
  }
}

/* ------------------------------------------------------------------- env */
typedef struct {
  int switched, in_air;
  double t_takeoff, pose_takeoff[3], init_height, rpy_takeoff[3];
  double max_flight_time, max_fwd, max_pitch, rel_max_height, max_delta_x, max_height;
  double old_tau[12], new_tau[12];
  double pos[3], vel[3], rpy[3];
  double max_pitch_bf; /* BackFlip.max_pitch: set in __init__ only (robot_tasks.py:524) */
  double old_fwd, actual_fwd;
  /* continuous jumping (task_base.py:222-400) */
  int is_jumping, first_jump, end_jump, jump_counter, good_jump_counter, n_jumps;
  double cumulative_fwd, cumulative_flight_time, max_jump_height;
  double fwd_arr[QSO_MAX_JUMPS], perf_arr[QSO_MAX_JUMPS];
} TaskState;

struct QsoEnv {
  QsoEnvConfig cfg;
  RobotCfg rc;
  QsoWorld* w;
  long sim_steps, env_steps;
  double last_action[12], last_filtered[12];
  double tau_motor[12], tau_spring[12];
  /* contact summary of the last physics step (quadruped.py:224-258) */
  int n_valid, n_invalid, foot_contact[4];
  double foot_force[4];
  TaskState ts;
  /* Butterworth filter (utils/action_filter.py:110-127,191-213) */
  double fb[3], fa[3], xh[2][12], yh[2][12];
  /* landing controller (landing_wrapper.py, landing_wrapper_2.py; utils/timer.py) */
  int land_mode, land_gains;
  double land_timer, land_end, hold_action[12], kp_save[12], kd_save[12];
  /* imitation tasks (task_base.py:169-220) */
  double* demo;
  int demo_len, demo_counter, delta_demo, desired_reset;
  /* go-to-rest controller (go_to_rest_wrapper.py) */
  int rest_active;
  double rest_h, rest_t0, rest_start[12], init_action[12];
};
enum { LAND_POLICY = 0, LAND_HOLD = 1, LAND_LANDING = 2, LAND_SPENT = 3, LAND_TAKEOFF_BF = 4 };

void qso_env_default_config(QsoEnvConfig* c) {
  c->enable_springs = 0;
  c->control_mode = QSO_CTRL_PD;
  c->action_mode = QSO_ACT_SYMMETRIC;
  c->task = QSO_TASK_NO_TASK;
  c->obs_mode = QSO_OBS_ENCODER;
  c->action_repeat = 10;
  c->is_rl_interface = 1;
  c->enable_action_interpolation = 0;
  c->enable_action_filter = 0;
  c->settling_steps = 2500;
  c->landing_mode = 0;
  c->rest_mode = 0;
  c->time_step = 0.001;
}

int qso_env_action_dim(const QsoEnv* e) {
  if (!e->cfg.is_rl_interface) return 12;
  return e->cfg.action_mode == QSO_ACT_DEFAULT ? 12 : (e->cfg.action_mode == QSO_ACT_SYMMETRIC ? 6 : 4);
}

static const int OBS_DIM[] = {24, 30, 24, 27, 28, 29, 28, 29, 32, 27, 28, 29};
int qso_env_obs_dim(const QsoEnv* e) { return OBS_DIM[e->cfg.obs_mode]; }

QsoEnv* qso_env_create(const QsoEnvConfig* c) {
  QsoEnv* e = (QsoEnv*)calloc(1, sizeof(QsoEnv));
  e->cfg = *c;
  robot_cfg(c->enable_springs, c->task, &e->rc);
  e->w = qso_world_create();
  QsoWorldParams p;
  qso_world_get_params(e->w, &p);
  p.dt = c->time_step;
  p.num_iterations = 300 / c->action_repeat; /* quadruped_gym_env.py:113 */
  qso_world_set_params(e->w, &p);
  /* scipy.signal.butter(2, 3/(fs/2)) in closed form; fs = 1/env_time_step */
  double fs = 1.0 / (c->action_repeat * c->time_step);
  double K = tan(PI * 3.0 / fs), nrm = 1.0 / (1 + sqrt(2.0) * K + K * K);
  e->fb[0] = K * K * nrm; e->fb[1] = 2 * e->fb[0]; e->fb[2] = e->fb[0];
  e->fa[0] = 1; e->fa[1] = 2 * (K * K - 1) * nrm; e->fa[2] = (1 - sqrt(2.0) * K + K * K) * nrm;
  return e;
}
void qso_env_destroy(QsoEnv* e) { qso_world_destroy(e->w); free(e->demo); free(e); }
QsoWorld* qso_env_world(QsoEnv* e) { return e->w; }
void qso_env_set_gains(QsoEnv* e, const double* kp, const double* kd) {
  memcpy(e->rc.kp, kp, sizeof e->rc.kp);
  memcpy(e->rc.kd, kd, sizeof e->rc.kd);
}
void qso_env_set_springs(QsoEnv* e, const double* k, const double* b, const double* rest) {
  memcpy(e->rc.spring_k, k, 3 * sizeof(double));
  memcpy(e->rc.spring_b, b, 3 * sizeof(double));
  memcpy(e->rc.spring_rest, rest, 3 * sizeof(double));
}
void qso_env_get_last_action(const QsoEnv* e, double* a12) { memcpy(a12, e->last_action, sizeof e->last_action); }
void qso_env_get_torques(const QsoEnv* e, double* tm, double* tsp) {
  memcpy(tm, e->tau_motor, sizeof e->tau_motor);
  memcpy(tsp, e->tau_spring, sizeof e->tau_spring);
}

/* quadruped.py:224-258 on the contact set of the last stepSimulation */
static void contact_info(QsoEnv* e) {
  e->n_valid = e->n_invalid = 0;
  for (int k = 0; k < 4; k++) { e->foot_contact[k] = 0; e->foot_force[k] = 0; }
  int n = qso_world_num_contacts(e->w);
  for (int i = 0; i < n; i++) {
    int link; double nf, dist, pos[3];
    qso_world_get_contact(e->w, i, &link, &nf, &dist, pos);
    /* pybullet link ids: feet 5,9,13,17; thighs 3,7,11,15 (quadruped.py:545-596) */
    int is_foot = (link >= 5 && (link - 5) % 4 == 0);
    if (!is_foot) { e->n_invalid++; continue; }
    int k = (link - 5) / 4;
    e->n_valid++;
    e->foot_force[k] += nf;
    e->foot_contact[k] = 1;
  }
  /* self collisions count when a calf is involved (quadruped.py:236-241; calf link ids 4, 8, 12, 16) */
  n = qso_world_num_self_contacts(e->w);
  for (int i = 0; i < n; i++) {
    int a, b; double dist;
    qso_world_get_self_contact(e->w, i, &a, &b, &dist);
    const int a_calf = a >= 4 && (a - 4) % 4 == 0, b_calf = b >= 4 && (b - 4) % 4 == 0;
    if (a_calf || b_calf) e->n_invalid++;
  }
}
static int is_flying(const QsoEnv* e) {
  return !(e->foot_contact[0] || e->foot_contact[1] || e->foot_contact[2] || e->foot_contact[3]);
}

/* quadruped.py:288-320 */
static void apply_action(QsoEnv* e, const double* cmd, int torque_mode) {
  double st[QSO_NSTATE];
  qso_world_get_state(e->w, st);
  const double *q = st + 13, *qd = st + 25;
  qso_pd_torque(e->rc.kp, e->rc.kd, e->rc.tau_max, cmd, q, qd, torque_mode, e->tau_motor);
  qso_world_add_torque(e->w, e->tau_motor);
  if (e->cfg.enable_springs) {
    qso_spring_torque(e->rc.spring_k, e->rc.spring_b, e->rc.spring_rest, q, qd, e->tau_spring);
    qso_world_add_torque(e->w, e->tau_spring);
  } else {
    memset(e->tau_spring, 0, sizeof e->tau_spring);
  }
}

static double sim_time(const QsoEnv* e) { return e->sim_steps * e->cfg.time_step; }

/* ---- task (tasks/task_base.py:34-166, tasks/robot_tasks.py) ---- */
static int is_jump_task(int t) { return t != QSO_TASK_NO_TASK; }

static double jumping_distance(const QsoEnv* e) { /* task_base.py:109-116 */
  const TaskState* t = &e->ts;
  double yaw = t->rpy_takeoff[2];
  double dx = t->pos[0] - t->pose_takeoff[0], dy = t->pos[1] - t->pose_takeoff[1];
  double x = cos(yaw) * dx - sin(yaw) * dy;
  return x > 0 ? x : 0;
}

/* task families: 0 = TaskJumping, 1 = TaskContinuousJumping, 2 = TaskContinuousJumping2 */
static int task_family(int t) {
  if (t == QSO_TASK_CONTINUOUS_JUMPING_FORWARD || t == QSO_TASK_CONTINUOUS_JUMPING_FORWARD2) return 1;
  if (t == QSO_TASK_CONTINUOUS_JUMPING_FORWARD3 || t == QSO_TASK_CONTINUOUS_JUMPING_FORWARD_PPO ||
      t == QSO_TASK_CONTINUOUS_JUMPING_FORWARD_DEMO) return 2; /* TaskJumpingDemo2(TaskContinuousJumping2), task_base.py:402 */
  return 0;
}
static void cont2_params(int t, double* jump_limit, double* height_limit, double* bound) {
  if (t == QSO_TASK_CONTINUOUS_JUMPING_FORWARD3) { *jump_limit = 0.6; *height_limit = 0.45; *bound = 0.7; } /* robot_tasks.py:172-177 */
  else { *jump_limit = 0.6; *height_limit = 0.5; *bound = 0.85; } /* robot_tasks.py:553-561, task_base.py:280-287 */
}
static double entropy_fwd(const TaskState* t) { /* task_base.py:365-373 */
  double sum = 0;
  for (int i = 0; i < t->n_jumps; i++) sum += t->fwd_arr[i];
  if (t->jump_counter == 0 || sum < 0.05) return 0;
  int size = t->n_jumps < 3 ? 3 : t->n_jumps;
  double h = 0;
  for (int i = 0; i < t->n_jumps; i++) {
    double p = t->fwd_arr[i] / sum;
    if (p != 0) h -= p * log2(p);
  }
  return h / log2((double)size);
}

static void task_on_step(QsoEnv* e) { /* task_base.py:61-107 */
  TaskState* t = &e->ts;
  if (!is_jump_task(e->cfg.task)) return;
  double st[QSO_NSTATE];
  qso_world_get_state(e->w, st);
  /* _task_jump_take_off :152-160 */
  if (!t->switched && is_flying(e) && st[9] / 9.81 > 0.06) t->switched = 1;
  memcpy(t->old_tau, t->new_tau, sizeof t->old_tau);
  memcpy(t->new_tau, e->tau_motor, sizeof t->new_tau);
  memcpy(t->pos, st, sizeof t->pos);
  memcpy(t->vel, st + 7, sizeof t->vel);
  qso_rpy_from_quat(st + 3, t->rpy);
  double z = t->pos[2], dh = z - t->init_height;
  if (dh < 0) dh = 0;
  if (dh > t->rel_max_height) t->rel_max_height = dh;
  if (fabs(z) > t->max_height) t->max_height = fabs(z);
  if (fabs(t->pos[0]) > t->max_delta_x) t->max_delta_x = fabs(t->pos[0]);
  if (fabs(t->rpy[1]) > t->max_pitch) t->max_pitch = fabs(t->rpy[1]);
  const int fam = task_family(e->cfg.task);
  const int jumping_now = is_flying(e) && st[9] / 9.81 > 0.06; /* detect_jumping, task_base.py:236-241 */
  if (fam == 0) {
    if (is_flying(e)) {
      if (!t->in_air) {
        t->in_air = 1;
        t->t_takeoff = sim_time(e);
        memcpy(t->pose_takeoff, t->pos, sizeof t->pos);
        memcpy(t->rpy_takeoff, t->rpy, sizeof t->rpy);
      } else {
        double d = jumping_distance(e);
        if (d > t->max_fwd) t->max_fwd = d;
      }
    } else {
      if (t->in_air) {
        double ft = sim_time(e) - t->t_takeoff;
        if (ft > t->max_flight_time) t->max_flight_time = ft;
        double d = jumping_distance(e);
        if (d > t->max_fwd) t->max_fwd = d;
        t->in_air = 0;
      } else {
        t->max_fwd = 0; /* task_base.py:106-107 */
      }
    }
  } else if (fam == 1) { /* TaskContinuousJumping._compute_jumping_info, task_base.py:243-262 */
    const double jump_limit = 0.5;
    const double time_limit = e->cfg.task == QSO_TASK_CONTINUOUS_JUMPING_FORWARD ? 0.15 : 0.35; /* robot_tasks.py:105-106,139-140 */
    if (is_flying(e)) {
      if (!t->in_air) {
        t->in_air = 1;
        t->t_takeoff = sim_time(e);
        memcpy(t->pose_takeoff, t->pos, sizeof t->pos);
        memcpy(t->rpy_takeoff, t->rpy, sizeof t->rpy);
        t->is_jumping = jumping_now;
      }
    } else if (t->in_air) {
      double ft = sim_time(e) - t->t_takeoff;
      if (ft > t->max_flight_time) t->max_flight_time = ft;
      double d = jumping_distance(e);
      if (d > t->max_fwd) t->max_fwd = d;
      t->cumulative_fwd += t->max_fwd < jump_limit ? t->max_fwd : jump_limit;           /* update_end_jump :264-266 */
      t->cumulative_flight_time += t->max_flight_time < time_limit ? t->max_flight_time : time_limit;
      t->in_air = 0;
      t->is_jumping = 0;
    }
  } else { /* TaskContinuousJumping2._compute_jumping_info, task_base.py:319-338 */
    double jump_limit, height_limit, bound;
    cont2_params(e->cfg.task, &jump_limit, &height_limit, &bound);
    t->end_jump = 0;
    if (is_flying(e)) {
      if (!t->in_air) {
        t->in_air = 1;
        t->t_takeoff = sim_time(e);
        memcpy(t->pose_takeoff, t->pos, sizeof t->pos);
        memcpy(t->rpy_takeoff, t->rpy, sizeof t->rpy);
        t->is_jumping = jumping_now;
        t->max_jump_height = 0; /* set to z, then restart_jump_performance_variables() zeroes it (:326-330) */
      } else if (t->pos[2] > t->max_jump_height) {
        t->max_jump_height = t->pos[2];
      }
    } else if (t->in_air) {
      double ft = sim_time(e) - t->t_takeoff;
      if (ft > t->max_flight_time) t->max_flight_time = ft;
      if (!t->first_jump) { /* update_end_jump :340-353 */
        t->jump_counter++;
        double d = jumping_distance(e);
        double fwd = d < jump_limit ? d : jump_limit;
        double hh = t->max_jump_height < height_limit ? t->max_jump_height : height_limit;
        double perf = 0.7 * fwd / jump_limit + 0.3 * hh / height_limit;
        if (perf >= bound) t->good_jump_counter++;
        if (t->n_jumps < QSO_MAX_JUMPS) { t->fwd_arr[t->n_jumps] = fwd; t->perf_arr[t->n_jumps] = perf; t->n_jumps++; }
        t->end_jump = 1;
      } else {
        t->first_jump = 0;
      }
      t->in_air = 0;
      t->is_jumping = 0;
    }
  }
  int task = e->cfg.task;
  if (task == QSO_TASK_BACKFLIP) { /* robot_tasks.py:527-530 */
    double p = qso_backflip_pitch(st + 3, t->switched);
    if (p > t->max_pitch_bf) t->max_pitch_bf = p;
  }
  if (task == QSO_TASK_BACKFLIP_PPO) { /* robot_tasks.py:745-747 */
    double p = qso_backflip_pitch(st + 3, t->switched);
    if (p > t->max_pitch) t->max_pitch = p;
  }
  if (task == QSO_TASK_JUMPING_FORWARD_PPO || task == QSO_TASK_JUMPING_FORWARD_PPO_HP) {
    t->old_fwd = t->actual_fwd; /* robot_tasks.py:420-426 */
    t->actual_fwd = t->max_fwd;
  }
}

static int is_demo_task(int t) { return t >= QSO_TASK_JUMPING_IN_PLACE_DEMO && t <= QSO_TASK_CONTINUOUS_JUMPING_FORWARD_DEMO; }

static void task_reset(QsoEnv* e) { /* task_base.py:40-59 */
  TaskState* t = &e->ts;
  if (!is_jump_task(e->cfg.task)) return;
  if (is_demo_task(e->cfg.task)) { /* TaskJumpingDemo._reset, task_base.py:178-183 */
    if (!e->desired_reset) e->demo_counter = 0;
    e->delta_demo = e->demo_len - e->demo_counter;
  }
  double st[QSO_NSTATE];
  qso_world_get_state(e->w, st);
  double keep_bf = t->max_pitch_bf;
  memset(t, 0, sizeof *t);
  t->max_pitch_bf = keep_bf;
  t->first_jump = 1;
  t->t_takeoff = sim_time(e);
  memcpy(t->pose_takeoff, st, sizeof t->pose_takeoff);
  t->init_height = st[2];
  qso_rpy_from_quat(st + 3, t->rpy_takeoff);
  memcpy(t->old_tau, e->tau_motor, sizeof t->old_tau);
  memcpy(t->new_tau, e->tau_motor, sizeof t->new_tau);
  task_on_step(e);
}

static int task_terminated(QsoEnv* e) {
  int task = e->cfg.task;
  if (!is_jump_task(task)) return 0;
  double st[QSO_NSTATE], R[9];
  qso_world_get_state(e->w, st);
  quat_R(st + 3, R);
  int fallen_ground = e->ts.pos[2] < e->rc.fallen_height; /* task_base.py:123-124 */
  int fallen_orient = R[8] < 0.85;                        /* task_base.py:126-130 */
  if (task == QSO_TASK_BACKFLIP) return fallen_ground || e->n_invalid > 0; /* robot_tasks.py:532-533 */
  if (task == QSO_TASK_BACKFLIP_DEMO) return fallen_ground || e->n_invalid > 0 || e->demo_len == e->demo_counter; /* :239-241 */
  if (is_demo_task(task)) /* task_base.py:211-212 */
    return (fallen_orient && fallen_ground) || e->n_invalid > 0 || e->demo_counter == e->demo_len;
  return (fallen_orient && fallen_ground) || e->n_invalid > 0;          /* task_base.py:146-147 */
}

static double norm12(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < 12; i++) s += (a[i] - b[i]) * (a[i] - b[i]);
  return sqrt(s);
}

static double task_reward(QsoEnv* e) {
  const TaskState* t = &e->ts;
  int task = e->cfg.task;
  double max_h_task, k_h;
  if (is_demo_task(task)) { /* TaskJumpingDemo._reward, task_base.py:194-209 */
    const int A = qso_env_action_dim(e);
    const double* da = e->demo + (size_t)e->demo_counter * A;
    double s = 0;
    for (int i = 0; i < A; i++) s += (da[i] - e->last_action[i]) * (da[i] - e->last_action[i]);
    e->demo_counter += 1;
    return exp(-0.35 * sqrt(s)) / e->delta_demo;
  }
  switch (task) {
    case QSO_TASK_JUMPING_IN_PLACE_PPO: max_h_task = 1.0; k_h = 0.023; break;   /* robot_tasks.py:259,268 */
    case QSO_TASK_JUMPING_IN_PLACE_PPO_HP: max_h_task = 1.25; k_h = 0.023; break; /* :493 */
    case QSO_TASK_JUMPING_FORWARD_PPO: max_h_task = 0.9; k_h = 0.026; break;     /* :370,380 */
    case QSO_TASK_JUMPING_FORWARD_PPO_HP: max_h_task = 1.1; k_h = 0.026; break;  /* :507 */
    case QSO_TASK_BACKFLIP_PPO: max_h_task = 0.7; k_h = 0.026; break;            /* :713,723 */
    default: return 0;
  }
  double z = t->pos[2];
  double h_clip = (z < 0.29 || z > max_h_task) ? 0 : z;
  double F = e->foot_force[0] + e->foot_force[1] + e->foot_force[2] + e->foot_force[3];
  double over = F > 800 ? F : 0;
  double rew_h = k_h * h_clip;
  double rew_smooth = 0.015 * exp(-0.1 * norm12(t->old_tau, t->new_tau));
  double rew_contact = -3e-4 * over;
  double rew_pitch = 0.014 * exp(-26 * fabs(t->rpy[1]));
  if (task == QSO_TASK_JUMPING_IN_PLACE_PPO || task == QSO_TASK_JUMPING_IN_PLACE_PPO_HP) {
    double rew_pos = 0.013 * exp(-40.0 * fabs(t->pos[0]));
    return 0.05 * rew_pos + 0.5 * rew_contact + 0.2 * rew_smooth + 0.45 * rew_h + 0.3 * rew_pitch; /* :334-344 */
  }
  if (task == QSO_TASK_JUMPING_FORWARD_PPO || task == QSO_TASK_JUMPING_FORWARD_PPO_HP) {
    double max_fwd = task == QSO_TASK_JUMPING_FORWARD_PPO ? 1.3 : 1.4;
    double fwd = (t->actual_fwd > max_fwd || t->actual_fwd == t->old_fwd) ? 0 : t->actual_fwd; /* :411-416 */
    return 0.4 * rew_contact + 0.2 * rew_smooth + 0.25 * rew_h + 0.3 * rew_pitch + 0.4 * (0.038 * fwd); /* :461-471 */
  }
  /* BackflipPPO :783-800 */
  double st[QSO_NSTATE];
  qso_world_get_state(e->w, st);
  double pbf = z > 0.5 ? qso_backflip_pitch(st + 3, t->switched) : 0;
  return 0.4 * rew_contact + 0.2 * rew_smooth + 0.25 * rew_h + 0.3 * (0.014 * pbf);
}

static double task_reward_end(QsoEnv* e) {
  const TaskState* t = &e->ts;
  int term = task_terminated(e);
  double r = 0;
  switch (e->cfg.task) {
    case QSO_TASK_JUMPING_IN_PLACE: { /* robot_tasks.py:31-57 */
      double h = t->rel_max_height > 0.9 ? 1.0 : t->rel_max_height / 0.9;
      r += 0.7 * h;
      r += h * 0.3 * exp(-t->max_pitch * t->max_pitch / (0.15 * 0.15));
      r += h * 0.05 * exp(-t->max_delta_x * t->max_delta_x / 0.05);
      if (!term) r += 0.1 * h; else r -= 0.08 * (1 + 0.8 * h);
      return r;
    }
    case QSO_TASK_JUMPING_FORWARD: { /* robot_tasks.py:70-99 */
      double h = t->rel_max_height > 0.3 ? 1.0 : t->rel_max_height / 0.3;
      double d = t->max_fwd > 1.3 ? 1.0 : t->max_fwd / 1.3;
      double bm = (h + d) / 2;
      r += 0.25 * h;
      r += 0.5 * d * h;
      r += h * 0.25 * exp(-t->max_pitch * t->max_pitch / (0.15 * 0.15));
      if (!term) r += 0.1 * bm; else r -= 0.08 * (1 + 1.2 * bm);
      return r;
    }
    case QSO_TASK_BACKFLIP: { /* robot_tasks.py:535-550 */
      double h = clampd(t->max_height - 0.3, 0, 0.4) / 0.4;
      double p = t->max_pitch_bf / (2 * PI);
      r += p * 0.4; r += h * 0.4; r += h * p;
      if (t->switched && !term) r += 0.2;
      return r;
    }
    case QSO_TASK_JUMPING_IN_PLACE_PPO:
    case QSO_TASK_JUMPING_IN_PLACE_PPO_HP: /* robot_tasks.py:349-358 */
      return term ? -0.25 * t->max_height : 0.0;
    case QSO_TASK_JUMPING_FORWARD_PPO:
    case QSO_TASK_JUMPING_FORWARD_PPO_HP: /* robot_tasks.py:476-485 */
      return term ? 0.0 : 0.05 * (t->max_fwd + t->max_height) / 2;
    case QSO_TASK_BACKFLIP_PPO: /* robot_tasks.py:802-809 */
      return term ? 0.0 : 0.2 * (0.7 * t->max_pitch / 5 + 0.3 * t->max_height) / 2;
    case QSO_TASK_CONTINUOUS_JUMPING_FORWARD: { /* robot_tasks.py:112-131 */
      double a = t->cumulative_flight_time / 0.15, b = t->cumulative_fwd / 0.5, bm = (a + b) / 2;
      r += 0.25 * a; r += 0.5 * b;
      r += a * 0.25 * exp(-t->max_pitch * t->max_pitch / (0.15 * 0.15));
      if (!term) r += 0.1 * bm;
      return r;
    }
    case QSO_TASK_CONTINUOUS_JUMPING_FORWARD2: { /* robot_tasks.py:146-166 */
      double a = (t->max_flight_time < 0.35 ? t->max_flight_time : 0.35) / 0.35;
      double b = (t->max_fwd < 0.5 ? t->max_fwd : 0.5) / 0.5, bm = (a + b) / 2;
      r += 0.25 * a; r += 0.5 * b;
      r += b * 0.15 * exp(-(t->max_pitch * t->max_pitch / (0.15 * 0.15)));
      r += 0.4 * (sim_time(e) / 10.0) * bm;
      if (!term) r += 0.2 * bm;
      return r;
    }
    case QSO_TASK_CONTINUOUS_JUMPING_FORWARD3: { /* robot_tasks.py:183-212 */
      int size = t->n_jumps < 3 ? 3 : t->n_jumps;
      double sum = 0, mx = t->n_jumps < 3 ? 0.0 : t->perf_arr[0]; /* np.max over the zero-padded array */
      for (int i = 0; i < t->n_jumps; i++) { sum += t->perf_arr[i]; if (t->perf_arr[i] > mx) mx = t->perf_arr[i]; }
      double avg = sum / size;
      double rew_entropy = exp((entropy_fwd(t) - 1) / 0.3), rew_avg = 0;
      rew_avg += avg * 0.15 * exp(-t->max_pitch * t->max_pitch / (0.15 * 0.15));
      rew_avg += avg * 0.4 * (sim_time(e) / 10.0);
      rew_avg += avg * rew_entropy * 0.2;
      rew_avg += avg * 0.25;
      r = 0.8 * rew_avg + 0.2 * mx;
      r += 0.1 * t->good_jump_counter;
      if (!term) r += 0.2 * avg;
      return r;
    }
    case QSO_TASK_CONTINUOUS_JUMPING_FORWARD_PPO: { /* robot_tasks.py:687-698 */
      int size = t->n_jumps < 3 ? 3 : t->n_jumps;
      double sum = 0;
      for (int i = 0; i < t->n_jumps; i++) sum += t->perf_arr[i];
      double rr = (sum / size) * exp((entropy_fwd(t) - 1) / 0.3);
      return term ? rr - 1 : rr;
    }
    default: return 0;
  }
}

void qso_env_get_task_state(const QsoEnv* e, double* o) {
  const TaskState* t = &e->ts;
  memset(o, 0, 32 * sizeof(double));
  o[0] = t->switched; o[1] = t->in_air; o[2] = t->t_takeoff;
  o[3] = t->pose_takeoff[0]; o[4] = t->pose_takeoff[1]; o[5] = t->pose_takeoff[2];
  o[6] = t->init_height; o[7] = t->rpy_takeoff[2];
  o[8] = t->max_flight_time; o[9] = t->max_fwd; o[10] = t->max_pitch;
  o[11] = t->rel_max_height; o[12] = t->max_delta_x; o[13] = t->max_height;
  o[14] = t->max_pitch_bf; o[15] = t->old_fwd; o[16] = t->actual_fwd;
  o[17] = e->n_valid; o[18] = e->n_invalid;
  for (int k = 0; k < 4; k++) { o[19 + k] = e->foot_contact[k]; o[23 + k] = e->foot_force[k]; }
  o[27] = (double)e->sim_steps; o[28] = (double)e->env_steps;
  o[29] = t->is_jumping; o[30] = t->cumulative_fwd; o[31] = t->cumulative_flight_time;
}
void qso_env_get_jump_arrays(const QsoEnv* e, double* o) {
  const TaskState* t = &e->ts;
  o[0] = t->n_jumps;
  o[1] = t->jump_counter; o[2] = t->good_jump_counter; o[3] = t->first_jump; o[4] = t->max_jump_height; o[5] = t->end_jump;
  for (int i = 0; i < QSO_MAX_JUMPS; i++) { o[6 + i] = t->fwd_arr[i]; o[6 + QSO_MAX_JUMPS + i] = t->perf_arr[i]; }
}

/* ---- sensors (sensors/robot_sensors.py, sensors/sensor_collection.py:18-105) ---- */
static void observe(QsoEnv* e, double* obs) {
  double st[QSO_NSTATE], rpy[3], R[9];
  qso_world_get_state(e->w, st);
  const double *pos = st, *quat = st + 3, *v = st + 7, *wv = st + 10, *q = st + 13, *qd = st + 25;
  qso_rpy_from_quat(quat, rpy);
  quat_R(quat, R);
  double wl[3]; /* quadruped.py:141-170: R^T omega */
  for (int i = 0; i < 3; i++) wl[i] = R[i] * wv[0] + R[3 + i] * wv[1] + R[6 + i] * wv[2];
  int n = 0;
#define PUT(x) obs[n++] = (x)
#define PUTN(p, c) do { for (int _i = 0; _i < (c); _i++) obs[n++] = (p)[_i]; } while (0)
  int m = e->cfg.obs_mode;
  double landing = e->ts.switched;
  switch (m) {
    case QSO_OBS_ENCODER: PUTN(q, 12); PUTN(qd, 12); break;
    case QSO_OBS_ENCODER_2: PUTN(v, 3); PUTN(wv, 3); PUTN(q, 12); PUTN(qd, 12); break;
    case QSO_OBS_CARTESIAN_NO_IMU: {
      double fp[12], fv[12];
      for (int k = 0; k < 4; k++) { /* quadruped.py:440-449 */
        double J[9];
        qso_fk_jacobian(q + 3 * k, k, fp + 3 * k, J);
        for (int a = 0; a < 3; a++)
          fv[3 * k + a] = J[3 * a] * qd[3 * k] + J[3 * a + 1] * qd[3 * k + 1] + J[3 * a + 2] * qd[3 * k + 2];
      }
      PUTN(fp, 12); PUTN(fv, 12);
      break;
    }
    case QSO_OBS_ARS_BASIC: PUTN(q, 12); PUTN(qd, 12); PUT(rpy[1]); PUT(pos[2]); PUT(v[2]); break;
    case QSO_OBS_ARS_SENSOR: PUTN(q, 12); PUTN(qd, 12); PUT(rpy[1]); PUT(wl[1]); PUT(pos[2]); PUT(v[2]); break;
    case QSO_OBS_LANDING_SENSOR:
      PUTN(q, 12); PUTN(qd, 12); PUT(rpy[1]); PUT(wl[1]); PUT(pos[2]); PUT(v[2]); PUT(landing); break;
    case QSO_OBS_PPO_BASIC: PUTN(q, 12); PUTN(qd, 12); PUT(rpy[1]); PUT(pos[2]); PUT(v[2]); PUT(landing); break;
    case QSO_OBS_PPO_BASIC_X:
      PUTN(q, 12); PUTN(qd, 12); PUT(rpy[1]); PUT(pos[2]); PUT(v[2]); PUT(v[0]); PUT(landing); break;
    case QSO_OBS_PPO_BASIC_CONTACT:
      PUTN(q, 12); PUTN(qd, 12); PUT(rpy[1]); PUT(pos[2]); PUT(v[2]); PUT(landing);
      for (int k = 0; k < 4; k++) PUT(e->foot_contact[k]);
      break;
    case QSO_OBS_ARS_BACKFLIP:
      PUTN(q, 12); PUTN(qd, 12); PUT(pos[2]); PUT(v[2]); PUT(qso_backflip_pitch(quat, e->ts.switched)); break;
    case QSO_OBS_PPO_BACKFLIP:
      PUTN(q, 12); PUTN(qd, 12); PUT(pos[2]); PUT(v[2]); PUT(qso_backflip_pitch(quat, e->ts.switched));
      PUT(landing); break;
    case QSO_OBS_PPO_CONTINUOUS_JUMPING_FORWARD:
      PUTN(q, 12); PUTN(qd, 12); PUT(pos[2]); PUT(v[2]); PUT(rpy[1]); PUT(landing); PUT((double)e->ts.is_jumping); break;
  }
#undef PUT
#undef PUTN
}

void qso_env_set_demo(QsoEnv* e, const double* actions, int length) {
  const int A = qso_env_action_dim(e);
  free(e->demo);
  e->demo = (double*)malloc(sizeof(double) * (size_t)length * A);
  memcpy(e->demo, actions, sizeof(double) * (size_t)length * A);
  e->demo_len = length;
  e->demo_counter = 0;
}
void qso_env_set_demo_counter(QsoEnv* e, int value) { e->demo_counter = value; }
int qso_env_get_demo_counter(const QsoEnv* e) { return e->demo_counter; }

static void reset_impl(QsoEnv* e, double mu, const double* desired, double* obs);
void qso_env_reset(QsoEnv* e, double mu, double* obs) { reset_impl(e, mu, NULL, obs); }
void qso_env_reset_to_state(QsoEnv* e, double mu, const double* state37, double* obs) { reset_impl(e, mu, state37, obs); }

/* ---- reset (quadruped_gym_env.py:278-329, interface_base.py:182-200) ---- */
static void reset_impl(QsoEnv* e, double mu, const double* desired, double* obs) {
  double st[QSO_NSTATE];
  memset(st, 0, sizeof st);
  st[2] = 0.32; st[6] = 1.0; /* configs:23,26 */
  memcpy(st + 13, e->rc.init_angles, 12 * sizeof(double));
  if (desired) memcpy(st, desired, sizeof st); /* Quadruped.reset_desired_state, quadruped.py:521-525 */
  e->desired_reset = desired != NULL;
  qso_world_set_state(e->w, st);
  QsoWorldParams p;
  qso_world_get_params(e->w, &p);
  p.mu_ground = mu; /* env_randomizer.py:287-289 */
  qso_world_set_params(e->w, &p);
  e->sim_steps = e->env_steps = 0;
  if (e->land_gains) { /* a new Quadruped has the default gains (quadruped_gym_env.py:299-319) */
    memcpy(e->rc.kp, e->kp_save, sizeof e->kp_save);
    memcpy(e->rc.kd, e->kd_save, sizeof e->kd_save);
    e->land_gains = 0;
  }
  e->land_mode = LAND_POLICY;
  e->land_timer = e->land_end = 0;
  e->rest_active = 0;
  memset(e->last_action, 0, sizeof e->last_action);
  memset(e->last_filtered, 0, sizeof e->last_filtered);
  memset(e->tau_motor, 0, sizeof e->tau_motor);
  memset(e->tau_spring, 0, sizeof e->tau_spring);
  int adim = qso_env_action_dim(e);
  if (desired) {
    /* robot_desired_state is set: no settle, _last_action stays zero (quadruped_gym_env.py:284,288-289) */
  } else if (e->cfg.is_rl_interface) {
    /* _settle_robot_by_reference(get_init_pose(), 2500) */
    const RobotCfg* c = &e->rc;
    double a12[12], act[12], cmd[12];
    int cart = e->cfg.control_mode == QSO_CTRL_CARTESIAN_PD;
    int sidx = cart ? 1 : 0;
    if (cart) unscale_command(c->cart_lo, c->cart_hi, c->nominal_foot, a12);
    else unscale_command(c->ang_lo, c->ang_hi, c->init_angles, a12);
    /* _convert_to_actual_action_space (action_interface.py:17-18,41-44,67-74) */
    if (e->cfg.action_mode == QSO_ACT_DEFAULT) memcpy(act, a12, sizeof a12);
    else if (e->cfg.action_mode == QSO_ACT_SYMMETRIC) { memcpy(act, a12, 3 * sizeof(double)); memcpy(act + 3, a12 + 6, 3 * sizeof(double)); }
    else {
      int s = 0;
      for (int j = 0; j < 3; j++) if (j != sidx) { act[s] = a12[j]; act[2 + s] = a12[6 + j]; s++; }
    }
    double b12[12];
    expand_action(e->cfg.action_mode, sidx, act, b12);
    command_from_a12(c, e->cfg.control_mode, b12, cmd);
    for (int i = 0; i < e->cfg.settling_steps; i++) {
      apply_action(e, cmd, 0);
      qso_world_step(e->w);
    }
    /* settling_action = _transform_motor_command_to_action(settling_command)
     * (interface_base.py:196-200).  In CARTESIAN_PD mode settling_command holds
     * JOINT ANGLES (it went through IK) but is scaled with the CARTESIAN limits,
     * so the stored _last_action is (0, 1, -1) per leg: reproduced as is. */
    double s12[12], sact[12];
    if (cart) unscale_command(c->cart_lo, c->cart_hi, cmd, s12);
    else unscale_command(c->ang_lo, c->ang_hi, cmd, s12);
    if (e->cfg.action_mode == QSO_ACT_DEFAULT) memcpy(sact, s12, sizeof s12);
    else if (e->cfg.action_mode == QSO_ACT_SYMMETRIC) { memcpy(sact, s12, 3 * sizeof(double)); memcpy(sact + 3, s12 + 6, 3 * sizeof(double)); }
    else {
      int s2 = 0;
      for (int j = 0; j < 3; j++) if (j != sidx) { sact[s2] = s12[j]; sact[2 + s2] = s12[6 + j]; s2++; }
    }
    memset(e->last_action, 0, sizeof e->last_action);
    memcpy(e->last_action, sact, adim * sizeof(double));
    memcpy(e->init_action, e->last_action, sizeof e->init_action); /* ac_interface.get_init_action(), interface_base.py:74-78 */
  } else {
    /* settle_robot_by_pd (control_interface/utils.py:22-30): PD, DEFAULT space, 1500 ticks */
    double a12[12], cmd[12];
    unscale_command(e->rc.ang_lo, e->rc.ang_hi, e->rc.init_angles, a12);
    scale_action(e->rc.ang_lo, e->rc.ang_hi, a12, cmd);
    for (int i = 0; i < 1500; i++) {
      apply_action(e, cmd, 0);
      qso_world_step(e->w);
    }
  }
  contact_info(e);
  task_reset(e);
  if (e->cfg.enable_action_filter) {
    for (int h = 0; h < 2; h++)
      for (int i = 0; i < 12; i++) { e->xh[h][i] = e->last_action[i]; e->yh[h][i] = e->last_action[i]; }
  }
  {
    double st[QSO_NSTATE];
    qso_world_get_state(e->w, st);
    e->rest_h = st[2]; /* GoToRestWrapper.reset: h_old = h_actual = z (go_to_rest_wrapper.py:83-87) */
  }
  if (obs) observe(e, obs);
}

/* ---- step (quadruped_gym_env.py:227-256) ---- */
/* env.get_landing_action() (quadruped_gym_env.py:375-379): landing pose -> action space */
static void landing_action(const QsoEnv* e, double* act) {
  const RobotCfg* c = &e->rc;
  double a12[12], pose[12];
  int cart = e->cfg.control_mode == QSO_CTRL_CARTESIAN_PD, sidx = cart ? 1 : 0;
  if (cart) { /* CARTESIAN_LANDING_POSE: nominal foot position with z = LANDING_Z (configs:67-72) */
    memcpy(pose, c->nominal_foot, sizeof pose);
    for (int k = 0; k < 4; k++) pose[3 * k + 2] = -0.29;
    unscale_command(c->cart_lo, c->cart_hi, pose, a12);
  } else {    /* ANGLE_LANDING_POSE = INIT_MOTOR_ANGLES (configs:38) */
    unscale_command(c->ang_lo, c->ang_hi, c->init_angles, a12);
  }
  memset(act, 0, 12 * sizeof(double));
  if (e->cfg.action_mode == QSO_ACT_DEFAULT) memcpy(act, a12, sizeof a12);
  else if (e->cfg.action_mode == QSO_ACT_SYMMETRIC) { memcpy(act, a12, 3 * sizeof(double)); memcpy(act + 3, a12 + 6, 3 * sizeof(double)); }
  else {
    int s = 0;
    for (int j = 0; j < 3; j++) if (j != sidx) { act[s] = a12[j]; act[2 + s] = a12[6 + j]; s++; }
  }
}
void qso_env_get_landing_state(const QsoEnv* e, double* o) { o[0] = e->land_mode; o[1] = e->land_timer; o[2] = e->land_end; }
void qso_env_get_rest_state(const QsoEnv* e, double* o) { o[0] = e->rest_active; o[1] = e->rest_h; o[2] = e->rest_t0; }
/* ActionWrapper._transform_motor_command_to_action (action_interface.py): a 12-vector scaled with the interface's
 * limits (joint angles, or -- CARTESIAN_PD -- foot positions), reduced to the action space */
static void command_to_action_space(const QsoEnv* e, const double* cmd12, double* act) {
  const RobotCfg* c = &e->rc;
  double a12[12];
  int cart = e->cfg.control_mode == QSO_CTRL_CARTESIAN_PD, sidx = cart ? 1 : 0;
  if (cart) unscale_command(c->cart_lo, c->cart_hi, cmd12, a12);
  else unscale_command(c->ang_lo, c->ang_hi, cmd12, a12);
  memset(act, 0, 12 * sizeof(double));
  if (e->cfg.action_mode == QSO_ACT_DEFAULT) memcpy(act, a12, sizeof a12);
  else if (e->cfg.action_mode == QSO_ACT_SYMMETRIC) { memcpy(act, a12, 3 * sizeof(double)); memcpy(act + 3, a12 + 6, 3 * sizeof(double)); }
  else {
    int s = 0;
    for (int j = 0; j < 3; j++) if (j != sidx) { act[s] = a12[j]; act[2 + s] = a12[6 + j]; s++; }
  }
}

void qso_env_step(QsoEnv* e, const double* action, double* obs, double* reward, int* done, int* truncated) {
  int adim = qso_env_action_dim(e);
  double cur[12], scripted[12], ramp[12];
  if (e->cfg.rest_mode && e->rest_active) {
    /* go_to_rest (go_to_rest_wrapper.py:58-81): ramp from the pose at activation to the init action */
    const double T = e->cfg.enable_springs ? 1.0 : 0.3, t = sim_time(e), t0 = e->rest_t0, t1 = t0 + T;
    for (int i = 0; i < 12; i++) { /* generate_ramp, interface_base.py:112-119 */
      if (t < t0) ramp[i] = e->rest_start[i];
      else if (t > t1) ramp[i] = e->init_action[i];
      else ramp[i] = e->rest_start[i] + (e->init_action[i] - e->rest_start[i]) * (t - t0) / (t1 - t0);
    }
    action = ramp;
  }
  if (e->cfg.landing_mode) {
    /* take_off_phase (landing_wrapper.py:47-54): repeat the action until the timer is up, then landing_phase */
    if (e->land_mode == LAND_HOLD) {
      if (e->land_timer > e->land_end) { /* Timer.time_up, utils/timer.py:39-43 */
        e->land_mode = LAND_LANDING;
        if (e->cfg.landing_mode == 1) { /* temporary_switch_motor_control_gain, landing_wrapper.py:18-36 */
          memcpy(e->kp_save, e->rc.kp, sizeof e->kp_save);
          memcpy(e->kd_save, e->rc.kd, sizeof e->kd_save);
          for (int i = 0; i < 12; i++) { e->rc.kp[i] = 60.0; e->rc.kd[i] = 1.5; }
          e->land_gains = 1;
        }
      } else {
        e->land_timer += e->cfg.action_repeat * e->cfg.time_step; /* step_timer */
        action = e->hold_action;
      }
    }
    if (e->land_mode == LAND_TAKEOFF_BF) { /* take_off_action, landing_wrapper_backflip.py:21 */
      static const double bf[12] = {0, 1, -1, 0, 1, -1, 0, 0, 0, 0, 0, 0};
      memcpy(scripted, bf, sizeof bf);
      action = scripted;
    }
    if (e->land_mode == LAND_LANDING) { landing_action(e, scripted); action = scripted; }
  }
  memset(cur, 0, sizeof cur);
  memcpy(cur, action, adim * sizeof(double));
  memcpy(e->last_action, cur, sizeof cur);
  if (e->cfg.enable_action_filter) { /* action_filter.py:110-121 */
    for (int i = 0; i < adim; i++) {
      double y = cur[i] * e->fb[0] + e->xh[0][i] * e->fb[1] + e->xh[1][i] * e->fb[2] -
                 (e->yh[0][i] * e->fa[1] + e->yh[1][i] * e->fa[2]);
      e->xh[1][i] = e->xh[0][i]; e->xh[0][i] = cur[i];
      e->yh[1][i] = e->yh[0][i]; e->yh[0][i] = y;
      cur[i] = y;
    }
    memcpy(e->last_filtered, cur, sizeof cur);
  }
  int cart = e->cfg.control_mode == QSO_CTRL_CARTESIAN_PD;
  for (int s = 0; s < e->cfg.action_repeat; s++) {
    double cmd[12];
    if (e->cfg.is_rl_interface) {
      /* _interpolate_actions (:187-205) reads _last_action, which step() has
       * already overwritten with the current action (:229-234): a no-op. */
      double a12[12];
      expand_action(e->cfg.action_mode, cart ? 1 : 0, cur, a12);
      command_from_a12(&e->rc, e->cfg.control_mode, a12, cmd);
      apply_action(e, cmd, 0);
    } else {
      apply_action(e, cur, e->cfg.control_mode == QSO_CTRL_TORQUE);
    }
    qso_world_step(e->w);
    e->sim_steps++;
  }
  e->env_steps++;
  contact_info(e);
  task_on_step(e);
  double r = task_reward(e);
  int term = task_terminated(e);
  int d = 0, tr = 0;
  if (term || sim_time(e) > 10.0) { tr = !term; d = 1; }
  if (d) r += task_reward_end(e);
  if (obs) observe(e, obs);
  *reward = r; *done = d; *truncated = tr;
  if (e->cfg.landing_mode && !d) {
    double st[QSO_NSTATE];
    qso_world_get_state(e->w, st);
    const int lm = e->cfg.landing_mode, flying = is_flying(e);
    /* what starts the scripted phase: the take-off switch (:58-66), or -- continuous variant -- a detected jump */
    const int trigger = lm == 3 ? e->ts.is_jumping : e->ts.switched;
    if (e->land_mode == LAND_POLICY && trigger) {
      if (lm == 4 || lm == 5) {
        e->land_mode = LAND_TAKEOFF_BF; /* landing_wrapper_backflip.py:54-60,72-73 */
      } else { /* take_off_phase with the apex timer, start_jumping_timer (landing_wrapper.py:47-60) */
        e->land_mode = LAND_HOLD;
        memset(e->hold_action, 0, sizeof e->hold_action);
        memcpy(e->hold_action, e->last_action, adim * sizeof(double));
        e->land_timer = sim_time(e);
        e->land_end = e->land_timer + st[9] / 9.81; /* task.compute_time_for_peak_heihgt, task_base.py:157-160 */
      }
    } else if (e->land_mode == LAND_TAKEOFF_BF) {
      /* until PitchBackFlip._get_pitch >= 5 pi / 8 (landing_wrapper_backflip.py:22-23,57-60) */
      if (qso_backflip_pitch(st + 3, e->ts.switched) >= 5.0 * PI / 8.0)
        e->land_mode = (lm == 5 && !flying) ? LAND_SPENT : LAND_LANDING; /* backflip2: `while ... and is_flying` :50 */
    } else if (e->land_mode == LAND_LANDING) {
      if ((lm == 2 || lm == 5) && !flying) e->land_mode = LAND_SPENT;      /* landing_wrapper_2.py:39-46,71 */
      else if (lm == 3 && !e->ts.is_jumping) e->land_mode = LAND_POLICY;    /* landing_wrapper_continuous.py:39-46 */
    }
  }
  /* GoToRestWrapper.step (:43-52) runs where the wrapper below it returns: after a step that leaves the landing
   * controller unscripted */
  if (e->cfg.rest_mode && !e->rest_active && (e->land_mode == LAND_POLICY || e->land_mode == LAND_SPENT)) {
    double st[QSO_NSTATE];
    qso_world_get_state(e->w, st);
    const double h_old = e->rest_h;
    e->rest_h = st[2];
    const int ground = e->foot_contact[0] && e->foot_contact[1] && e->foot_contact[2] && e->foot_contact[3];
    if (!d && e->ts.switched && ground && e->rest_h - h_old > 0) { /* rest_condition :89-95 */
      e->rest_active = 1;
      e->rest_t0 = sim_time(e);
      command_to_action_space(e, st + 13, e->rest_start); /* get_start_action :54-56 */
      if (!e->land_gains) {
        memcpy(e->kp_save, e->rc.kp, sizeof e->kp_save);
        memcpy(e->kd_save, e->rc.kd, sizeof e->kd_save);
        e->land_gains = 1;
      }
      for (int i = 0; i < 12; i++) { e->rc.kp[i] = 60.0; e->rc.kd[i] = e->cfg.enable_springs ? 0.8 : 1.5; } /* :21-41 */
    }
  }
}
