// qs_model_host.h -- host-side (double precision) construction of the constants
// the kernels take as parameters: the merged 13-body Go1 model, the robot-level
// limits/gains of the reference's config modules, and the sensor-noise table.
// Every number restates a reference file:line (paths under
// /root/reference/quadruped_spring/).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

#include "qs_types.h"

namespace qs {
namespace host {

struct Sym3 { double v[6]; };  // xx xy xz yy yz zz

struct RigidBody {
  double m;
  double c[3];     // com in the frame the body is expressed in
  double Ic[6];    // about its own com, same axes
};

inline void add_body(RigidBody& acc, double m, const double* c, const double* Idiag) {
  // merge (m, c, diag inertia about c) into acc (both in the same frame)
  const double M = acc.m + m;
  double cn[3];
  for (int i = 0; i < 3; i++) cn[i] = M > 0 ? (acc.m * acc.c[i] + m * c[i]) / M : 0.0;
  auto shift = [&](double mm, const double* from, double* I6) {
    const double d[3] = {from[0] - cn[0], from[1] - cn[1], from[2] - cn[2]};
    const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    I6[0] += mm * (dd - d[0] * d[0]); I6[1] += -mm * d[0] * d[1]; I6[2] += -mm * d[0] * d[2];
    I6[3] += mm * (dd - d[1] * d[1]); I6[4] += -mm * d[1] * d[2]; I6[5] += mm * (dd - d[2] * d[2]);
  };
  double I6[6];
  for (int i = 0; i < 6; i++) I6[i] = acc.Ic[i];
  shift(acc.m, acc.c, I6);
  I6[0] += Idiag[0]; I6[3] += Idiag[1]; I6[5] += Idiag[2];
  shift(m, c, I6);
  acc.m = M;
  for (int i = 0; i < 3; i++) acc.c[i] = cn[i];
  for (int i = 0; i < 6; i++) acc.Ic[i] = I6[i];
}

// Bullet's inertia for a link loaded without URDF_USE_INERTIA_FROM_FILE
// (quadruped.py:534-539): collision AABB extents ex,ey,ez
inline void aabb_inertia(double m, double ex, double ey, double ez, double* I) {
  I[0] = m / 12.0 * (ey * ey + ez * ez);
  I[1] = m / 12.0 * (ex * ex + ez * ez);
  I[2] = m / 12.0 * (ex * ex + ey * ey);
}

// btCollisionShape::getContactBreakingThreshold in relative mode: (|aabb diag|/2 +
// |aabb centre|) * gContactBreakingThreshold, aabb in the link's inertial frame
inline double shape_thresh(const double* half, const double* centre_minus_com, double gthr) {
  const double r = std::sqrt(half[0] * half[0] + half[1] * half[1] + half[2] * half[2]);
  const double c = std::sqrt(centre_minus_com[0] * centre_minus_com[0] + centre_minus_com[1] * centre_minus_com[1] +
                             centre_minus_com[2] * centre_minus_com[2]);
  return (r + c) * gthr;
}

template <typename T> inline void build_model(ModelConstT<T>& M, double gthr) {
  std::memset(&M, 0, sizeof(M));
  // ---- body 0 = base + trunk + imu_link (go1.urdf:47-111), base frame
  RigidBody tr{0, {0, 0, 0}, {0, 0, 0, 0, 0, 0}};
  const double zero3[3] = {0, 0, 0};
  add_body(tr, 0.00001, zero3, zero3);  // `base`: no collision shape -> zero inertia
  double It[3];
  aabb_inertia(5.204, 0.3762, 0.0935, 0.114, It);
  const double ctr[3] = {0.0223, 0.0, -0.0005};
  add_body(tr, 5.204, ctr, It);
  double Ii[3];
  aabb_inertia(0.001, 0.001, 0.001, 0.001, Ii);
  const double cimu[3] = {-0.01592, -0.06659, -0.00617};
  add_body(tr, 0.001, cimu, Ii);
  M.trunk_m = T(tr.m);
  const double cc = tr.c[0] * tr.c[0] + tr.c[1] * tr.c[1] + tr.c[2] * tr.c[2];
  for (int i = 0; i < 3; i++) M.trunk_h[i] = T(tr.m * tr.c[i]);
  M.trunk_I[0] = T(tr.Ic[0] + tr.m * (cc - tr.c[0] * tr.c[0]));
  M.trunk_I[1] = T(tr.Ic[1] - tr.m * tr.c[0] * tr.c[1]);
  M.trunk_I[2] = T(tr.Ic[2] - tr.m * tr.c[0] * tr.c[2]);
  M.trunk_I[3] = T(tr.Ic[3] + tr.m * (cc - tr.c[1] * tr.c[1]));
  M.trunk_I[4] = T(tr.Ic[4] - tr.m * tr.c[1] * tr.c[2]);
  M.trunk_I[5] = T(tr.Ic[5] + tr.m * (cc - tr.c[2] * tr.c[2]));
  // ---- legs FR FL RR RL (go1.urdf:112-241 and mirrored copies)
  const double sx[4] = {1, 1, -1, -1}, sy[4] = {-1, 1, -1, 1};
  double Ih[3], Ith[3], Ica[3];
  aabb_inertia(0.591, 0.092, 0.04, 0.092, Ih);     // hip cylinder r 0.046, len 0.04 rolled 90 deg
  aabb_inertia(0.92, 0.034, 0.0245, 0.213, Ith);   // thigh box pitched 90 deg
  aabb_inertia(0.131, 0.016, 0.016, 0.213, Ica);   // calf box pitched 90 deg
  const double Ifoot = 0.4 * 0.06 * 0.02 * 0.02;   // foot sphere, single child at identity
  for (int k = 0; k < 4; k++) {
    M.hip_pos[k][0] = T(sx[k] * 0.1881); M.hip_pos[k][1] = T(sy[k] * 0.04675); M.hip_pos[k][2] = T(0);
    M.thigh_off_y[k] = T(sy[k] * 0.08);
    const double chip[3] = {-sx[k] * 0.00541, -sy[k] * 0.00074, 6e-06};
    const double cth[3] = {-0.003468, -sy[k] * 0.018947, -0.032736};
    const double cca[3] = {0.006286, 0.001307, -0.122269};  // not mirrored in the URDF
    const double cfo[3] = {0, 0, -0.213};
    RigidBody hip{0, {0, 0, 0}, {0, 0, 0, 0, 0, 0}}, th = hip, ca = hip;
    add_body(hip, 0.591, chip, Ih);
    add_body(th, 0.92, cth, Ith);
    add_body(ca, 0.131, cca, Ica);
    const double If3[3] = {Ifoot, Ifoot, Ifoot};
    add_body(ca, 0.06, cfo, If3);
    // leg_dynamics relies on it (body_spi_hip / body_spi_diag): hip and thigh are single links, diagonal about their com
    for (int i : {1, 2, 4}) if (hip.Ic[i] != 0.0 || th.Ic[i] != 0.0) std::abort();
    const RigidBody* B[3] = {&hip, &th, &ca};
    for (int b = 0; b < 3; b++) {
      M.body_m[k][b] = T(B[b]->m);
      for (int i = 0; i < 3; i++) M.body_com[k][b][i] = T(B[b]->c[i]);
      for (int i = 0; i < 6; i++) M.body_Ic[k][b][i] = T(B[b]->Ic[i]);
    }
  }
  M.link_len = T(0.213);
  M.foot_radius = T(0.02);
  {
    const double h[3] = {0.02, 0.02, 0.02};
    M.foot_thresh = T(shape_thresh(h, zero3, gthr));
  }
  {
    const double h[3] = {0.3762 / 2, 0.0935 / 2, 0.114 / 2}, c[3] = {-ctr[0], -ctr[1], -ctr[2]};
    for (int i = 0; i < 3; i++) M.trunk_half[i] = T(h[i]);
    M.trunk_thresh = T(shape_thresh(h, c, gthr));
  }
  {
    const double h[3] = {0.0005, 0.0005, 0.0005};
    for (int i = 0; i < 3; i++) M.imu_pos[i] = T(cimu[i]);
    M.imu_half = T(0.0005);
    M.imu_thresh = T(shape_thresh(h, zero3, gthr));
  }
  {
    const double h[3] = {0.046, 0.02, 0.046}, c[3] = {0.00541, 0.00074, -6e-06};
    M.hip_r = T(0.046); M.hip_hl = T(0.02);
    M.hip_thresh = T(shape_thresh(h, c, gthr));
  }
  {
    const double h[3] = {0.017, 0.01225, 0.1065}, c[3] = {0.003468, 0.018947, -0.1065 + 0.032736};
    for (int i = 0; i < 3; i++) M.thigh_half[i] = T(h[i]);
    M.thigh_c[0] = T(0); M.thigh_c[1] = T(0); M.thigh_c[2] = T(-0.1065);
    M.thigh_thresh = T(shape_thresh(h, c, gthr));
  }
  {
    const double h[3] = {0.008, 0.008, 0.1065}, c[3] = {-0.006286, -0.001307, -0.1065 + 0.122269};
    for (int i = 0; i < 3; i++) M.calf_half[i] = T(h[i]);
    M.calf_c[0] = T(0); M.calf_c[1] = T(0); M.calf_c[2] = T(-0.1065);
    M.calf_thresh = T(shape_thresh(h, c, gthr));
  }
  M.joint_lo[0] = T(-1.0471975512); M.joint_hi[0] = T(1.0471975512);
  M.joint_lo[1] = T(-0.663225115758); M.joint_hi[1] = T(2.96705972839);
  M.joint_lo[2] = T(-2.72271363311); M.joint_hi[2] = T(-0.837758040957);
}

inline int obs_dim_of(int mode) {
  static const int d[] = {24, 30, 24, 27, 28, 29, 28, 29, 32, 27, 28, 29};
  return (mode >= 0 && mode < 12) ? d[mode] : -1;
}
inline int action_dim_of(int is_rl, int action_mode) {
  if (!is_rl) return 12;
  return action_mode == QS_ACT_DEFAULT ? 12 : (action_mode == QS_ACT_SYMMETRIC ? 6 : 4);
}

inline void build_robot(const qs_config& c, RobotConst& R) {
  std::memset(&R, 0, sizeof(R));
  const double PI = 3.14159265358979323846;
  const bool sp = c.enable_springs != 0;
  const double side[4] = {-1, 1, -1, 1};
  for (int k = 0; k < 4; k++) {
    R.init_angles[3 * k] = 0.f; R.init_angles[3 * k + 1] = float(PI / 4); R.init_angles[3 * k + 2] = float(-PI / 2);
    R.ang_hi[3 * k] = 0.2f; R.ang_hi[3 * k + 1] = float(PI / 4 + 0.5); R.ang_hi[3 * k + 2] = -0.95f;
    R.ang_lo[3 * k] = -0.2f; R.ang_lo[3 * k + 1] = float(PI / 4 - 0.5); R.ang_lo[3 * k + 2] = sp ? -2.5f : -2.12f;
    const double nom[3] = {0.0, side[k] * 0.0847, -0.32};
    const double up[3] = {0.2, 0.05, sp ? 0.18 : 0.11}, dn[3] = {0.2, 0.05, 0.07};
    for (int j = 0; j < 3; j++) {
      R.nominal_foot[3 * k + j] = float(nom[j]);
      R.cart_hi[3 * k + j] = float(nom[j] + up[j]);
      R.cart_lo[3 * k + j] = float(nom[j] - dn[j]);
    }
    R.tau_max[3 * k] = 23.7f; R.tau_max[3 * k + 1] = 23.7f; R.tau_max[3 * k + 2] = 33.55f;
    if (sp) { R.kp[3 * k] = R.kp[3 * k + 1] = R.kp[3 * k + 2] = 75.f; }
    else { R.kp[3 * k] = 55.f; R.kp[3 * k + 1] = 60.f; R.kp[3 * k + 2] = 60.f; }
    R.kd[3 * k] = 0.8f; R.kd[3 * k + 1] = 1.0f; R.kd[3 * k + 2] = 1.0f;
  }
  if (c.task == QS_TASK_BACKFLIP) { R.ang_hi[7] = float(PI / 2); R.ang_hi[10] = float(PI / 2); }  // motor_interface.py:20-22
  R.spring_k[0] = 20.f; R.spring_k[1] = 20.f; R.spring_k[2] = 30.f;
  R.spring_b[0] = R.spring_b[1] = R.spring_b[2] = 0.3f;
  R.spring_rest[0] = 0.f; R.spring_rest[1] = float(PI / 4); R.spring_rest[2] = float(-PI / 2 + 0.3);
  R.fallen_height = sp ? 0.10f : 0.12f;
  // ---- sensor noise (configs:215-230), in the element order of the obs mode
  const double STD = 0.01;
  double qn[12];
  for (int k = 0; k < 4; k++) {
    qn[3 * k] = 0.2 * STD * 0.1;
    qn[3 * k + 1] = (PI / 4 + 0.5) * STD * 0.1;
    qn[3 * k + 2] = (sp ? 2.5 : 2.12) * STD * 0.1;
  }
  const double qdn = 10.0 * STD * 0.6, hn = 0.4 * STD * 0.8, pn = PI * STD * 0.9, vln = 5.0 * STD * 0.8,
               van = 3.0 * STD, prn = 5.0 * STD, fvn = 10.0 * STD;
  const double fpn[3] = {0.1 * STD, 0.05 * STD, 0.1 * STD};
  int n = 0;
  auto put = [&](double v) { R.obs_noise[n++] = float(v); };
  auto put_q = [&]() { for (int i = 0; i < 12; i++) put(qn[i]); for (int i = 0; i < 12; i++) put(qdn); };
  switch (c.obs_mode) {
    case QS_OBS_ENCODER: put_q(); break;
    case QS_OBS_ENCODER_2: for (int i = 0; i < 3; i++) put(vln); for (int i = 0; i < 3; i++) put(van); put_q(); break;
    case QS_OBS_CARTESIAN_NO_IMU: for (int i = 0; i < 12; i++) put(fpn[i % 3]); for (int i = 0; i < 12; i++) put(fvn); break;
    case QS_OBS_ARS_BASIC: put_q(); put(pn); put(hn); put(vln); break;
    case QS_OBS_ARS_SENSOR: put_q(); put(pn); put(prn); put(hn); put(vln); break;
    case QS_OBS_LANDING_SENSOR: put_q(); put(pn); put(prn); put(hn); put(vln); put(0); break;
    case QS_OBS_PPO_BASIC: put_q(); put(pn); put(hn); put(vln); put(0); break;
    case QS_OBS_PPO_BASIC_X: put_q(); put(pn); put(hn); put(vln); put(vln); put(0); break;
    case QS_OBS_PPO_BASIC_CONTACT: put_q(); put(pn); put(hn); put(vln); put(0); for (int i = 0; i < 4; i++) put(0); break;
    case QS_OBS_ARS_BACKFLIP: put_q(); put(hn); put(vln); put(pn); break;
    case QS_OBS_PPO_BACKFLIP: put_q(); put(hn); put(vln); put(pn); put(0); break;
    default: put_q(); put(hn); put(vln); put(pn); put(0); put(0); break;
  }
  // ---- scipy.signal.butter(2, 3 Hz / (fs/2)) in closed form, fs = 1 / env_time_step
  const double fs = 1.0 / (c.action_repeat * c.time_step);
  const double K = std::tan(PI * 3.0 / fs), nrm = 1.0 / (1 + std::sqrt(2.0) * K + K * K);
  R.filt_b[0] = float(K * K * nrm); R.filt_b[1] = float(2 * K * K * nrm); R.filt_b[2] = float(K * K * nrm);
  R.filt_a[0] = 1.f; R.filt_a[1] = float(2 * (K * K - 1) * nrm); R.filt_a[2] = float((1 - std::sqrt(2.0) * K + K * K) * nrm);
  // ---- env.get_landing_action() (quadruped_gym_env.py:375-379): landing pose -> [-1, 1] -> action space
  // ANGLE_LANDING_POSE = INIT_MOTOR_ANGLES (configs:38); CARTESIAN_LANDING_POSE = nominal foot, z = -0.29 (configs:67-72)
  const bool cart = c.control_mode == QS_CTRL_CARTESIAN_PD;
  double a12[12];
  for (int i = 0; i < 12; i++) {
    const double lo = cart ? R.cart_lo[i] : R.ang_lo[i], hi = cart ? R.cart_hi[i] : R.ang_hi[i];
    double pose = cart ? (i % 3 == 2 ? -0.29 : double(R.nominal_foot[i])) : double(R.init_angles[i]);
    pose = std::min(std::max(pose, lo), hi);
    a12[i] = std::min(std::max(-1.0 + 2.0 * (pose - lo) / (hi - lo), -1.0), 1.0);  // interface_base.py:92-100
  }
  const int sidx = cart ? 1 : 0;  // _convert_to_actual_action_space (action_interface.py:17-18,41-44,67-74)
  if (!c.is_rl_interface || c.action_mode == QS_ACT_DEFAULT) {
    for (int i = 0; i < 12; i++) R.landing_action[i] = float(a12[i]);
  } else if (c.action_mode == QS_ACT_SYMMETRIC) {
    for (int j = 0; j < 3; j++) { R.landing_action[j] = float(a12[j]); R.landing_action[3 + j] = float(a12[6 + j]); }
  } else {
    int s2 = 0;
    for (int j = 0; j < 3; j++) if (j != sidx) { R.landing_action[s2] = float(a12[j]); R.landing_action[2 + s2] = float(a12[6 + j]); s2++; }
  }
  R.env_dt = float(c.action_repeat * c.time_step);
}

}  // namespace host
}  // namespace qs
