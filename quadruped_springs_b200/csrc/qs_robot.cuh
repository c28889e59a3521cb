// qs_robot.cuh -- device functions for the reference's own arithmetic:
// action mapping, PD / PEA torque, leg FK / Jacobian / IK, orientation
// quantities.  Templated on the scalar type so the same code serves the fp32
// product kernels and the fp64 algorithm-check instantiation.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "qs_types.h"

#define QS_DEV __host__ __device__ __forceinline__
#define QS_DEVONLY __device__ __forceinline__
#define QS_PI 3.14159265358979323846

namespace qs {

template <typename T> QS_DEV T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> QS_DEV T tmax(T a, T b) { return a > b ? a : b; }
// min / max as ONE instruction (FMNMX) where the select's NaN behaviour is not needed: `a > b ? a : b` has to be a compare
// and a select (it returns b when either is NaN, FMNMX returns the other operand), and a tick has ~130 of them.  In a
// clamp tmin(tmax(x, lo), hi) the two agree even for a NaN x (both give lo); a max-reduction differs only in dropping a NaN.
template <typename T> QS_DEV T fmin_t(T a, T b) { return tmin(a, b); }
template <typename T> QS_DEV T fmax_t(T a, T b) { return tmax(a, b); }
QS_DEV float fmin_t(float a, float b) { return fminf(a, b); }
QS_DEV float fmax_t(float a, float b) { return fmaxf(a, b); }
template <typename T> QS_DEV T clampt(T x, T lo, T hi) { return fmin_t(fmax_t(x, lo), hi); }

QS_DEV void sincos_t(float x, float* s, float* c) { sincosf(x, s, c); }
QS_DEV void sincos_t(double x, double* s, double* c) { sincos(x, s, c); }
QS_DEV float sqrt_t(float x) { return sqrtf(x); }
QS_DEV double sqrt_t(double x) { return sqrt(x); }
QS_DEV float rsqrt_t(float x) { return rsqrtf(x); }
QS_DEV double rsqrt_t(double x) { return 1.0 / sqrt(x); }
QS_DEV float atan2_t(float y, float x) { return atan2f(y, x); }
QS_DEV double atan2_t(double y, double x) { return atan2(y, x); }
QS_DEV float asin_t(float x) { return asinf(x); }
QS_DEV double asin_t(double x) { return asin(x); }
QS_DEV float exp_t(float x) { return expf(x); }
QS_DEV double exp_t(double x) { return exp(x); }
QS_DEV float abs_t(float x) { return fabsf(x); }
// division on the hot path: fp32 device code multiplies by the hardware reciprocal (one MUFU.RCP and one FMUL, <= 2 ulp).
// The divisors of the tick are masses, inertias, Delassus diagonals and dt, all far inside the normal range, so neither
// the IEEE slow-path subroutine (~10% of the tick's code size once) nor the range scaling of div.approx (four more
// instructions per division, ~40 divisions per tick) buys anything.  The fp64 check instantiation and host code divide exactly.
QS_DEV float div_t(float a, float b) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return a * r;
#else
  return a / b;
#endif
}
// 1/sqrt of a quantity that is O(1) and positive by construction (Cholesky pivots, quaternion norm): the bare MUFU.RSQ
QS_DEV float rsqrt_pos(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}
QS_DEV double rsqrt_pos(double x) { return 1.0 / sqrt(x); }
// sqrt of a non-negative quantity inside the tick (|omega|): the bare MUFU.SQRT (0 -> 0, ~1 ulp)
QS_DEV float sqrt_pos(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
QS_DEV double sqrt_pos(double x) { return sqrt(x); }
// sin x / x and cos x for |x| <= pi/8 (half the rotation of one tick, clamped there by the integrator): Taylor series to
// x^10 / x^12, exact to fp32 rounding on that range -- a dozen multiply-adds instead of sincosf's range reduction
QS_DEV void sinc_cos_small(float x, float* sinc, float* c) {
  const float x2 = x * x;
  *sinc = 1.f + x2 * (-1.f / 6.f + x2 * (1.f / 120.f + x2 * (-1.f / 5040.f + x2 * (1.f / 362880.f + x2 * (-1.f / 39916800.f)))));
  *c = 1.f + x2 * (-0.5f + x2 * (1.f / 24.f + x2 * (-1.f / 720.f + x2 * (1.f / 40320.f + x2 * (-1.f / 3628800.f + x2 * (1.f / 479001600.f))))));
}
QS_DEV void sinc_cos_small(double x, double* sinc, double* c) {
  *c = cos(x);
  *sinc = fabs(x) < 1e-8 ? 1.0 - x * x / 6.0 : sin(x) / x;
}
QS_DEV double div_t(double a, double b) { return a / b; }
// sin/cos of a joint angle inside the physics tick: |x| <= pi, so the 2-MUFU approximation (abs error
// < 5e-7) is within fp32 rounding of the dynamics; the analytic accessor kernels keep the exact sincosf
QS_DEV void sincos_tick(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
  __sincosf(x, s, c);
#else
  sincosf(x, s, c);
#endif
}
QS_DEV void sincos_tick(double x, double* s, double* c) { sincos(x, s, c); }
QS_DEV double abs_t(double x) { return fabs(x); }

// leg geometry used by the reference's analytic kinematics
// (go1/configs_go1_with_springs.py:56-58)
template <typename T> struct LegLen {
  static constexpr T l1 = T(0.0847), l2 = T(0.213), l3 = T(0.213);
};

// quadruped.py:360-362: -1 for the right legs (0 FR, 2 RR), +1 for the left
QS_DEV float side_sign(int leg) { return (leg & 1) ? 1.f : -1.f; }

// quadruped.py:348-392 (_compute_jacobian_and_position)
template <typename T>
QS_DEV void fk_jacobian(const T* q, int leg, T* pos, T* J /*9 row-major or nullptr*/) {
  const T l1 = LegLen<T>::l1, l2 = LegLen<T>::l2, l3 = LegLen<T>::l3;
  const T sg = T(side_sign(leg));
  T s1, c1, s2, c2, s3, c3;
  sincos_t(q[0], &s1, &c1);
  sincos_t(q[1], &s2, &c2);
  sincos_t(q[2], &s3, &c3);
  const T c23 = c2 * c3 - s2 * s3, s23 = s2 * c3 + c2 * s3;
  if (J) {
    J[0] = T(0);
    J[3] = -sg * l1 * s1 + l2 * c2 * c1 + l3 * c23 * c1;
    J[6] = sg * l1 * c1 + l2 * c2 * s1 + l3 * c23 * s1;
    J[1] = -l3 * c23 - l2 * c2;
    J[4] = -l2 * s2 * s1 - l3 * s23 * s1;
    J[7] = l2 * s2 * c1 + l3 * s23 * c1;
    J[2] = -l3 * c23;
    J[5] = -l3 * s23 * s1;
    J[8] = l3 * s23 * c1;
  }
  pos[0] = -l3 * s23 - l2 * s2;
  pos[1] = l1 * sg * c1 + l3 * (s1 * c23) + l2 * c2 * s1;
  pos[2] = l1 * sg * s1 - l3 * (c1 * c23) - l2 * c1 * c2;
}

// quadruped.py:399-438 (ComputeInverseKinematics)
template <typename T> QS_DEV void leg_ik(const T* xyz, int leg, T* q) {
  const T l1 = LegLen<T>::l1, l2 = LegLen<T>::l2, l3 = LegLen<T>::l3;
  const T x = xyz[0], y = xyz[1], z = xyz[2];
  T D = (y * y + z * z - l1 * l1 + x * x - l2 * l2 - l3 * l3) / (T(2) * l3 * l2);
  D = clampt(D, T(-1), T(1));
  const T sg = T(side_sign(leg));
  const T sw = -sqrt_t(T(1) - D * D);  // sin(wrist), cos(wrist) = D
  const T wrist = atan2_t(sw, D);
  T sc = y * y + z * z - l1 * l1;
  sc = tmax(sc, T(0));
  const T rt = sqrt_t(sc);
  const T shoulder = -atan2_t(z, y) - atan2_t(rt, sg * l1);
  T swr, cwr;
  sincos_t(wrist, &swr, &cwr);
  const T elbow = atan2_t(-x, rt) - atan2_t(l3 * swr, l2 + l3 * cwr);
  q[0] = -shoulder;
  q[1] = elbow;
  q[2] = wrist;
}

// action_interface.py:14-15,29-39,58-65 (_convert_to_default_action_space)
template <typename T> QS_DEV void expand_action(int action_mode, int symm_idx, const T* a, T* a12) {
  if (action_mode == QS_ACT_DEFAULT) {
#pragma unroll
    for (int i = 0; i < 12; i++) a12[i] = a[i];
  } else if (action_mode == QS_ACT_SYMMETRIC) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      a12[j] = a[j];
      a12[3 + j] = (j == symm_idx) ? -a[j] : a[j];
      a12[6 + j] = a[3 + j];
      a12[9 + j] = (j == symm_idx) ? -a[3 + j] : a[3 + j];
    }
  } else {
    // np.insert(leg, symm_idx, 0); left = right (no negation)
    T fr[3], rr[3];
    if (symm_idx == 0) {
      fr[0] = T(0); fr[1] = a[0]; fr[2] = a[1];
      rr[0] = T(0); rr[1] = a[2]; rr[2] = a[3];
    } else {
      fr[0] = a[0]; fr[1] = T(0); fr[2] = a[1];
      rr[0] = a[2]; rr[1] = T(0); rr[2] = a[3];
    }
#pragma unroll
    for (int j = 0; j < 3; j++) { a12[j] = fr[j]; a12[3 + j] = fr[j]; a12[6 + j] = rr[j]; a12[9 + j] = rr[j]; }
  }
}

// interface_base.py:84-90 + motor_interface.py:35-37,70-80
template <typename T>
QS_DEV void action12_to_command(const RobotConst& rc, int control_mode, const T* a12, T* cmd) {
  if (control_mode == QS_CTRL_PD) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const T lo = T(rc.ang_lo[i]), hi = T(rc.ang_hi[i]);
      const T x = clampt(a12[i], T(-1), T(1));
      cmd[i] = clampt(lo + T(0.5) * (x + T(1)) * (hi - lo), lo, hi);
    }
  } else if (control_mode == QS_CTRL_CARTESIAN_PD) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      T foot[3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const T lo = T(rc.cart_lo[3 * k + j]), hi = T(rc.cart_hi[3 * k + j]);
        const T x = clampt(a12[3 * k + j], T(-1), T(1));
        foot[j] = clampt(lo + T(0.5) * (x + T(1)) * (hi - lo), lo, hi);
      }
      leg_ik(foot, k, cmd + 3 * k);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 12; i++) cmd[i] = a12[i];
  }
}

// interface_base.py:92-100 (_scale_helper_motor_command_to_action)
template <typename T> QS_DEV T command_to_action1(T c, T lo, T hi) {
  const T x = clampt(c, lo, hi);
  return clampt(T(-1) + T(2) * (x - lo) / (hi - lo), T(-1), T(1));
}

// quadruped_motor.py:45-99 for one joint
template <typename T> QS_DEV T pd_torque1(T kp, T kd, T tmax_, T cmd, T q, T qd, bool torque_mode) {
  const T t = torque_mode ? cmd : (T(-1) * (kp * (q - cmd)) - kd * (qd - T(0)));
  return clampt(t, -tmax_, tmax_);
}

// quadruped_motor.py:101-104 + springs.py:28-74 for one leg
template <typename T>
QS_DEV void spring_torque_leg(int leg, const T* k3, const T* b3, const T* rest3, const T* q, const T* qd, T* tau) {
  T k0 = k3[0], k1 = k3[1], k2 = k3[2], b0 = b3[0], b1 = b3[1], b2 = b3[2];
  const bool right = !(leg & 1);
  const bool hip_cond = right ? (q[0] > rest3[0]) : (q[0] < rest3[0]);
  if (hip_cond) { k0 = T(0); b0 = T(0); }
  if (q[1] < rest3[1]) { k1 = T(0); b1 = T(0); }
  if (q[2] > rest3[2]) { k2 = T(0); b2 = T(0); }
  tau[0] = -k0 * (q[0] - rest3[0]) - b0 * qd[0];
  tau[1] = -k1 * (q[1] - rest3[1]) - b1 * qd[1];
  tau[2] = -k2 * (q[2] - rest3[2]) - b2 * qd[2];
}

// rotation matrix (local -> world) of an xyzw quaternion
template <typename T> QS_DEV void quat_to_R(const T* q, T* R) {
  const T x = q[0], y = q[1], z = q[2], w = q[3];
  const T s = T(2) / (x * x + y * y + z * z + w * w);
  R[0] = T(1) - s * (y * y + z * z); R[1] = s * (x * y - w * z); R[2] = s * (x * z + w * y);
  R[3] = s * (x * y + w * z); R[4] = T(1) - s * (x * x + z * z); R[5] = s * (y * z - w * x);
  R[6] = s * (x * z - w * y); R[7] = s * (y * z + w * x); R[8] = T(1) - s * (x * x + y * y);
}

// pybullet getEulerFromQuaternion as used by quadruped.py:131-139
template <typename T> QS_DEV void rpy_from_quat(const T* q, T* rpy) {
  const T sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], sqw = q[3] * q[3];
  const T sarg = T(-2) * (q[0] * q[2] - q[3] * q[1]);
  if (sarg <= T(-0.99999)) {
    rpy[0] = T(0); rpy[1] = T(-0.5 * QS_PI); rpy[2] = T(2) * atan2_t(q[0], -q[1]);
  } else if (sarg >= T(0.99999)) {
    rpy[0] = T(0); rpy[1] = T(0.5 * QS_PI); rpy[2] = T(2) * atan2_t(-q[0], q[1]);
  } else {
    rpy[0] = atan2_t(T(2) * (q[1] * q[2] + q[3] * q[0]), sqw - sqx - sqy + sqz);
    rpy[1] = asin_t(sarg);
    rpy[2] = atan2_t(T(2) * (q[0] * q[1] + q[3] * q[2]), sqw + sqx - sqy - sqz);
  }
}

// robot_sensors.py:333-340 (PitchBackFlip._get_pitch): -as_euler("yxz")[0] = atan2(R20, R22)
template <typename T> QS_DEV T backflip_pitch(const T* R, bool switched) {
  T pitch = -atan2_t(-R[6], R[8]);
  if (pitch < T(0) && switched) pitch = T(2 * QS_PI) + pitch;
  return pitch;
}

}  // namespace qs
