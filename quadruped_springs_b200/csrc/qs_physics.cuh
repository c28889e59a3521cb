// qs_physics.cuh -- one physics tick for one env held in one thread's registers.
//
// Replaces pybullet.stepSimulation() (reference: quadruped_gym_env.py:218-219)
// for the Go1 URDF.  Same model and step semantics as the CPU oracle
// (oracle/qso_physics.c), different formulation, chosen for a one-env-per-thread
// GPU mapping:
//   * merged 13-body model (fixed links folded), everything expressed in the
//     BASE frame about the base origin, so composite inertias are plain sums;
//   * per leg: 3x3 joint-space inertia M_kk, coupling F_k (3x6), bias by
//     Newton-Euler; the base sees the Schur complement
//     S = M_bb - sum_k F_k^T M_kk^-1 F_k  (= its articulated-body inertia);
//   * contacts: PGS in a reduced space.  With S = L L^T, G_k = J_b - J_kk B_k,
//     Y_k = L^-1 G_k^T and H_k = J_kk M_kk^-1 J_kk^T the Delassus block is
//     A_kl = Y_k^T Y_l + delta_kl H_k, so a sweep needs one 6-vector z = sum Y lam
//     instead of Bullet's 18-vector; row order, clamps, cone projection, warm
//     start and early exit are Bullet's (see the oracle for the restatement).
#pragma once
#include "qs_robot.cuh"

namespace qs {

template <typename T> QS_DEV void cross3(const T* a, const T* b, T* o) {
  const T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
template <typename T> QS_DEV void cross3_add(const T* a, const T* b, T* o) {
  const T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] += x; o[1] += y; o[2] += z;
}
template <typename T> QS_DEV T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> QS_DEV void m3_v(const T* R, const T* v, T* o) {  // o = R v
  const T x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  const T y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  const T z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
template <typename T> QS_DEV void m3t_v(const T* R, const T* v, T* o) {  // o = R^T v
  const T x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  const T y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  const T z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
template <typename T> QS_DEV void sym3_mul(const T* I, const T* w, T* o) {
  o[0] = I[0] * w[0] + I[1] * w[1] + I[2] * w[2];
  o[1] = I[1] * w[0] + I[3] * w[1] + I[4] * w[2];
  o[2] = I[2] * w[0] + I[4] * w[1] + I[5] * w[2];
}

// rigid-body (or composite) spatial inertia about the base origin, base axes
template <typename T> struct SpI {
  T m, h[3], I[6];
};
// momentum of velocity (w, v):  L = I w + h x v,  p = m v - h x w
template <typename T> QS_DEV void spi_apply(const SpI<T>& A, const T* w, const T* v, T* L, T* p) {
  sym3_mul(A.I, w, L);
  cross3_add(A.h, v, L);
  T t[3];
  cross3(A.h, w, t);
  p[0] = A.m * v[0] - t[0]; p[1] = A.m * v[1] - t[1]; p[2] = A.m * v[2] - t[2];
}
template <typename T> QS_DEV void spi_add(SpI<T>& a, const SpI<T>& b) {
  a.m += b.m;
#pragma unroll
  for (int i = 0; i < 3; i++) a.h[i] += b.h[i];
#pragma unroll
  for (int i = 0; i < 6; i++) a.I[i] += b.I[i];
}
// inertia of a body given in its link frame (mass, com, Ic about com) placed at
// rotation R (link->base) and origin r (base coords)
template <typename T>
QS_DEV void body_spi(T m, const T* com, const T* Ic, const T* R, const T* r, SpI<T>& o) {
  T c[3];
  m3_v(R, com, c);
  c[0] += r[0]; c[1] += r[1]; c[2] += r[2];
  T t[9];  // t = R * Ic
#pragma unroll
  for (int i = 0; i < 3; i++) {
    t[3 * i + 0] = R[3 * i] * Ic[0] + R[3 * i + 1] * Ic[1] + R[3 * i + 2] * Ic[2];
    t[3 * i + 1] = R[3 * i] * Ic[1] + R[3 * i + 1] * Ic[3] + R[3 * i + 2] * Ic[4];
    t[3 * i + 2] = R[3 * i] * Ic[2] + R[3 * i + 1] * Ic[4] + R[3 * i + 2] * Ic[5];
  }
  const T cc = dot3(c, c);
  o.I[0] = t[0] * R[0] + t[1] * R[1] + t[2] * R[2] + m * (cc - c[0] * c[0]);
  o.I[1] = t[0] * R[3] + t[1] * R[4] + t[2] * R[5] - m * c[0] * c[1];
  o.I[2] = t[0] * R[6] + t[1] * R[7] + t[2] * R[8] - m * c[0] * c[2];
  o.I[3] = t[3] * R[3] + t[4] * R[4] + t[5] * R[5] + m * (cc - c[1] * c[1]);
  o.I[4] = t[3] * R[6] + t[4] * R[7] + t[5] * R[8] - m * c[1] * c[2];
  o.I[5] = t[6] * R[6] + t[7] * R[7] + t[8] * R[8] + m * (cc - c[2] * c[2]);
  o.m = m;
  o.h[0] = m * c[0]; o.h[1] = m * c[1]; o.h[2] = m * c[2];
}

// upper-triangular index of a symmetric 6x6 stored in 21 entries
__host__ __device__ constexpr int s6(int i, int j) { return i <= j ? i * (11 - i) / 2 + j : j * (11 - j) / 2 + i; }

template <typename T> struct EnvState {
  T pos[3], quat[4], vlin[3], vang[3], q[12], qd[12];
};
template <typename T> struct ContactState {
  T lam_n[4];   // normal impulse of the last tick per foot (N s); force = lam/dt
  int mask;     // bits 0-3: foot manifold point exists (quadruped.py:250-257)
  int invalid;  // number of non-foot shapes touching the ground (quadruped.py:243-249)
  int work_contacts;   // sum over ticks of active foot contacts        } algorithmic-work counters
  int work_row_iters;  // sum over ticks of contacts x PGS sweeps run   } (bench.py roofline)
};

template <typename T> struct LegKin {
  T s1, c1, s2, c2, s23, c23;
  T a2[3];                     // thigh/calf joint axis (base coords); hip axis is x
  T r1[3], r2[3], r3[3], r4[3];  // joint origins and foot centre (base coords)
};

template <typename T> QS_DEV void leg_kin(int k, const T* q, const ModelConstT<T>& M, LegKin<T>& K) {
  T s3, c3;
  sincos_t(q[0], &K.s1, &K.c1);
  sincos_t(q[1], &K.s2, &K.c2);
  sincos_t(q[2], &s3, &c3);
  K.c23 = K.c2 * c3 - K.s2 * s3;
  K.s23 = K.s2 * c3 + K.c2 * s3;
  K.a2[0] = T(0); K.a2[1] = K.c1; K.a2[2] = K.s1;
  const T l = M.link_len, dy = M.thigh_off_y[k];
#pragma unroll
  for (int i = 0; i < 3; i++) K.r1[i] = M.hip_pos[k][i];
  K.r2[0] = K.r1[0]; K.r2[1] = K.r1[1] + dy * K.c1; K.r2[2] = K.r1[2] + dy * K.s1;
  K.r3[0] = K.r2[0] - l * K.s2; K.r3[1] = K.r2[1] + l * K.s1 * K.c2; K.r3[2] = K.r2[2] - l * K.c1 * K.c2;
  K.r4[0] = K.r3[0] - l * K.s23; K.r4[1] = K.r3[1] + l * K.s1 * K.c23; K.r4[2] = K.r3[2] - l * K.c1 * K.c23;
}

// contact Jacobian rows of foot k for direction d (base coords) at point pc:
// base part (pc x d, d), joint part a_j . ((pc - r_j) x d)
template <typename T>
QS_DEV void foot_jac_dir(const LegKin<T>& K, const T* pc, const T* d, T* Jb /*6*/, T* Jk /*3*/) {
  cross3(pc, d, Jb);
  Jb[3] = d[0]; Jb[4] = d[1]; Jb[5] = d[2];
  T u[3], t[3];
  u[0] = pc[0] - K.r1[0]; u[1] = pc[1] - K.r1[1]; u[2] = pc[2] - K.r1[2];
  cross3(u, d, t);
  Jk[0] = t[0];  // hip axis = x
  u[0] = pc[0] - K.r2[0]; u[1] = pc[1] - K.r2[1]; u[2] = pc[2] - K.r2[2];
  cross3(u, d, t);
  Jk[1] = dot3(K.a2, t);
  u[0] = pc[0] - K.r3[0]; u[1] = pc[1] - K.r3[1]; u[2] = pc[2] - K.r3[2];
  cross3(u, d, t);
  Jk[2] = dot3(K.a2, t);
}

// f += I A + V x* (I V) for one body; V = (Vw, Vv), A = (Aw, Av)
template <typename T>
QS_DEV void body_force(const SpI<T>& I, const T* Vw, const T* Vv, const T* Aw, const T* Av, T* fn, T* fl) {
  T L[3], p[3], La[3], pa[3];
  spi_apply(I, Vw, Vv, L, p);
  spi_apply(I, Aw, Av, La, pa);
  // crf(V)(L,p) = (w x L + v x p, w x p)
  cross3_add(Vw, L, La);
  cross3_add(Vv, p, La);
  cross3_add(Vw, p, pa);
#pragma unroll
  for (int i = 0; i < 3; i++) { fn[i] += La[i]; fl[i] += pa[i]; }
}

template <typename T> QS_DEV T clamp_vel(T v, T mx) { return tmin(tmax(v, -mx), mx); }

// ------------------------------------------------------------------------------------------
// One tick.  tau = joint torques applied this tick (motor + spring, already summed).
// cs: in = previous tick's contact impulses (warm start), out = this tick's.
// ------------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ void physics_tick(EnvState<T>& st, const T* tau, T mu, ContactState<T>& cs,
                             const ModelConstT<T>& M, const SolverConst& SC, bool detect_invalid) {
  const T dt = T(SC.dt);
  const T mcv = T(SC.max_coord_vel);
  T Rb[9];
  quat_to_R(st.quat, Rb);
  T wb[3], vb[3];
  m3t_v(Rb, st.vang, wb);
  m3t_v(Rb, st.vlin, vb);
  const T nb[3] = {Rb[6], Rb[7], Rb[8]};  // world z in base coords
  const T gacc = T(-SC.gravity_z);
  const T A0[3] = {gacc * nb[0], gacc * nb[1], gacc * nb[2]};  // fictitious base acceleration
  const T zero3[3] = {T(0), T(0), T(0)};
  T wxv[3];
  cross3(wb, vb, wxv);

  // composite inertia of the whole robot and Newton-Euler base force, trunk first
  SpI<T> tot;
  tot.m = M.trunk_m;
#pragma unroll
  for (int i = 0; i < 3; i++) tot.h[i] = M.trunk_h[i];
#pragma unroll
  for (int i = 0; i < 6; i++) tot.I[i] = M.trunk_I[i];
  T fb[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  body_force(tot, wb, vb, zero3, A0, fb, fb + 3);

  T S6[21];
#pragma unroll
  for (int i = 0; i < 21; i++) S6[i] = T(0);

  // per-leg results kept for the later phases
  T Bm[4][18];  // M_kk^-1 F_k            (3x6)
  T ev[4][3];   // M_kk^-1 (tau - h) + B_lin (w x v)
  T G[4][18];   // contact rows G, later Y = L^-1 G^T   (3 dirs x 6)
  T H[4][6];    // J_kk M_kk^-1 J_kk^T (sym 3x3: nn n1 n2 11 12 22)
  T W[4][9];    // M_kk^-1 J_kk^T, W[j*3+dir]
  T cvel[4][3]; // J nu of the part that does not depend on the base solve
  T gap[4];
  int active = 0, invalid = 0;

  const T tdir[3][3] = {{nb[0], nb[1], nb[2]}, {-Rb[3], -Rb[4], -Rb[5]}, {Rb[0], Rb[1], Rb[2]}};

#pragma unroll
  for (int k = 0; k < 4; k++) {
    const T* q = st.q + 3 * k;
    const T* qd = st.qd + 3 * k;
    LegKin<T> K;
    leg_kin(k, q, M, K);
    // link rotations (link -> base)
    const T RH[9] = {T(1), T(0), T(0), T(0), K.c1, -K.s1, T(0), K.s1, K.c1};
    const T RT[9] = {K.c2, T(0), K.s2, K.s1 * K.s2, K.c1, -K.s1 * K.c2, -K.c1 * K.s2, K.s1, K.c1 * K.c2};
    const T RC[9] = {K.c23, T(0), K.s23, K.s1 * K.s23, K.c1, -K.s1 * K.c23, -K.c1 * K.s23, K.s1, K.c1 * K.c23};
    SpI<T> Ih, It, Ic;
    body_spi(M.body_m[k][0], M.body_com[k][0], M.body_Ic[k][0], RH, K.r1, Ih);
    body_spi(M.body_m[k][1], M.body_com[k][1], M.body_Ic[k][1], RT, K.r2, It);
    body_spi(M.body_m[k][2], M.body_com[k][2], M.body_Ic[k][2], RC, K.r3, Ic);

    // motion subspaces S_j = (a_j, r_j x a_j)
    const T a1[3] = {T(1), T(0), T(0)};
    T S1v[3], S2v[3], S3v[3];
    cross3(K.r1, a1, S1v);
    cross3(K.r2, K.a2, S2v);
    cross3(K.r3, K.a2, S3v);

    // ---- Newton-Euler bias (velocity products + gravity), individual bodies
    T m1w[3], m1v[3], m2w[3], m2v[3], m3w[3], m3v[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      m1w[i] = a1[i] * qd[0]; m1v[i] = S1v[i] * qd[0];
      m2w[i] = K.a2[i] * qd[1]; m2v[i] = S2v[i] * qd[1];
      m3w[i] = K.a2[i] * qd[2]; m3v[i] = S3v[i] * qd[2];
    }
    T V1w[3], V1v[3], A1w[3], A1v[3];
    cross3(wb, m1w, A1w);
    cross3(wb, m1v, A1v);
    cross3_add(vb, m1w, A1v);
#pragma unroll
    for (int i = 0; i < 3; i++) { V1w[i] = wb[i] + m1w[i]; V1v[i] = vb[i] + m1v[i]; A1v[i] += A0[i]; }
    T V2w[3], V2v[3], A2w[3], A2v[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { A2w[i] = A1w[i]; A2v[i] = A1v[i]; }
    cross3_add(V1w, m2w, A2w);
    cross3_add(V1w, m2v, A2v);
    cross3_add(V1v, m2w, A2v);
#pragma unroll
    for (int i = 0; i < 3; i++) { V2w[i] = V1w[i] + m2w[i]; V2v[i] = V1v[i] + m2v[i]; }
    T V3w[3], V3v[3], A3w[3], A3v[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { A3w[i] = A2w[i]; A3v[i] = A2v[i]; }
    cross3_add(V2w, m3w, A3w);
    cross3_add(V2w, m3v, A3v);
    cross3_add(V2v, m3w, A3v);
#pragma unroll
    for (int i = 0; i < 3; i++) { V3w[i] = V2w[i] + m3w[i]; V3v[i] = V2v[i] + m3v[i]; }

    T f3n[3] = {T(0), T(0), T(0)}, f3l[3] = {T(0), T(0), T(0)};
    body_force(Ic, V3w, V3v, A3w, A3v, f3n, f3l);
    const T h3 = dot3(K.a2, f3n) + dot3(S3v, f3l);
    body_force(It, V2w, V2v, A2w, A2v, f3n, f3l);  // now thigh + calf
    const T h2 = dot3(K.a2, f3n) + dot3(S2v, f3l);
    body_force(Ih, V1w, V1v, A1w, A1v, f3n, f3l);  // now the whole leg
    const T h1 = f3n[0] + dot3(S1v, f3l);

    // ---- composite inertias and joint-space blocks
    spi_add(It, Ic);  // thigh + calf
    spi_add(Ih, It);  // whole leg
    T F1[6], F2[6], F3[6];
    spi_apply(Ic, K.a2, S3v, F3, F3 + 3);
    spi_apply(It, K.a2, S2v, F2, F2 + 3);
    spi_apply(Ih, a1, S1v, F1, F1 + 3);
    const T M33 = dot3(K.a2, F3) + dot3(S3v, F3 + 3);
    const T M23 = dot3(K.a2, F3) + dot3(S2v, F3 + 3);
    const T M13 = F3[0] + dot3(S1v, F3 + 3);
    const T M22 = dot3(K.a2, F2) + dot3(S2v, F2 + 3);
    const T M12 = F2[0] + dot3(S1v, F2 + 3);
    const T M11 = F1[0] + dot3(S1v, F1 + 3);
    // inverse of the symmetric 3x3 (adjugate)
    const T c00 = M22 * M33 - M23 * M23, c01 = M13 * M23 - M12 * M33, c02 = M12 * M23 - M13 * M22;
    const T c11 = M11 * M33 - M13 * M13, c12 = M12 * M13 - M11 * M23, c22 = M11 * M22 - M12 * M12;
    const T idet = T(1) / (M11 * c00 + M12 * c01 + M13 * c02);
    const T Mi[6] = {c00 * idet, c01 * idet, c02 * idet, c11 * idet, c12 * idet, c22 * idet};  // 00 01 02 11 12 22
    const T t1 = tau[3 * k] - h1, t2 = tau[3 * k + 1] - h2, t3 = tau[3 * k + 2] - h3;
    const T d0 = Mi[0] * t1 + Mi[1] * t2 + Mi[2] * t3;
    const T d1 = Mi[1] * t1 + Mi[3] * t2 + Mi[4] * t3;
    const T d2 = Mi[2] * t1 + Mi[4] * t2 + Mi[5] * t3;
#pragma unroll
    for (int c = 0; c < 6; c++) {
      Bm[k][c] = Mi[0] * F1[c] + Mi[1] * F2[c] + Mi[2] * F3[c];
      Bm[k][6 + c] = Mi[1] * F1[c] + Mi[3] * F2[c] + Mi[4] * F3[c];
      Bm[k][12 + c] = Mi[2] * F1[c] + Mi[4] * F2[c] + Mi[5] * F3[c];
    }
    ev[k][0] = d0 + Bm[k][3] * wxv[0] + Bm[k][4] * wxv[1] + Bm[k][5] * wxv[2];
    ev[k][1] = d1 + Bm[k][9] * wxv[0] + Bm[k][10] * wxv[1] + Bm[k][11] * wxv[2];
    ev[k][2] = d2 + Bm[k][15] * wxv[0] + Bm[k][16] * wxv[1] + Bm[k][17] * wxv[2];
    // Schur complement and base right-hand side
#pragma unroll
    for (int a = 0; a < 6; a++) {
#pragma unroll
      for (int b = a; b < 6; b++)
        S6[s6(a, b)] -= F1[a] * Bm[k][b] + F2[a] * Bm[k][6 + b] + F3[a] * Bm[k][12 + b];
      fb[a] += (a < 3 ? f3n[a] : f3l[a - 3]) + F1[a] * d0 + F2[a] * d1 + F3[a] * d2;
    }
    spi_add(tot, Ih);

    // ---- collision detection on the poses at the start of the tick
    gap[k] = st.pos[2] + dot3(nb, K.r4) - M.foot_radius;
    if (gap[k] < M.foot_thresh) {
      active |= 1 << k;
      const T pc[3] = {K.r4[0] - M.foot_radius * nb[0], K.r4[1] - M.foot_radius * nb[1],
                       K.r4[2] - M.foot_radius * nb[2]};
      T Jk[3][3];
#pragma unroll
      for (int dd = 0; dd < 3; dd++) {
        T Jb[6];
        foot_jac_dir(K, pc, tdir[dd], Jb, Jk[dd]);
#pragma unroll
        for (int c = 0; c < 6; c++)
          G[k][6 * dd + c] = Jb[c] - (Jk[dd][0] * Bm[k][c] + Jk[dd][1] * Bm[k][6 + c] + Jk[dd][2] * Bm[k][12 + c]);
        W[k][0 * 3 + dd] = Mi[0] * Jk[dd][0] + Mi[1] * Jk[dd][1] + Mi[2] * Jk[dd][2];
        W[k][1 * 3 + dd] = Mi[1] * Jk[dd][0] + Mi[3] * Jk[dd][1] + Mi[4] * Jk[dd][2];
        W[k][2 * 3 + dd] = Mi[2] * Jk[dd][0] + Mi[4] * Jk[dd][1] + Mi[5] * Jk[dd][2];
        cvel[k][dd] = Jb[0] * wb[0] + Jb[1] * wb[1] + Jb[2] * wb[2] + Jb[3] * vb[0] + Jb[4] * vb[1] + Jb[5] * vb[2] +
                      Jk[dd][0] * (qd[0] + dt * ev[k][0]) + Jk[dd][1] * (qd[1] + dt * ev[k][1]) +
                      Jk[dd][2] * (qd[2] + dt * ev[k][2]);
      }
#define QS_H(a, b) (Jk[a][0] * W[k][0 * 3 + b] + Jk[a][1] * W[k][1 * 3 + b] + Jk[a][2] * W[k][2 * 3 + b])
      H[k][0] = QS_H(0, 0); H[k][1] = QS_H(0, 1); H[k][2] = QS_H(0, 2);
      H[k][3] = QS_H(1, 1); H[k][4] = QS_H(1, 2); H[k][5] = QS_H(2, 2);
#undef QS_H
    }
    if (detect_invalid) {
      // non-foot shapes vs the plane: support-function distance below the link's
      // contact breaking threshold (quadruped.py:243-249 -> invalid contact)
      const T ch[3] = {K.r1[0], K.r1[1], K.r1[2]};
      const T nz = dot3(nb, K.a2);
      const T zh = st.pos[2] + dot3(nb, ch) - (abs_t(nz) * M.hip_hl + M.hip_r * sqrt_t(tmax(T(1) - nz * nz, T(0))));
      invalid += zh < M.hip_thresh;
      T c[3], zc;
      m3_v(RT, M.thigh_c, c);
      zc = st.pos[2] + dot3(nb, K.r2) + dot3(nb, c);
      zc -= abs_t(nb[0] * RT[0] + nb[1] * RT[3] + nb[2] * RT[6]) * M.thigh_half[0] +
            abs_t(nb[0] * RT[1] + nb[1] * RT[4] + nb[2] * RT[7]) * M.thigh_half[1] +
            abs_t(nb[0] * RT[2] + nb[1] * RT[5] + nb[2] * RT[8]) * M.thigh_half[2];
      invalid += zc < M.thigh_thresh;
      m3_v(RC, M.calf_c, c);
      zc = st.pos[2] + dot3(nb, K.r3) + dot3(nb, c);
      zc -= abs_t(nb[0] * RC[0] + nb[1] * RC[3] + nb[2] * RC[6]) * M.calf_half[0] +
            abs_t(nb[0] * RC[1] + nb[1] * RC[4] + nb[2] * RC[7]) * M.calf_half[1] +
            abs_t(nb[0] * RC[2] + nb[1] * RC[5] + nb[2] * RC[8]) * M.calf_half[2];
      invalid += zc < M.calf_thresh;
    }
  }
  if (detect_invalid) {
    T zt = st.pos[2] - (abs_t(nb[0]) * M.trunk_half[0] + abs_t(nb[1]) * M.trunk_half[1] + abs_t(nb[2]) * M.trunk_half[2]);
    invalid += zt < M.trunk_thresh;
    T zi = st.pos[2] + dot3(nb, M.imu_pos) - (abs_t(nb[0]) + abs_t(nb[1]) + abs_t(nb[2])) * M.imu_half;
    invalid += zi < M.imu_thresh;
  }

  // ---- base: S = M_bb - sum F^T B, Cholesky, solve
  {
    const T* I = tot.I;
    const T* h = tot.h;
    S6[s6(0, 0)] += I[0]; S6[s6(0, 1)] += I[1]; S6[s6(0, 2)] += I[2];
    S6[s6(1, 1)] += I[3]; S6[s6(1, 2)] += I[4]; S6[s6(2, 2)] += I[5];
    // upper-right block [h]x
    S6[s6(0, 4)] += -h[2]; S6[s6(0, 5)] += h[1];
    S6[s6(1, 3)] += h[2];  S6[s6(1, 5)] += -h[0];
    S6[s6(2, 3)] += -h[1]; S6[s6(2, 4)] += h[0];
    S6[s6(3, 3)] += tot.m; S6[s6(4, 4)] += tot.m; S6[s6(5, 5)] += tot.m;
  }
  T Ld[6];  // reciprocal diagonal of L; S6 now holds L (L(i,j), j<=i at s6(j,i))
#pragma unroll
  for (int j = 0; j < 6; j++) {
    T dj = S6[s6(j, j)];
#pragma unroll
    for (int k2 = 0; k2 < j; k2++) dj -= S6[s6(k2, j)] * S6[s6(k2, j)];
    const T inv = rsqrt_t(dj);
    Ld[j] = inv;
    S6[s6(j, j)] = dj * inv;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      T v = S6[s6(j, i)];
#pragma unroll
      for (int k2 = 0; k2 < j; k2++) v -= S6[s6(k2, i)] * S6[s6(k2, j)];
      S6[s6(j, i)] = v * inv;
    }
  }
  T ab[6];
  {
    T y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      T v = -fb[i];
#pragma unroll
      for (int k2 = 0; k2 < i; k2++) v -= S6[s6(k2, i)] * y[k2];
      y[i] = v * Ld[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
      T v = y[i];
#pragma unroll
      for (int k2 = i + 1; k2 < 6; k2++) v -= S6[s6(i, k2)] * ab[k2];
      ab[i] = v * Ld[i];
    }
  }
  // classical acceleration of the base origin: spatial + w x v
  ab[3] += wxv[0]; ab[4] += wxv[1]; ab[5] += wxv[2];

  // ---- v += dt a, clamped like btMultiBody::applyDeltaVeeMultiDof
  bool base_clamped = false;
  T wb1[3], vb1[3];
  {
    T aw[3], av[3];
    m3_v(Rb, ab, aw);
    m3_v(Rb, ab + 3, av);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const T w1 = st.vang[i] + dt * aw[i], v1 = st.vlin[i] + dt * av[i];
      st.vang[i] = clamp_vel(w1, mcv);
      st.vlin[i] = clamp_vel(v1, mcv);
      base_clamped |= (st.vang[i] != w1) | (st.vlin[i] != v1);
    }
    m3t_v(Rb, st.vang, wb1);
    m3t_v(Rb, st.vlin, vb1);
  }
  int leg_clamped = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      T acc = ev[k][j];
#pragma unroll
      for (int c = 0; c < 6; c++) acc -= Bm[k][6 * j + c] * ab[c];
      const T v1 = st.qd[3 * k + j] + dt * acc;
      st.qd[3 * k + j] = clamp_vel(v1, mcv);
      if (st.qd[3 * k + j] != v1) leg_clamped |= 1 << k;
    }
  }

  // ---- contact rows: right-hand sides, Y = L^-1 G^T, diagonal
  T rhs[4][3], dinv[4][3], lam[4][3];
  T z[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  const T db[6] = {wb1[0] - wb[0], wb1[1] - wb[1], wb1[2] - wb[2], vb1[0] - vb[0], vb1[1] - vb[1], vb1[2] - vb[2]};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    lam[k][0] = lam[k][1] = lam[k][2] = T(0);
    if (!(active & (1 << k))) continue;
    T rel[3];
    if (base_clamped || (leg_clamped & (1 << k))) {
      // rare: a velocity clamp fired, evaluate J nu* directly
      LegKin<T> K;
      leg_kin(k, st.q + 3 * k, M, K);
      const T pc[3] = {K.r4[0] - M.foot_radius * nb[0], K.r4[1] - M.foot_radius * nb[1],
                       K.r4[2] - M.foot_radius * nb[2]};
#pragma unroll
      for (int dd = 0; dd < 3; dd++) {
        T Jb[6], Jk[3];
        foot_jac_dir(K, pc, tdir[dd], Jb, Jk);
        rel[dd] = Jb[0] * wb1[0] + Jb[1] * wb1[1] + Jb[2] * wb1[2] + Jb[3] * vb1[0] + Jb[4] * vb1[1] + Jb[5] * vb1[2] +
                  Jk[0] * st.qd[3 * k] + Jk[1] * st.qd[3 * k + 1] + Jk[2] * st.qd[3 * k + 2];
      }
    } else {
#pragma unroll
      for (int dd = 0; dd < 3; dd++) {
        T r = cvel[k][dd];
#pragma unroll
        for (int c = 0; c < 6; c++) r += G[k][6 * dd + c] * db[c];
        rel[dd] = r;
      }
    }
    const T dist = gap[k] + T(SC.linear_slop);
    T pos_err = T(0), vel_err = -rel[0];
    if (dist > T(0)) vel_err -= dist / dt; else pos_err = -dist * T(SC.contact_erp) / dt;
    rhs[k][0] = pos_err + vel_err;
    rhs[k][1] = -rel[1];
    rhs[k][2] = -rel[2];
    const T Hd[3] = {H[k][0], H[k][3], H[k][5]};
#pragma unroll
    for (int dd = 0; dd < 3; dd++) {
      T* g = G[k] + 6 * dd;
      T nn = T(0);
#pragma unroll
      for (int i = 0; i < 6; i++) {
        T v = g[i];
#pragma unroll
        for (int k2 = 0; k2 < i; k2++) v -= S6[s6(k2, i)] * g[k2];
        g[i] = v * Ld[i];
        nn += g[i] * g[i];
      }
      dinv[k][dd] = T(1) / (nn + Hd[dd]);
    }
    // warm start of the normal impulse (Bullet m_warmstartingFactor)
    if (cs.mask & (1 << k)) {
      const T imp = cs.lam_n[k] * T(SC.warmstart);
      lam[k][0] = imp;
#pragma unroll
      for (int i = 0; i < 6; i++) z[i] += G[k][i] * imp;
    }
  }

  // ---- projected Gauss-Seidel (rows: normals of all feet, then friction cones)
  if (active) {
    const T thr = T(SC.residual_threshold);
    const int iters = SC.num_iterations;
    const int nact = (active & 1) + ((active >> 1) & 1) + ((active >> 2) & 1) + ((active >> 3) & 1);
    cs.work_contacts += nact;
    for (int it = 0; it < iters; it++) {
      T res = T(0);
      cs.work_row_iters += nact;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (!(active & (1 << k))) continue;
        const T* Y = G[k];
        T w = H[k][0] * lam[k][0] + H[k][1] * lam[k][1] + H[k][2] * lam[k][2];
#pragma unroll
        for (int i = 0; i < 6; i++) w += Y[i] * z[i];
        T dI = (rhs[k][0] - w) * dinv[k][0];
        const T sum = lam[k][0] + dI;
        if (sum < T(0)) { dI = -lam[k][0]; lam[k][0] = T(0); } else lam[k][0] = sum;
#pragma unroll
        for (int i = 0; i < 6; i++) z[i] += Y[i] * dI;
        const T dv = dI / dinv[k][0];
        res = tmax(res, dv * dv);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (!(active & (1 << k))) continue;
        const T* Ya = G[k] + 6;
        const T* Yb = G[k] + 12;
        T wa = H[k][1] * lam[k][0] + H[k][3] * lam[k][1] + H[k][4] * lam[k][2];
        T wbb = H[k][2] * lam[k][0] + H[k][4] * lam[k][1] + H[k][5] * lam[k][2];
#pragma unroll
        for (int i = 0; i < 6; i++) { wa += Ya[i] * z[i]; wbb += Yb[i] * z[i]; }
        T sa = lam[k][1] + (rhs[k][1] - wa) * dinv[k][1];
        T sb = lam[k][2] + (rhs[k][2] - wbb) * dinv[k][2];
        const T lim = mu * T(SC.mu_link) * lam[k][0];
        const T r2 = sa * sa + sb * sb;
        if (r2 >= lim * lim) {
          const T sc = r2 > T(0) ? lim * rsqrt_t(r2) : T(0);
          sa *= sc; sb *= sc;
        }
        const T dIa = sa - lam[k][1], dIb = sb - lam[k][2];
        lam[k][1] = sa; lam[k][2] = sb;
#pragma unroll
        for (int i = 0; i < 6; i++) z[i] += Ya[i] * dIa + Yb[i] * dIb;
        const T ra = dIa / dinv[k][1], rb = dIb / dinv[k][2];
        res = tmax(res, ra * ra + rb * rb);
      }
      if (res <= thr) break;
    }
    // ---- apply: base twist change L^-T z, joint change W lam - B dnu_b
    T dnu[6];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
      T v = z[i];
#pragma unroll
      for (int k2 = i + 1; k2 < 6; k2++) v -= S6[s6(i, k2)] * dnu[k2];
      dnu[i] = v * Ld[i];
    }
    T dw[3], dv[3];
    m3_v(Rb, dnu, dw);
    m3_v(Rb, dnu + 3, dv);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      st.vang[i] = clamp_vel(st.vang[i] + dw[i], mcv);
      st.vlin[i] = clamp_vel(st.vlin[i] + dv[i], mcv);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        T acc = T(0);
#pragma unroll
        for (int c = 0; c < 6; c++) acc -= Bm[k][6 * j + c] * dnu[c];
        if (active & (1 << k)) acc += W[k][3 * j] * lam[k][0] + W[k][3 * j + 1] * lam[k][1] + W[k][3 * j + 2] * lam[k][2];
        st.qd[3 * k + j] = clamp_vel(st.qd[3 * k + j] + acc, mcv);
      }
    }
  }
  cs.mask = active;
  cs.invalid = invalid;
#pragma unroll
  for (int k = 0; k < 4; k++) cs.lam_n[k] = lam[k][0];

  // ---- integrate positions with the new velocities (btMultiBody::stepPositionsMultiDof)
#pragma unroll
  for (int i = 0; i < 3; i++) st.pos[i] += dt * st.vlin[i];
  {
    const T* om = st.vang;
    T ang = sqrt_t(dot3(om, om));
    if (ang * dt > T(0.25 * QS_PI)) ang = T(0.25 * QS_PI) / dt;
    T sc, cw;
    if (ang < T(0.001)) {
      sc = T(0.5) * dt - dt * dt * dt * T(0.020833333333) * ang * ang;
      T sdummy;
      sincos_t(T(0.5) * ang * dt, &sdummy, &cw);
    } else {
      T sn;
      sincos_t(T(0.5) * ang * dt, &sn, &cw);
      sc = sn / ang;
    }
    const T ax = om[0] * sc, ay = om[1] * sc, az = om[2] * sc;
    const T* q = st.quat;
    const T nw = cw * q[3] - ax * q[0] - ay * q[1] - az * q[2];
    const T nx = cw * q[0] + ax * q[3] + ay * q[2] - az * q[1];
    const T ny = cw * q[1] - ax * q[2] + ay * q[3] + az * q[0];
    const T nz = cw * q[2] + ax * q[1] - ay * q[0] + az * q[3];
    const T inv = rsqrt_t(nx * nx + ny * ny + nz * nz + nw * nw);
    st.quat[0] = nx * inv; st.quat[1] = ny * inv; st.quat[2] = nz * inv; st.quat[3] = nw * inv;
  }
#pragma unroll
  for (int i = 0; i < 12; i++) st.q[i] += dt * st.qd[i];
}

}  // namespace qs
