/*
 * qso_physics.c -- CPU ORACLE (test infrastructure, not the product).
 *
 * Double-precision restatement of what the reference obtains from
 * pybullet.stepSimulation() (quadruped_gym_env.py:219): one 1 ms step of
 * btMultiBodyDynamicsWorld for the Go1 URDF (go1/go1_description/urdf/go1.urdf)
 * standing on pybullet_data/plane.urdf.
 *
 * PARITY UNPINNED: pybullet==3.2.5 (setup.py:8) is a third-party wheel that is
 * neither vendored under /root/reference nor installable in this container, and
 * the reference holds no golden vectors at that boundary.  The
 * algorithm below follows Bullet's published structure:
 *   - model: 19 links kept un-merged (fixed joints are 0-DoF links), inertia
 *     tensors recomputed from collision geometry because the reference loads the
 *     URDF without URDF_USE_INERTIA_FROM_FILE (quadruped.py:534-539);
 *   - forward dynamics: Featherstone articulated-body recursion over link-local
 *     frames (btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof);
 *   - per-row impulse responses M^-1 J^T through the cached articulated
 *     inertias (btMultiBody::calcAccelerationDeltasMultiDof);
 *   - velocity-space projected Gauss-Seidel in Bullet's row order: joint limits,
 *     contact normals, then implicit-cone friction pairs
 *     (btMultiBodyConstraintSolver::solveSingleIteration);
 *   - velocity clamp in applyDeltaVeeMultiDof, semi-implicit Euler with an
 *     exponential-map quaternion update (btMultiBody::stepPositionsMultiDof).
 * The CUDA product uses a different formulation (merged 13-body model, composite
 * inertias + per-leg Schur complement, reduced-space PGS) so agreement between
 * the two is a real check.
 */
#include "qso.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NL QSO_NLINKS
#define ND QSO_NDOF
#define MAXROWS (12 + 3 * QSO_MAX_CONTACTS)
#define QSO_MAX_SELF 64
#define MAXPTS 8

enum { SH_NONE = 0, SH_BOX, SH_CYL_Y, SH_SPHERE };
enum { JT_ROOT = -1, JT_FIXED = 0, JT_REVOLUTE = 1 };

typedef struct {
  int parent;
  int jtype;
  double axis[3];
  double jxyz[3];
  double mass, com[3], idiag[3];
  int dof;
  double lower, upper;
  int shape;
  double sdim[3];
  double sxyz[3];
  double thresh;
  int is_foot;
} Link;

typedef struct {
  int link, pt;
  double pos[3];   /* world point on the robot shape */
  double dist;
  double lambda_n; /* applied normal impulse of this step */
  int constrained;
} Contact;

typedef struct {
  double J[ND], MinvJ[ND];
  double dinv, rhs, lo, hi, applied, friction;
  int contact; /* index into contacts or -1 */
} Row;

struct QsoWorld {
  Link L[NL];
  QsoWorldParams P;
  double payload_m, payload_pos[3]; /* block rigidly attached to the trunk (quadruped.py:778-819) */
  /* state */
  double pos[3], quat[4], vlin[3], vang[3], q[12], qd[12];
  double tau[12];
  /* kinematics cache */
  double Rw[NL][9], pw[NL][3];
  double Xup[NL][36], S[NL][6], I6[NL][36];
  /* ABA factor cache */
  double U[NL][6], d[NL], IA0[36];
  int factored;
  /* contacts */
  Contact C[QSO_MAX_CONTACTS];
  int nC;
  /* self contacts of the last collision phase: link pairs (oracle indices) and their distance; detection only */
  int selfA[QSO_MAX_SELF], selfB[QSO_MAX_SELF];
  double selfD[QSO_MAX_SELF];
  int nSelf;
  double prev_lambda[NL][MAXPTS];
  int prev_valid[NL][MAXPTS];
  Row rows[MAXROWS];
  int last_iters;
  int cone_clamped;
  int last_nlim, last_nnorm;
  double max_fric_ratio; /* running max of |f_t| / (lambda_n) demanded before the cone projection */
};

/* ------------------------------------------------------------------ utils */
static void m3_mul(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j];
      T[3 * i + j] = s;
    }
  memcpy(C, T, sizeof T);
}
static void m3_v(const double* A, const double* v, double* o) {
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
static void m3t_v(const double* A, const double* v, double* o) {
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
static void cross(const double* a, const double* b, double* o) {
  double t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
  o[0] = t0; o[1] = t1; o[2] = t2;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void skew(const double* v, double* M) {
  M[0] = 0; M[1] = -v[2]; M[2] = v[1];
  M[3] = v[2]; M[4] = 0; M[5] = -v[0];
  M[6] = -v[1]; M[7] = v[0]; M[8] = 0;
}
static void quat_to_R(const double* q, double* R) { /* xyzw, local->world */
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double n = x * x + y * y + z * z + w * w, s = 2.0 / n;
  R[0] = 1 - s * (y * y + z * z); R[1] = s * (x * y - w * z); R[2] = s * (x * z + w * y);
  R[3] = s * (x * y + w * z); R[4] = 1 - s * (x * x + z * z); R[5] = s * (y * z - w * x);
  R[6] = s * (x * z - w * y); R[7] = s * (y * z + w * x); R[8] = 1 - s * (x * x + y * y);
}
static void axis_rot(const double* a, double th, double* R) {
  double c = cos(th), s = sin(th), t = 1 - c;
  R[0] = t * a[0] * a[0] + c; R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
  R[3] = t * a[0] * a[1] + s * a[2]; R[4] = t * a[1] * a[1] + c; R[5] = t * a[1] * a[2] - s * a[0];
  R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = t * a[2] * a[2] + c;
}
static void m6_v(const double* A, const double* v, double* o) {
  double t[6];
  for (int i = 0; i < 6; i++) {
    double s = 0;
    for (int k = 0; k < 6; k++) s += A[6 * i + k] * v[k];
    t[i] = s;
  }
  memcpy(o, t, sizeof t);
}
static void m6t_v(const double* A, const double* v, double* o) {
  double t[6];
  for (int i = 0; i < 6; i++) {
    double s = 0;
    for (int k = 0; k < 6; k++) s += A[6 * k + i] * v[k];
    t[i] = s;
  }
  memcpy(o, t, sizeof t);
}
/* spatial cross products: v = [w; u] */
static void crm_v(const double* v, const double* m, double* o) { /* v x m (motion) */
  double a[3], b[3], c[3];
  cross(v, m, a);
  cross(v + 3, m, b);
  cross(v, m + 3, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
static void crf_v(const double* v, const double* f, double* o) { /* v x* f (force) */
  double a[3], b[3], c[3];
  cross(v, f, a);
  cross(v + 3, f + 3, b);
  cross(v, f + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
/* Pluecker motion transform A->B; B origin at r (A coords), E rotates A coords into B coords */
static void xform(const double* E, const double* r, double* X) {
  double rx[9], Erx[9];
  skew(r, rx);
  m3_mul(E, rx, Erx);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      X[6 * i + j] = E[3 * i + j];
      X[6 * i + j + 3] = 0;
      X[6 * (i + 3) + j] = -Erx[3 * i + j];
      X[6 * (i + 3) + j + 3] = E[3 * i + j];
    }
}
/* C += X^T A X */
static void xtax_add(const double* X, const double* A, double* C) {
  double T[36];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += A[6 * i + k] * X[6 * k + j];
      T[6 * i + j] = s;
    }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += X[6 * k + i] * T[6 * k + j];
      C[6 * i + j] += s;
    }
}
/* solve A x = b, A 6x6 SPD (Gaussian elimination with partial pivoting) */
static void solve6(const double* A, const double* b, double* x) {
  double M[6][7];
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < 6; j++) M[i][j] = A[6 * i + j];
    M[i][6] = b[i];
  }
  for (int c = 0; c < 6; c++) {
    int p = c;
    for (int r = c + 1; r < 6; r++)
      if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
    if (p != c)
      for (int j = 0; j < 7; j++) { double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; }
    for (int r = c + 1; r < 6; r++) {
      double f = M[r][c] / M[c][c];
      for (int j = c; j < 7; j++) M[r][j] -= f * M[c][j];
    }
  }
  for (int i = 5; i >= 0; i--) {
    double s = M[i][6];
    for (int j = i + 1; j < 6; j++) s -= M[i][j] * x[j];
    x[i] = s / M[i][i];
  }
}

/* ------------------------------------------------------------------ model */
/* go1.urdf restated.  Line numbers refer to
 * quadruped_spring/go1/go1_description/urdf/go1.urdf. */
static void set3(double* d, double a, double b, double c) { d[0] = a; d[1] = b; d[2] = c; }

static void link_init(Link* l, int parent, int jtype, double ax, double ay, double az, double jx,
                      double jy, double jz, double mass, double cx, double cy, double cz) {
  memset(l, 0, sizeof *l);
  l->parent = parent; l->jtype = jtype;
  set3(l->axis, ax, ay, az); set3(l->jxyz, jx, jy, jz);
  l->mass = mass; set3(l->com, cx, cy, cz);
  l->dof = -1; l->shape = SH_NONE;
}

/* Bullet's inertia for a link loaded without URDF_USE_INERTIA_FROM_FILE:
 * the collision compound's axis-aligned box about the inertial frame
 * (btCompoundShape::calculateLocalInertia), or the child shape's own formula
 * when it is the only child at identity; zero when there is no collision
 * shape.  ext = full extents of that box. */
static void inertia_from_aabb(Link* l, double ex, double ey, double ez) {
  double m = l->mass / 12.0;
  l->idiag[0] = m * (ey * ey + ez * ez);
  l->idiag[1] = m * (ex * ex + ez * ez);
  l->idiag[2] = m * (ex * ex + ey * ey);
}

static void finish_shape(Link* l, double gthr) {
  /* btCollisionShape::getContactBreakingThreshold = angularMotionDisc * factor,
   * angularMotionDisc = |aabb diag|/2 + |aabb centre| in the inertial frame */
  if (l->shape == SH_NONE) { l->thresh = 0; return; }
  double h[3];
  if (l->shape == SH_BOX) { h[0] = l->sdim[0]; h[1] = l->sdim[1]; h[2] = l->sdim[2]; }
  else if (l->shape == SH_CYL_Y) { h[0] = l->sdim[0]; h[1] = l->sdim[1]; h[2] = l->sdim[0]; }
  else { h[0] = h[1] = h[2] = l->sdim[0]; }
  double c[3] = {l->sxyz[0] - l->com[0], l->sxyz[1] - l->com[1], l->sxyz[2] - l->com[2]};
  double disc = sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]) + sqrt(dot3(c, c));
  l->thresh = disc * gthr;
}

static void build_go1(QsoWorld* w) {
  Link* L = w->L;
  const double PI_6 = 1.0471975512, TH_LO = -0.663225115758, TH_HI = 2.96705972839;
  const double CA_LO = -2.72271363311, CA_HI = -0.837758040957;
  /* 0: base (go1.urdf:47-60), root, no collision -> zero inertia */
  link_init(&L[0], -1, JT_ROOT, 0, 0, 0, 0, 0, 0, 0.00001, 0, 0, 0);
  /* 1: trunk (go1.urdf:61-86) */
  link_init(&L[1], 0, JT_FIXED, 0, 0, 0, 0, 0, 0, 5.204, 0.0223, 0.0, -0.0005);
  L[1].shape = SH_BOX; set3(L[1].sdim, 0.3762 / 2, 0.0935 / 2, 0.114 / 2); set3(L[1].sxyz, 0, 0, 0);
  inertia_from_aabb(&L[1], 0.3762, 0.0935, 0.114);
  /* 2: imu_link (go1.urdf:87-111) */
  link_init(&L[2], 1, JT_FIXED, 0, 0, 0, -0.01592, -0.06659, -0.00617, 0.001, 0, 0, 0);
  L[2].shape = SH_BOX; set3(L[2].sdim, 0.0005, 0.0005, 0.0005); set3(L[2].sxyz, 0, 0, 0);
  inertia_from_aabb(&L[2], 0.001, 0.001, 0.001);
  /* legs FR FL RR RL (go1.urdf:112-241 and the three mirrored copies) */
  const double sx[4] = {1, 1, -1, -1}, sy[4] = {-1, 1, -1, 1};
  for (int k = 0; k < 4; k++) {
    int hip = 3 + 4 * k, thigh = hip + 1, calf = hip + 2, foot = hip + 3;
    link_init(&L[hip], 1, JT_REVOLUTE, 1, 0, 0, sx[k] * 0.1881, sy[k] * 0.04675, 0, 0.591,
              -sx[k] * 0.00541, -sy[k] * 0.00074, 6e-06);
    L[hip].dof = 3 * k; L[hip].lower = -PI_6; L[hip].upper = PI_6;
    L[hip].shape = SH_CYL_Y; set3(L[hip].sdim, 0.046, 0.02, 0); set3(L[hip].sxyz, 0, 0, 0);
    inertia_from_aabb(&L[hip], 0.092, 0.04, 0.092);
    link_init(&L[thigh], hip, JT_REVOLUTE, 0, 1, 0, 0, sy[k] * 0.08, 0, 0.92, -0.003468,
              -sy[k] * 0.018947, -0.032736);
    L[thigh].dof = 3 * k + 1; L[thigh].lower = TH_LO; L[thigh].upper = TH_HI;
    L[thigh].shape = SH_BOX; set3(L[thigh].sdim, 0.034 / 2, 0.0245 / 2, 0.213 / 2);
    set3(L[thigh].sxyz, 0, 0, -0.1065);
    inertia_from_aabb(&L[thigh], 0.034, 0.0245, 0.213);
    /* NB: every calf has com y = +0.001307 in the URDF (not mirrored). */
    link_init(&L[calf], thigh, JT_REVOLUTE, 0, 1, 0, 0, 0, -0.213, 0.131, 0.006286, 0.001307,
              -0.122269);
    L[calf].dof = 3 * k + 2; L[calf].lower = CA_LO; L[calf].upper = CA_HI;
    L[calf].shape = SH_BOX; set3(L[calf].sdim, 0.016 / 2, 0.016 / 2, 0.213 / 2);
    set3(L[calf].sxyz, 0, 0, -0.1065);
    inertia_from_aabb(&L[calf], 0.016, 0.016, 0.213);
    link_init(&L[foot], calf, JT_FIXED, 0, 0, 0, 0, 0, -0.213, 0.06, 0, 0, 0);
    L[foot].shape = SH_SPHERE; set3(L[foot].sdim, 0.02, 0, 0); set3(L[foot].sxyz, 0, 0, 0);
    L[foot].idiag[0] = L[foot].idiag[1] = L[foot].idiag[2] = 0.4 * 0.06 * 0.02 * 0.02;
    L[foot].is_foot = 1;
  }
  for (int i = 0; i < NL; i++) finish_shape(&L[i], w->P.breaking_threshold);
}

void qso_default_params(QsoWorldParams* p) {
  p->dt = 1e-3;
  p->num_iterations = 30;
  p->gravity_z = -9.8;
  p->mu_ground = 1.0;
  p->mu_link = 1.0;
  p->contact_erp = 0.08;
  p->limit_erp = 0.2;
  p->linear_slop = 1e-5;
  p->warmstart = 0.1;
  p->residual_threshold = 1e-7;
  p->max_coord_vel = 30.1;
  p->breaking_threshold = 0.02;
  p->self_collision = 1;           /* URDF_USE_SELF_COLLISION, quadruped.py:530-543 */
  p->enable_limits = 1;
  p->body_contact_response = 1;
}

QsoWorld* qso_world_create(void) {
  QsoWorld* w = (QsoWorld*)calloc(1, sizeof(QsoWorld));
  qso_default_params(&w->P);
  build_go1(w);
  w->quat[3] = 1.0;
  w->pos[2] = 0.32;
  for (int k = 0; k < 4; k++) { w->q[3 * k] = 0; w->q[3 * k + 1] = M_PI / 4; w->q[3 * k + 2] = -M_PI / 2; }
  return w;
}
void qso_world_destroy(QsoWorld* w) { free(w); }
void qso_world_set_params(QsoWorld* w, const QsoWorldParams* p) {
  w->P = *p;
  for (int i = 0; i < NL; i++) finish_shape(&w->L[i], w->P.breaking_threshold);
}
void qso_world_get_params(const QsoWorld* w, QsoWorldParams* p) { *p = w->P; }

void qso_world_set_state(QsoWorld* w, const double* s) {
  memcpy(w->pos, s, 3 * sizeof(double));
  memcpy(w->quat, s + 3, 4 * sizeof(double));
  memcpy(w->vlin, s + 7, 3 * sizeof(double));
  memcpy(w->vang, s + 10, 3 * sizeof(double));
  memcpy(w->q, s + 13, 12 * sizeof(double));
  memcpy(w->qd, s + 25, 12 * sizeof(double));
  memset(w->prev_valid, 0, sizeof w->prev_valid);
  memset(w->tau, 0, sizeof w->tau);
  w->nC = 0;
  w->factored = 0;
}
void qso_world_get_state(const QsoWorld* w, double* s) {
  memcpy(s, w->pos, 3 * sizeof(double));
  memcpy(s + 3, w->quat, 4 * sizeof(double));
  memcpy(s + 7, w->vlin, 3 * sizeof(double));
  memcpy(s + 10, w->vang, 3 * sizeof(double));
  memcpy(s + 13, w->q, 12 * sizeof(double));
  memcpy(s + 25, w->qd, 12 * sizeof(double));
}
void qso_world_add_torque(QsoWorld* w, const double* tau) {
  for (int i = 0; i < 12; i++) w->tau[i] += tau[i];
}
void qso_world_get_dynamics(const QsoWorld* w, int pyb, double* mass, double* idiag, double* com) {
  const Link* l = &w->L[pyb + 1];
  *mass = l->mass;
  memcpy(idiag, l->idiag, sizeof l->idiag);
  memcpy(com, l->com, sizeof l->com);
}
void qso_world_set_mass(QsoWorld* w, int pyb, double mass) {
  /* changeDynamics(mass=) keeps the inertia diagonal (Bullet only rescales it
   * when localInertiaDiagonal is passed). */
  w->L[pyb + 1].mass = mass;
  w->factored = 0;
}
void qso_world_set_payload(QsoWorld* w, double mass, const double* pos3) {
  w->payload_m = mass;
  memcpy(w->payload_pos, pos3, sizeof w->payload_pos);
  w->factored = 0;
}
int qso_world_last_iterations(const QsoWorld* w) { return w->last_iters; }
int qso_world_cone_clamped(const QsoWorld* w) { return w->cone_clamped; }
double qso_world_max_fric_ratio(QsoWorld* w, int reset) { double r = w->max_fric_ratio; if (reset) w->max_fric_ratio = 0; return r; }
int qso_world_last_rows(const QsoWorld* w, int* nlim, int* nnorm) { *nlim = w->last_nlim; *nnorm = w->last_nnorm; return w->last_nlim + 3 * w->last_nnorm; }

/* ------------------------------------------------------------- kinematics */
static void kinematics(QsoWorld* w) {
  Link* L = w->L;
  quat_to_R(w->quat, w->Rw[0]);
  memcpy(w->pw[0], w->pos, sizeof w->pos);
  for (int i = 0; i < NL; i++) {
    Link* l = &L[i];
    /* spatial inertia about the link origin, link coordinates */
    double cx[9], cxcxT[9], cxT[9];
    skew(l->com, cx);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) cxT[3 * a + b] = cx[3 * b + a];
    m3_mul(cx, cxT, cxcxT);
    double* I = w->I6[i];
    memset(I, 0, 36 * sizeof(double));
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        I[6 * a + b] = (a == b ? l->idiag[a] : 0.0) + l->mass * cxcxT[3 * a + b];
        I[6 * a + b + 3] = l->mass * cx[3 * a + b];
        I[6 * (a + 3) + b] = l->mass * cxT[3 * a + b];
      }
    for (int a = 0; a < 3; a++) I[6 * (a + 3) + a + 3] = l->mass;
    if (i == 1 && w->payload_m > 0) {
      /* _add_base_mass_offset (quadruped.py:778-819): a 0.1 m cube of mass m held to the trunk by a JOINT_FIXED
       * constraint at `pos` in the base frame.  Bullet solves that constraint with the contacts (a stiff but not
       * rigid link between two bodies); here the block is welded on: its spatial inertia is added to the trunk's. */
      const double mp = w->payload_m, ib = mp * (0.1 * 0.1) / 6.0;
      skew(w->payload_pos, cx);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) cxT[3 * a + b] = cx[3 * b + a];
      m3_mul(cx, cxT, cxcxT);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          I[6 * a + b] += (a == b ? ib : 0.0) + mp * cxcxT[3 * a + b];
          I[6 * a + b + 3] += mp * cx[3 * a + b];
          I[6 * (a + 3) + b] += mp * cxT[3 * a + b];
        }
      for (int a = 0; a < 3; a++) I[6 * (a + 3) + a + 3] += mp;
    }
    if (i == 0) continue;
    double Rj[9], E[9];
    if (l->jtype == JT_REVOLUTE) {
      axis_rot(l->axis, w->q[l->dof], Rj);
      for (int a = 0; a < 3; a++) { w->S[i][a] = l->axis[a]; w->S[i][a + 3] = 0; }
    } else {
      memset(Rj, 0, sizeof Rj); Rj[0] = Rj[4] = Rj[8] = 1;
      memset(w->S[i], 0, sizeof w->S[i]);
    }
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) E[3 * a + b] = Rj[3 * b + a];
    xform(E, l->jxyz, w->Xup[i]);
    m3_mul(w->Rw[l->parent], Rj, w->Rw[i]);
    double o[3];
    m3_v(w->Rw[l->parent], l->jxyz, o);
    for (int a = 0; a < 3; a++) w->pw[i][a] = w->pw[l->parent][a] + o[a];
  }
}

/* articulated-body inertias for the current q (depend on q only) */
static void aba_factor(QsoWorld* w) {
  static double IA[NL][36];
  memcpy(IA, w->I6, sizeof IA);
  for (int i = NL - 1; i >= 1; i--) {
    Link* l = &w->L[i];
    double Ia[36];
    memcpy(Ia, IA[i], sizeof Ia);
    if (l->jtype == JT_REVOLUTE) {
      m6_v(IA[i], w->S[i], w->U[i]);
      double d = 0;
      for (int a = 0; a < 6; a++) d += w->S[i][a] * w->U[i][a];
      w->d[i] = d;
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) Ia[6 * a + b] -= w->U[i][a] * w->U[i][b] / d;
    }
    xtax_add(w->Xup[i], Ia, IA[l->parent]);
  }
  memcpy(w->IA0, IA[0], sizeof w->IA0);
  w->factored = 1;
}

/* x = M^-1 f for a generalized force f = [n_b(3) f_b(3) (base coords, about
 * the base origin), tau(12)]; the O(n) propagation Bullet uses for impulse
 * responses (calcAccelerationDeltasMultiDof). */
static void aba_solve(QsoWorld* w, const double* f, double* x) {
  double pA[NL][6], u[NL], a[NL][6];
  memset(pA, 0, sizeof pA);
  for (int i = NL - 1; i >= 1; i--) {
    Link* l = &w->L[i];
    double pa[6];
    memcpy(pa, pA[i], sizeof pa);
    if (l->jtype == JT_REVOLUTE) {
      double sp = 0;
      for (int k = 0; k < 6; k++) sp += w->S[i][k] * pA[i][k];
      u[i] = f[6 + l->dof] - sp;
      for (int k = 0; k < 6; k++) pa[k] += w->U[i][k] * u[i] / w->d[i];
    }
    double t[6];
    m6t_v(w->Xup[i], pa, t);
    for (int k = 0; k < 6; k++) pA[l->parent][k] += t[k];
  }
  double rhs[6];
  for (int k = 0; k < 6; k++) rhs[k] = f[k] - pA[0][k];
  solve6(w->IA0, rhs, a[0]);
  for (int k = 0; k < 6; k++) x[k] = a[0][k];
  for (int i = 1; i < NL; i++) {
    Link* l = &w->L[i];
    m6_v(w->Xup[i], a[l->parent], a[i]);
    if (l->jtype == JT_REVOLUTE) {
      double ua = 0;
      for (int k = 0; k < 6; k++) ua += w->U[i][k] * a[i][k];
      double qdd = (u[i] - ua) / w->d[i];
      x[6 + l->dof] = qdd;
      for (int k = 0; k < 6; k++) a[i][k] += w->S[i][k] * qdd;
    }
  }
}

/* recursive Newton-Euler: h = M nudot + C(q,nu) nu + g(q) in generalized
 * coordinates nu = [omega_b, v_b (base coords), qd]. */
static void rnea(QsoWorld* w, const double* nu, const double* nudot, int with_gravity, double* h) {
  double v[NL][6], a[NL][6], f[NL][6];
  for (int k = 0; k < 6; k++) { v[0][k] = nu ? nu[k] : 0; a[0][k] = nudot ? nudot[k] : 0; }
  if (with_gravity) {
    double g[3] = {0, 0, w->P.gravity_z}, gb[3];
    m3t_v(w->Rw[0], g, gb);
    for (int k = 0; k < 3; k++) a[0][3 + k] -= gb[k];
  }
  for (int i = 0; i < NL; i++) {
    Link* l = &w->L[i];
    if (i > 0) {
      double vJ[6] = {0, 0, 0, 0, 0, 0}, t[6];
      m6_v(w->Xup[i], v[l->parent], v[i]);
      m6_v(w->Xup[i], a[l->parent], a[i]);
      if (l->jtype == JT_REVOLUTE) {
        double qd = nu ? nu[6 + l->dof] : 0, qdd = nudot ? nudot[6 + l->dof] : 0;
        for (int k = 0; k < 6; k++) vJ[k] = w->S[i][k] * qd;
        for (int k = 0; k < 6; k++) v[i][k] += vJ[k];
        crm_v(v[i], vJ, t);
        for (int k = 0; k < 6; k++) a[i][k] += t[k] + w->S[i][k] * qdd;
      }
    }
    double Iv[6], Ia[6], t[6];
    m6_v(w->I6[i], v[i], Iv);
    m6_v(w->I6[i], a[i], Ia);
    crf_v(v[i], Iv, t);
    for (int k = 0; k < 6; k++) f[i][k] = Ia[k] + t[k];
  }
  for (int i = NL - 1; i >= 1; i--) {
    Link* l = &w->L[i];
    if (l->jtype == JT_REVOLUTE) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += w->S[i][k] * f[i][k];
      h[6 + l->dof] = s;
    }
    double t[6];
    m6t_v(w->Xup[i], f[i], t);
    for (int k = 0; k < 6; k++) f[l->parent][k] += t[k];
  }
  for (int k = 0; k < 6; k++) h[k] = f[0][k];
}

static void gen_vel(const QsoWorld* w, double* nu) {
  m3t_v(w->Rw[0], w->vang, nu);
  m3t_v(w->Rw[0], w->vlin, nu + 3);
  memcpy(nu + 6, w->qd, 12 * sizeof(double));
}

/* ------------------------------------------------------------ test hooks */
void qso_world_mass_matrix(QsoWorld* w, double* M) {
  kinematics(w);
  for (int j = 0; j < ND; j++) {
    double e[ND], h[ND];
    memset(e, 0, sizeof e);
    e[j] = 1;
    rnea(w, NULL, e, 0, h);
    for (int i = 0; i < ND; i++) M[ND * i + j] = h[i];
  }
}
void qso_world_bias(QsoWorld* w, const double* tau12, double* nudot) {
  double nu[ND], h[ND], f[ND];
  kinematics(w);
  aba_factor(w);
  gen_vel(w, nu);
  rnea(w, nu, NULL, 1, h);
  for (int k = 0; k < 6; k++) f[k] = -h[k];
  for (int k = 0; k < 12; k++) f[6 + k] = (tau12 ? tau12[k] : 0) - h[6 + k];
  aba_solve(w, f, nudot);
}
void qso_world_link_pose(QsoWorld* w, int link, double* R9, double* p3) {
  kinematics(w);
  memcpy(R9, w->Rw[link], 9 * sizeof(double));
  memcpy(p3, w->pw[link], 3 * sizeof(double));
}
double qso_world_energy(QsoWorld* w, double* lin_mom, double* ang_mom) {
  /* direct per-link sum in world coordinates (independent of the spatial code) */
  kinematics(w);
  double E = 0, P[3] = {0, 0, 0}, Lm[3] = {0, 0, 0};
  double wl[NL][3], vl[NL][3]; /* world angular velocity and velocity of link origin */
  memcpy(wl[0], w->vang, sizeof wl[0]);
  memcpy(vl[0], w->vlin, sizeof vl[0]);
  for (int i = 0; i < NL; i++) {
    Link* l = &w->L[i];
    if (i > 0) {
      int p = l->parent;
      double r[3], t[3];
      for (int k = 0; k < 3; k++) r[k] = w->pw[i][k] - w->pw[p][k];
      cross(wl[p], r, t);
      for (int k = 0; k < 3; k++) { vl[i][k] = vl[p][k] + t[k]; wl[i][k] = wl[p][k]; }
      if (l->jtype == JT_REVOLUTE) {
        double aw[3];
        m3_v(w->Rw[i], l->axis, aw);
        for (int k = 0; k < 3; k++) wl[i][k] += aw[k] * w->qd[l->dof];
      }
    }
    double cw[3], vc[3], t[3], wloc[3], Iw[3];
    m3_v(w->Rw[i], l->com, cw);
    cross(wl[i], cw, t);
    for (int k = 0; k < 3; k++) vc[k] = vl[i][k] + t[k];
    m3t_v(w->Rw[i], wl[i], wloc);
    for (int k = 0; k < 3; k++) Iw[k] = l->idiag[k] * wloc[k];
    E += 0.5 * l->mass * dot3(vc, vc) + 0.5 * dot3(wloc, Iw);
    E += -l->mass * w->P.gravity_z * (w->pw[i][2] + cw[2]);
    double Lw[3], pc[3], rxp[3];
    m3_v(w->Rw[i], Iw, Lw);
    for (int k = 0; k < 3; k++) { pc[k] = w->pw[i][k] + cw[k]; t[k] = l->mass * vc[k]; }
    cross(pc, t, rxp);
    for (int k = 0; k < 3; k++) { P[k] += t[k]; Lm[k] += Lw[k] + rxp[k]; }
  }
  if (lin_mom) memcpy(lin_mom, P, sizeof P);
  if (ang_mom) memcpy(ang_mom, Lm, sizeof Lm);
  return E;
}

/* -------------------------------------------------------------- collision */
static void add_contact(QsoWorld* w, int link, int pt, const double* p, double dist) {
  if (w->nC >= QSO_MAX_CONTACTS) return;
  Contact* c = &w->C[w->nC++];
  c->link = link; c->pt = pt; c->dist = dist; c->lambda_n = 0; c->constrained = 0;
  memcpy(c->pos, p, sizeof c->pos);
}

/* convex-vs-plane (btConvexPlaneCollisionAlgorithm): ONE manifold point per
 * shape and frame, the support vertex, while it is closer than the link's
 * contact breaking threshold.  (Bullet's persistent manifold would also keep up
 * to three older points of a box; every task of the reference ends the episode
 * at the first non-foot contact, so that refinement is left out.) */
static void collide(QsoWorld* w) {
  w->nC = 0;
  for (int i = 1; i < NL; i++) {
    Link* l = &w->L[i];
    if (l->shape == SH_NONE) continue;
    double c[3], t[3];
    m3_v(w->Rw[i], l->sxyz, t);
    for (int k = 0; k < 3; k++) c[k] = w->pw[i][k] + t[k];
    if (l->shape == SH_SPHERE) {
      double d = c[2] - l->sdim[0];
      if (d < l->thresh) {
        double p[3] = {c[0], c[1], d};
        add_contact(w, i, 0, p, d);
      }
    } else if (l->shape == SH_BOX) {
      /* support vertex in -z: corner with sign_i = -sign(R[2][i]) */
      double loc[3], p[3];
      for (int a = 0; a < 3; a++) loc[a] = (w->Rw[i][6 + a] > 0 ? -1.0 : 1.0) * l->sdim[a];
      m3_v(w->Rw[i], loc, t);
      for (int k = 0; k < 3; k++) p[k] = c[k] + t[k];
      if (p[2] < l->thresh) add_contact(w, i, 0, p, p[2]);
    } else { /* cylinder, axis = link y: lowest point of the lower rim */
      double a[3] = {w->Rw[i][1], w->Rw[i][4], w->Rw[i][7]};
      double dn[3] = {-a[2] * a[0], -a[2] * a[1], 1 - a[2] * a[2]}; /* z - (z.a)a */
      double n = sqrt(dot3(dn, dn));
      double rad[3];
      if (n > 1e-9) { for (int k = 0; k < 3; k++) rad[k] = -dn[k] / n * l->sdim[0]; }
      else { rad[0] = w->Rw[i][0] * l->sdim[0]; rad[1] = w->Rw[i][3] * l->sdim[0]; rad[2] = w->Rw[i][6] * l->sdim[0]; }
      double sgn = a[2] > 0 ? -1.0 : 1.0;
      double p[3];
      for (int k = 0; k < 3; k++) p[k] = c[k] + sgn * l->sdim[1] * a[k] + rad[k];
      if (p[2] < l->thresh) add_contact(w, i, 0, p, p[2]);
    }
  }
}

/* ------------------------------------------------------- self collision (detection only)
 * The reference loads the robot with URDF_USE_SELF_COLLISION (quadruped.py:530-543) and counts a self contact as
 * invalid only when a calf is involved (quadruped.py:237-241); every task ends the episode on an invalid contact
 * (task_base.py:146-147), so the contact RESPONSE of such a pair never outlives the control step and is left out.
 * Pairs tested per calf: trunk, imu_link, the four hips (its own hip is a grandparent: Bullet's default filter only
 * drops parent-child pairs, i.e. calf-thigh and calf-foot of the same leg), the thighs, calves and feet of the other
 * legs.  A pair is in contact while the distance between the two shapes is below the smaller of their contact
 * breaking thresholds (btPersistentManifold).  Geometry (stated approximation of Bullet's GJK on the URDF
 * primitives, the same in the CUDA kernels): the long boxes of calf and thigh are capsules -- the box's axis inset
 * by the radius, radius = half the (mean) side: calf 0.008, thigh 0.0146 --, feet and imu_link are spheres, the hip
 * is its exact cylinder and the trunk its exact box; segment-to-shape distances by a 20-step ternary search of the
 * (convex) distance along the segment. */
#define SELF_IT 20
static const double CALF_R = 0.008, THIGH_R = 0.0146, LINK_LEN = 0.213;

static double seg_seg_dist(const double* p0, const double* p1, const double* q0, const double* q1) {
  double d1[3], d2[3], r[3];
  for (int k = 0; k < 3; k++) { d1[k] = p1[k] - p0[k]; d2[k] = q1[k] - q0[k]; r[k] = p0[k] - q0[k]; }
  const double a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), c = dot3(d1, r), b = dot3(d1, d2);
  const double den = a * e - b * b;
  double sN = den > 1e-12 ? (b * f - c * e) / den : 0.0;
  sN = sN < 0 ? 0 : (sN > 1 ? 1 : sN);
  double tN = (b * sN + f) / e;
  if (tN < 0) { tN = 0; sN = -c / a; sN = sN < 0 ? 0 : (sN > 1 ? 1 : sN); }
  else if (tN > 1) { tN = 1; sN = (b - c) / a; sN = sN < 0 ? 0 : (sN > 1 ? 1 : sN); }
  double d[3];
  for (int k = 0; k < 3; k++) d[k] = r[k] + sN * d1[k] - tN * d2[k];
  return sqrt(dot3(d, d));
}
static double pt_seg_dist(const double* c, const double* p0, const double* p1) {
  double d[3], r[3];
  for (int k = 0; k < 3; k++) { d[k] = p1[k] - p0[k]; r[k] = c[k] - p0[k]; }
  double t = dot3(r, d) / dot3(d, d);
  t = t < 0 ? 0 : (t > 1 ? 1 : t);
  for (int k = 0; k < 3; k++) r[k] -= t * d[k];
  return sqrt(dot3(r, r));
}
/* distance of a point (box frame) to the box of half extents h; 0 inside */
static double pt_box_dist(const double* p, const double* h) {
  double s = 0;
  for (int k = 0; k < 3; k++) { double e = fabs(p[k]) - h[k]; if (e > 0) s += e * e; }
  return sqrt(s);
}
/* distance of a point to the cylinder (centre c, unit axis a, radius r, half length hl); 0 inside */
static double pt_cyl_dist(const double* p, const double* c, const double* a, double r, double hl) {
  double v[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
  const double ax = dot3(v, a);
  double rho2 = dot3(v, v) - ax * ax;
  const double er = sqrt(rho2 > 0 ? rho2 : 0) - r, ea = fabs(ax) - hl;
  return sqrt((er > 0 ? er * er : 0) + (ea > 0 ? ea * ea : 0));
}
static void self_add(QsoWorld* w, int a, int b, double d) {
  if (w->nSelf >= QSO_MAX_SELF) return;
  w->selfA[w->nSelf] = a; w->selfB[w->nSelf] = b; w->selfD[w->nSelf] = d; w->nSelf++;
}
static void link_point(const QsoWorld* w, int link, double z, double* out) {
  const double loc[3] = {0, 0, z};
  double t[3];
  m3_v(w->Rw[link], loc, t);
  for (int k = 0; k < 3; k++) out[k] = w->pw[link][k] + t[k];
}
static void collide_self(QsoWorld* w) {
  w->nSelf = 0;
  double c0[4][3], c1[4][3], t0[4][3], t1[4][3];
  for (int k = 0; k < 4; k++) {
    const int thigh = 4 + 4 * k, calf = 5 + 4 * k;
    link_point(w, calf, -CALF_R, c0[k]); link_point(w, calf, -(LINK_LEN - CALF_R), c1[k]);
    link_point(w, thigh, -THIGH_R, t0[k]); link_point(w, thigh, -(LINK_LEN - THIGH_R), t1[k]);
  }
  for (int i = 0; i < 4; i++) {
    const int calf = 5 + 4 * i;
    const double thc = w->L[calf].thresh;
#define THR(other) (w->L[other].thresh < thc ? w->L[other].thresh : thc)
    /* trunk box and hip cylinders: ternary search of the convex distance along the calf's axis */
    for (int s = 0; s < 5; s++) {
      const int other = s == 0 ? 1 : 3 + 4 * (s - 1);
      double lo = 0, hi = 1, best = 1e30;
      for (int it = 0; it <= SELF_IT; it++) {
        const double ta = lo + (hi - lo) / 3, tb = hi - (hi - lo) / 3;
        double da, db, pa[3], pb[3];
        for (int k = 0; k < 3; k++) { pa[k] = c0[i][k] + ta * (c1[i][k] - c0[i][k]); pb[k] = c0[i][k] + tb * (c1[i][k] - c0[i][k]); }
        if (s == 0) {
          double la[3], lb[3], va[3], vb[3];
          for (int k = 0; k < 3; k++) { va[k] = pa[k] - w->pw[1][k]; vb[k] = pb[k] - w->pw[1][k]; }
          m3t_v(w->Rw[1], va, la); m3t_v(w->Rw[1], vb, lb);
          da = pt_box_dist(la, w->L[1].sdim); db = pt_box_dist(lb, w->L[1].sdim);
        } else {
          const double ax[3] = {w->Rw[other][1], w->Rw[other][4], w->Rw[other][7]};
          da = pt_cyl_dist(pa, w->pw[other], ax, w->L[other].sdim[0], w->L[other].sdim[1]);
          db = pt_cyl_dist(pb, w->pw[other], ax, w->L[other].sdim[0], w->L[other].sdim[1]);
        }
        if (da < best) best = da;
        if (db < best) best = db;
        if (da <= db) hi = tb; else lo = ta;
      }
      const double d = best - CALF_R;
      if (d < THR(other)) self_add(w, calf, other, d);
    }
    { /* imu_link: a 1 mm cube, taken as a sphere of its half side */
      const double d = pt_seg_dist(w->pw[2], c0[i], c1[i]) - CALF_R - w->L[2].sdim[0];
      if (d < THR(2)) self_add(w, calf, 2, d);
    }
    for (int j = 0; j < 4; j++) {
      if (j == i) continue;
      const int thigh = 4 + 4 * j, calf2 = 5 + 4 * j, foot = 6 + 4 * j;
      double d = seg_seg_dist(c0[i], c1[i], t0[j], t1[j]) - CALF_R - THIGH_R;
      if (d < THR(thigh)) self_add(w, calf, thigh, d);
      if (j > i) {
        d = seg_seg_dist(c0[i], c1[i], c0[j], c1[j]) - 2 * CALF_R;
        if (d < THR(calf2)) self_add(w, calf, calf2, d);
      }
      d = pt_seg_dist(w->pw[foot], c0[i], c1[i]) - CALF_R - w->L[foot].sdim[0];
      if (d < THR(foot)) self_add(w, calf, foot, d);
    }
#undef THR
  }
}

int qso_world_num_self_contacts(const QsoWorld* w) { return w->nSelf; }
void qso_world_get_self_contact(const QsoWorld* w, int i, int* pyb_link_a, int* pyb_link_b, double* dist) {
  *pyb_link_a = w->selfA[i] - 1; *pyb_link_b = w->selfB[i] - 1; *dist = w->selfD[i];
}
/* the same detection on the current state without stepping (tests) */
int qso_world_detect_self_contacts(QsoWorld* w) {
  kinematics(w);
  collide_self(w);
  return w->nSelf;
}

int qso_world_num_contacts(const QsoWorld* w) { return w->nC; }
void qso_world_get_contact(const QsoWorld* w, int i, int* pyb_link, double* nf, double* dist, double* pos) {
  const Contact* c = &w->C[i];
  *pyb_link = c->link - 1;
  *nf = c->lambda_n / w->P.dt;
  *dist = c->dist;
  memcpy(pos, c->pos, sizeof c->pos);
}

/* Jacobian row of world point p on link `link`, world direction dir, in
 * generalized coordinates [omega_b, v_b (base coords), qd]. */
static void point_jacobian(const QsoWorld* w, int link, const double* p, const double* dir, double* J) {
  memset(J, 0, ND * sizeof(double));
  double r[3], rxd[3];
  for (int k = 0; k < 3; k++) r[k] = p[k] - w->pw[0][k];
  cross(r, dir, rxd);
  m3t_v(w->Rw[0], rxd, J);
  m3t_v(w->Rw[0], dir, J + 3);
  for (int i = link; i > 0; i = w->L[i].parent) {
    const Link* l = &w->L[i];
    if (l->jtype != JT_REVOLUTE) continue;
    double aw[3], t[3];
    m3_v(w->Rw[i], l->axis, aw);
    for (int k = 0; k < 3; k++) r[k] = p[k] - w->pw[i][k];
    cross(aw, r, t);
    J[6 + l->dof] = dot3(dir, t);
  }
}

static double row_setup(QsoWorld* w, Row* r, const double* nu) {
  aba_solve(w, r->J, r->MinvJ);
  double A = 0, rel = 0;
  for (int k = 0; k < ND; k++) { A += r->J[k] * r->MinvJ[k]; rel += r->J[k] * nu[k]; }
  r->dinv = 1.0 / A;
  r->applied = 0;
  return rel;
}

static double resolve_row(Row* r, double* dV) {
  double dvn = 0;
  for (int k = 0; k < ND; k++) dvn += r->J[k] * dV[k];
  double dI = r->rhs - dvn * r->dinv;
  double sum = r->applied + dI;
  if (sum < r->lo) { dI = r->lo - r->applied; r->applied = r->lo; }
  else if (sum > r->hi) { dI = r->hi - r->applied; r->applied = r->hi; }
  else r->applied = sum;
  for (int k = 0; k < ND; k++) dV[k] += r->MinvJ[k] * dI;
  return dI / r->dinv;
}

static void apply_delta(QsoWorld* w, const double* dnu, double mult) {
  /* btMultiBody::applyDeltaVeeMultiDof: add, then clamp every stored velocity
   * coordinate (world base omega/vel + joint rates) to +-maxCoordinateVelocity */
  double dw[3], dv[3];
  m3_v(w->Rw[0], dnu, dw);
  m3_v(w->Rw[0], dnu + 3, dv);
  double mx = w->P.max_coord_vel;
  for (int k = 0; k < 3; k++) {
    w->vang[k] += dw[k] * mult;
    w->vlin[k] += dv[k] * mult;
    if (w->vang[k] > mx) w->vang[k] = mx; if (w->vang[k] < -mx) w->vang[k] = -mx;
    if (w->vlin[k] > mx) w->vlin[k] = mx; if (w->vlin[k] < -mx) w->vlin[k] = -mx;
  }
  for (int k = 0; k < 12; k++) {
    w->qd[k] += dnu[6 + k] * mult;
    if (w->qd[k] > mx) w->qd[k] = mx; if (w->qd[k] < -mx) w->qd[k] = -mx;
  }
}

void qso_world_step(QsoWorld* w) {
  const QsoWorldParams* P = &w->P;
  const double dt = P->dt;
  double nu[ND], h[ND], f[ND], acc[ND];

  /* 1. collision detection on the current poses */
  kinematics(w);
  collide(w);
  if (w->P.self_collision) collide_self(w); else w->nSelf = 0;

  /* 2. unconstrained forward dynamics; v += dt * a (clamped) */
  aba_factor(w);
  gen_vel(w, nu);
  rnea(w, nu, NULL, 1, h);
  for (int k = 0; k < 6; k++) f[k] = -h[k];
  for (int k = 0; k < 12; k++) f[6 + k] = w->tau[k] - h[6 + k];
  aba_solve(w, f, acc);
  {
    /* classical base-origin acceleration = spatial + omega x v */
    double t[3];
    cross(nu, nu + 3, t);
    for (int k = 0; k < 3; k++) acc[3 + k] += t[k];
  }
  apply_delta(w, acc, dt);
  gen_vel(w, nu);

  /* 3. constraint rows */
  Row* R = w->rows;
  int nlim = 0, nnorm = 0;
  Row* lim = R;
  if (P->enable_limits) {
    for (int i = 1; i < NL; i++) {
      Link* l = &w->L[i];
      if (l->jtype != JT_REVOLUTE) continue;
      for (int side = 0; side < 2; side++) {
        double pen = side ? l->upper - w->q[l->dof] : w->q[l->dof] - l->lower;
        if (pen > 0) continue;
        Row* r = &lim[nlim++];
        memset(r->J, 0, sizeof r->J);
        r->J[6 + l->dof] = side ? -1.0 : 1.0;
        double rel = row_setup(w, r, nu);
        r->rhs = (-pen * P->limit_erp / dt - rel) * r->dinv;
        r->lo = 0; r->hi = 100.0; r->contact = -1;
      }
    }
  }
  Row* nrm = lim + nlim;
  const double n[3] = {0, 0, 1}, t1[3] = {0, -1, 0}, t2[3] = {1, 0, 0};
  for (int c = 0; c < w->nC; c++) {
    Contact* ct = &w->C[c];
    if (!w->L[ct->link].is_foot && !P->body_contact_response) continue;
    ct->constrained = 1;
    Row* r = &nrm[nnorm++];
    point_jacobian(w, ct->link, ct->pos, n, r->J);
    double rel = row_setup(w, r, nu);
    double dist = ct->dist + P->linear_slop;
    double pos_err = 0, vel_err = -rel;
    if (dist > 0) vel_err -= dist / dt; else pos_err = -dist * P->contact_erp / dt;
    r->rhs = (pos_err + vel_err) * r->dinv;
    r->lo = 0; r->hi = 1e10; r->contact = c;
    r->friction = P->mu_ground * P->mu_link;
  }
  Row* fr = nrm + nnorm;
  for (int j = 0; j < nnorm; j++) {
    Contact* ct = &w->C[nrm[j].contact];
    for (int s = 0; s < 2; s++) {
      Row* r = &fr[2 * j + s];
      point_jacobian(w, ct->link, ct->pos, s ? t2 : t1, r->J);
      double rel = row_setup(w, r, nu);
      r->rhs = -rel * r->dinv;
      r->friction = nrm[j].friction; r->contact = nrm[j].contact;
      r->lo = r->hi = 0;
    }
  }

  /* 4. projected Gauss-Seidel in velocity space */
  double dV[ND];
  memset(dV, 0, sizeof dV);
  for (int j = 0; j < nnorm; j++) { /* warm start of contact normals */
    Contact* ct = &w->C[nrm[j].contact];
    /* warm start (m_warmstartingFactor) is kept for the foot contacts only */
    if (w->L[ct->link].is_foot && w->prev_valid[ct->link][ct->pt]) {
      double imp = w->prev_lambda[ct->link][ct->pt] * P->warmstart;
      nrm[j].applied = imp;
      for (int k = 0; k < ND; k++) dV[k] += nrm[j].MinvJ[k] * imp;
    }
  }
  w->cone_clamped = 0;
  int it = 0;
  if (nlim + nnorm > 0) {
    for (it = 0; it < P->num_iterations; it++) {
      double res = 0;
      for (int j = 0; j < nlim; j++) {
        int idx = (it & 1) ? j : nlim - 1 - j;
        double dv = resolve_row(&lim[idx], dV);
        if (dv * dv > res) res = dv * dv;
      }
      for (int j = 0; j < nnorm; j++) {
        double dv = resolve_row(&nrm[j], dV);
        if (dv * dv > res) res = dv * dv;
      }
      for (int j = 0; j < nnorm; j++) {
        /* resolveConeFrictionConstraintRows: both rows from the same velocity,
         * then projection of (sumA,sumB) onto the disc of radius mu*lambda_n */
        Row *a = &fr[2 * j], *b = &fr[2 * j + 1];
        double lim_imp = a->friction * nrm[j].applied;
        double dva = 0, dvb = 0;
        for (int k = 0; k < ND; k++) { dva += a->J[k] * dV[k]; dvb += b->J[k] * dV[k]; }
        double dIa = a->rhs - dva * a->dinv, dIb = b->rhs - dvb * b->dinv;
        double sa = a->applied + dIa, sb = b->applied + dIb;
        {
          double dem = sqrt(sa * sa + sb * sb);
          double rat = nrm[j].applied > 0 ? dem / nrm[j].applied : (dem > 0 ? 1e30 : 0);
          if (rat > w->max_fric_ratio) w->max_fric_ratio = rat;
        }
        if (sa * sa + sb * sb >= lim_imp * lim_imp) {
          double rr = sqrt(sa * sa + sb * sb);
          double sc = rr > 0 ? lim_imp / rr : 0;
          sa *= sc; sb *= sc;
          w->cone_clamped = 1;
        }
        dIa = sa - a->applied; dIb = sb - b->applied;
        a->applied = sa; b->applied = sb;
        for (int k = 0; k < ND; k++) dV[k] += a->MinvJ[k] * dIa + b->MinvJ[k] * dIb;
        double ra = dIa / a->dinv, rb = dIb / b->dinv;
        if (ra * ra + rb * rb > res) res = ra * ra + rb * rb;
      }
      if (res <= P->residual_threshold || it >= P->num_iterations - 1) { it++; break; }
    }
  }
  w->last_iters = it;
  w->last_nlim = nlim;
  w->last_nnorm = nnorm;

  /* 5. apply solver delta (clamped), remember impulses for reporting/warm start */
  apply_delta(w, dV, 1.0);
  memset(w->prev_valid, 0, sizeof w->prev_valid);
  for (int j = 0; j < nnorm; j++) {
    Contact* ct = &w->C[nrm[j].contact];
    ct->lambda_n = nrm[j].applied;
    w->prev_lambda[ct->link][ct->pt] = nrm[j].applied;
    w->prev_valid[ct->link][ct->pt] = 1;
  }

  /* 6. integrate positions with the new velocities (stepPositionsMultiDof) */
  for (int k = 0; k < 3; k++) w->pos[k] += dt * w->vlin[k];
  {
    double* om = w->vang;
    double ang = sqrt(dot3(om, om));
    if (ang * dt > 0.25 * M_PI) ang = 0.25 * M_PI / dt; /* ANGULAR_MOTION_THRESHOLD */
    double ax[3], sc;
    if (ang < 0.001) sc = 0.5 * dt - dt * dt * dt * 0.020833333333 * ang * ang;
    else sc = sin(0.5 * ang * dt) / ang;
    for (int k = 0; k < 3; k++) ax[k] = om[k] * sc;
    double cw = cos(0.5 * ang * dt);
    /* q_new = dq (x) q, local->world convention, xyzw */
    double* q = w->quat;
    double nq[4];
    nq[3] = cw * q[3] - ax[0] * q[0] - ax[1] * q[1] - ax[2] * q[2];
    nq[0] = cw * q[0] + ax[0] * q[3] + ax[1] * q[2] - ax[2] * q[1];
    nq[1] = cw * q[1] - ax[0] * q[2] + ax[1] * q[3] + ax[2] * q[0];
    nq[2] = cw * q[2] + ax[0] * q[1] - ax[1] * q[0] + ax[2] * q[3];
    double nn = sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
    for (int k = 0; k < 4; k++) q[k] = nq[k] / nn;
  }
  for (int k = 0; k < 12; k++) w->q[k] += dt * w->qd[k];

  /* 7. forces and joint torques are cleared (clearMultiBodyForces) */
  memset(w->tau, 0, sizeof w->tau);
}
