// tick_host_check.cu -- the fast tick (csrc/qs_physics.cuh) instantiated on the HOST in fp64 and fp32 and run over
// (a) 40 random airborne / touching states for 20 ticks under random torques and (b) four 400-tick standing sequences
// (PD law, four feet down, sliding and spinning starts: warm starts, friction cones, many sweeps); prints every state.
// Built twice by tests/test_tick_host.py -- with the per-leg passes on PAIRS of legs (default, qs_packed.cuh) and with
// -DQS_PACK_LEGS=0 (one leg at a time, the plain scalar templates) -- the two outputs must agree to fp64 rounding: the
// pair instantiation is the same arithmetic.  No GPU needed: the pair type falls back to two scalar operations per
// half on the host.
//   nvcc -std=c++17 -w [-DQS_PACK_LEGS=0] -o /tmp/tick_host_check tools/tick_host_check.cu && /tmp/tick_host_check out64.txt out32.txt
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../quadruped_springs_b200/csrc/qs_physics.cuh"
#include "../quadruped_springs_b200/csrc/qs_model_host.h"
using namespace qs;

template <typename T> static void dump(FILE* f, int rc, const EnvState<T>& st, const ContactState<T>& cs) {
  std::fprintf(f, "%d %d %d", rc, cs.mask, cs.invalid);
  for (int i = 0; i < 3; i++) std::fprintf(f, " %.17g", double(st.pos[i]));
  for (int i = 0; i < 4; i++) std::fprintf(f, " %.17g", double(st.quat[i]));
  for (int i = 0; i < 3; i++) std::fprintf(f, " %.17g %.17g", double(st.vlin[i]), double(st.vang[i]));
  for (int i = 0; i < 12; i++) std::fprintf(f, " %.17g %.17g", double(st.q[i]), double(st.qd[i]));
  for (int i = 0; i < 4; i++) std::fprintf(f, " %.17g", double(cs.lam_n[i]));
  std::fprintf(f, " %d %d\n", cs.work_contacts, cs.work_row_iters);
}

template <typename T> static void run(const char* out) {
  ModelConstT<T> M;
  host::build_model<T>(M, 0.02);
  ModelLegPairsT<T> M2;
  make_leg_pairs(M, M2);
  SolverConst SC;
  std::memset(&SC, 0, sizeof SC);
  SC.dt = 1e-3f; SC.gravity_z = -9.8f; SC.contact_erp = 0.08f; SC.limit_erp = 0.2f; SC.linear_slop = 1e-5f;
  SC.warmstart = 0.1f; SC.residual_threshold = 1e-7f; SC.max_coord_vel = 30.1f; SC.mu_link = 1.f;
  SC.num_iterations = 30; SC.enable_limits = 1; SC.body_response = 1;
  FILE* f = std::fopen(out, "w");
  if (!f) { std::perror(out); std::exit(1); }
  static T scratch[QS_TICK_SCRATCH];
  std::srand(1);
  auto r = []() { return double(std::rand()) / RAND_MAX * 2 - 1; };
  for (int trial = 0; trial < 40; trial++) {
    EnvState<T> st; ContactState<T> cs;
    std::memset(&st, 0, sizeof st); std::memset(&cs, 0, sizeof cs);
    double qq[4] = {r() * 0.2, r() * 0.2, r() * 0.2, 1.0};
    const double n = std::sqrt(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
    for (int i = 0; i < 4; i++) st.quat[i] = T(qq[i] / n);
    st.pos[2] = T(trial % 2 ? 0.30 + 0.02 * r() : 0.6);
    for (int k = 0; k < 4; k++) { st.q[3 * k] = T(0.2 * r()); st.q[3 * k + 1] = T(0.8 + 0.3 * r()); st.q[3 * k + 2] = T(-1.6 + 0.3 * r()); }
    for (int i = 0; i < 12; i++) st.qd[i] = T(2 * r());
    for (int i = 0; i < 3; i++) { st.vlin[i] = T(0.5 * r()); st.vang[i] = T(r()); }
    T tau[12];
    for (int i = 0; i < 12; i++) tau[i] = T(10 * r());
    for (int t = 0; t < 20; t++) {
      const int rc = physics_tick<T>(st, tau, T(0.8), cs, M, SC, true, Scratch<T>{scratch, 1}, EnvModelRef{nullptr, 0, 0}, &M2);
      dump(f, rc, st, cs);
      if (rc) break;
    }
  }
  for (int trial = 0; trial < 4; trial++) {
    EnvState<T> st; ContactState<T> cs;
    std::memset(&st, 0, sizeof st); std::memset(&cs, 0, sizeof cs);
    st.quat[3] = T(1); st.pos[2] = T(0.32);
    const double q0[3] = {0.0, 0.785, -1.57};
    for (int k = 0; k < 4; k++) for (int j = 0; j < 3; j++) st.q[3 * k + j] = T(q0[j] + 0.05 * trial * (j == 1));
    st.vlin[0] = T(0.3 * trial); st.vang[2] = T(0.5 * trial);
    for (int t = 0; t < 400; t++) {
      T tau[12];
      for (int i = 0; i < 12; i++) tau[i] = T(75.0) * (T(q0[i % 3]) - st.q[i]) - T(1.0) * st.qd[i];
      const int rc = physics_tick<T>(st, tau, T(0.6), cs, M, SC, true, Scratch<T>{scratch, 1}, EnvModelRef{nullptr, 0, 0}, &M2);
      if (t % 10 == 0 || rc) dump(f, rc, st, cs);
      if (rc) break;
    }
  }
  std::fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s out64.txt out32.txt\n", argv[0]); return 2; }
  run<double>(argv[1]);
  run<float>(argv[2]);
  std::printf("tick_host_check ok (QS_PACK_LEGS=%d)\n", int(QS_PACK_LEGS));
  return 0;
}
