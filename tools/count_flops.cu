// count_flops.cu -- instrumented operation count of the product's physics tick.
//
// Instantiates the SAME templated code the kernels run (csrc/qs_physics.cuh,
// csrc/qs_robot.cuh) on the host with a counting scalar type, for a robot with
// 0 / 2 / 4 feet on the ground and 1 / 2 PGS sweeps, and fits
//     W_tick = W0 + [n_c > 0] * Wany + n_c * Wc + sweeps * n_c * Wrow
// (add, sub, mul, div, sqrt, rsqrt = 1 FLOP each; an FMA is 2; sin/cos counted
// apart).  Output: quadruped_springs_b200/flop_model.json, which bench.py uses
// for roofline.achieved, and which replaces the provisional estimate of
// BASELINE.md section 3.  Build+run:  nvcc -std=c++17 -o /tmp/count_flops tools/count_flops.cu && /tmp/count_flops
#include <cmath>
#include <cstdio>
#include <cstring>

static long long g_add = 0, g_mul = 0, g_div = 0, g_sqrt = 0, g_trig = 0;

struct Cnt {
  double v;
  Cnt() = default;
  Cnt(double x) : v(x) {}
  Cnt(float x) : v(x) {}
  Cnt(int x) : v(x) {}
  explicit operator double() const { return v; }
};
inline Cnt operator+(Cnt a, Cnt b) { g_add++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { g_add++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { g_mul++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { g_div++; return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt& operator+=(Cnt& a, Cnt b) { g_add++; a.v += b.v; return a; }
inline Cnt& operator-=(Cnt& a, Cnt b) { g_add++; a.v -= b.v; return a; }
inline Cnt& operator*=(Cnt& a, Cnt b) { g_mul++; a.v *= b.v; return a; }
inline bool operator<(Cnt a, Cnt b) { return a.v < b.v; }
inline bool operator>(Cnt a, Cnt b) { return a.v > b.v; }
inline bool operator<=(Cnt a, Cnt b) { return a.v <= b.v; }
inline bool operator>=(Cnt a, Cnt b) { return a.v >= b.v; }
inline bool operator==(Cnt a, Cnt b) { return a.v == b.v; }
inline bool operator!=(Cnt a, Cnt b) { return a.v != b.v; }
inline void sincos_t(Cnt x, Cnt* s, Cnt* c) { g_trig += 2; s->v = std::sin(x.v); c->v = std::cos(x.v); }
inline Cnt sqrt_t(Cnt x) { g_sqrt++; return Cnt(std::sqrt(x.v)); }
inline Cnt rsqrt_t(Cnt x) { g_sqrt++; return Cnt(1.0 / std::sqrt(x.v)); }
inline Cnt abs_t(Cnt x) { return Cnt(std::fabs(x.v)); }
inline Cnt atan2_t(Cnt y, Cnt x) { g_trig++; return Cnt(std::atan2(y.v, x.v)); }
inline Cnt asin_t(Cnt x) { g_trig++; return Cnt(std::asin(x.v)); }
inline Cnt exp_t(Cnt x) { g_trig++; return Cnt(std::exp(x.v)); }

namespace qs {  // overloads of the hot-path helpers for the counting type (declared before the templates)
inline void sincos_tick(Cnt x, Cnt* s, Cnt* c) { sincos_t(x, s, c); }
inline Cnt div_t(Cnt a, Cnt b) { return a / b; }
inline Cnt rsqrt_pos(Cnt x) { return rsqrt_t(x); }
inline Cnt sqrt_pos(Cnt x) { return sqrt_t(x); }
inline void sinc_cos_small(Cnt x, Cnt* sinc, Cnt* c) { g_trig += 2; c->v = std::cos(x.v); sinc->v = std::fabs(x.v) < 1e-8 ? 1.0 : std::sin(x.v) / x.v; }
}  // namespace qs

#include "../quadruped_springs_b200/csrc/qs_physics.cuh"
#include "../quadruped_springs_b200/csrc/qs_model_host.h"

using namespace qs;

struct Counts { long long add, mul, div, sq, trig; long long flops() const { return add + mul + div + sq; } };

static Counts run(int n_contacts, int sweeps, bool springs, Counts* torque_only) {
  ModelConstT<Cnt> M;
  host::build_model<Cnt>(M, 0.02);
  qs_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.enable_springs = 1; cfg.action_repeat = 10; cfg.time_step = 0.001; cfg.obs_mode = QS_OBS_ARS_BASIC;
  RobotConst RC;
  host::build_robot(cfg, RC);
  SolverConst SC;
  SC.dt = 1e-3f; SC.gravity_z = -9.8f; SC.contact_erp = 0.08f; SC.limit_erp = 0.2f; SC.linear_slop = 1e-5f;
  SC.warmstart = 0.1f; SC.residual_threshold = -1.f /* never exit early */; SC.max_coord_vel = 30.1f; SC.mu_link = 1.f;
  SC.num_iterations = sweeps; SC.enable_limits = 0;
  EnvState<Cnt> st;
  ContactState<Cnt> cs;
  std::memset(&st, 0, sizeof st);
  std::memset(&cs, 0, sizeof cs);
  st.quat[3] = Cnt(1.0);
  for (int k = 0; k < 4; k++) { st.q[3 * k] = Cnt(0.01); st.q[3 * k + 1] = Cnt(0.8); st.q[3 * k + 2] = Cnt(-1.6); }
  if (n_contacts == 2) { st.q[2] = Cnt(-2.2); st.q[5] = Cnt(-2.2); }  // front feet lifted
  st.pos[2] = Cnt(n_contacts == 0 ? 1.0 : 0.3165);
  for (int i = 0; i < 12; i++) st.qd[i] = Cnt(0.3 * (i % 3) - 0.2);
  st.vlin[0] = Cnt(0.1); st.vlin[2] = Cnt(-0.05); st.vang[1] = Cnt(0.2);
  cs.mask = n_contacts == 4 ? 15 : (n_contacts == 2 ? 12 : 0);
  for (int k = 0; k < 4; k++) cs.lam_n[k] = Cnt(0.03);
  // torques: PD + PEA exactly as run_ticks evaluates them
  g_add = g_mul = g_div = g_sqrt = g_trig = 0;
  Cnt tau[12];
  const Cnt sk[3] = {Cnt(20.), Cnt(20.), Cnt(30.)}, sb[3] = {Cnt(.3), Cnt(.3), Cnt(.3)}, sr[3] = {Cnt(0.), Cnt(0.785), Cnt(-1.27)};
  for (int i = 0; i < 12; i++)
    tau[i] = pd_torque1(Cnt(75.), Cnt(1.), Cnt(RC.tau_max[i]), Cnt(RC.init_angles[i]), st.q[i], st.qd[i], false);
  if (springs)
    for (int k = 0; k < 4; k++) {
      Cnt ts[3];
      spring_torque_leg(k, sk, sb, sr, st.q + 3 * k, st.qd + 3 * k, ts);
      for (int j = 0; j < 3; j++) tau[3 * k + j] += ts[j];
    }
  if (torque_only) *torque_only = {g_add, g_mul, g_div, g_sqrt, g_trig};
  static Cnt scratch[QS_TICK_SCRATCH];
  physics_tick<Cnt>(st, tau, Cnt(0.8), cs, M, SC, false, Scratch<Cnt>{scratch, 1});
  int got = __builtin_popcount(cs.mask);
  if (got != n_contacts) std::fprintf(stderr, "WARNING: wanted %d contacts, got %d\n", n_contacts, got);
  return {g_add, g_mul, g_div, g_sqrt, g_trig};
}

int main(int argc, char** argv) {
  Counts tq;
  Counts c0 = run(0, 1, true, &tq);
  Counts c21 = run(2, 1, true, nullptr), c22 = run(2, 2, true, nullptr);
  Counts c41 = run(4, 1, true, nullptr), c42 = run(4, 2, true, nullptr);
  const double wrow = double(c42.flops() - c41.flops()) / 4.0;
  const double wrow2 = double(c22.flops() - c21.flops()) / 2.0;
  // c(n,1) = W0 + Wany + n*Wc + n*Wrow
  const double wc = (double(c41.flops() - c21.flops()) / 2.0) - wrow;
  const double wany = double(c21.flops()) - double(c0.flops()) - 2 * (wc + wrow);
  const char* out = argc > 1 ? argv[1] : "quadruped_springs_b200/flop_model.json";
  FILE* f = std::fopen(out, "w");
  if (!f) { std::perror(out); return 1; }
  std::fprintf(f,
    "{\n \"how\": \"tools/count_flops.cu: host instantiation of csrc/qs_physics.cuh with a counting scalar; add/sub/mul/div/sqrt = 1 FLOP, FMA = 2, sin/cos apart\",\n"
    " \"W0_flight_tick\": %lld,\n \"W0_breakdown\": {\"add\": %lld, \"mul\": %lld, \"div\": %lld, \"sqrt\": %lld, \"sincos\": %lld},\n"
    " \"torque_pd_pea_in_W0\": %lld,\n \"W_any_contact\": %.1f,\n \"W_per_contact\": %.1f,\n \"W_per_contact_sweep\": %.1f,\n"
    " \"W_per_contact_sweep_check_2feet\": %.1f,\n \"tick_4feet_1sweep\": %lld,\n \"tick_4feet_30sweeps\": %.0f,\n"
    " \"epilogue_per_control_step_estimate\": 400\n}\n",
    c0.flops(), c0.add, c0.mul, c0.div, c0.sq, c0.trig, tq.flops(), wany, wc, wrow, wrow2, c41.flops(),
    double(c41.flops()) + 29 * 4 * wrow);
  std::fclose(f);
  std::printf("W0=%lld Wany=%.1f Wc=%.1f Wrow=%.1f (check %.1f) trig=%lld\n", c0.flops(), wany, wc, wrow, wrow2, c0.trig);
  return 0;
}
