// qs_packed.cuh -- a pair of scalars that the tick's per-leg templates are instantiated with to work on TWO LEGS at
// once: sm_100's packed FP32 instructions (FFMA2 / FMUL2 / FADD2: two multiply-adds per issue slot and lane, operand
// negation and scalar-broadcast operands for free) halve the instruction count of the per-leg dynamics, which is the
// larger half of a tick, in kernels that are bound by instruction issue.
//
// The leg templates are written as ordinary arithmetic (a * b + c * d ...) and rely on the compiler's contraction into
// FFMA for scalars.  Intrinsics are never contracted, so the pair type does it itself: a product is a lazy object
// (PkProd) that fuses with the addition or subtraction it meets;  x + a*b, a*b - x, a*b + c*d ... each become ONE fused
// multiply-add (plus one multiply for the second product), which is what nvcc makes of the scalar code.
// There are deliberately no comparisons: the callers look at the halves.
#pragma once
#include <cmath>
#include <type_traits>

#include "qs_robot.cuh"

namespace qs {

template <typename S> struct PkProd;

template <typename S> struct alignas(2 * sizeof(S)) PkT {
  S x, y;
  PkT() = default;
  template <typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type>
  QS_DEV PkT(U s) : x(S(s)), y(S(s)) {}            // broadcast: a scalar operand in SASS (R.F32), no second register
  QS_DEV PkT(S a, S b) : x(a), y(b) {}
  QS_DEV PkT(const PkProd<S>& p);

  // ---- the three machine operations
  static QS_DEV PkT fma(const PkT& a, const PkT& b, const PkT& c) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    if constexpr (std::is_same<S, float>::value) {
      const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
      return PkT(r.x, r.y);
    } else {
      return PkT(::fma(a.x, b.x, c.x), ::fma(a.y, b.y, c.y));
    }
#else
    return PkT(S(std::fma(a.x, b.x, c.x)), S(std::fma(a.y, b.y, c.y)));
#endif
  }
  static QS_DEV PkT mul(const PkT& a, const PkT& b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    if constexpr (std::is_same<S, float>::value) {
      const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
      return PkT(r.x, r.y);
    } else
#endif
      return PkT(a.x * b.x, a.y * b.y);
  }
  static QS_DEV PkT add(const PkT& a, const PkT& b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    if constexpr (std::is_same<S, float>::value) {
      const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
      return PkT(r.x, r.y);
    } else
#endif
      return PkT(a.x + b.x, a.y + b.y);
  }
  static QS_DEV PkT neg(const PkT& a) { return PkT(-a.x, -a.y); }   // folds into the consumer's operand modifier

  friend QS_DEV PkProd<S> operator*(const PkT& a, const PkT& b) { return PkProd<S>{a, b}; }
  friend QS_DEV PkT operator+(const PkT& a, const PkT& b) { return add(a, b); }
  friend QS_DEV PkT operator-(const PkT& a, const PkT& b) { return add(a, neg(b)); }
  friend QS_DEV PkT operator-(const PkT& a) { return neg(a); }
  QS_DEV PkT& operator+=(const PkT& b) { *this = add(*this, b); return *this; }
  QS_DEV PkT& operator-=(const PkT& b) { *this = add(*this, neg(b)); return *this; }
  QS_DEV PkT& operator*=(const PkT& b) { *this = mul(*this, b); return *this; }
  QS_DEV PkT& operator+=(const PkProd<S>& p);
  QS_DEV PkT& operator-=(const PkProd<S>& p);
};

template <typename S> struct PkProd {
  PkT<S> a, b;
  QS_DEV PkT<S> value() const { return PkT<S>::mul(a, b); }
  friend QS_DEV PkT<S> operator+(const PkProd& p, const PkProd& q) { return PkT<S>::fma(p.a, p.b, q.value()); }
  friend QS_DEV PkT<S> operator+(const PkProd& p, const PkT<S>& c) { return PkT<S>::fma(p.a, p.b, c); }
  friend QS_DEV PkT<S> operator+(const PkT<S>& c, const PkProd& p) { return PkT<S>::fma(p.a, p.b, c); }
  friend QS_DEV PkT<S> operator-(const PkProd& p, const PkProd& q) { return PkT<S>::fma(p.a, p.b, PkT<S>::neg(q.value())); }
  friend QS_DEV PkT<S> operator-(const PkProd& p, const PkT<S>& c) { return PkT<S>::fma(p.a, p.b, PkT<S>::neg(c)); }
  friend QS_DEV PkT<S> operator-(const PkT<S>& c, const PkProd& p) { return PkT<S>::fma(PkT<S>::neg(p.a), p.b, c); }
  friend QS_DEV PkProd operator*(const PkProd& p, const PkT<S>& c) { return PkProd{p.value(), c}; }
  friend QS_DEV PkProd operator*(const PkT<S>& c, const PkProd& p) { return PkProd{c, p.value()}; }
  friend QS_DEV PkProd operator*(const PkProd& p, const PkProd& q) { return PkProd{p.value(), q.value()}; }
  friend QS_DEV PkProd operator-(const PkProd& p) { return PkProd{PkT<S>::neg(p.a), p.b}; }
};

template <typename S> QS_DEV PkT<S>::PkT(const PkProd<S>& p) { *this = p.value(); }
template <typename S> QS_DEV PkT<S>& PkT<S>::operator+=(const PkProd<S>& p) { *this = fma(p.a, p.b, *this); return *this; }
template <typename S> QS_DEV PkT<S>& PkT<S>::operator-=(const PkProd<S>& p) { *this = fma(neg(p.a), p.b, *this); return *this; }

// the scalar helpers of qs_robot.cuh, half by half (no packed forms of these exist)
template <typename S> QS_DEV PkT<S> tmin(PkT<S> a, PkT<S> b) { return PkT<S>(tmin(a.x, b.x), tmin(a.y, b.y)); }
template <typename S> QS_DEV PkT<S> tmax(PkT<S> a, PkT<S> b) { return PkT<S>(tmax(a.x, b.x), tmax(a.y, b.y)); }
template <typename S> QS_DEV PkT<S> fmin_t(PkT<S> a, PkT<S> b) { return PkT<S>(fmin_t(a.x, b.x), fmin_t(a.y, b.y)); }
template <typename S> QS_DEV PkT<S> fmax_t(PkT<S> a, PkT<S> b) { return PkT<S>(fmax_t(a.x, b.x), fmax_t(a.y, b.y)); }
template <typename S> QS_DEV PkT<S> abs_t(PkT<S> a) { return PkT<S>(abs_t(a.x), abs_t(a.y)); }
template <typename S> QS_DEV PkT<S> sqrt_t(PkT<S> a) { return PkT<S>(sqrt_t(a.x), sqrt_t(a.y)); }
template <typename S> QS_DEV PkT<S> rsqrt_t(PkT<S> a) { return PkT<S>(rsqrt_t(a.x), rsqrt_t(a.y)); }
template <typename S> QS_DEV PkT<S> div_t(PkT<S> a, PkT<S> b) { return PkT<S>(div_t(a.x, b.x), div_t(a.y, b.y)); }
template <typename S> QS_DEV PkT<S> div_t(PkT<S> a, const PkProd<S>& b) { return div_t(a, b.value()); }
template <typename S> QS_DEV void sincos_tick(PkT<S> a, PkT<S>* s, PkT<S>* c) {
  sincos_tick(a.x, &s->x, &c->x);
  sincos_tick(a.y, &s->y, &c->y);
}

// accumulation of a per-leg contribution into a per-robot sum: both halves of a pair go into the one scalar sum
template <typename T> QS_DEV void acc_add(T& a, const T& v) { a += v; }
template <typename S> QS_DEV void acc_add(S& a, const PkT<S>& v) { a += v.x; a += v.y; }
template <typename S> QS_DEV void acc_add(S& a, const PkProd<S>& p) { const PkT<S> v = p.value(); a += v.x; a += v.y; }

// The leg tables of ModelConstT for the leg pairs (0, 1) and (2, 3): what leg_kin / leg_dynamics read when they are
// instantiated for a pair (same member names; the rolled loop indexes them by the pair, so they live in memory).
template <typename T> struct ModelLegPairsT {
  PkT<T> hip_pos[2][3], thigh_off_y[2], link_len, body_m[2][3], body_com[2][3][3], body_Ic[2][3][6];
};
template <typename T> __host__ __device__ inline void make_leg_pairs(const ModelConstT<T>& M, ModelLegPairsT<T>& P) {
  P.link_len = PkT<T>(M.link_len, M.link_len);
  for (int kp = 0; kp < 2; kp++) {
    const int k0 = 2 * kp, k1 = k0 + 1;
    P.thigh_off_y[kp] = PkT<T>(M.thigh_off_y[k0], M.thigh_off_y[k1]);
    for (int i = 0; i < 3; i++) P.hip_pos[kp][i] = PkT<T>(M.hip_pos[k0][i], M.hip_pos[k1][i]);
    for (int j = 0; j < 3; j++) {
      P.body_m[kp][j] = PkT<T>(M.body_m[k0][j], M.body_m[k1][j]);
      for (int i = 0; i < 3; i++) P.body_com[kp][j][i] = PkT<T>(M.body_com[k0][j][i], M.body_com[k1][j][i]);
      for (int i = 0; i < 6; i++) P.body_Ic[kp][j][i] = PkT<T>(M.body_Ic[k0][j][i], M.body_Ic[k1][j][i]);
    }
  }
}

template <typename T> struct is_pair : std::false_type {};
template <typename S> struct is_pair<PkT<S>> : std::true_type {};

}  // namespace qs
