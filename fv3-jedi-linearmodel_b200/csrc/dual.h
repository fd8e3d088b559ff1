// Forward-mode dual numbers used by the tangent-linear (TL) kernels and by the
// seeded evaluations of the gather-form adjoint (AD) kernels.  fp64 only.
#pragma once
#include <math.h>
#include "platform.h"

namespace fv3lm {

struct Dual {
  double v, d;
  HD Dual() : v(0.0), d(0.0) {}
  HD Dual(double v_) : v(v_), d(0.0) {}
  HD Dual(double v_, double d_) : v(v_), d(d_) {}
};

HD double val(double a) { return a; }
HD double val(const Dual& a) { return a.v; }
HD double tan_of(double) { return 0.0; }
HD double tan_of(const Dual& a) { return a.d; }

HD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
HD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
HD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
HD Dual operator/(Dual a, Dual b) {
  double r = 1.0 / b.v, q = a.v * r;
  return Dual(q, (a.d - q * b.d) * r);
}
HD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
HD Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
HD Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
HD Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
HD Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
HD Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
HD Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
HD Dual operator/(Dual a, double b) { double r = 1.0 / b; return Dual(a.v * r, a.d * r); }
HD Dual operator/(double a, Dual b) { double q = a / b.v; return Dual(q, -q * b.d / b.v); }
HD Dual& operator+=(Dual& a, Dual b) { a.v += b.v; a.d += b.d; return a; }
HD Dual& operator-=(Dual& a, Dual b) { a.v -= b.v; a.d -= b.d; return a; }
HD Dual& operator*=(Dual& a, Dual b) { a = a * b; return a; }
HD Dual& operator+=(Dual& a, double b) { a.v += b; return a; }
HD Dual& operator*=(Dual& a, double b) { a.v *= b; a.d *= b; return a; }

// math (value-branching functions differentiate the active branch, like Tapenade)
HD double m_sqrt(double a) { return sqrt(a); }
HD Dual m_sqrt(Dual a) { double s = sqrt(a.v); return Dual(s, s > 0.0 ? 0.5 * a.d / s : 0.0); }
HD double m_exp(double a) { return exp(a); }
HD Dual m_exp(Dual a) { double e = exp(a.v); return Dual(e, e * a.d); }
HD double m_log(double a) { return log(a); }
HD Dual m_log(Dual a) { return Dual(log(a.v), a.d / a.v); }
HD double m_abs(double a) { return fabs(a); }
HD Dual m_abs(Dual a) { return a.v >= 0.0 ? a : -a; }
HD double m_max(double a, double b) { return a > b ? a : b; }
HD double m_min(double a, double b) { return a < b ? a : b; }
HD Dual m_max(Dual a, Dual b) { return a.v > b.v ? a : b; }
HD Dual m_min(Dual a, Dual b) { return a.v < b.v ? a : b; }
HD Dual m_max(double a, Dual b) { return a > b.v ? Dual(a) : b; }
HD Dual m_max(Dual a, double b) { return a.v > b ? a : Dual(b); }
HD Dual m_min(double a, Dual b) { return a < b.v ? Dual(a) : b; }
HD Dual m_min(Dual a, double b) { return a.v < b ? a : Dual(b); }
// x**p with constant real exponent
HD double m_pow(double a, double p) { return pow(a, p); }
HD Dual m_pow(Dual a, double p) { double r = pow(a.v, p); return Dual(r, p * r / a.v * a.d); }

}  // namespace fv3lm

// ---------------------------------------------------------------------------------
// Multi-seed dual number: one value and M tangents.  Used by the gather-form adjoint to
// obtain, in ONE evaluation of a stage at an output cell, the derivatives with respect to
// all input fields read at the same stencil offset (engine.h, AdGroup).
// ---------------------------------------------------------------------------------
namespace fv3lm {

template <int M> struct DualN {
  double v, d[M];
  HD DualN() : v(0.0) {
#pragma unroll
    for (int m = 0; m < M; m++) d[m] = 0.0;
  }
  HD DualN(double v_) : v(v_) {
#pragma unroll
    for (int m = 0; m < M; m++) d[m] = 0.0;
  }
};
#define FV3LM_FORM for (int m = 0; m < M; m++)
template <int M> HD double val(const DualN<M>& a) { return a.v; }
template <int M> HD DualN<M> operator+(DualN<M> a, const DualN<M>& b) { a.v += b.v;
#pragma unroll
  FV3LM_FORM a.d[m] += b.d[m]; return a; }
template <int M> HD DualN<M> operator-(DualN<M> a, const DualN<M>& b) { a.v -= b.v;
#pragma unroll
  FV3LM_FORM a.d[m] -= b.d[m]; return a; }
template <int M> HD DualN<M> operator*(const DualN<M>& a, const DualN<M>& b) { DualN<M> r; r.v = a.v * b.v;
#pragma unroll
  FV3LM_FORM r.d[m] = a.d[m] * b.v + a.v * b.d[m]; return r; }
template <int M> HD DualN<M> operator/(const DualN<M>& a, const DualN<M>& b) { DualN<M> r; double rb = 1.0 / b.v, q = a.v * rb; r.v = q;
#pragma unroll
  FV3LM_FORM r.d[m] = (a.d[m] - q * b.d[m]) * rb; return r; }
template <int M> HD DualN<M> operator-(DualN<M> a) { a.v = -a.v;
#pragma unroll
  FV3LM_FORM a.d[m] = -a.d[m]; return a; }
template <int M> HD DualN<M> operator+(DualN<M> a, double b) { a.v += b; return a; }
template <int M> HD DualN<M> operator+(double a, DualN<M> b) { b.v += a; return b; }
template <int M> HD DualN<M> operator-(DualN<M> a, double b) { a.v -= b; return a; }
template <int M> HD DualN<M> operator-(double a, const DualN<M>& b) { DualN<M> r; r.v = a - b.v;
#pragma unroll
  FV3LM_FORM r.d[m] = -b.d[m]; return r; }
template <int M> HD DualN<M> operator*(DualN<M> a, double b) { a.v *= b;
#pragma unroll
  FV3LM_FORM a.d[m] *= b; return a; }
template <int M> HD DualN<M> operator*(double a, DualN<M> b) { b.v *= a;
#pragma unroll
  FV3LM_FORM b.d[m] *= a; return b; }
template <int M> HD DualN<M> operator/(DualN<M> a, double b) { double r = 1.0 / b; a.v *= r;
#pragma unroll
  FV3LM_FORM a.d[m] *= r; return a; }
template <int M> HD DualN<M> operator/(double a, const DualN<M>& b) { DualN<M> r; double q = a / b.v, s = -q / b.v; r.v = q;
#pragma unroll
  FV3LM_FORM r.d[m] = s * b.d[m]; return r; }
template <int M> HD DualN<M>& operator+=(DualN<M>& a, const DualN<M>& b) { a = a + b; return a; }
template <int M> HD DualN<M>& operator-=(DualN<M>& a, const DualN<M>& b) { a = a - b; return a; }
template <int M> HD DualN<M>& operator*=(DualN<M>& a, const DualN<M>& b) { a = a * b; return a; }
template <int M> HD DualN<M>& operator+=(DualN<M>& a, double b) { a.v += b; return a; }
template <int M> HD DualN<M>& operator*=(DualN<M>& a, double b) { a = a * b; return a; }
template <int M> HD DualN<M> scaled(const DualN<M>& a, double v, double f) { DualN<M> r; r.v = v;   // value v, tangents f * a.d
#pragma unroll
  FV3LM_FORM r.d[m] = f * a.d[m]; return r; }
template <int M> HD DualN<M> m_sqrt(const DualN<M>& a) { double s = sqrt(a.v); return scaled(a, s, s > 0.0 ? 0.5 / s : 0.0); }
template <int M> HD DualN<M> m_exp(const DualN<M>& a) { double e = exp(a.v); return scaled(a, e, e); }
template <int M> HD DualN<M> m_log(const DualN<M>& a) { return scaled(a, log(a.v), 1.0 / a.v); }
template <int M> HD DualN<M> m_abs(const DualN<M>& a) { return a.v >= 0.0 ? a : -a; }
template <int M> HD DualN<M> m_max(const DualN<M>& a, const DualN<M>& b) { return a.v > b.v ? a : b; }
template <int M> HD DualN<M> m_min(const DualN<M>& a, const DualN<M>& b) { return a.v < b.v ? a : b; }
template <int M> HD DualN<M> m_max(double a, const DualN<M>& b) { return a > b.v ? DualN<M>(a) : b; }
template <int M> HD DualN<M> m_max(const DualN<M>& a, double b) { return a.v > b ? a : DualN<M>(b); }
template <int M> HD DualN<M> m_min(double a, const DualN<M>& b) { return a < b.v ? DualN<M>(a) : b; }
template <int M> HD DualN<M> m_min(const DualN<M>& a, double b) { return a.v < b ? a : DualN<M>(b); }
template <int M> HD DualN<M> m_pow(const DualN<M>& a, double p) { double r = pow(a.v, p); return scaled(a, r, p * r / a.v); }
#undef FV3LM_FORM

}  // namespace fv3lm
