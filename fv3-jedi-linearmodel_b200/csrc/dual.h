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
