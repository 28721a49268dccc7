// oracle/cpu_dual.h — forward-mode derivatives for the CPU port.  TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.
//
// Written for oracle/bump_cpu.cpp only, so that the CPU checker shares no source with the CUDA library it checks
// (round 1 included the product's csrc/bump_dual.cuh).  A value with N partial derivatives; every rule below is
// the textbook one (sum, product, quotient, chain), spelled out per operation.
#ifndef ORACLE_CPU_DUAL_H_
#define ORACLE_CPU_DUAL_H_

#include <array>
#include <cmath>

namespace cpuad {

template <int N>
struct Fwd {
    double v = 0.0;              // value
    std::array<double, N> d{};   // partial derivatives (zero-initialised)

    Fwd() = default;
    Fwd(double x) : v(x) {}      // a constant
    static Fwd seed(double x, int k) {   // independent variable number k
        Fwd r(x);
        r.d[k] = 1.0;
        return r;
    }
    // f(a) with known f'(a.v)
    Fwd through(double fv, double dfda) const {
        Fwd r(fv);
        for (int i = 0; i < N; ++i) r.d[i] = dfda * d[i];
        return r;
    }
};

template <int N>
inline Fwd<N> operator+(const Fwd<N>& a, const Fwd<N>& b) {
    Fwd<N> r(a.v + b.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int N>
inline Fwd<N> operator-(const Fwd<N>& a, const Fwd<N>& b) {
    Fwd<N> r(a.v - b.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int N>
inline Fwd<N> operator-(const Fwd<N>& a) {
    return a.through(-a.v, -1.0);
}
template <int N>
inline Fwd<N> operator*(const Fwd<N>& a, const Fwd<N>& b) {
    Fwd<N> r(a.v * b.v);
    for (int i = 0; i < N; ++i) r.d[i] = b.v * a.d[i] + a.v * b.d[i];
    return r;
}
template <int N>
inline Fwd<N> operator/(const Fwd<N>& a, const Fwd<N>& b) {
    const double q = a.v / b.v;
    Fwd<N> r(q);
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - q * b.d[i]) / b.v;
    return r;
}
// mixed with plain doubles
template <int N> inline Fwd<N> operator+(const Fwd<N>& a, double s) { return a.through(a.v + s, 1.0); }
template <int N> inline Fwd<N> operator+(double s, const Fwd<N>& a) { return a.through(a.v + s, 1.0); }
template <int N> inline Fwd<N> operator-(const Fwd<N>& a, double s) { return a.through(a.v - s, 1.0); }
template <int N> inline Fwd<N> operator-(double s, const Fwd<N>& a) { return a.through(s - a.v, -1.0); }
template <int N> inline Fwd<N> operator*(const Fwd<N>& a, double s) { return a.through(a.v * s, s); }
template <int N> inline Fwd<N> operator*(double s, const Fwd<N>& a) { return a.through(a.v * s, s); }
template <int N> inline Fwd<N> operator/(const Fwd<N>& a, double s) { return a.through(a.v / s, 1.0 / s); }
template <int N> inline Fwd<N> operator/(double s, const Fwd<N>& a) { return a.through(s / a.v, -s / (a.v * a.v)); }

template <int N> inline Fwd<N> log(const Fwd<N>& a) { return a.through(std::log(a.v), 1.0 / a.v); }
template <int N> inline Fwd<N> log1p(const Fwd<N>& a) { return a.through(std::log1p(a.v), 1.0 / (1.0 + a.v)); }
template <int N> inline Fwd<N> exp(const Fwd<N>& a) { const double e = std::exp(a.v); return a.through(e, e); }
template <int N> inline Fwd<N> sqrt(const Fwd<N>& a) { const double s = std::sqrt(a.v); return a.through(s, 0.5 / s); }
template <int N> inline Fwd<N> square(const Fwd<N>& a) { return a.through(a.v * a.v, 2.0 * a.v); }

// log(e^a + e^b): the derivative is the softmax-weighted mean of the inputs' derivatives; an input at -inf has
// weight exactly zero (jnp.logaddexp's JVP), and logaddexp(-inf, -inf) = -inf with zero derivative.
template <int N>
inline Fwd<N> logaddexp(const Fwd<N>& a, const Fwd<N>& b) {
    const double hi = std::fmax(a.v, b.v);
    if (hi == -INFINITY) return Fwd<N>(-INFINITY);
    const double ea = std::exp(a.v - hi), eb = std::exp(b.v - hi), tot = ea + eb;
    Fwd<N> r(hi + std::log(tot));
    const double pa = ea / tot, pb = eb / tot;
    for (int i = 0; i < N; ++i) r.d[i] = (pa > 0.0 ? pa * a.d[i] : 0.0) + (pb > 0.0 ? pb * b.d[i] : 0.0);
    return r;
}

}  // namespace cpuad
#endif  // ORACLE_CPU_DUAL_H_
