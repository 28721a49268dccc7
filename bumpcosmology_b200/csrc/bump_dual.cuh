// Forward-mode dual numbers for the prologue kernels (theta-dependent tables and their tangents).
// The streaming kernel does NOT use these: its per-sample gradient is hand-derived (bump_stream.cuh).
#pragma once
#include <math.h>

namespace bump {

template <int N>
struct Dual {
    double v;
    double d[N];

    __host__ __device__ Dual() {}
    __host__ __device__ Dual(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = 0.0;
    }
    __host__ __device__ static Dual var(double x, int k) {
        Dual r(x);
        r.d[k] = 1.0;
        return r;
    }
};

#define BUMP_DUAL_BIN(op, expr_v, expr_d)                                                  \
    template <int N>                                                                        \
    __host__ __device__ inline Dual<N> operator op(const Dual<N>& a, const Dual<N>& b) {    \
        Dual<N> r;                                                                          \
        r.v = expr_v;                                                                       \
        _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = expr_d;                      \
        return r;                                                                           \
    }
BUMP_DUAL_BIN(+, a.v + b.v, a.d[i] + b.d[i])
BUMP_DUAL_BIN(-, a.v - b.v, a.d[i] - b.d[i])
BUMP_DUAL_BIN(*, a.v* b.v, a.d[i] * b.v + a.v * b.d[i])
#undef BUMP_DUAL_BIN

template <int N>
__host__ __device__ inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    double ib = 1.0 / b.v;
    r.v = a.v * ib;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
template <int N>
__host__ __device__ inline Dual<N> operator-(const Dual<N>& a) {
    Dual<N> r;
    r.v = -a.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
    return r;
}
// scalar mixes
template <int N>
__host__ __device__ inline Dual<N> operator+(const Dual<N>& a, double s) { Dual<N> r = a; r.v += s; return r; }
template <int N>
__host__ __device__ inline Dual<N> operator+(double s, const Dual<N>& a) { return a + s; }
template <int N>
__host__ __device__ inline Dual<N> operator-(const Dual<N>& a, double s) { Dual<N> r = a; r.v -= s; return r; }
template <int N>
__host__ __device__ inline Dual<N> operator-(double s, const Dual<N>& a) { return (-a) + s; }
template <int N>
__host__ __device__ inline Dual<N> operator*(const Dual<N>& a, double s) {
    Dual<N> r;
    r.v = a.v * s;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s;
    return r;
}
template <int N>
__host__ __device__ inline Dual<N> operator*(double s, const Dual<N>& a) { return a * s; }
template <int N>
__host__ __device__ inline Dual<N> operator/(const Dual<N>& a, double s) { return a * (1.0 / s); }
template <int N>
__host__ __device__ inline Dual<N> operator/(double s, const Dual<N>& a) { return Dual<N>(s) / a; }

template <int N>
__host__ __device__ inline Dual<N> chain(const Dual<N>& a, double fv, double fprime) {
    Dual<N> r;
    r.v = fv;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * fprime;
    return r;
}
template <int N>
__host__ __device__ inline Dual<N> dlog(const Dual<N>& a) { return chain(a, log(a.v), 1.0 / a.v); }
template <int N>
__host__ __device__ inline Dual<N> dexp(const Dual<N>& a) { double e = exp(a.v); return chain(a, e, e); }
template <int N>
__host__ __device__ inline Dual<N> dsqrt(const Dual<N>& a) { double s = sqrt(a.v); return chain(a, s, 0.5 / s); }
template <int N>
__host__ __device__ inline Dual<N> dlog1p(const Dual<N>& a) { return chain(a, log1p(a.v), 1.0 / (1.0 + a.v)); }
template <int N>
__host__ __device__ inline Dual<N> dsquare(const Dual<N>& a) { return a * a; }
template <int N>
__host__ __device__ inline Dual<N> dselect(bool c, const Dual<N>& a, const Dual<N>& b) { return c ? a : b; }

// logaddexp with JAX/torch JVP semantics: tangent = sum softmax * tangent; -inf inputs carry zero weight.
template <int N>
__host__ __device__ inline Dual<N> dlogaddexp(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    double m = fmax(a.v, b.v);
    if (m == -INFINITY) {
        r = Dual<N>(-INFINITY);
        return r;
    }
    double ea = exp(a.v - m), eb = exp(b.v - m);
    double s = ea + eb;
    r.v = m + log(s);
    double wa = ea / s, wb = eb / s;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (wa != 0.0 ? wa * a.d[i] : 0.0) + (wb != 0.0 ? wb * b.d[i] : 0.0);
    return r;
}

}  // namespace bump
