// Forward-mode dual numbers for the device-evaluation kernels.
//
// Replaces the reference's pycppad/CppAD tapes (devices/cppaddev.py:113-157):
// a Dual<P> carries a value and P tangent lanes (P = number of control
// ports), every operation propagates exact first derivatives, so one pass
// through a model's equations yields out[] and the full Jacobian.  P = 0 is
// the plain-double instantiation used where only values are needed
// (charge update, reference cppaddev.eval :160).
//
// Semantics that the models rely on (reference cppaddev.py:56-96 and CppAD):
//   select(b, c, d)  = (b > 0) ? c : d   strict, applied to value AND tangents
//                      (the un-selected branch may be NaN/inf and must not
//                      leak; never multiply by 0/1 masks)
//   safe_exp(x)      = exp(x) for x < 50, exp(50)*(x-49) above
//   d|x|/dx          = sign(x) with sign(0) = 0
//   comparisons with NaN are false (=> "d" branch)
#pragma once
#include <math.h>

#define CDN_HD __device__ __forceinline__

template <int P>
struct Dual {
    double v;
    double d[P > 0 ? P : 1];
    CDN_HD Dual() {}
    CDN_HD Dual(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < P; ++i) d[i] = 0.0;
    }
};

// ---- construction -----------------------------------------------------------
template <int P>
CDN_HD Dual<P> make_var(double x, int lane) {
    Dual<P> r; r.v = x;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = (i == lane) ? 1.0 : 0.0;
    return r;
}

// ---- addition / subtraction ---------------------------------------------------
template <int P>
CDN_HD Dual<P> operator-(const Dual<P>& a) {
    Dual<P> r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = -a.d[i];
    return r;
}
template <int P>
CDN_HD Dual<P> operator+(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int P>
CDN_HD Dual<P> operator+(const Dual<P>& a, double b) {
    Dual<P> r = a; r.v = a.v + b; return r;
}
template <int P>
CDN_HD Dual<P> operator+(double a, const Dual<P>& b) { return b + a; }
template <int P>
CDN_HD Dual<P> operator-(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int P>
CDN_HD Dual<P> operator-(const Dual<P>& a, double b) {
    Dual<P> r = a; r.v = a.v - b; return r;
}
template <int P>
CDN_HD Dual<P> operator-(double a, const Dual<P>& b) {
    Dual<P> r; r.v = a - b.v;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = -b.d[i];
    return r;
}

// ---- multiplication ---------------------------------------------------------
template <int P>
CDN_HD Dual<P> operator*(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
template <int P>
CDN_HD Dual<P> operator*(const Dual<P>& a, double b) {
    Dual<P> r; r.v = a.v * b;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] * b;
    return r;
}
template <int P>
CDN_HD Dual<P> operator*(double a, const Dual<P>& b) { return b * a; }

// ---- division ---------------------------------------------------------------
// One reciprocal per dual division: MUFU seed + two Newton steps (~1 ulp), then
// the quotient gets the usual residual correction, which makes it the
// correctly rounded a/b in all but rare half-way cases.  Branch-free: for a
// zero, denormal, huge, infinite or NaN divisor the refinement would turn the
// seed into NaN, so the seed itself is returned -- it is the IEEE result for
// 0, inf and NaN (+-inf, +-0, NaN propagate exactly as in the host arithmetic
// the models were written for); a denormal divisor gives inf and one above
// 2^1022 gives 0, off from IEEE by less than 2^-1022 absolute.  (An IEEE
// `1.0 / x` slow path behind a branch was measured first: ptxas lays it out
// as the fall-through, so each of the ~110 divisions of a BSIM3 evaluation
// took a jump over 20 dead instructions -- 20 % of the stall samples.)
CDN_HD double cdn_rcp(double x) {
    const unsigned ex = (unsigned)__double2hiint(x) & 0x7ff00000u;
    const bool special = ex - 0x00100000u >= 0x7fc00000u;
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    double e = fma(-x, r0, 1.0);
    double r = fma(r0, e, r0);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return special ? r0 : r;
}
// Reciprocal for the dual DIVISIONS: one Newton step (relative error ~2^-46).
// cdn_div's residual correction squares that error, so the quotient is as good
// as with cdn_rcp; the tangent lanes of a quotient, which are multiplied by the
// reciprocal itself, carry ~1e-14 of relative error (Jacobian entries are
// compared at 1e-8).  Two dependent DFMAs fewer on the chain of each of the
// ~110 divisions of a BSIM3 evaluation.
CDN_HD double cdn_rcp1(double x) {
    const unsigned ex = (unsigned)__double2hiint(x) & 0x7ff00000u;
    const bool special = ex - 0x00100000u >= 0x7fc00000u;
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double e = fma(-x, r0, 1.0);
    const double r = fma(r0, e, r0);
    return special ? r0 : r;
}
// a/b given inv ~ 1/b
CDN_HD double cdn_div(double a, double b, double inv) {
    double q = a * inv;
    const double rem = fma(-b, q, a);
    if (fabs(rem) <= 1.79769313486231570815e308) q = fma(rem, inv, q);   // finite
    return q;
}
template <int P>
CDN_HD Dual<P> operator/(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r;
    const double inv = cdn_rcp1(b.v);
    r.v = cdn_div(a.v, b.v, inv);
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
template <int P>
CDN_HD Dual<P> operator/(const Dual<P>& a, double b) {
    Dual<P> r;
    const double inv = cdn_rcp1(b);
    r.v = cdn_div(a.v, b, inv);
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] * inv;
    return r;
}
template <int P>
CDN_HD Dual<P> operator/(double a, const Dual<P>& b) {
    Dual<P> r;
    const double inv = cdn_rcp1(b.v);
    r.v = cdn_div(a, b.v, inv);
    const double k = -r.v * inv;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = k * b.d[i];
    return r;
}

// ---- elementary functions -----------------------------------------------------
template <int P>
CDN_HD Dual<P> d_scale(const Dual<P>& x, double z, double dz) {
    Dual<P> r; r.v = z;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = dz * x.d[i];
    return r;
}
template <int P> CDN_HD Dual<P> exp(const Dual<P>& x) {
    const double z = ::exp(x.v); return d_scale(x, z, z);
}
template <int P> CDN_HD Dual<P> log(const Dual<P>& x) {
    return d_scale(x, ::log(x.v), cdn_rcp(x.v));
}
// sqrt(x) and its derivative 1/(2 sqrt(x)) from ONE reciprocal square root:
// MUFU seed + two Newton steps, z = x*y with a residual correction (within
// 1 ulp of the IEEE result).  Branch-free like cdn_rcp: +-0, denormals (treated
// as 0), +inf, negative and NaN arguments take the values of the host
// arithmetic through selects (sqrt: 0 / inf / NaN, derivative: inf / 0 / NaN);
// CUDA's ::sqrt carries an out-of-line slow path per call site and the
// derivative needed a second reciprocal sequence.
CDN_HD void cdn_sqrt_rsqrt(double x, double& z, double& half_y) {
    const int hi = __double2hiint(x);
    const bool normal = (unsigned)(hi - 0x00100000) < 0x7fe00000u;      // positive, normal, finite
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // y <- y (1.5 - 0.5 x y^2), twice
    const double hx = 0.5 * x;
    double t = fma(-hx * y, y, 0.5);
    y = fma(y, t, y);
    t = fma(-hx * y, y, 0.5);
    y = fma(y, t, y);
    double zz = x * y;
    const double hy = 0.5 * y;
    zz = fma(fma(-zz, zz, x), hy, zz);
    // special arguments
    const bool zero = (hi & 0x7ff00000) == 0;                          // +-0 or denormal
    const bool pinf = (hi == 0x7ff00000) && (__double2loint(x) == 0);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    z = normal ? zz : (zero ? 0.0 * x : (pinf ? x : nan));
    half_y = normal ? hy : (zero ? inf : (pinf ? 0.0 : nan));
}
template <int P> CDN_HD Dual<P> sqrt(const Dual<P>& x) {
    double z, dz;
    cdn_sqrt_rsqrt(x.v, z, dz);
    return d_scale(x, z, dz);
}
template <int P> CDN_HD Dual<P> tan(const Dual<P>& x) {
    const double z = ::tan(x.v); return d_scale(x, z, 1.0 + z * z);
}
template <int P> CDN_HD Dual<P> tanh(const Dual<P>& x) {
    const double z = ::tanh(x.v); return d_scale(x, z, 1.0 - z * z);
}
template <int P> CDN_HD Dual<P> sin(const Dual<P>& x) {
    return d_scale(x, ::sin(x.v), ::cos(x.v));
}
template <int P> CDN_HD Dual<P> cos(const Dual<P>& x) {
    return d_scale(x, ::cos(x.v), -::sin(x.v));
}
template <int P> CDN_HD Dual<P> cosh(const Dual<P>& x) {
    return d_scale(x, ::cosh(x.v), ::sinh(x.v));
}
template <int P> CDN_HD Dual<P> dabs(const Dual<P>& x) {
    const double s = (x.v > 0.0) ? 1.0 : ((x.v < 0.0) ? -1.0 : 0.0);
    return d_scale(x, ::fabs(x.v), s);
}
// x**p, constant exponent
template <int P> CDN_HD Dual<P> dpow(const Dual<P>& x, double p) {
    const double z = ::pow(x.v, p); return d_scale(x, z, p * z * cdn_rcp(x.v));
}
template <int P> CDN_HD Dual<P> sq(const Dual<P>& x) { return x * x; }
// value of a record entry that is a double (isothermal record) or a dual
// (electro-thermal record, models/thermal.cuh)
CDN_HD double val(double x) { return x; }
template <int P> CDN_HD double val(const Dual<P>& x) { return x.v; }
CDN_HD double sq(double x) { return x * x; }

// ---- selection ------------------------------------------------------------------
template <int P>
CDN_HD Dual<P> select(double b, const Dual<P>& c, const Dual<P>& d) {
    Dual<P> r;
    const bool m = b > 0.0;
    r.v = m ? c.v : d.v;
#pragma unroll
    for (int i = 0; i < P; ++i) r.d[i] = m ? c.d[i] : d.d[i];
    return r;
}
template <int P>
CDN_HD Dual<P> select(const Dual<P>& b, const Dual<P>& c, const Dual<P>& d) {
    return select(b.v, c, d);
}
template <int P>
CDN_HD Dual<P> select(const Dual<P>& b, const Dual<P>& c, double d) {
    return select(b.v, c, Dual<P>(d));
}
template <int P>
CDN_HD Dual<P> select(const Dual<P>& b, double c, const Dual<P>& d) {
    return select(b.v, Dual<P>(c), d);
}
template <int P>
CDN_HD Dual<P> select(double b, const Dual<P>& c, double d) {
    return select(b, c, Dual<P>(d));
}
template <int P>
CDN_HD Dual<P> select(double b, double c, const Dual<P>& d) {
    return select(b, Dual<P>(c), d);
}
template <int P>
CDN_HD double selectd(const Dual<P>& b, double c, double d) {
    return (b.v > 0.0) ? c : d;
}

template <int P> CDN_HD Dual<P> safe_exp(const Dual<P>& x) {
    // reference cppaddev.safe_exp: condexp_lt(x, 50, exp(x), exp(50)(x-49))
    const double e50 = 5.184705528587072e21;   // exp(50.)
    const Dual<P> c = exp(x);
    const Dual<P> d = e50 * (x - 50.0 + 1.0);
    return select(50.0 - x.v, c, d);
}
