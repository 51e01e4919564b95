// P-N junction current / charge laws.
// Reference: devices/diode.py Junction.get_id :74-80, get_qd :83-98;
// devices/svdiode.py SVJunction.get_idvd :80-97, get_qd :99-111.
// The record layout is param_layout.JUNCTION (J_* indexes).
#pragma once
#include "../dual.cuh"
#include "../param_index.h"

// Strided view of one instance's parameter column: p[k] = base[k*ld].
struct PV {
    const double* base;
    long ld;
    CDN_HD double operator[](int k) const { return __ldg(base + (long)k * ld); }
    CDN_HD PV sub(int off) const { PV r; r.base = base + (long)off * ld; r.ld = ld; return r; }
    // streaming hint of the ring-staged accessor (dev_eval.cu PVRing): "rows
    // below `first` are not read any more"; nothing to do for on-demand loads
    CDN_HD void mark(int) {}
};

template <int P, class A>
CDN_HD Dual<P> junction_id(const A& j, const Dual<P>& vd) {
    return j[J_t_is] * (safe_exp(vd / j[J_kexp]) - 1.0);
}

// Depletion charge; callers guard on the cj0 != 0 flag (the reference returns
// 0*vd otherwise).
template <int P, class A>
CDN_HD Dual<P> junction_qd(const A& j, const Dual<P>& vd) {
    // `auto`: doubles for the packed isothermal record, duals (d/dT lane) for
    // the electro-thermal one
    const auto fc = j[J_fc], t_vj = j[J_t_vj], k5 = j[J_k5];
    const auto k6 = j[J_k6], k7 = j[J_k7], m = j[J_m];
    const double k4 = val(j[J_k4]);               // 1 - m: never temperature dependent
    const Dual<P> b = fc * t_vj - vd;
    const Dual<P> u = 1.0 - vd / t_vj;
    // (1 - vd/vj)**(1 - m): the default grading m = 1/2 is a square root
    const Dual<P> c = k5 * (1.0 - ((k4 == 0.5) ? sqrt(u) : dpow(u, k4)));
    const Dual<P> d = k6 * ((1.0 - fc * (1.0 + m)) * vd
                            + .5 * m * vd * vd / t_vj - k7)
                      + j[J_k8];          // k5 (1 - (1-fc)**k4), folded on the host
    return select(b, c, d);
}

// State-variable junction: current and voltage as functions of x.
template <int P, class A>
CDN_HD void svjunction_idvd(const A& j, const Dual<P>& x, Dual<P>& iD, Dual<P>& vD) {
    const auto t_is = j[J_t_is], alpha = j[J_alpha], svth = j[J_svth];
    const auto kexp = j[J_kexp];
    const Dual<P> b = svth - x;
    const Dual<P> c = t_is * (exp(alpha * x) - 1.0);
    const Dual<P> lin = 1.0 + alpha * (x - svth);
    const Dual<P> d = t_is * kexp * lin - t_is;
    iD = select(b, c, d);
    const Dual<P> dv = svth + log(lin) / alpha;
    vD = select(b, x, dv);
}
