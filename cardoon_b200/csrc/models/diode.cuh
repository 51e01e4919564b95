// Junction diode.  Reference: devices/diode.py eval_cqs :275-307.
// in: vd.  out: [id] (+ [qd] when F_DIODE_QD).
#pragma once
#include "junction.cuh"

// `A`: PV (packed isothermal record) or DualRec (thermal.cuh)
template <int P, class A>
CDN_HD void diode_core(const A& p, unsigned flags, const Dual<P>& vd, Dual<P>& iD, Dual<P>& qD) {
    const auto jtn = p.sub(P_DIODE_jtn_t_is);
    iD = junction_id(jtn, vd);
    qD = Dual<P>(0.0);
    if (flags & F_DIODE_CJ0) qD = junction_qd(jtn, vd);
    if (flags & F_DIODE_TT) qD = qD + p[P_DIODE_tt] * iD;
    if (flags & F_DIODE_BV)
        iD = iD - p[P_DIODE_ibv] * safe_exp(-(vd + p[P_DIODE_bv]) / p[P_DIODE_n] / p[P_DIODE_vt]);
    const auto area = p[P_DIODE_area];
    iD = iD * area;
    qD = qD * area;
}

template <int P, class S>
CDN_HD void diode_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    Dual<P> iD, qD;
    diode_core(p, flags, v[0], iD, qD);
    out.put(0, iD);
    if (flags & F_DIODE_QD) out.put(1, qD);
}

// State-variable diode.  Reference: devices/svdiode.py eval_cqs :300-333.
// in: x (state variable).  out: [id, gyr*vd] (+ [qd] when F_SVDIODE_QD).
// (The reference adds tt*iD to an unassigned qD when cj0 = 0 and tt != 0 --
// a NameError there; here qD starts from zero as in diode.py.)
template <int P, class A>
CDN_HD void svdiode_core(const A& p, unsigned flags, const Dual<P>& x, Dual<P>& iD, Dual<P>& vD,
                         Dual<P>& qD) {
    const auto jtn = p.sub(P_SVDIODE_jtn_t_is);
    svjunction_idvd(jtn, x, iD, vD);
    qD = Dual<P>(0.0);
    if (flags & F_SVDIODE_CJ0) qD = junction_qd(jtn, vD);
    if (flags & F_SVDIODE_TT) qD = qD + p[P_SVDIODE_tt] * iD;
    if (flags & F_SVDIODE_BV)
        iD = iD - p[P_SVDIODE_ibv]
                  * safe_exp(-(vD + p[P_SVDIODE_bv]) / p[P_SVDIODE_n] / p[P_SVDIODE_vt]);
    const auto area = p[P_SVDIODE_area];
    iD = iD * area;
    vD = vD * val(p[P_SVDIODE_gyr]);
    qD = qD * area;
}

template <int P, class S>
CDN_HD void svdiode_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    Dual<P> iD, vD, qD;
    svdiode_core(p, flags, v[0], iD, vD, qD);
    out.put(0, iD);
    out.put(1, vD);
    if (flags & F_SVDIODE_QD) out.put(2, qD);
}
