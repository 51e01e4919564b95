// Junction diode.  Reference: devices/diode.py eval_cqs :275-307.
// in: vd.  out: [id] (+ [qd] when F_DIODE_QD).
#pragma once
#include "junction.cuh"

template <int P, class S>
CDN_HD void diode_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    const PV jtn = p.sub(P_DIODE_jtn_t_is);
    const Dual<P> vd = v[0];
    Dual<P> iD = junction_id(jtn, vd);
    Dual<P> qD(0.0);
    if (flags & F_DIODE_CJ0) qD = junction_qd(jtn, vd);
    if (flags & F_DIODE_TT) qD = qD + p[P_DIODE_tt] * iD;
    if (flags & F_DIODE_BV)
        iD = iD - p[P_DIODE_ibv] * safe_exp(-(vd + p[P_DIODE_bv]) / p[P_DIODE_n] / p[P_DIODE_vt]);
    const double area = p[P_DIODE_area];
    out.put(0, iD * area);
    if (flags & F_DIODE_QD) out.put(1, qD * area);
}
