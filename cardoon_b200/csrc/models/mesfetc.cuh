// Curtice cubic MESFET.  Reference: devices/mesfetc.py eval_cqs :192-226.
// in: v = [vgs, vgd, vgs(t-tau), vgd(t-tau)].
// out: i = [igs, igd, ids], q = [qgs, qgd].
#pragma once
#include "junction.cuh"

// `A`: PV or the dual record of the electro-thermal variant (thermal.cuh);
// temperature-independent entries are read through val().
template <int P, class S, class A>
CDN_HD void mesfetc_eval(const A& p, unsigned flags, const Dual<P>* v, S& out) {
    typedef Dual<P> D;
    const auto gs = p.sub(P_MESFETC_diogs_t_is), gd = p.sub(P_MESFETC_diogd_t_is);
    D igs = junction_id(gs, v[0]);
    D igd = junction_id(gd, v[1]);
    D qgs = (flags & F_MESFETC_CGS0) ? junction_qd(gs, v[0]) : 0.0 * v[0];
    D qgd = (flags & F_MESFETC_CGD0) ? junction_qd(gd, v[1]) : 0.0 * v[1];
    const double ib0 = val(p[P_MESFETC_ib0]), vbd = val(p[P_MESFETC_vbd]);
    const auto k6 = p[P_MESFETC_k6];
    igs = igs - ib0 * exp(-(v[0] + vbd) / k6);
    igd = igd - ib0 * exp(-(v[1] + vbd) / k6);
    const double swap = ((v[0].v - v[1].v) > 0.0) ? 1.0 : -1.0;
    const D vds = swap * (v[0] - v[1]);
    const D vgsi = select(swap, v[0], v[1]);
    const D vgsiT = select(swap, v[2], v[3]);
    const D vx = vgsiT * (1.0 + p[P_MESFETC_Beta] * (val(p[P_MESFETC_vds0]) - vds));
    D ids = (val(p[P_MESFETC_a0]) + vx * (val(p[P_MESFETC_a1]) + vx * (val(p[P_MESFETC_a2]) + vx * val(p[P_MESFETC_a3]))))
            * tanh(val(p[P_MESFETC_gama]) * vds) * p[P_MESFETC_idsFac];
    ids = select(vgsi - p[P_MESFETC_Vt0], ids, 0.0);
    ids = select(ids, ids, 0.0);
    const double area = val(p[P_MESFETC_area]);
    out.put(0, igs * area);
    out.put(1, igd * area);
    out.put(2, (ids * swap) * area);
    out.put(3, qgs * area);
    out.put(4, qgd * area);
}
