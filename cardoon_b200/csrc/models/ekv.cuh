// EPFL EKV 2.6 intrinsic MOSFET.  Reference: devices/mosEKV.py f_simple
// :19-33, f_accurate :35-58 (always three Newton refinements, exp(v) below
// v = -20), eval_cqs :391-555.  Parameter-only sub-expressions are folded by
// the host packer (mosEKV.IntEKV.cuda_pack).
// in: v = [vdb, vgb, vsb].  out: i = [ids, idb, isb], q = [qd, qg, qs].
#pragma once
#include "junction.cuh"

template <int P>
CDN_HD Dual<P> ekv_f_simple(const Dual<P>& v) {
    const Dual<P> d = exp(.5 * v);
    const Dual<P> c = log(1.0 + d);
    const Dual<P> f = select(v.v + 20.0, c, d);
    return f * f;
}

template <int P>
CDN_HD Dual<P> ekv_f_accurate(const Dual<P>& v) {
    Dual<P> i = ekv_f_simple(v);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const Dual<P> k1 = sqrt(0.25 + i);
        const Dual<P> f = v - (2.0 * k1 - 1.0 + log(k1 - .5));
        const Dual<P> vprime = -1.0 / (2.0 * k1 * (0.5 - k1)) + 1.0 / k1;
        i = i + f / vprime;
    }
    return select(v.v + 20.0, i, exp(v));
}

template <int P>
CDN_HD Dual<P> ekv_interp(const Dual<P>& v, bool simple) {
    return simple ? ekv_f_simple(v) : ekv_f_accurate(v);
}

// `A`: PV or the dual record of the electro-thermal variant (thermal.cuh);
// temperature-independent entries are read through val().
template <int P, class A>
CDN_HD void ekv_intrinsic(const A& p, unsigned flags, const Dual<P>* v,
                          Dual<P>* iv, Dual<P>* qv) {
    typedef Dual<P> D;
    const bool simple = (flags & F_EKV_SIMPLE) != 0;
    const double tf = val(p[P_EKV_tf]);
    const D vd0 = tf * v[0], vg = tf * v[1], vs0 = tf * v[2];
    const double swap = ((vd0.v - vs0.v) > 0.0) ? 1.0 : -1.0;
    const D vd = select(swap, vd0, vs0);
    const D vs = select(swap, vs0, vd0);
    const auto Vt = p[P_EKV_Vt], phiT = p[P_EKV_phiT];
    const double ga = val(p[P_EKV_gammaa]), go2 = val(p[P_EKV_go2]);
    // effective gate voltage, pinch-off voltage (:409-439)
    const D vgprime = vg - p[P_EKV_vt0a] - val(p[P_EKV_deltavRSCE]) + phiT + p[P_EKV_gsqphi];
    D vp0 = vgprime - phiT - ga * (sqrt(vgprime + val(p[P_EKV_g2o4])) - go2);
    vp0 = select(vgprime, vp0, D(-phiT));
    const auto vt16 = p[P_EKV_vt16];
    const D vsprime = 0.5 * (vs + phiT + sqrt(sq(vs + phiT) + vt16));
    const D vdprime = 0.5 * (vd + phiT + sqrt(sq(vd + phiT) + vt16));
    D tmp = val(p[P_EKV_letaleff]) * (sqrt(vsprime) + sqrt(vdprime));
    tmp = tmp - val(p[P_EKV_weta3]) * sqrt(vp0 + phiT) / val(p[P_EKV_weff]);
    const D gammao = ga - val(p[P_EKV_epscox]) * tmp;
    const D gammaprime = 0.5 * (gammao + sqrt(gammao * gammao + p[P_EKV_vt01]));
    D vp = vgprime - phiT;
    vp = vp - gammaprime * (sqrt(vgprime + sq(.5 * gammaprime)) - gammaprime / 2.0);
    vp = select(vgprime, vp, D(-phiT));
    const auto vt4 = p[P_EKV_vt4];
    const D n = 1.0 + ga / (2.0 * sqrt(vp + phiT + vt4));
    // forward current, saturation voltages (:441-449)
    const D i_f = ekv_interp((vp - vs) / Vt, simple);
    const auto vc = p[P_EKV_vc];
    const D sqif = sqrt(i_f);
    D vdss = sqrt(0.25 + Vt * sqif / vc) - 0.5;
    vdss = vdss * vc;
    tmp = sqif - 0.75 * log(i_f);
    D vdssprime = sqrt(0.25 + Vt * tmp / vc) - 0.5;
    vdssprime = vdssprime * vc;
    vdssprime = vdssprime + p[P_EKV_vdsslog];
    // channel-length modulation (:450-463)
    const double Lambda = val(p[P_EKV_Lambda]), lc = val(p[P_EKV_lc]);
    const auto ucrit = p[P_EKV_t_ucrit];
    tmp = sqif - vdss / Vt;
    D deltav = sqrt(Lambda * tmp + 1.0 / 64);
    deltav = deltav * vt4;
    const D vds = .5 * (vd - vs);
    D vip = sqrt(vdss * vdss + deltav * deltav);
    vip = vip - sqrt(sq(vds - vdss) + deltav * deltav);
    D deltal = log(1.0 + (vds - vip) / lc / ucrit);
    deltal = deltal * val(p[P_EKV_lamlc]);
    const D lprime = val(p[P_EKV_nsleff]) - deltal + (vds + vip) / ucrit;
    const D leq = 0.5 * (lprime + sqrt(lprime * lprime + val(p[P_EKV_lmin2])));
    // reverse currents (:464-471)
    tmp = vp - vds - vs;
    tmp = tmp - sqrt(vdssprime * vdssprime + deltav * deltav);
    tmp = tmp + sqrt(sq(vds - vdssprime) + deltav * deltav);
    const D irprime = ekv_interp(tmp / Vt, simple);
    const D ir = ekv_interp((vp - vd) / Vt, simple);
    // quasi-static charges (:474-493)
    const D sqvpphi = sqrt(vp + phiT + 1.e-6);
    const D nq = 1.0 + go2 / sqvpphi;
    const D xf = sqrt(0.25 + i_f);
    const D xr = sqrt(0.25 + ir);
    tmp = sq(xf + xr);
    D qd = 3.0 * (xr * xr * xr) + 6.0 * xr * xr * xf + 4.0 * xr * xf * xf + 2.0 * (xf * xf * xf);
    qd = qd * (4.0 / 15.0 / tmp);
    qd = -nq * (qd - .5);
    D qs = 3.0 * (xf * xf * xf) + 6.0 * xf * xf * xr + 4.0 * xf * xr * xr + 2.0 * (xr * xr * xr);
    qs = qs * (4.0 / 15.0 / tmp);
    qs = -nq * (qs - .5);
    const D qi = qs + qd;
    D qb1 = -ga * sqvpphi / Vt;
    qb1 = qb1 - (nq - 1.0) * qi / nq;
    const D qb = select(vgprime, qb1, -vgprime / Vt);
    const D qg = -qi - qb - val(p[P_EKV_qox]);
    // mobility reduction, drain current (:495-526)
    const D betao = p[P_EKV_kpanpw] / leq;
    const D betaoprime = p[P_EKV_bop] * betao;
    tmp = dabs(qb + val(p[P_EKV_eta]) * qi);
    tmp = 1.0 + p[P_EKV_coxvt] * tmp / val(p[P_EKV_e0]) / val(p[P_EKV_epSi]);
    const D beta = betaoprime / tmp;
    const D IS = 2.0 * n * beta * Vt * Vt;
    const D ids = IS * (i_f - irprime);
    // impact ionisation (:529-533)
    const D vib = vd - vs - val(p[P_EKV_ibn2]) * vdss;
    D idb1 = ids * val(p[P_EKV_iba]) * vib / p[P_EKV_t_ibb];
    idb1 = idb1 * safe_exp(p[P_EKV_ntibblc] / vib);
    const D idb = select(vib, idb1, 0.0);
    const auto qscale = p[P_EKV_qscale];
    qv[0] = select(swap, qd, qs) * qscale;
    qv[1] = qg * qscale;
    qv[2] = select(swap, qs, qd) * qscale;
    iv[0] = (swap * ids) * tf;
    iv[1] = select(swap, idb, 0.0) * tf;
    iv[2] = select(swap, D(0.0), idb) * tf;
}
