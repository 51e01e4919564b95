// Electro-thermal device variants (`<dev>_t`).  Reference:
// devices/autoThermal.py eval_cqs :97-124 -- the last control port is the
// temperature rise over ambient; the wrapper calls the base device's
// set_temp_vars with that (AD-typed) temperature, evaluates the base
// eval_cqs and appends the dissipated power as one more current output.
//
// Here the temperature-dependent entries of the base model's record are
// recomputed per evaluation as duals (the T lane carries d/dT) from a raw,
// temperature-independent record appended by the host packer
// (cuda_pack_thermal), and the base model code runs unchanged on that dual
// record (the model functions are templates over the record accessor).
#pragma once
#include "diode.cuh"
#include "bjt.cuh"
#include "mesfetc.cuh"
#include "mosext.cuh"
#include "ekv.cuh"

#define CDN_CONST_K 1.3806488e-23        // globalVars.py:26-39
#define CDN_CONST_Q 1.602176565e-19
#define CDN_CONST_T0 273.15

// base-model record with dual entries
template <int P, int N>
struct DualRec {
    Dual<P> r[N];
    struct Sub {
        const DualRec* q;
        int off;
        CDN_HD const Dual<P>& operator[](int k) const { return q->r[off + k]; }
    };
    CDN_HD const Dual<P>& operator[](int k) const { return r[k]; }
    CDN_HD Sub sub(int off) const { Sub s; s.q = this; s.off = off; return s; }
    CDN_HD void mark(int) {}
};

template <int P>
struct TempState {
    Dual<P> Tabs, tnratio, vt, egap_t;
    double Tnomabs, egapn;
};

// device temperature from the thermal port (autoThermal.py:110) and the
// quantities every set_temp_vars starts from (diode.py:263-270)
template <int P>
CDN_HD TempState<P> temp_state(const Dual<P>& dT, double temp_amb, double Tnomabs, double egapn,
                               double eg0) {
    TempState<P> t;
    const Dual<P> temp = dT + temp_amb;
    t.Tabs = temp + CDN_CONST_T0;
    t.vt = CDN_CONST_K * t.Tabs / CDN_CONST_Q;
    t.egap_t = eg0 - .000702 * (t.Tabs * t.Tabs) / (t.Tabs + 1108.);
    t.tnratio = t.Tabs / Tnomabs;
    t.Tnomabs = Tnomabs;
    t.egapn = egapn;
    return t;
}

// Junction.set_temp_vars (diode.py:53-72) on the raw record `raw` (JR_*),
// written into the J_* entries of `rec` starting at `off`
// (SV: SVJunction.set_temp_vars, svdiode.py:53-78)
template <bool SV, int P, class R>
CDN_HD void junction_tempvars(const PV& raw, const TempState<P>& t, bool cap, R& rec, int off) {
    typedef Dual<P> D;
    rec.r[off + J_t_is] = raw[JR_isat] * dpow(t.tnratio, raw[JR_k1])
                          * exp(raw[JR_k2] - raw[JR_k3] / t.Tabs);
    if (SV) {
        const double n = raw[JR_n];
        const D alpha = 1. / n / t.vt;
        rec.r[off + J_alpha] = alpha;
        rec.r[off + J_svth] = log(5e10 / alpha) / alpha;
        rec.r[off + J_kexp] = n * t.vt * 5e10;
    } else {
        rec.r[off + J_kexp] = t.vt * raw[JR_n];
    }
    if (cap) {
        const double vj = raw[JR_vj], m = raw[JR_m], fc = raw[JR_fc];
        const D t_vj = vj * t.tnratio - 3. * t.vt * log(t.tnratio) - t.tnratio * t.egapn + t.egap_t;
        const D t_cj0 = raw[JR_cj0] * (1. + m * (.0004 * (t.Tabs - t.Tnomabs) + 1. - t_vj / vj));
        const D k5 = t_vj * t_cj0 / raw[JR_k4];
        rec.r[off + J_t_vj] = t_vj;
        rec.r[off + J_k5] = k5;
        rec.r[off + J_k6] = t_cj0 * raw[JR_c6];
        rec.r[off + J_k7] = (1. - fc * (1. + m)) * t_vj * fc + .5 * m * t_vj * fc * fc;
        rec.r[off + J_k8] = k5 * raw[JR_c8];
    }
}

// diode_t: in [vd, dT]; out [id, power] (+ [qd]).  diode.py:254-272 (set_temp_vars),
// :310-316 (power = vd * id).
template <int P, class S>
CDN_HD void diode_t_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    DualRec<P, NP_DIODE> rec;
#pragma unroll
    for (int k = 0; k < NP_DIODE; ++k) rec.r[k] = Dual<P>(p[k]);
    const TempState<P> t = temp_state(v[1], p[P_DIODE_T_temp_amb], p[P_DIODE_T_Tnomabs],
                                      p[P_DIODE_T_egapn], p[P_DIODE_T_eg0]);
    rec.r[P_DIODE_vt] = t.vt;
    junction_tempvars<false>(p.sub(P_DIODE_T_jraw_isat), t, (flags & F_DIODE_CJ0) != 0, rec,
                      P_DIODE_jtn_t_is);
    Dual<P> iD, qD;
    diode_core(rec, flags, v[0], iD, qD);
    out.put(0, iD);
    out.put(1, v[0] * iD);
    if (flags & F_DIODE_QD) out.put(2, qD);
}

// collects the base model's outputs so that the power can be inserted between
// currents and charges (autoThermal.py:121-123)
template <int P>
struct BufSink {
    Dual<P> o[CDN_MAX_OUT_BUF];
    CDN_HD void put(int slot, const Dual<P>& d) { o[slot] = d; }
};

template <int P>
CDN_HD Dual<P> pick(const Dual<P>* v, int idx) {       // v[idx] without local memory
    Dual<P> r = v[0];
#pragma unroll
    for (int c = 1; c < CDN_MAX_XIN; ++c)
        if (c == idx) r = v[c];
    return r;
}

// bjt_t / svbjt_t: in [vbe|x1, vbc|x2, (v_ib), (v_sub), dT]; out = the base
// model's currents, the power, then the charges.  set_temp_vars: bjt.py:453-482,
// svbjt.py:276-305, extrinsic :170-180; power: bjt.py:607-625, svbjt.py:438-458,
// extrinsic :208-222 (RE, RC not counted).
template <int P, bool SV, class S>
CDN_HD void bjt_t_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    typedef Dual<P> D;
    DualRec<P, NP_BJT> rec;
#pragma unroll
    for (int k = 0; k < NP_BJT; ++k) rec.r[k] = D(p[k]);
    const int nrb = (flags & F_BJT_RB) ? 1 : 0, nsub = (flags & F_BJT_SUB) ? 1 : 0;
    const D dT = pick(v, 2 + nrb + nsub);
    const TempState<P> t = temp_state(dT, p[P_BJT_T_temp_amb], p[P_BJT_T_Tnomabs],
                                      p[P_BJT_T_egapn], p[P_BJT_T_eg]);
    const D tnXTB = dpow(t.tnratio, p[P_BJT_T_xtb]);
    junction_tempvars<SV>(p.sub(P_BJT_T_rawjif_isat), t, (flags & F_BJT_CJE) != 0, rec,
                          P_BJT_jif_t_is);
    junction_tempvars<SV>(p.sub(P_BJT_T_rawjir_isat), t, (flags & F_BJT_CJC) != 0, rec,
                          P_BJT_jir_t_is);
    if (flags & F_BJT_ISE) {
        junction_tempvars<false>(p.sub(P_BJT_T_rawjile_isat), t, false, rec, P_BJT_jile_t_is);
        rec.r[P_BJT_jile_t_is] = rec.r[P_BJT_jile_t_is] / tnXTB;
    }
    if (flags & F_BJT_ISC) {
        junction_tempvars<false>(p.sub(P_BJT_T_rawjilc_isat), t, false, rec, P_BJT_jilc_t_is);
        rec.r[P_BJT_jilc_t_is] = rec.r[P_BJT_jilc_t_is] / tnXTB;
    }
    if (nsub)
        junction_tempvars<false>(p.sub(P_BJT_T_rawcbj_isat), t, (flags & F_BJT_CJS) != 0, rec,
                                 P_BJT_cbj_t_is);
    rec.r[P_BJT_bf_t] = p[P_BJT_T_bf] * tnXTB;
    rec.r[P_BJT_br_t] = p[P_BJT_T_br] * tnXTB;
    BufSink<P> buf;
    const int n = bjt_eval<P, SV>(rec, flags, v, buf);
    const int ncur = (SV ? 5 : 3) + nrb + nsub;
    D pout;
    if (SV)
        pout = (buf.o[0] * buf.o[1] + buf.o[2] * buf.o[3] + buf.o[4] * (buf.o[1] - buf.o[3]))
               / p[P_BJT_gyr];
    else
        pout = buf.o[0] * v[0] + buf.o[1] * v[1] + buf.o[2] * (v[0] - v[1]);
    if (nrb) pout = pout + buf.o[SV ? 5 : 3] * v[2];
    if (nsub) pout = pout + pick(v, 2 + nrb) * buf.o[ncur - 1];
#pragma unroll
    for (int k = 0; k < CDN_MAX_OUT_BUF; ++k) {
        if (k < ncur) out.put(k, buf.o[k]);
        else if (k < n) out.put(k + 1, buf.o[k]);
    }
    out.put(ncur, pout);
}

// mesfetc_t: in [vgs, vgd, dT, vgs(t-tau), vgd(t-tau)] (the thermal port sits
// before the delayed copies, autoThermal.py:114-116); out [igs, igd, ids,
// power], q [qgs, qgd].  set_temp_vars: mesfetc.py:160-189; power :233-242.
template <int P, class S>
CDN_HD void mesfetc_t_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    typedef Dual<P> D;
    DualRec<P, NP_MESFETC> rec;
#pragma unroll
    for (int k = 0; k < NP_MESFETC; ++k) rec.r[k] = D(p[k]);
    const double temp_amb = p[P_MESFETC_T_temp_amb];
    const TempState<P> t = temp_state(v[2], temp_amb, p[P_MESFETC_T_Tnomabs],
                                      p[P_MESFETC_T_egapn], p[P_MESFETC_T_eg0]);
    junction_tempvars<false>(p.sub(P_MESFETC_T_rawgs_isat), t, (flags & F_MESFETC_CGS0) != 0, rec,
                             P_MESFETC_diogs_t_is);
    junction_tempvars<false>(p.sub(P_MESFETC_T_rawgd_isat), t, (flags & F_MESFETC_CGD0) != 0, rec,
                             P_MESFETC_diogd_t_is);
    const D delta_T = (v[2] + temp_amb) - p[P_MESFETC_T_tnom];
    rec.r[P_MESFETC_k6] = p[P_MESFETC_T_nr] * t.vt;
    rec.r[P_MESFETC_Vt0] = p[P_MESFETC_T_vt0] * (1. + (delta_T * p[P_MESFETC_T_avt0])
                                               + (delta_T * delta_T * p[P_MESFETC_T_bvt0]));
    const double tbet = p[P_MESFETC_T_tbet], tm = p[P_MESFETC_T_tm], tme = p[P_MESFETC_T_tme];
    if (tbet != 0.)          // beta * 1.01 ** (delta_T * tbet)
        rec.r[P_MESFETC_Beta] = p[P_MESFETC_T_beta] * exp((delta_T * tbet) * 0.009950330853168083);
    if (tme * tm != 0.)
        rec.r[P_MESFETC_idsFac] = dpow(1. + delta_T * tm, tme);
    const D v1[4] = {v[0], v[1], v[3], v[4]};
    BufSink<P> buf;
    mesfetc_eval(rec, flags, v1, buf);
    const D pout = (v[0] - v[1]) * buf.o[2] + v[0] * buf.o[0] + v[1] * buf.o[1];
#pragma unroll
    for (int k = 0; k < 3; ++k) out.put(k, buf.o[k]);
    out.put(3, pout);
    out.put(4, buf.o[3]);
    out.put(5, buf.o[4]);
}

// svdiode_t: in [x, dT]; out [id, gyr*vd, power] (+ [qd]).  svdiode.py:280-295
// (set_temp_vars), :336-345 (power = id * (gyr vd) / gyr).
template <int P, class S>
CDN_HD void svdiode_t_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    DualRec<P, NP_SVDIODE> rec;
#pragma unroll
    for (int k = 0; k < NP_SVDIODE; ++k) rec.r[k] = Dual<P>(p[k]);
    const TempState<P> t = temp_state(v[1], p[P_SVDIODE_T_temp_amb], p[P_SVDIODE_T_Tnomabs],
                                      p[P_SVDIODE_T_egapn], p[P_SVDIODE_T_eg0]);
    rec.r[P_SVDIODE_vt] = t.vt;
    junction_tempvars<true>(p.sub(P_SVDIODE_T_jraw_isat), t, (flags & F_SVDIODE_CJ0) != 0, rec,
                            P_SVDIODE_jtn_t_is);
    Dual<P> iD, vD, qD;
    svdiode_core(rec, flags, v[0], iD, vD, qD);
    out.put(0, iD);
    out.put(1, vD);
    out.put(2, iD * vD / p[P_SVDIODE_gyr]);
    if (flags & F_SVDIODE_QD) out.put(3, qD);
}

// res_t: in [v, dT]; out [i, power].  resistor.py:107-131: the conductance
// 1/r (or the sheet-resistance value) divided by 1 + (tc1 + tc2 dT') dT',
// dT' = temp - tnom.  (The reference's `self.g /= ...` compounds on repeated
// calls; it is executed once, while the AD tape is recorded, which is what is
// restated here.)
template <int P, class S>
CDN_HD void res_t_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    const Dual<P> deltaT = (v[1] + p[P_RES_T_temp_amb]) - p[P_RES_T_tnom];
    const Dual<P> g = p[P_RES_T_g0] / (1. + (p[P_RES_T_tc1] + p[P_RES_T_tc2] * deltaT) * deltaT);
    const Dual<P> i = g * v[0];
    out.put(0, i);
    out.put(1, v[0] * i);
}

// mosekv_t: in [vdb, vgb, vsb, dT]; out [ids, idb, isb, power], q [qd, qg, qs].
// set_temp_vars: mosEKV.py:329-357 and the extrinsic wrapper mosExt.py:252-278;
// the folded record entries (mosEKV.IntEKV.cuda_pack) are rebuilt from them.
// power: mosEKV.py:586-595.
template <int P, class S>
CDN_HD void ekv_t_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    typedef Dual<P> D;
    DualRec<P, NP_EKV> rec;
#pragma unroll
    for (int k = 0; k < NP_EKV; ++k) rec.r[k] = D(p[k]);
    const D temp = v[3] + p[P_EKV_T_temp_amb];
    const D T = CDN_CONST_T0 + temp;
    const D Vt = CDN_CONST_K * T / CDN_CONST_Q;
    const double Tn = p[P_EKV_T_Tn], Tref = p[P_EKV_T_Tref], sga = p[P_EKV_T_sga];
    const double tf = p[P_EKV_tf], gammaa = p[P_EKV_gammaa], cox = p[P_EKV_T_cox];
    const D vt0T = p[P_EKV_T_vt0] - p[P_EKV_T_tcv] * (T - Tn);
    const D kpT = p[P_EKV_T_kp] * dpow(T / Tn, p[P_EKV_T_bex]);
    D kpa = kpT * (1. + p[P_EKV_T_akp] / sga);
    kpa = select(kpa, kpa, 0.0);
    const D t_ucrit = p[P_EKV_T_ucrit] * dpow(T / Tn, p[P_EKV_T_ucex]);
    const D egT = 1.16 - 0.000702 * T * T / (T + 1108.);
    const D phiT = p[P_EKV_T_phi] * T / Tref - 3. * Vt * log(T / Tref)
                   - p[P_EKV_T_egTref] * T / Tref + egT;
    const D t_ibb = p[P_EKV_T_ibb] * (1. + p[P_EKV_T_ibbt] * (T - Tref));
    const D vc = t_ucrit * p[P_EKV_T_leff] * p[P_EKV_T_ns];
    const D qb0 = gammaa * sqrt(phiT);
    rec.r[P_EKV_Vt] = Vt;
    rec.r[P_EKV_vt0a] = vt0T + p[P_EKV_T_avto] / sga;
    rec.r[P_EKV_phiT] = phiT;
    rec.r[P_EKV_vc] = vc;
    rec.r[P_EKV_t_ucrit] = t_ucrit;
    rec.r[P_EKV_t_ibb] = t_ibb;
    rec.r[P_EKV_gsqphi] = gammaa * sqrt(phiT);
    rec.r[P_EKV_vt16] = 16. * Vt * Vt;
    rec.r[P_EKV_vt01] = 0.1 * Vt;
    rec.r[P_EKV_vt4] = 4. * Vt;
    rec.r[P_EKV_vdsslog] = Vt * (log(vc / (2. * Vt)) - 0.6);
    rec.r[P_EKV_kpanpw] = kpa * p[P_EKV_T_np] * p[P_EKV_weff];
    rec.r[P_EKV_bop] = 1. + cox * qb0 / p[P_EKV_e0] / p[P_EKV_epSi];
    rec.r[P_EKV_coxvt] = cox * Vt;
    rec.r[P_EKV_ntibblc] = -t_ibb * p[P_EKV_lc];
    rec.r[P_EKV_qscale] = p[P_EKV_T_Cox] * Vt * tf;
    // extrinsic junctions (mosExt.py:260-278)
    const TempState<P> t = temp_state(v[3], p[P_EKV_T_temp_amb], p[P_EKV_T_Tnabs],
                                      p[P_EKV_T_egapn], p[P_EKV_T_eg0]);
    if (flags & F_MOSEXT_AD)
        junction_tempvars<false>(p.sub(P_EKV_T_rawdj_isat), t, (flags & F_MOSEXT_CJ_D) != 0, rec,
                                 P_EKV_dj_t_is);
    if (flags & F_MOSEXT_PD)
        junction_tempvars<false>(p.sub(P_EKV_T_rawdjsw_isat), t, (flags & F_MOSEXT_CJSW_D) != 0,
                                 rec, P_EKV_djsw_t_is);
    if (flags & F_MOSEXT_ASRC)
        junction_tempvars<false>(p.sub(P_EKV_T_rawsj_isat), t, (flags & F_MOSEXT_CJ_S) != 0, rec,
                                 P_EKV_sj_t_is);
    if (flags & F_MOSEXT_PS)
        junction_tempvars<false>(p.sub(P_EKV_T_rawsjsw_isat), t, (flags & F_MOSEXT_CJSW_S) != 0,
                                 rec, P_EKV_sjsw_t_is);
    D iv[3], qv[3];
    ekv_intrinsic(rec, flags, v, iv, qv);
    mosext_apply(rec, P_EKV_m, flags, tf, v, iv, qv);
#pragma unroll
    for (int k = 0; k < 3; ++k) out.put(k, iv[k]);
    out.put(3, (v[0] - v[2]) * iv[0] + v[0] * iv[1] + v[2] * iv[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) out.put(4 + k, qv[k]);
}
