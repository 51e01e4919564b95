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
template <int P, class R>
CDN_HD void junction_tempvars(const PV& raw, const TempState<P>& t, bool cap, R& rec, int off) {
    typedef Dual<P> D;
    rec.r[off + J_t_is] = raw[JR_isat] * dpow(t.tnratio, raw[JR_k1])
                          * exp(raw[JR_k2] - raw[JR_k3] / t.Tabs);
    rec.r[off + J_kexp] = t.vt * raw[JR_n];
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
    junction_tempvars(p.sub(P_DIODE_T_jraw_isat), t, (flags & F_DIODE_CJ0) != 0, rec,
                      P_DIODE_jtn_t_is);
    Dual<P> iD, qD;
    diode_core(rec, flags, v[0], iD, qD);
    out.put(0, iD);
    out.put(1, v[0] * iD);
    if (flags & F_DIODE_QD) out.put(2, qD);
}
