// Gummel-Poon BJT, plain and state-variable formulations, with the extrinsic
// collector-bulk junction.
// Reference: devices/bjt.py eval_cqs :485-607 (+ extrinsic :183-202),
//            devices/svbjt.py eval_cqs :308-435.
// Quirks kept: the reverse current of the plain model uses the *forward*
// junction (bjt.py:510); rb(ib) tan formula selected by |ib| > 0 (:557-563);
// everything times area*typef at the end, gyrator output pre-divided by
// area^2 (:569, :604).
//
// Fixed output slots (compacted by the caller according to flags):
//   plain: i = [ibe, ibc, ice, rb, sub]           q = [qbe, qbc, qbx, qsub]
//   sv:    i = [ibe, gyr*vbe, ibc, gyr*vbc, ice, rb, sub]
#pragma once
#include "junction.cuh"

//
// `A` is the record accessor: PV (packed isothermal record) or the dual
// record of the electro-thermal variant (thermal.cuh); entries that never
// depend on temperature are read through val().
template <int P, bool SV, class S, class A>
CDN_HD int bjt_eval(const A& p, unsigned flags, const Dual<P>* v, S& out) {
    const double typef = val(p[P_BJT_typef]), area = val(p[P_BJT_area]), gyr = val(p[P_BJT_gyr]);
    const auto jif = p.sub(P_BJT_jif_t_is), jir = p.sub(P_BJT_jir_t_is);
    const Dual<P> x1 = typef * v[0];
    const Dual<P> x2 = typef * v[1];
    Dual<P> ibf, ibr, vbe, vbc;
    if (SV) {
        svjunction_idvd(jif, x1, ibf, vbe);
        svjunction_idvd(jir, x2, ibr, vbc);
    } else {
        vbe = x1; vbc = x2;
        ibf = junction_id(jif, vbe);
        ibr = junction_id(jif, vbc);
    }
    Dual<P> ile(0.0), ilc(0.0);
    if (flags & F_BJT_ISE) ile = junction_id(p.sub(P_BJT_jile_t_is), vbe);
    if (flags & F_BJT_ISC) ilc = junction_id(p.sub(P_BJT_jilc_t_is), vbc);
    Dual<P> q1m1(1.0);
    if (flags & F_BJT_VAR) q1m1 = q1m1 - vbe / val(p[P_BJT_var]);
    if (flags & F_BJT_VAF) q1m1 = q1m1 - vbc / val(p[P_BJT_vaf]);
    Dual<P> kqb = 1.0 / q1m1;
    if (flags & (F_BJT_IKF | F_BJT_IKR)) {
        Dual<P> q2(0.0);
        if (flags & F_BJT_IKF) q2 = q2 + ibf / val(p[P_BJT_ikf]);
        if (flags & F_BJT_IKR) q2 = q2 + ibr / val(p[P_BJT_ikr]);
        kqb = kqb * (.5 * (1.0 + sqrt(1.0 + 4.0 * q2)));
    }
    const double k = area * typef;
    int n = 0;
    out.put(n++, (ibf / p[P_BJT_bf_t] + ile) * k);
    if (SV) out.put(n++, (gyr * vbe) * k);
    out.put(n++, (ibr / p[P_BJT_br_t] + ilc) * k);
    if (SV) out.put(n++, (gyr * vbc) * k);
    out.put(n++, ((ibf - ibr) / kqb) * k);
    Dual<P> vbcx(0.0);
    int nv = 2;
    if (flags & F_BJT_RB) {
        const Dual<P> ib = (typef * v[2]) * gyr;
        nv = 3;
        Dual<P> rb;
        const double rb0 = val(p[P_BJT_rb]), rbm = val(p[P_BJT_rbm]);
        if (flags & F_BJT_IRB) {
            const Dual<P> ib1 = dabs(ib);
            Dual<P> x = sqrt(1.0 + val(p[P_BJT_ck1]) * ib1) - 1.0;
            x = x * (val(p[P_BJT_ck2]) / sqrt(ib1));
            const Dual<P> tx = tan(x);
            const Dual<P> c = rbm + 3.0 * (rb0 - rbm) * (tx - x) / (x * tx * tx);
            rb = select(ib1, c, rb0);
        } else {
            rb = rbm + (rb0 - rbm) / kqb;
        }
        out.put(n++, (gyr * ib * rb / (area * area)) * k);
        vbcx = ib * rb / area + vbc;
    }
    if (flags & F_BJT_SUB) {
        const auto cbj = p.sub(P_BJT_cbj_t_is);
        const Dual<P> vs = v[nv] * typef;
        out.put(n++, junction_id(cbj, vs) * area * typef);
    }
    // ---- charges ----
    if (flags & F_BJT_QBE) {
        Dual<P> qbe(0.0);
        if (flags & F_BJT_TF) {
            Dual<P> tfeff(val(p[P_BJT_tf]));
            if (flags & F_BJT_VTF) {
                const Dual<P> x = ibf / (ibf + val(p[P_BJT_itf]));
                const Dual<P> arg = vbc / 1.44 / val(p[P_BJT_vtf]);
                const Dual<P> ex = SV ? exp(arg) : safe_exp(arg);
                tfeff = tfeff * (1.0 + val(p[P_BJT_xtf]) * x * x * ex);
            }
            qbe = tfeff * ibf;
        }
        if (flags & F_BJT_CJE) qbe = qbe + junction_qd(jif, vbe);
        out.put(n++, qbe * k);
    }
    if (flags & F_BJT_QBC) {
        Dual<P> qbc(0.0);
        if (flags & F_BJT_TR) qbc = val(p[P_BJT_tr]) * ibr;
        if (flags & F_BJT_QBX) {
            const double xcjc = val(p[P_BJT_xcjc]);
            Dual<P> qbx(0.0);
            if (flags & F_BJT_CJC) {
                qbc = qbc + junction_qd(jir, vbc) * xcjc;
                qbx = junction_qd(jir, vbcx) * (1.0 - xcjc);
            }
            out.put(n++, qbc * k);
            out.put(n++, qbx * k);
        } else {
            if (flags & F_BJT_CJC) qbc = qbc + junction_qd(jir, vbc);
            out.put(n++, qbc * k);
        }
    }
    if ((flags & F_BJT_SUB) && (flags & F_BJT_CJS)) {
        const auto cbj = p.sub(P_BJT_cbj_t_is);
        const Dual<P> vs = v[nv] * typef;
        out.put(n++, junction_qd(cbj, vs) * area * typef);
    }
    return n;
}
