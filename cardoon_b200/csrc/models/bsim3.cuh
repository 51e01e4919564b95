// BSIM3 v3.2.4 intrinsic MOSFET: drain/substrate currents and terminal
// charges.  Reference: devices/mosBSIM3v3.py eval_cqs :403-911 (line numbers
// in comments refer to that file).  Parameter-only sub-expressions arrive
// folded by the host packer (mosBSIM3v3.BSIM3.cuda_pack); the operation order
// of everything bias-dependent is the reference's.
//
// Quirks kept on purpose: D/S swap flag is -1 at Vds = 0
// (:418); the first Vgsteff assignment (:529-531) is dead; the Weff guard
// uses sqrt(1+nlx/leff) (:545); the substrate-current condassign (:720-723)
// is discarded so Isub = T2(:648) * Idsa; vfbzb's T5 uses the exponent
// argument (:761); VdseffCV += 1e-200 (:873); k_temp = 2.0137...*T0 (:851).
//
// in: v = [vdb, vgb, vsb].  out: i = [ids, idb, isb], q = [qd, qg, qs].
#pragma once
#include "junction.cuh"

#define BSIM3_MAX_EXP 5.834617425e14
#define BSIM3_MIN_EXP 1.713908431e-15
#define BSIM3_EXP_THRESHOLD 34.0

template <int P>
CDN_HD Dual<P> log1pexp(const Dual<P>& x) {           // :24-33
    const Dual<P> ex = exp(x);
    const Dual<P> y1 = select(x.v - BSIM3_EXP_THRESHOLD, x, log(1.0 + ex));
    return select(-x.v - BSIM3_EXP_THRESHOLD, ex, y1);
}

// `A` is the record accessor: PV (on-demand loads) or the evaluation kernel's
// PVRing, which streams the rows through shared memory in layout order; the
// p.mark(r) calls tell it that rows below r have been consumed (the BSIM3 list
// of param_layout.py is sorted by first use in this function).
template <int P, class A>
CDN_HD void bsim3_intrinsic(A& p, unsigned flags, const Dual<P>* v,
                            Dual<P>* iv, Dual<P>* qv) {
    typedef Dual<P> D;
    p.mark(0);
    const double tf = p[P_BSIM3_tf];
    const D vd0 = tf * v[0], vg = tf * v[1], vs0 = tf * v[2];
    // ---- D/S swap (:418-423)
    const double swap = ((vd0.v - vs0.v) > 0.0) ? 1.0 : -1.0;
    const D vd = select(swap, vd0, vs0);
    const D vs = select(swap, vs0, vd0);
    const D VDS = vd - vs;
    const D VGS = vg - vs;
    const D VBS = -vs;
    // ---- effective body bias (:429-435)
    const double vbsc = p[P_BSIM3_vbsc], phi = p[P_BSIM3_phi];
    const double sqrtPhi = p[P_BSIM3_sqrtPhi];
    D Vbseff;
    {
        const D T0 = VBS - vbsc - 0.001;
        const D T1 = sqrt(T0 * T0 - 0.004 * vbsc);
        const D t = vbsc + .5 * (T0 + T1);
        Vbseff = select(-t + VBS, VBS, t);
    }
    const D Phis = select(Vbseff, (phi * phi) / (phi + Vbseff), phi - Vbseff);
    const D sqrtPhis = select(Vbseff, p[P_BSIM3_phis3] / (phi + 0.5 * Vbseff), sqrt(Phis));
    const D Xdep = p[P_BSIM3_Xdep0] * sqrtPhis / sqrtPhi;
    // ---- threshold voltage (:446-500)
    const double V0 = p[P_BSIM3_V0];
    D Theta0, Vth;
    {
        const D T3 = sqrt(Xdep);
        const double dvt2 = p[P_BSIM3_dvt2], dvt2w = p[P_BSIM3_dvt2w];
        const double factor1 = p[P_BSIM3_factor1];
        D T1 = select(dvt2 * Vbseff + .5, 1.0 + dvt2 * Vbseff,
                      (1.0 + 3.0 * dvt2 * Vbseff) / (3.0 + 8.0 * dvt2 * Vbseff));
        const D ltl = factor1 * T3 * T1;
        T1 = select(dvt2w * Vbseff + .5, 1.0 + dvt2w * Vbseff,
                    (1.0 + 3.0 * dvt2w * Vbseff) / (3.0 + 8.0 * dvt2w * Vbseff));
        const D ltw = factor1 * T3 * T1;
        D k_temp = p[P_BSIM3_c_dvt1] / ltl;
        D T2 = select(k_temp + BSIM3_EXP_THRESHOLD, safe_exp(k_temp), D(BSIM3_MIN_EXP));
        Theta0 = T2 * (1.0 + 2.0 * T2);
        const D Delt_vth = p[P_BSIM3_dvt0] * Theta0 * V0;
        k_temp = p[P_BSIM3_c_dvt1w] / ltw;
        T2 = select(k_temp + BSIM3_EXP_THRESHOLD, safe_exp(k_temp), D(BSIM3_MIN_EXP));
        T2 = T2 * (1.0 + 2.0 * T2);
        const D T0 = p[P_BSIM3_dvt0w] * T2;
        T2 = T0 * V0;
        p.mark(P_BSIM3_t1a);
        T1 = p[P_BSIM3_t1a] + (p[P_BSIM3_kt1eff] + p[P_BSIM3_kt2] * Vbseff) * p[P_BSIM3_ToTnm1];
        const D T3b = p[P_BSIM3_eta0] + p[P_BSIM3_etab] * Vbseff;
        const D DIBL_Sft = (T3b * p[P_BSIM3_theta0vb0]) * VDS;
        Vth = p[P_BSIM3_vth0mk1] + p[P_BSIM3_k1ox] * sqrtPhis
              - p[P_BSIM3_k2ox] * Vbseff - Delt_vth - T2
              + (p[P_BSIM3_k3] + p[P_BSIM3_k3b] * Vbseff) * p[P_BSIM3_tmp2] + T1 - DIBL_Sft;
    }
    // ---- subthreshold slope factor n (:502-509)
    p.mark(P_BSIM3_cox);
    const double cox = p[P_BSIM3_cox], Vt = p[P_BSIM3_Vt];
    D n;
    {
        const D tmp2 = p[P_BSIM3_nfepSi] / Xdep;
        const D tmp3 = p[P_BSIM3_cdsc] + p[P_BSIM3_cdscb] * Vbseff + p[P_BSIM3_cdscd] * VDS;
        const D tmp4 = (tmp2 + tmp3 * Theta0 + p[P_BSIM3_cit]) / cox;
        n = select(tmp4 + .5, 1.0 + tmp4, (1.0 + 3.0 * tmp4) * (1.0 / (3.0 + 8.0 * tmp4)));
    }
    // ---- effective Vgst (:511-537)
    const D Vgs_eff = VGS;
    const D Vgst = Vgs_eff - Vth;
    D Vgsteff;
    {
        const D T10 = 2.0 * n * Vt;
        const D VgstNVt = Vgst / T10;
        const D ExpArg = -(2.0 * 0.08 + Vgst) / T10;
        const D T1 = T10 * log1pexp(VgstNVt);
        const D T2 = 1.0 + T10 * cox * safe_exp(ExpArg) / Vt / p[P_BSIM3_cdep0];
        Vgsteff = select(ExpArg - BSIM3_EXP_THRESHOLD,
                         p[P_BSIM3_vtcdepcox] / safe_exp((Vgst + 0.08) / n / Vt),
                         T1 / T2);
    }
    // ---- effective channel geometry, Rds (:539-551)
    p.mark(P_BSIM3_weff);
    const double weff = p[P_BSIM3_weff], leff = p[P_BSIM3_leff];
    D Weff, Rds;
    {
        const D T9 = sqrtPhis - sqrtPhi;
        const D k_temp = weff - 2.0 * (p[P_BSIM3_dwg] * Vgsteff + p[P_BSIM3_dwb] * T9);
        Weff = select(-k_temp + 2.0e-8, 2e-8 * (4.0e-8 - k_temp) * p[P_BSIM3_sqrt1nlx], k_temp);
        const D T0 = p[P_BSIM3_prwg] * Vgsteff + p[P_BSIM3_prwb] * (sqrtPhis - sqrtPhi);
        const double rds0 = p[P_BSIM3_rds0];
        Rds = select(T0 + 0.9, rds0 * (1.0 + T0), rds0 * (.8 + T0) / (17.0 + 20.0 * T0));
    }
    // ---- bulk-charge effect (:552-574)
    D Abulk, Abulk0;
    {
        const D T1 = 0.5 * p[P_BSIM3_k1ox] / sqrtPhis;
        const D T9 = sqrt(p[P_BSIM3_xj] * Xdep);
        const D T5 = leff / (leff + 2.0 * T9);
        const D T2 = (p[P_BSIM3_a0] * T5) + p[P_BSIM3_b0b1];
        const D T6 = T5 * T5;
        const D T7 = T5 * T6;
        Abulk0 = 1.0 + T1 * T2;
        const D T8 = p[P_BSIM3_agsa0] * T7;
        Abulk = Abulk0 - T1 * T8 * Vgsteff;
        Abulk0 = select(-Abulk0 + .1, (.2 - Abulk0) / (3.0 - 20.0 * Abulk0), Abulk0);
        Abulk = select(-Abulk + .1, (.2 - Abulk) / (3.0 - 20.0 * Abulk), Abulk);
        const D T2k = p[P_BSIM3_keta] * Vbseff;
        const D T0 = select(T2k + 0.9, 1.0 / (1.0 + T2k), (17.0 + 20.0 * T2k) / (0.8 + T2k));
        Abulk = Abulk * T0;
        Abulk0 = Abulk0 * T0;
    }
    // ---- mobility, Esat (:575-584)
    p.mark(P_BSIM3_tox);
    const double tox = p[P_BSIM3_tox], vsattemp = p[P_BSIM3_vsattemp];
    D ueff, Esat;
    {
        const D T0 = Vgsteff + 2.0 * Vth;
        const D T2 = p[P_BSIM3_ua] + p[P_BSIM3_uc] * Vbseff;
        const D T3 = T0 / tox;
        const D T5 = T3 * (T2 + p[P_BSIM3_ub] * T3);
        const D Denomi = select(T5 + .8, 1.0 + T5, (.6 + T5) / (7.0 + 10.0 * T5));
        ueff = p[P_BSIM3_u0temp] / Denomi;
        Esat = 2.0 * vsattemp / ueff;
    }
    // ---- Vdsat, Vdseff (:585-603)
    const D WVCoxRds = (Weff * vsattemp * cox) * Rds;
    const D EsatL = Esat * leff;
    const double Lambda = p[P_BSIM3_a2];
    const D Vgst2Vtm = Vgsteff + 2.0 * Vt;
    D Vdsat, Vdseff;
    {
        const D T0 = 1.0 / (Abulk * EsatL + Vgst2Vtm);
        const D T3 = EsatL * Vgst2Vtm;
        Vdsat = T3 * T0;
        const double delta = p[P_BSIM3_delta];
        const D T1 = Vdsat - VDS - delta;
        const D T2 = sqrt(sq(T1) + 4.0 * delta * Vdsat);
        const D k_temp = Vdsat - .5 * (T1 + T2);
        Vdseff = select(k_temp - VDS, VDS, k_temp);
    }
    // ---- Vasat (:610-617)
    D Vasat;
    {
        const D T6 = 1.0 - .5 * Abulk * Vdsat / Vgst2Vtm;
        D T9 = WVCoxRds * Vgsteff;
        const D T0 = EsatL + Vdsat + 2.0 * T9 * T6;
        T9 = WVCoxRds * Abulk;
        const D T1 = 2.0 / Lambda - 1.0 + T9;
        Vasat = T0 / T1;
    }
    const D diffVds = VDS - Vdseff;
    // ---- VACLM, VADIBL, Va (:619-668)
    D Va;
    {
        p.mark(P_BSIM3_pclm);
        const double pclm = p[P_BSIM3_pclm], litl = p[P_BSIM3_litl];
        const D VACLM = select((diffVds - 1.0e-10) * pclm,
                               leff * (Abulk + Vgsteff / EsatL) * diffVds / (pclm * Abulk * litl),
                               D(BSIM3_MAX_EXP));
        const double thetaRout = p[P_BSIM3_thetaRout], pdiblb = p[P_BSIM3_pdiblb];
        D VADIBL = select(thetaRout,
                          (Vgst2Vtm - Vgst2Vtm * Abulk * Vdsat / (Vgst2Vtm + Abulk * Vdsat)) / thetaRout,
                          D(BSIM3_MAX_EXP));
        VADIBL = select(pdiblb * Vbseff + 0.9, VADIBL / (1.0 + pdiblb * Vbseff),
                        VADIBL * (17.0 + 20.0 * pdiblb * Vbseff) / (.8 + pdiblb * Vbseff));
        const D T8 = p[P_BSIM3_pvag] / EsatL;
        const D T9 = T8 * Vgsteff;
        const D T0 = select(T9 + 0.9, 1.0 + T9, (.8 + T9) / (17.0 + 20.0 * T9));
        const D T3 = VACLM + VADIBL;
        const D T1 = VACLM * VADIBL / T3;
        Va = Vasat + T0 * T1;
    }
    // ---- substrate-current-induced body effect (:670-679)
    D rcpVASCBE;
    if (flags & F_BSIM3_PSCBE1)
        rcpVASCBE = select(dabs(diffVds),
                           p[P_BSIM3_pscbe2] * safe_exp(p[P_BSIM3_npscbe1litl] / diffVds) / leff,
                           0.0);
    else
        rcpVASCBE = D(p[P_BSIM3_pscbe2] / leff);
    // ---- drain current (:686-704)
    D Ids, Idsa;
    {
        const D CoxWovL = cox * Weff / leff;
        const D beta = ueff * CoxWovL;
        D T0 = 1.0 - .5 * Abulk * Vdseff / Vgst2Vtm;
        const D fgche1 = Vgsteff * T0;
        D T9 = Vdseff / EsatL;
        const D fgche2 = 1.0 + T9;
        const D gche = beta * fgche1 / fgche2;
        T0 = 1.0 + gche * Rds;
        T9 = Vdseff / T0;
        const D Idl = gche * T9;
        T9 = diffVds / Va;
        T0 = 1.0 + T9;
        Idsa = Idl * T0;
        T9 = diffVds * rcpVASCBE;
        T0 = 1.0 + T9;
        Ids = Idsa * T0;
    }
    // ---- substrate current (:706-726)
    D Isub(0.0);
    if (flags & F_BSIM3_ISUB) Isub = p[P_BSIM3_t2sub] * Idsa;
    iv[0] = (swap * Ids) * tf;
    iv[1] = select(swap, Isub, 0.0) * tf;
    iv[2] = select(swap, D(0.0), Isub) * tf;
    // ---- charges (:738-907)
    p.mark(P_BSIM3_vfbzb);
    const double vfbzb = p[P_BSIM3_vfbzb], Tox = p[P_BSIM3_Tox];
    const double ldeb = p[P_BSIM3_ldeb], epSi = p[P_BSIM3_epSi];
    const double k1ox = p[P_BSIM3_k1ox];
    const D VbseffCV = select(Vbseff, phi - Phis, Vbseff);
    D VgsteffCV;
    {
        const D T0 = n * p[P_BSIM3_noff] * Vt;
        const D T1 = (Vgs_eff - Vth) / T0;
        VgsteffCV = select(T1 - BSIM3_EXP_THRESHOLD, Vgs_eff - Vth - p[P_BSIM3_voffcv],
                           T0 * log1pexp(T1));
    }
    D Vfbeff;
    {
        const D V3 = vfbzb - Vgs_eff + VbseffCV - .02;
        const D T0 = sqrt(sq(V3) + p[P_BSIM3_avfbzb08]);
        Vfbeff = vfbzb - 0.5 * (V3 + T0);
    }
    D CoxWLcen;
    {
        const D T0 = (Vgs_eff - VbseffCV - vfbzb) / Tox;
        const D T1 = T0 * p[P_BSIM3_acde];
        D Tcen = select(BSIM3_EXP_THRESHOLD + T1, ldeb * exp(T1), D(ldeb * BSIM3_MIN_EXP));
        Tcen = select(-BSIM3_EXP_THRESHOLD + T1, D(ldeb * BSIM3_MAX_EXP), Tcen);
        const D V3 = ldeb - Tcen - p[P_BSIM3_c1tox];
        const D V4 = sqrt(sq(V3) + p[P_BSIM3_c2toxldeb]);
        Tcen = ldeb - .5 * (V3 + V4);
        const D Ccen = epSi / Tcen;
        const D Coxeff = Ccen * cox / (Ccen + cox);
        CoxWLcen = Weff * leff * Coxeff;
    }
    const D Qac0 = CoxWLcen * (Vfbeff - vfbzb);
    D Qsub0;
    {
        const double T0 = .5 * k1ox;
        const D T3 = Vgs_eff - Vfbeff - VbseffCV - VgsteffCV;
        const D T1 = select(-T3, T0 + T3 / k1ox, sqrt(T0 * T0 + T3));
        Qsub0 = CoxWLcen * k1ox * (T1 - T0);
    }
    D DeltaPhi;
    {
        const D T1 = 2.0 * p[P_BSIM3_t0dp] + VgsteffCV;
        DeltaPhi = Vt * log(1.0 + T1 * VgsteffCV / p[P_BSIM3_t2dp]);
    }
    {
        const D T3 = 4.0 * (Vth - vfbzb - phi);
        const D T0 = select(T3, .5 * (VgsteffCV + T3) / Tox, .5 * (VgsteffCV + 1.0e-20) / Tox);
        const D k_temp = 2.01375270747048 * T0;
        const D T1 = 1.0 + k_temp;
        const D Tcen = 1.9e-9 / T1;
        const D Ccen = epSi / Tcen;
        const D Coxeff = Ccen * cox / (Ccen + cox);
        CoxWLcen = Weff * leff * Coxeff;
    }
    const D AbulkCV = Abulk0 * p[P_BSIM3_AbulkCVfactor];
    const D VdsatCV = (VgsteffCV - DeltaPhi) / AbulkCV;
    D VdseffCV;
    {
        const D T0 = VdsatCV - VDS - .02;
        const D T1 = sqrt(sq(T0) + .08 * VdsatCV);
        VdseffCV = VdsatCV - .5 * (T0 + T1);
        VdseffCV = VdseffCV + 1e-200;
    }
    D qgate, qbulk, qsrc;
    {
        const D T0 = AbulkCV * VdseffCV;
        const D T1 = VgsteffCV - DeltaPhi;
        D T2 = 12.0 * (T1 - .5 * T0 + 1e-20);
        D T3 = T0 / T2;
        qgate = CoxWLcen * (T1 - T0 * (.5 - T3));
        qbulk = CoxWLcen * (1.0 - AbulkCV) * (.5 * VdseffCV - T0 * VdseffCV / T2);
        T2 = T2 / 12.0;
        T3 = .5 * CoxWLcen / sq(T2);
        const D T4 = T1 * (2.0 * sq(T0) / 3.0 + T1 * (T1 - 4.0 * T0 / 3.0))
                     - 2.0 * (T0 * T0 * T0) / 15.0;
        qsrc = -T3 * T4;
    }
    qgate = qgate + (Qac0 + Qsub0 - qbulk);
    qbulk = qbulk - (Qac0 + Qsub0);
    const D qdrn = -(qbulk + qgate + qsrc);
    qv[0] = select(swap, qdrn, qsrc) * tf;
    qv[1] = qgate * tf;
    qv[2] = select(swap, qsrc, qdrn) * tf;
}
