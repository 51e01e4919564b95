"""
Device equations of the reference, restated on numpy / ``Dual`` numbers
(oracle, test infrastructure).

Each ``*_cqs(dev, v)`` follows the reference's ``eval_cqs`` of that model
operation by operation (file:line cited per function) and reads the raw
post-``process_params``/``set_temp_vars`` attributes of a device instance,
like the reference does -- none of the host-side constant folding the CUDA
packers perform is used here, so the two paths are independent from the
attributes downwards.  ``v`` is a list of control voltages (numpy arrays of
shape ``(n,)`` or ``Dual``); the result is ``(currents, charges)`` as lists.

``device_eval(dev, xin, deriv)`` is the stand-in for
cppaddev.eval / eval_and_deriv (cppaddev.py:134-170).
"""
import numpy as np

from . import dual as D
from .dual import condassign, safe_exp, exp, log, sqrt

EPSI = 104.5e-12      # globalVars.py:38
EPOX = 34.5e-12       # globalVars.py:39


# ---- junctions (diode.py:74-98, svdiode.py:80-111) ------------------------
def junction_id(j, vd):
    return j._t_is * (safe_exp(vd / j._kexp) - 1.)


def junction_qd(j, vd, guarded=True):
    if guarded and j.cj0 == 0.:
        return 0. * vd
    b = j.fc * j._t_vj - vd
    c = j._k5 * (1. - D.power(1. - vd / j._t_vj, j._k4))
    d = j._k6 * ((1. - j.fc * (1. + j.m)) * vd
                 + .5 * j.m * vd * vd / j._t_vj - j._k7) \
        + j._k5 * (1. - pow(1. - j.fc, j._k4))
    return condassign(b, c, d)


def svjunction_idvd(j, x):
    b = j._svth - x
    c = j._t_is * (exp(j._alpha * x) - 1.)
    d = j._t_is * j._kexp * (1. + j._alpha * (x - j._svth)) - j._t_is
    iD = condassign(b, c, d)
    d = j._svth + log(1. + j._alpha * (x - j._svth)) / j._alpha
    vD = condassign(b, x, d)
    return iD, vD


# ---- diode (diode.py:275-307) ---------------------------------------------
def diode_cqs(dev, v):
    vd = v[0]
    iD = junction_id(dev.jtn, vd)
    qD = junction_qd(dev.jtn, vd) if dev.cj0 != 0. else 0.
    if dev.tt != 0.:
        qD = qD + dev.tt * iD
    if dev.bv < np.inf:
        iD = iD - dev.ibv * safe_exp(-(vd + dev.bv) / dev.n / dev.vt)
    iD = iD * dev.area
    if dev._qd:
        return [iD], [qD * dev.area]
    return [iD], []


# ---- state-variable diode (svdiode.py:300-333) ------------------------------
def svdiode_cqs(dev, v):
    iD, vD = svjunction_idvd(dev.jtn, v[0])
    qD = junction_qd(dev.jtn, vD, guarded=False) if dev.cj0 != 0. else 0.
    if dev.tt != 0.:
        qD = qD + dev.tt * iD
    if dev.bv < np.inf:
        iD = iD - dev.ibv * safe_exp(-(vD + dev.bv) / dev.n / dev.vt)
    iD = iD * dev.area
    vD = vD * dev_gyr()
    if dev._qd:
        return [iD, vD], [qD * dev.area]
    return [iD, vD], []


# ---- Gummel-Poon (bjt.py:485-607) and state-variable GP (svbjt.py:308-435)
def _gp_core(dev, ibf, ibr, vbe, vbc, vib, sv):
    """Common part once junction currents/voltages are known."""
    gyr = dev_gyr()
    ile = junction_id(dev.jile, vbe) if dev.ise != 0. else 0.
    ilc = junction_id(dev.jilc, vbc) if dev.isc != 0. else 0.
    q1m1 = 1.
    if dev.var != 0.:
        q1m1 = q1m1 - vbe / dev.var
    if dev.vaf != 0.:
        q1m1 = q1m1 - vbc / dev.vaf
    kqb = 1. / q1m1
    if dev.ikf + dev.ikr != 0.:
        q2 = 0.
        if dev.ikf != 0.:
            q2 = q2 + ibf / dev.ikf
        if dev.ikr != 0.:
            q2 = q2 + ibr / dev.ikr
        kqb = kqb * (.5 * (1. + sqrt(1. + 4. * q2)))
    ibe = ibf / dev._bf_t + ile
    ibc = ibr / dev._br_t + ilc
    ice = (ibf - ibr) / kqb
    if sv:
        iout = [ibe, gyr * vbe, ibc, gyr * vbc, ice]
    else:
        iout = [ibe, ibc, ice]
    vbcx = None
    if dev.rb != 0.:
        ib = vib * gyr
        if dev.irb != 0.:
            ib1 = D.dabs(ib)
            x = sqrt(1. + dev._ck1 * ib1) - 1.
            x = x * (dev._ck2 / sqrt(ib1))
            tx = D.tan(x)
            c = dev.rbm + 3. * (dev.rb - dev.rbm) * (tx - x) / (x * tx * tx)
            rb = condassign(ib1, c, dev.rb)
        else:
            rb = dev.rbm + (dev.rb - dev.rbm) / kqb
        iout.append(gyr * ib * rb / pow(dev.area, 2))
        vbcx = ib * rb / dev.area + vbc
    nq = dev.ncharges
    qout = [0.] * nq
    if dev.tf != 0.:
        tfeff = dev.tf
        if dev.vtf != 0.:
            x = ibf / (ibf + dev.itf)
            ex = (exp if sv else safe_exp)(vbc / 1.44 / dev.vtf)
            tfeff = tfeff * (1. + dev.xtf * x * x * ex)
        qout[0] = tfeff * ibf
    if dev.cje != 0.:
        qout[0] = qout[0] + junction_qd(dev.jif, vbe, guarded=not sv)
    if dev._qbx:
        if dev.tr != 0.:
            qout[-2] = dev.tr * ibr
        if dev.cjc != 0.:
            qout[-2] = qout[-2] \
                + junction_qd(dev.jir, vbc, guarded=not sv) * dev.xcjc
            qout[-1] = junction_qd(dev.jir, vbcx, guarded=not sv) \
                * (1. - dev.xcjc)
    else:
        if dev.tr != 0.:
            qout[-1] = dev.tr * ibr
        if dev.cjc != 0.:
            qout[-1] = qout[-1] + junction_qd(dev.jir, vbc,
                                              guarded=not sv)
    k = dev.area * dev._typef
    return [i * k for i in iout], [q * k for q in qout]


_GYR = [1e-3]


def dev_gyr():
    return _GYR[0]


def bjt_cqs(dev, v):
    v1 = [dev._typef * x for x in v]
    ibf = junction_id(dev.jif, v1[0])
    # the reverse current also uses the *forward* junction (bjt.py:510)
    ibr = junction_id(dev.jif, v1[1])
    vib = v1[2] if dev.rb != 0. else None
    return _gp_core(dev, ibf, ibr, v1[0], v1[1], vib, False)


def svbjt_cqs(dev, v):
    v1 = [dev._typef * x for x in v]
    ibf, vbe = svjunction_idvd(dev.jif, v1[0])
    ibr, vbc = svjunction_idvd(dev.jir, v1[1])
    vib = v1[2] if dev.rb != 0. else None
    return _gp_core(dev, ibf, ibr, vbe, vbc, vib, True)


def extrinsic_bjt_cqs(dev, v, intrinsic):
    """Substrate junction added by the wrapper (bjt.py:183-202)."""
    iout, qout = intrinsic(dev, v)
    if dev._addCBjtn:
        v1 = v[dev._bccn] * dev._typef
        iout = iout + [junction_id(dev.cbjtn, v1) * dev.area * dev._typef]
        if dev.cjs != 0.:
            qout = qout + [junction_qd(dev.cbjtn, v1) * dev.area
                           * dev._typef]
    return iout, qout


# ---- BSIM3 v3.2.4 (mosBSIM3v3.py:403-911) ---------------------------------
MAX_EXP = 5.834617425e14
MIN_EXP = 1.713908431e-15
MIN_1 = MIN_EXP * (1. + 2. * MIN_EXP)
EXP_THRESHOLD = 34.0


def log1pexp(x):
    """mosBSIM3v3.py:24-33"""
    y1 = condassign(x - EXP_THRESHOLD, x, log(1. + exp(x)))
    return condassign(-x - EXP_THRESHOLD, exp(x), y1)


def bsim3_cqs(s, v):
    tf = s._tf
    vd, vg, vs = tf * v[0], tf * v[1], tf * v[2]
    # swap drain and source when vds <= 0 (:418-423)
    swap = condassign(vd - vs, 1., -1.)
    vd, vs = condassign(swap, vd, vs), condassign(swap, vs, vd)
    VDS = vd - vs
    VGS = vg - vs
    VBS = -vs
    # effective body bias (:429-435)
    T0 = VBS - s.vbsc - 0.001
    T1 = sqrt(T0 * T0 - 0.004 * s.vbsc)
    Vbseff = s.vbsc + .5 * (T0 + T1)
    Vbseff = condassign(-Vbseff + VBS, VBS, Vbseff)
    Phis = condassign(Vbseff, s.phi**2 / (s.phi + Vbseff), s.phi - Vbseff)
    sqrtPhis = condassign(Vbseff, s.phis3 / (s.phi + 0.5 * Vbseff),
                          sqrt(Phis))
    Xdep = s.Xdep0 * sqrtPhis / s.sqrtPhi
    # threshold voltage (:446-500)
    T3 = sqrt(Xdep)
    T1 = condassign(s.dvt2 * Vbseff + .5, 1. + s.dvt2 * Vbseff,
                    (1. + 3. * s.dvt2 * Vbseff) / (3. + 8. * s.dvt2 * Vbseff))
    ltl = s.factor1 * T3 * T1
    T1 = condassign(s.dvt2w * Vbseff + .5, 1. + s.dvt2w * Vbseff,
                    (1. + 3. * s.dvt2w * Vbseff)
                    / (3. + 8. * s.dvt2w * Vbseff))
    ltw = s.factor1 * T3 * T1
    k_temp = -.5 * s.dvt1 * s.leff / ltl
    T2 = condassign(k_temp + EXP_THRESHOLD, safe_exp(k_temp), MIN_EXP)
    Theta0 = T2 * (1. + 2. * T2)
    Delt_vth = s.dvt0 * Theta0 * s.V0
    k_temp = -.5 * s.dvt1w * s.weff * s.leff / ltw
    T2 = condassign(k_temp + EXP_THRESHOLD, safe_exp(k_temp), MIN_EXP)
    T2 = T2 * (1. + 2. * T2)
    T0 = s.dvt0w * T2
    T2 = T0 * s.V0
    T0 = np.sqrt(1. + s.nlx / s.leff)
    sqrt1nlx = T0
    T1 = s.k1ox * (T0 - 1.) * s.sqrtPhi \
        + (s.kt1 + s.kt1l / s.leff + s.kt2 * Vbseff) * s._ToTnm1
    TMP2 = s.tox * s.phi / (s.weff + s.w0)
    T3 = s.eta0 + s.etab * Vbseff
    dDIBL_Sft_dVd = T3 * s.theta0vb0
    DIBL_Sft = dDIBL_Sft_dVd * VDS
    Vth = s._vth0 - s._k1 * s.sqrtPhi + s.k1ox * sqrtPhis \
        - s.k2ox * Vbseff - Delt_vth - T2 \
        + (s.k3 + s.k3b * Vbseff) * TMP2 + T1 - DIBL_Sft
    # subthreshold slope factor (:502-509)
    tmp2 = s.nfactor * EPSI / Xdep
    tmp3 = s.cdsc + s.cdscb * Vbseff + s.cdscd * VDS
    tmp4 = (tmp2 + tmp3 * Theta0 + s.cit) / s.cox
    n = condassign(tmp4 + .5, 1. + tmp4,
                   (1. + 3. * tmp4) * (1. / (3. + 8. * tmp4)))
    # effective Vgst (:511-537); the first assignment (:529-531) is
    # overwritten by the second
    Vgs_eff = VGS
    Vgst = Vgs_eff - Vth
    T10 = 2. * n * s._Vt
    VgstNVt = Vgst / T10
    ExpArg = -(2. * 0.08 + Vgst) / T10
    T1 = T10 * log1pexp(VgstNVt)
    T2 = 1. + T10 * s.cox * safe_exp(ExpArg) / s._Vt / s.cdep0
    Vgsteff = condassign(
        ExpArg - EXP_THRESHOLD,
        s._Vt * s.cdep0 / s.cox / safe_exp((Vgst + 0.08) / n / s._Vt),
        T1 / T2)
    # effective channel geometry (:539-551)
    T9 = sqrtPhis - s.sqrtPhi
    k_temp = s.weff - 2. * (s.dwg * Vgsteff + s.dwb * T9)
    Weff = condassign(-k_temp + 2.0e-8, 2e-8 * (4.0e-8 - k_temp) * sqrt1nlx,
                      k_temp)
    T0 = s.prwg * Vgsteff + s.prwb * (sqrtPhis - s.sqrtPhi)
    Rds = condassign(T0 + 0.9, s.rds0 * (1. + T0),
                     s.rds0 * (.8 + T0) / (17. + 20. * T0))
    # bulk charge effect (:552-574)
    T1 = 0.5 * s.k1ox / sqrtPhis
    T9 = sqrt(s.xj * Xdep)
    T5 = s.leff / (s.leff + 2. * T9)
    T2 = (s.a0 * T5) + s.b0 / (s.weff + s.b1)
    T6 = T5 * T5
    T7 = T5 * T6
    Abulk0 = 1. + T1 * T2
    T8 = s.ags * s.a0 * T7
    Abulk = Abulk0 - T1 * T8 * Vgsteff
    Abulk0 = condassign(-Abulk0 + .1, (.2 - Abulk0) / (3. - 20. * Abulk0),
                        Abulk0)
    Abulk = condassign(-Abulk + .1, (.2 - Abulk) / (3. - 20. * Abulk),
                       Abulk)
    T2 = s.keta * Vbseff
    T0 = condassign(T2 + 0.9, 1. / (1. + T2), (17. + 20. * T2) / (0.8 + T2))
    Abulk = Abulk * T0
    Abulk0 = Abulk0 * T0
    # mobility (:575-584)
    T0 = Vgsteff + 2. * Vth
    T2 = s._ua + s._uc * Vbseff
    T3 = T0 / s.tox
    T5 = T3 * (T2 + s._ub * T3)
    Denomi = condassign(T5 + .8, 1. + T5, (.6 + T5) / (7. + 10. * T5))
    ueff = s.u0temp / Denomi
    Esat = 2. * s.vsattemp / ueff
    # saturation voltage (:585-594)
    WVCox = Weff * s.vsattemp * s.cox
    WVCoxRds = WVCox * Rds
    EsatL = Esat * s.leff
    Lambda = s.a2
    Vgst2Vtm = Vgsteff + 2. * s._Vt
    T0 = 1. / (Abulk * EsatL + Vgst2Vtm)
    T3 = EsatL * Vgst2Vtm
    Vdsat = T3 * T0
    # effective Vds (:595-603)
    T1 = Vdsat - VDS - s.delta
    T2 = sqrt(T1**2 + 4. * s.delta * Vdsat)
    k_temp = Vdsat - .5 * (T1 + T2)
    Vdseff = condassign(k_temp - VDS, VDS, k_temp)
    # Vasat (:610-617)
    T6 = 1. - .5 * Abulk * Vdsat / Vgst2Vtm
    T9 = WVCoxRds * Vgsteff
    T0 = EsatL + Vdsat + 2. * T9 * T6
    T9 = WVCoxRds * Abulk
    T1 = 2. / Lambda - 1. + T9
    Vasat = T0 / T1
    diffVds = VDS - Vdseff
    # VACLM, VADIBL, Va (:619-668)
    VACLM = condassign(
        (diffVds - 1.0e-10) * s.pclm,
        s.leff * (Abulk + Vgsteff / EsatL) * diffVds
        / (s.pclm * Abulk * s.litl),
        MAX_EXP)
    T1 = np.sqrt(EPSI / EPOX * s.tox * s.Xdep0)
    T0 = float(safe_exp(-.5 * s.drout * s.leff / T1))
    T2rout = T0 + 2. * T0**2
    thetaRout = s.pdibl1 * T2rout + s.pdibl2
    VADIBL = condassign(
        thetaRout,
        (Vgst2Vtm - Vgst2Vtm * Abulk * Vdsat / (Vgst2Vtm + Abulk * Vdsat))
        / thetaRout,
        MAX_EXP)
    VADIBL = condassign(s.pdiblb * Vbseff + 0.9,
                        VADIBL / (1. + s.pdiblb * Vbseff),
                        VADIBL * (17. + 20. * s.pdiblb * Vbseff)
                        / (.8 + s.pdiblb * Vbseff))
    T8 = s.pvag / EsatL
    T9 = T8 * Vgsteff
    T0 = condassign(T9 + 0.9, 1. + T9, (.8 + T9) / (17. + 20. * T9))
    T3 = VACLM + VADIBL
    T1 = VACLM * VADIBL / T3
    Va = Vasat + T0 * T1
    # substrate-current-induced body effect (:670-679)
    if s.pscbe1 != 0.:
        rcpVASCBE = condassign(
            D.dabs(diffVds),
            s.pscbe2 * safe_exp(-s.pscbe1 * s.litl / diffVds) / s.leff,
            0.)
    else:
        rcpVASCBE = s.pscbe2 / s.leff
    # drain current (:686-704)
    CoxWovL = s.cox * Weff / s.leff
    beta = ueff * CoxWovL
    T0 = 1. - .5 * Abulk * Vdseff / Vgst2Vtm
    fgche1 = Vgsteff * T0
    T9 = Vdseff / EsatL
    fgche2 = 1. + T9
    gche = beta * fgche1 / fgche2
    T0 = 1. + gche * Rds
    T9 = Vdseff / T0
    Idl = gche * T9
    T9 = diffVds / Va
    T0 = 1. + T9
    Idsa = Idl * T0
    T9 = diffVds * rcpVASCBE
    T0 = 1. + T9
    Ids = Idsa * T0
    # substrate current (:706-726): the condassign at :720-723 is discarded,
    # so the factor is still T2 of the output-resistance block
    T1 = s.alpha0 + s.alpha1 * s.leff
    k_temp = condassign(T1, T2rout * Idsa, 0.)
    Isub = condassign(s.beta0, k_temp, 0.)
    iout = [swap * Ids * tf,
            condassign(swap, Isub, 0.) * tf,
            condassign(swap, 0., Isub) * tf]
    # ---- charges (:738-907) ----
    T0 = -.5 * s.dvt1w * s.weff * s.leff / s.factor1 / np.sqrt(s.Xdep0)
    T2 = float(condassign(T0 + EXP_THRESHOLD,
                          safe_exp(T0) * (1. + 2. * safe_exp(T0)), MIN_1))
    T0 = s.dvt0w * T2
    T2 = T0 * (s.vbi - s.phi)
    T0 = -.5 * s.dvt1 * s.leff / (s.factor1 * np.sqrt(s.Xdep0))
    T3 = float(condassign(T0 + EXP_THRESHOLD,
                          safe_exp(T0) * (1. + 2. * safe_exp(T0)), MIN_1))
    T3 = s.dvt0 * T3 * (s.vbi - s.phi)
    T4 = s.tox * s.phi / (s.weff + s.w0)
    # T0 is still the exponent argument here (:761)
    T5 = s.k1ox * (T0 - 1.) * s.sqrtPhi \
        + (s.kt1 + s.kt1l / s.leff) * s._ToTnm1
    T6 = s._vth0 - T2 - T3 + s.k3 * T4 + T5
    vfbzb = T6 - s.phi - s._k1 * s.sqrtPhi
    VbseffCV = condassign(Vbseff, s.phi - Phis, Vbseff)
    T0 = n * s.noff * s._Vt
    T1 = (Vgs_eff - Vth) / T0
    Vgsteff = condassign(T1 - EXP_THRESHOLD, Vgs_eff - Vth - s.voffcv,
                         T0 * log1pexp(T1))
    V3 = vfbzb - Vgs_eff + VbseffCV - .02
    T0 = sqrt(V3**2 + .08 * abs(vfbzb))
    Vfbeff = vfbzb - 0.5 * (V3 + T0)
    T0 = (Vgs_eff - VbseffCV - vfbzb) / s._Tox
    T1 = T0 * s.acde
    Tcen = condassign(EXP_THRESHOLD + T1, s.ldeb * exp(T1), s.ldeb * MIN_EXP)
    Tcen = condassign(-EXP_THRESHOLD + T1, s.ldeb * MAX_EXP, Tcen)
    V3 = s.ldeb - Tcen - 1e-3 * s.tox
    V4 = sqrt(V3**2 + 4e-3 * s.tox * s.ldeb)
    Tcen = s.ldeb - .5 * (V3 + V4)
    Ccen = EPSI / Tcen
    Coxeff = Ccen * s.cox / (Ccen + s.cox)
    CoxWLcen = Weff * s.leff * Coxeff
    Qac0 = CoxWLcen * (Vfbeff - vfbzb)
    T0 = .5 * s.k1ox
    T3 = Vgs_eff - Vfbeff - VbseffCV - Vgsteff
    T1 = condassign(-T3, T0 + T3 / s.k1ox, sqrt(T0**2 + T3))
    Qsub0 = CoxWLcen * s.k1ox * (T1 - T0)
    T2 = float(condassign(s.k1ox, s.moin * s._Vt * s.k1ox**2,
                          .25 * s.moin * s._Vt))
    T0 = float(condassign(s.k1ox, s.k1ox * np.sqrt(s.phi),
                          .5 * np.sqrt(s.phi)))
    T1 = 2. * T0 + Vgsteff
    DeltaPhi = s._Vt * log(1. + T1 * Vgsteff / T2)
    T3 = 4. * (Vth - vfbzb - s.phi)
    T0 = condassign(T3, .5 * (Vgsteff + T3) / s._Tox,
                    .5 * (Vgsteff + 1.0e-20) / s._Tox)
    k_temp = 2.01375270747048 * T0
    T1 = 1. + k_temp
    Tcen = 1.9e-9 / T1
    Ccen = EPSI / Tcen
    Coxeff = Ccen * s.cox / (Ccen + s.cox)
    CoxWLcen = Weff * s.leff * Coxeff
    AbulkCV = Abulk0 * s.AbulkCVfactor
    VdsatCV = (Vgsteff - DeltaPhi) / AbulkCV
    T0 = VdsatCV - VDS - .02
    T1 = sqrt(T0**2 + .08 * VdsatCV)
    VdseffCV = VdsatCV - .5 * (T0 + T1)
    VdseffCV = VdseffCV + 1e-200
    T0 = AbulkCV * VdseffCV
    T1 = Vgsteff - DeltaPhi
    T2 = 12. * (T1 - .5 * T0 + 1e-20)
    T3 = T0 / T2
    qgate = CoxWLcen * (T1 - T0 * (.5 - T3))
    qbulk = CoxWLcen * (1. - AbulkCV) * (.5 * VdseffCV - T0 * VdseffCV / T2)
    T2 = T2 / 12.
    T3 = .5 * CoxWLcen / (T2**2)
    T4 = T1 * (2. * T0**2 / 3. + T1 * (T1 - 4. * T0 / 3.)) \
        - 2. * T0**3 / 15.
    qsrc = -T3 * T4
    qgate = qgate + (Qac0 + Qsub0 - qbulk)
    qbulk = qbulk - (Qac0 + Qsub0)
    qdrn = -(qbulk + qgate + qsrc)
    qout = [condassign(swap, qdrn, qsrc) * tf, qgate * tf,
            condassign(swap, qsrc, qdrn) * tf]
    return iout, qout


# ---- extrinsic MOS wrapper (mosExt.py:281-324) ----------------------------
def mosext_cqs(s, v, intrinsic):
    iout, qout = intrinsic(s, v)
    v1 = -v[0] * s._tf
    if s.ad != 0.:
        iout[1] = iout[1] - junction_id(s.dj, v1) * s._tf
        if s.cj != 0.:
            qout[0] = qout[0] - junction_qd(s.dj, v1) * s._tf
    if s.pd != 0.:
        iout[1] = iout[1] - junction_id(s.djsw, v1) * s._tf
        if s.cjsw != 0.:
            qout[0] = qout[0] - junction_qd(s.djsw, v1) * s._tf
    v1 = -v[2] * s._tf
    if s.asrc != 0.:
        iout[2] = iout[2] - junction_id(s.sj, v1) * s._tf
        if s.cj != 0.:
            qout[2] = qout[2] - junction_qd(s.sj, v1) * s._tf
    if s.ps != 0.:
        iout[2] = iout[2] - junction_id(s.sjsw, v1) * s._tf
        if s.cjsw:
            qout[2] = qout[2] - junction_qd(s.sjsw, v1) * s._tf
    return [i * s.m for i in iout], [q * s.m for q in qout]


# ---- EKV 2.6 (mosEKV.py:19-58, 391-555) -----------------------------------
def f_simple(v):
    b = v + 20.
    d = exp(.5 * v)
    c = log(1. + d)
    f = condassign(b, c, d)
    return f**2


def f_accurate(v):
    i = f_simple(v)
    for _ in range(3):
        k1 = sqrt(0.25 + i)
        f = v - (2. * k1 - 1. + log(k1 - .5))
        vprime = -1. / (2. * k1 * (0.5 - k1)) + 1. / k1
        i = i + f / vprime
    return condassign(v + 20., i, exp(v))


def ekv_cqs(s, v):
    interp = f_simple if s.ekvint else f_accurate
    tf = s._tf
    vd, vg, vs = tf * v[0], tf * v[1], tf * v[2]
    swap = condassign(vd - vs, 1., -1.)
    vd, vs = condassign(swap, vd, vs), condassign(swap, vs, vd)
    Vt = s._Vt
    ga = s._gammaa
    vgprime = vg - s._vt0a - s._deltavRSCE + s.phiT + ga * sqrt(s.phiT)
    vp0 = vgprime - s.phiT - ga * (sqrt(vgprime + ga * ga / 4.) - ga / 2.)
    vp0 = condassign(vgprime, vp0, -s.phiT)
    tmp = 16. * Vt * Vt
    vsprime = 0.5 * (vs + s.phiT + sqrt((vs + s.phiT)**2 + tmp))
    vdprime = 0.5 * (vd + s.phiT + sqrt((vd + s.phiT)**2 + tmp))
    tmp = s.leta / s._leff * (sqrt(vsprime) + sqrt(vdprime))
    tmp = tmp - 3. * s.weta * sqrt(vp0 + s.phiT) / s._weff
    gammao = ga - EPSI / s.cox * tmp
    gammaprime = 0.5 * (gammao + sqrt(gammao * gammao + 0.1 * Vt))
    vp = vgprime - s.phiT
    vp = vp - gammaprime * (sqrt(vgprime + (.5 * gammaprime)**2)
                            - gammaprime / 2.)
    vp = condassign(vgprime, vp, -s.phiT)
    n = 1 + ga / (2. * sqrt(vp + s.phiT + 4. * Vt))
    i_f = interp((vp - vs) / Vt)
    vdss = sqrt(0.25 + Vt * sqrt(i_f) / s._vc) - 0.5
    vdss = vdss * s._vc
    tmp = sqrt(i_f) - 0.75 * log(i_f)
    vdssprime = sqrt(0.25 + Vt * tmp / s._vc) - 0.5
    vdssprime = vdssprime * s._vc
    vdssprime = vdssprime + Vt * (log(s._vc / (2. * Vt)) - 0.6)
    tmp = sqrt(i_f) - vdss / Vt
    deltav = sqrt(s.Lambda * tmp + 1. / 64)
    deltav = deltav * (4. * Vt)
    vds = .5 * (vd - vs)
    vip = sqrt(vdss * vdss + deltav * deltav)
    vip = vip - sqrt((vds - vdss)**2 + deltav * deltav)
    deltal = log(1. + (vds - vip) / s._lc / s._t_ucrit)
    deltal = deltal * (s.Lambda * s._lc)
    lprime = s.ns * s._leff - deltal + (vds + vip) / s._t_ucrit
    leq = 0.5 * (lprime + sqrt(lprime * lprime + s._lmin * s._lmin))
    tmp = vp - vds - vs
    tmp = tmp - sqrt(vdssprime * vdssprime + deltav * deltav)
    tmp = tmp + sqrt((vds - vdssprime)**2 + deltav * deltav)
    irprime = interp(tmp / Vt)
    ir = interp((vp - vd) / Vt)
    sqvpphi = sqrt(vp + s.phiT + 1.e-6)
    nq = 1. + ga / 2. / sqvpphi
    xf = sqrt(0.25 + i_f)
    xr = sqrt(0.25 + ir)
    tmp = (xf + xr)**2
    qd = 3. * xr**3 + 6. * xr * xr * xf + 4. * xr * xf * xf + 2. * xf**3
    qd = qd * (4. / 15. / tmp)
    qd = -nq * (qd - .5)
    qs = 3. * xf**3 + 6. * xf * xf * xr + 4. * xf * xr * xr + 2. * xr**3
    qs = qs * (4. / 15. / tmp)
    qs = -nq * (qs - .5)
    qi = qs + qd
    qb1 = -ga * sqvpphi / Vt
    qb1 = qb1 - (nq - 1.) * qi / nq
    qb = condassign(vgprime, qb1, -vgprime / Vt)
    qg = -qi - qb - s.qox
    betao = s.kpa * s.np * s._weff / leq
    betaoprime = 1. + s.cox * s._qb0 / s.e0 / EPSI
    betaoprime = betaoprime * betao
    tmp = D.dabs(qb + s.eta * qi)
    tmp = 1. + s.cox * Vt * tmp / s.e0 / EPSI
    beta = betaoprime / tmp
    IS = 2. * n * beta * Vt * Vt
    ids = IS * (i_f - irprime)
    vib = vd - vs - 2. * s.ibn * vdss
    idb1 = ids * s.iba * vib / s._t_ibb
    idb1 = idb1 * safe_exp(-s._t_ibb * s._lc / vib)
    idb = condassign(vib, idb1, 0.)
    k = s._Cox * Vt * tf
    qout = [condassign(swap, qd, qs) * k, qg * k,
            condassign(swap, qs, qd) * k]
    iout = [swap * ids * tf, condassign(swap, idb, 0.) * tf,
            condassign(swap, 0., idb) * tf]
    return iout, qout


# ---- MESFET (mesfetc.py:192-226) ------------------------------------------
def mesfetc_cqs(s, v):
    igs = junction_id(s.diogs, v[0])
    qgs = junction_qd(s.diogs, v[0])
    igd = junction_id(s.diogd, v[1])
    qgd = junction_qd(s.diogd, v[1])
    igs = igs - s.ib0 * exp(-(v[0] + s.vbd) / s._k6)
    igd = igd - s.ib0 * exp(-(v[1] + s.vbd) / s._k6)
    swap = condassign(v[0] - v[1], 1., -1.)
    vds = swap * (v[0] - v[1])
    vgsi = condassign(swap, v[0], v[1])
    vgsiT = condassign(swap, v[2], v[3])
    vx = vgsiT * (1. + s._Beta * (s.vds0 - vds))
    ids = (s.a0 + vx * (s.a1 + vx * (s.a2 + vx * s.a3))) \
        * D.tanh(s.gama * vds) * s._idsFac
    ids = condassign(vgsi - s._Vt0, ids, 0.)
    ids = condassign(ids, ids, 0.)
    return ([igs * s.area, igd * s.area, ids * swap * s.area],
            [qgs * s.area, qgd * s.area])


# ---- memristor (memristor.py:178-197) -------------------------------------
class _NP(object):
    """``np`` seen by the expression string: dual-aware functions."""
    exp = staticmethod(exp)
    log = staticmethod(log)
    sqrt = staticmethod(sqrt)
    tanh = staticmethod(D.tanh)
    tan = staticmethod(D.tan)
    sin = staticmethod(D.sin)
    cos = staticmethod(D.cos)
    cosh = staticmethod(D.cosh)
    abs = staticmethod(D.dabs)


def memr_cqs(s, v):
    q = s.c * v[1]
    M = eval(s.m, {'np': _NP, 'abs': D.dabs, 'q': q})
    return [dev_gyr()**2 * M * v[0]], []


# ---- electro-thermal wrapping (autoThermal.py:97-124) -------------------------
K_BOLTZ = 1.3806488e-23     # globalVars.py:26-39
Q_ELEC = 1.602176565e-19
T_ZERO = 273.15


class _Proxy(object):
    """Copy of a device / junction whose temperature-dependent attributes
    are replaced by Dual values (what the reference gets by calling
    set_temp_vars with an AD-typed temperature)."""

    def __init__(self, obj):
        self.__dict__.update(obj.__dict__)


def junction_temp_vars(j, Tabs, Tnomabs, vt, egapn, egap_t):
    """Junction.set_temp_vars, diode.py:53-72."""
    jt = _Proxy(j)
    tnratio = Tabs / Tnomabs
    jt._t_is = j.isat * D.power(tnratio, j._k1) * exp(j._k2 - j._k3 / Tabs)
    jt._kexp = vt * j.n
    if j.cj0 != 0.:
        jt._t_vj = j.vj * tnratio - 3. * vt * log(tnratio) \
            - tnratio * egapn + egap_t
        jt._t_cj0 = j.cj0 * (1. + j.m * (.0004 * (Tabs - Tnomabs)
                                         + 1. - jt._t_vj / j.vj))
        jt._k5 = jt._t_vj * jt._t_cj0 / j._k4
        jt._k6 = jt._t_cj0 * pow(1. - j.fc, - j.m - 1.)
        jt._k7 = ((1. - j.fc * (1. + j.m)) * jt._t_vj * j.fc
                  + .5 * j.m * jt._t_vj * j.fc * j.fc)
    return jt


def diode_temp_vars(dev, temp):
    """Device.set_temp_vars of the diode, diode.py:254-272."""
    d = _Proxy(dev)
    Tabs = temp + T_ZERO
    d.vt = K_BOLTZ * Tabs / Q_ELEC
    egap_t = dev.eg0 - .000702 * (Tabs * Tabs) / (Tabs + 1108.)
    d.jtn = junction_temp_vars(dev.jtn, Tabs, dev.Tnomabs, d.vt, dev.egapn,
                               egap_t)
    return d


def svjunction_temp_vars(j, Tabs, Tnomabs, vt, egapn, egap_t):
    """SVJunction.set_temp_vars, svdiode.py:53-78."""
    jt = junction_temp_vars(j, Tabs, Tnomabs, vt, egapn, egap_t)
    jt._alpha = 1. / j.n / vt
    max_exp_arg = 5e10
    jt._svth = log(max_exp_arg / jt._alpha) / jt._alpha
    jt._kexp = j.n * vt * max_exp_arg
    return jt


def bjt_temp_vars(dev, temp):
    """set_temp_vars of the intrinsic models (bjt.py:453-482,
    svbjt.py:276-305) and of the extrinsic wrapper (bjt.py:170-180)."""
    d = _Proxy(dev)
    sv = dev.devType.startswith('svbjt')
    jtv = svjunction_temp_vars if sv else junction_temp_vars
    Tabs = T_ZERO + temp
    tnratio = Tabs / dev.Tnomabs
    tnXTB = D.power(tnratio, dev.xtb)
    d.vt = K_BOLTZ * Tabs / Q_ELEC
    egap_t = dev.eg - .000702 * (Tabs * Tabs) / (Tabs + 1108.)
    args = (Tabs, dev.Tnomabs, d.vt, dev.egapn, egap_t)
    d.jif = jtv(dev.jif, *args)
    d.jir = jtv(dev.jir, *args)
    if dev.ise != 0.:
        d.jile = junction_temp_vars(dev.jile, *args)
        d.jile._t_is = d.jile._t_is / tnXTB
    if dev.isc != 0.:
        d.jilc = junction_temp_vars(dev.jilc, *args)
        d.jilc._t_is = d.jilc._t_is / tnXTB
    d._bf_t = dev.bf * tnXTB
    d._br_t = dev.br * tnXTB
    d.cbjtn = junction_temp_vars(dev.cbjtn, *args)
    return d


def bjt_power(d, v, i):
    """bjt.py:607-625 + extrinsic :208-222."""
    pRb = i[3] * v[2] if d.rb != 0. else 0.
    pout = i[0] * v[0] + i[1] * v[1] + i[2] * (v[0] - v[1]) + pRb
    if d._addCBjtn:
        pout = pout + v[d._bccn] * i[d._bcon]
    return pout


def svbjt_power(d, v, i):
    """svbjt.py:438-458 + extrinsic bjt.py:208-222."""
    gyrvce = i[1] - i[3]
    pRb = i[5] * v[2] if d.rb != 0. else 0.
    pout = (i[0] * i[1] + i[2] * i[3] + i[4] * gyrvce) / dev_gyr() + pRb
    if d._addCBjtn:
        pout = pout + v[d._bccn] * i[d._bcon]
    return pout


def svdiode_temp_vars(dev, temp):
    """svdiode.py:280-295."""
    d = _Proxy(dev)
    Tabs = temp + T_ZERO
    d.vt = K_BOLTZ * Tabs / Q_ELEC
    egap_t = dev.eg0 - .000702 * (Tabs * Tabs) / (Tabs + 1108.)
    d.jtn = svjunction_temp_vars(dev.jtn, Tabs, dev.Tnomabs, d.vt, dev.egapn,
                                 egap_t)
    return d


def res_temp_vars(dev, temp):
    """resistor.py:107-120 as executed once while the tape is recorded:
    g = g(process_params) / (1 + (tc1 + tc2 dT) dT)."""
    d = _Proxy(dev)
    deltaT = temp - dev.tnom
    d.g = dev.g / (1. + (dev.tc1 + dev.tc2 * deltaT) * deltaT)
    return d


def ekv_temp_vars(dev, temp):
    """mosEKV.py:329-357 + extrinsic wrapper mosExt.py:252-278."""
    d = _Proxy(dev)
    T = T_ZERO + temp
    d._Vt = K_BOLTZ * T / Q_ELEC
    vt0T = dev._vt0 - dev._tcv * (T - dev._Tn)
    d._vt0a = vt0T + dev.avto / dev._sga
    kpT = dev.kp * D.power(T / dev._Tn, dev.bex)
    kpa = kpT * (1 + dev.akp / dev._sga)
    d.kpa = condassign(kpa, kpa, 0.)
    d._t_ucrit = dev.ucrit * D.power(T / dev._Tn, dev.ucex)
    egT = 1.16 - 0.000702 * T * T / (T + 1108.)
    d.phiT = dev.phi * T / dev.Tref - 3. * d._Vt * log(T / dev.Tref) \
        - dev.egTref * T / dev.Tref + egT
    d._t_ibb = dev.ibb * (1. + dev.ibbt * (T - dev.Tref))
    d._vc = d._t_ucrit * dev._leff * dev.ns
    d._qb0 = dev._gammaa * sqrt(d.phiT)
    egap_t = dev.eg0 - .000702 * (T * T) / (T + 1108.)
    for name in ('dj', 'djsw', 'sj', 'sjsw'):
        j = getattr(dev, name)
        if j is not None:
            setattr(d, name, junction_temp_vars(j, T, dev._Tnabs, d._Vt,
                                                dev._egapn, egap_t))
    return d


def ekv_power(d, v, i):
    """mosEKV.py:586-595."""
    return (v[0] - v[2]) * i[0] + v[0] * i[1] + v[2] * i[2]


def mesfetc_temp_vars(dev, temp):
    """mesfetc.py:160-189."""
    d = _Proxy(dev)
    Tabs = T_ZERO + temp
    d.vt = K_BOLTZ * Tabs / Q_ELEC
    egap_t = dev.eg0 - .000702 * (Tabs * Tabs) / (Tabs + 1108.)
    args = (Tabs, dev.Tnomabs, d.vt, dev.egapn, egap_t)
    d.diogs = junction_temp_vars(dev.diogs, *args)
    d.diogd = junction_temp_vars(dev.diogd, *args)
    delta_T = temp - dev.tnom
    d._k6 = dev.nr * d.vt
    d._Vt0 = dev.vt0 * (1. + (delta_T * dev.avt0)
                        + (delta_T * delta_T * dev.bvt0))
    d._Beta = dev.beta * pow(1.01, (delta_T * dev.tbet)) if dev.tbet \
        else dev.beta
    d._idsFac = D.power(1. + delta_T * dev.tm, dev.tme) \
        if dev.tme * dev.tm != 0. else 1.
    return d


def mesfetc_power(d, v, i):
    """mesfetc.py:233-242."""
    return (v[0] - v[1]) * i[2] + v[0] * i[0] + v[1] * i[1]


_THERMAL = {}       # base devType -> (set_temp_vars, eval_cqs, power)


def thermal_cqs(dev, v):
    """ThermalDevice.eval_cqs, autoThermal.py:97-124: the last regular
    control port is the temperature rise; outputs get the power appended."""
    tempvars, base_cqs, power = _THERMAL[dev.devType[:-2]]
    tpn = dev._tpn
    temp = v[tpn] + dev._tamb
    d = tempvars(dev, temp)
    v1 = [x for k, x in enumerate(v) if k != tpn]
    iV, qV = base_cqs(d, v1)
    return list(iV) + [power(d, v1, iV)], qV


# ---- dispatch ---------------------------------------------------------------
def eval_cqs(dev, v):
    t = dev.devType
    if t.endswith('_t') and t[:-2] in _THERMAL:
        return thermal_cqs(dev, v)
    if t == 'diode':
        return diode_cqs(dev, v)
    if t == 'svdiode':
        return svdiode_cqs(dev, v)
    if t == 'bjt':
        return extrinsic_bjt_cqs(dev, v, bjt_cqs)
    if t == 'svbjt':
        return extrinsic_bjt_cqs(dev, v, svbjt_cqs)
    if t == 'bsim3_i':
        return bsim3_cqs(dev, v)
    if t == 'mosbsim3':
        return mosext_cqs(dev, v, bsim3_cqs)
    if t == 'ekv_i':
        return ekv_cqs(dev, v)
    if t == 'mosekv':
        return mosext_cqs(dev, v, ekv_cqs)
    if t == 'mesfetc':
        return mesfetc_cqs(dev, v)
    if t == 'memr':
        return memr_cqs(dev, v)
    raise NotImplementedError('oracle: no model for ' + t)


_THERMAL['diode'] = (diode_temp_vars, diode_cqs,
                     lambda d, v, i: v[0] * i[0])       # diode.py:310-316
_THERMAL['svdiode'] = (svdiode_temp_vars, svdiode_cqs,
                       lambda d, v, i: i[0] * i[1] / dev_gyr())
_THERMAL['res'] = (res_temp_vars, lambda d, v: ([d.g * v[0]], []),
                   lambda d, v, i: v[0] * i[0])           # resistor.py:123-138
_THERMAL['mosekv'] = (ekv_temp_vars,
                      lambda d, v: mosext_cqs(d, v, ekv_cqs), ekv_power)
_THERMAL['mesfetc'] = (mesfetc_temp_vars, mesfetc_cqs, mesfetc_power)
_THERMAL['bjt'] = (bjt_temp_vars,
                   lambda d, v: extrinsic_bjt_cqs(d, v, bjt_cqs), bjt_power)
_THERMAL['svbjt'] = (bjt_temp_vars,
                     lambda d, v: extrinsic_bjt_cqs(d, v, svbjt_cqs),
                     svbjt_power)


def device_eval(dev, xin, deriv=True, gyr=1e-3):
    """out[nout, n] (and jac[nout, nxin, n]) at xin[nxin, n].

    Stand-in for cppaddev.eval (:160) / eval_and_deriv (:134).
    """
    _GYR[0] = gyr
    xin = np.asarray(xin, dtype=float)
    if xin.ndim == 1:
        xin = xin[:, None]
    P, n = xin.shape
    with np.errstate(all='ignore'):
        if deriv:
            iout, qout = eval_cqs(dev, D.seed(xin))
            return D.jac_rows(iout + qout, P, n)
        iout, qout = eval_cqs(dev, [xin[k].copy() for k in range(P)])
        outs = iout + qout
        vals = np.empty((len(outs), n))
        for r, o in enumerate(outs):
            vals[r] = o
        return vals, None
