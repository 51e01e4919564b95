"""
BSIM3 v3.2.4 MOSFET: intrinsic ``bsim3_i`` and extrinsic ``mosbsim3``.

Plugin interface and parameter processing follow the reference's
devices/mosBSIM3v3.py (BSIM3 :35, paramDict :110-223, process_params
:249-374, set_temp_vars :376-401).  The 500-line eval_cqs (:403-911) is
implemented by csrc/models/bsim3.cuh; ``cuda_pack`` folds every sub-expression
of it that depends on parameters only (same FP64 operation order as the
reference evaluates them, so the folded values are bit-identical) and hands
the kernel 78 scalars per instance.

Terminal order: 0 drain, 1 gate, 2 source, 3 bulk.
"""
import numpy as np

from ..globalVars import const
from .. import circuit as cir
from . import cudadev as ad
from . import mosExt
from .junction import EMPTY
from .param_layout import flag_bits, LAYOUTS

_F = flag_bits('bsim3')

MAX_EXP = 5.834617425e14
MIN_EXP = 1.713908431e-15
MIN_1 = MIN_EXP * (1. + 2. * MIN_EXP)
EXP_THRESHOLD = 34.0


def _safe_exp(x):
    """exp with linear continuation above 50 (reference cppaddev.py:56)."""
    if np.ndim(x) == 0:
        if x < 50.:
            return np.exp(x)
        return np.exp(50.) * (x - 50. + 1.)
    with np.errstate(over='ignore'):
        return np.where(x < 50., np.exp(x), np.exp(50.) * (x - 50. + 1.))


class BSIM3(ad.CudaModel, cir.Element):
    """
    Intrinsic BSIM3 MOSFET Model (version 3.2.4)
    --------------------------------------------

    Default parameters listed for NMOS type; ``u0`` defaults to 250 for
    PMOS.  Currents ``[ids, idb, isb]``, charges ``[qd, qg, qs]``, controls
    ``[vdb, vgb, vsb]``.

    Netlist examples::

        bsim3_i:m1 2 3 4 gnd w=10e-6 l=1e-6 type = n
        bsim3_i:m2 4 5 6 6 w=30e-6 l=1e-6 type = p
    """
    category = "Semiconductor devices"
    devType = "bsim3_i"
    cudaModel = 'bsim3'
    numTerms = 4
    isNonlinear = True
    csOutPorts = [(0, 2), (0, 3), (2, 3)]
    controlPorts = [(0, 3), (1, 3), (2, 3)]
    vPortGuess = np.array([0., 0., 0.])
    qsOutPorts = [(0, 3), (1, 3), (2, 3)]

    paramDict = dict(
        cir.Element.tempItem,
        type=('N- or P-channel MOS (n or p)', '', str, 'n'),
        k1enable=('Enable k1, k2 internal calculation', '', bool, False),
        vth0=('Threshold voltage of long channel device at Vbs=0 and small '
              'Vds', 'V', float, 0.7),
        l=('Length', 'm', float, 1e-06),
        w=('Width', 'm', float, 1e-06),
        tox=('Gate oxide thickness', 'm', float, 1.5e-08),
        toxm=('Gate oxide thickness used in extraction', '', float, 1.5e-08),
        cdsc=('Drain/Source and channel coupling capacitance', 'F/m^2',
              float, 0.00024),
        cdscb=('Body-bias dependence of cdsc', 'F/V/m^2', float, 0),
        cdscd=('Drain-bias dependence of cdsc', 'F/V/m^2', float, 0),
        cit=('Interface state capacitance', '', float, 0),
        nfactor=('Subthreshold swing coefficient', '', float, 1),
        xj=('Junction depth', 'm', float, 1.5e-07),
        vsat=('Saturationvelocity at tnom', 'm/s', float, 80000),
        at=('Temperature coefficient of vsat', 'm/s', float, 33000),
        a0=('Non-uniform depletion width effect coefficient', '', float, 1),
        ags=('Gate bias coefficient of Abulk', '', float, 0),
        a1=('Non-saturation effect coefficient', '', float, 0),
        a2=('Non-saturation effect coefficient', '', float, 1),
        keta=('Body-bias coefficient of non-uniform depletion width effect',
              '', float, -0.047),
        nsub=('Substrate doping concentration', 'cm^{-3}', float, 6e+16),
        nch=('Channel doping concentration', 'cm^{-3}', float, 1.7e+17),
        ngate=('Poly-gate doping concentration', 'cm^{-3}', float, 0),
        vbm=('Maximum body voltage', 'V', float, -3),
        xt=('Doping depth', 'm', float, 1.55e-07),
        kt1=('Temperature coefficient of Vth', 'V', float, -0.11),
        kt1l=('Temperature coefficient of Vth', 'V m', float, 0),
        kt2=('Body-coefficient of kt1', '', float, 0.022),
        k3=('Narrow width effect coefficient', '', float, 80),
        k3b=('Body effect coefficient of k3', '', float, 0),
        w0=('Narrow width effect parameter', 'm', float, 2.5e-06),
        nlx=('Lateral non-uniform doping effect', 'm', float, 1.74e-07),
        dvt0=('Short channel effect coefficient 0', '', float, 2.2),
        dvt1=('Short channel effect coefficient 1', '', float, 0.53),
        dvt2=('Short channel effect coefficient 2', 'V^{-1}', float, -0.032),
        dvt0w=('Narrow width effect coefficient 0', 'm^{-1}', float, 0),
        dvt1w=('Narrow width effect coefficient 1', 'm^{-1}', float,
               5.3e+06),
        dvt2w=('Narrow width effect coefficient 2', 'V^{-1}', float, -0.032),
        drout=('DIBL coefficient of output resistance', '', float, 0.56),
        dsub=('DIBL coefficient in the subthreshold region', '', float,
              0.56),
        ua=('Linear gate dependence of mobility', 'm/V', float, 2.25e-09),
        ub=('Quadratic gate dependence of mobility', '(m/V)^2', float,
            5.87e-19),
        uc=('Body-bias dependence of mobility', 'm/V^2', float, -4.65e-11),
        u0=('Low-field mobility at Tnom', 'cm^2/V/s', float, 670),
        voff=('Threshold voltage offset', 'V', float, -0.08),
        tnom=('Nominal temperature', 'C', float, 27.),
        elm=('Non-quasi-static Elmore Constant Parameter', '', float, 5),
        delta=('Effective Vds parameter', 'V', float, 0.01),
        rdsw=('Sorce-drain resistance per width', '', float, 0),
        prwg=('Gate-bias effect on parasitic resistance', '', float, 0),
        prwb=('Body-effect on parasitic resistance', '', float, 0),
        prt=('Temperature coefficient of parasitic resistance', '', float,
             0),
        eta0=('Subthreshold region DIBL coefficeint', '', float, 0.08),
        etab=('Subthreshold region DIBL coefficeint', '', float, -0.07),
        pclm=('Channel length modulation coefficient', '', float, 1.3),
        pdibl1=('Drain-induced barrier lowering oefficient', '', float,
                0.39),
        pdibl2=('Drain-induced barrier lowering oefficient', '', float,
                0.0086),
        pdiblb=('Body-effect on drain induced barrier lowering', '', float,
                0),
        pscbe1=('Substrate current body-effect coeffiecient 1', 'V/m',
                float, 4.24e+08),
        pscbe2=('Substrate current body-effect coeffiecient 2', 'V/m',
                float, 1e-05),
        pvag=('Gate dependence of output resistance parameter', '', float,
              0),
        vfb=('Flat band voltage', 'V', float, -1),
        acde=('Exponential coefficient for finite charge thickness', '',
              float, 1),
        moin=('Coefficient for gate-bias dependent surface potential', '',
              float, 15),
        noff=('C-V turn-on/off parameter', '', float, 1),
        voffcv=('C-V lateral shift parameter', '', float, 0),
        lint=('Length reduction parameter', 'm', float, 0),
        ll=('Length reduction parameter', '', float, 0),
        llc=('Length reduction parameter for CV', '', float, 0),
        lln=('Length reduction parameter', '', float, 1),
        lw=('Length reduction parameter', '', float, 0),
        lwc=('Length reduction parameter for CV', '', float, 0),
        lwn=('Length reduction parameter', '', float, 1),
        lwl=('Length reduction parameter', '', float, 0),
        lwlc=('Length reduction parameter for CV', '', float, 0),
        wr=('Width dependence of rds', '', float, 1),
        wint=('Width reduction parameter', 'm', float, 0),
        dwg=('Width reduction parameter', 'm/V', float, 0),
        dwb=('Width reduction parameter', 'm/V', float, 0),
        wl=('Width reduction parameter', '', float, 0),
        wlc=('Width reduction parameter for CV', '', float, 0),
        wln=('Width reduction parameter', '', float, 1),
        ww=('Width reduction parameter', '', float, 0),
        wwc=('Width reduction parameter for CV', '', float, 0),
        wwn=('Width reduction parameter', '', float, 1),
        wwl=('Width reduction parameter', '', float, 0),
        wwlc=('Width reduction parameter for CV', '', float, 0),
        b0=('Abulk narrow width parameter', '', float, 0),
        b1=('Abulk narrow width parameter', '', float, 0),
        clc=('Vdsat paramater for C-V model', '', float, 1e-07),
        cle=('Vdsat paramater for C-V model', '', float, 0.6),
        alpha0=('Substrate current model parameter', 'm/V', float, 0),
        alpha1=('Substrate current model parameter', 'V^{-1}', float, 0),
        beta0=('Diode limiting current', 'V', float, 30),
        ute=('Temperature coefficient of mobility', '', float, -1.5),
        k1=('First order body effect coefficient', 'V^{0.5}', float, 0.53),
        k2=('Second order body effect coefficient', '', float, -0.0186),
        ua1=('Temperature coefficient for ua', 'm/V', float, 4.31e-09),
        ub1=('Temperature coefficient for ub', '(m/V)^2', float, -7.61e-18),
        uc1=('Temperature coefficient for uc', 'm/V^2', float, -5.6e-11),
    )

    def __init__(self, instanceName):
        cir.Element.__init__(self, instanceName)

    def process_params(self, thermal=False):
        ad.delete_tape(self)
        if self.type == 'n':
            self._tf = 1.
        elif self.type == 'p':
            self._tf = -1.
            if not self.is_set('u0'):
                self.u0 = 250
        else:
            raise cir.CircuitError(
                '{0}: unrecognized type: {1}. Valid types are "n" or "p"'
                .format(self.instanceName, self.type))
        self._Tn = const.T0 + self.tnom
        self._Vtn = const.k * self._Tn / const.q
        self.factor1 = np.sqrt(const.epSi / const.epOx * self.tox)
        Eg0 = 1.16 - 7.02e-4 * (self._Tn**2) / (self._Tn + 1108.0)
        ni = 1.45e10 * (self._Tn / 300.15) * np.sqrt(self._Tn / 300.15) \
            * np.exp(21.5565981 - Eg0 / (2. * self._Vtn))
        # effective channel length / width
        self.ldrn = self.l
        self.wdrn = self.w
        t0 = ad.power(self.ldrn, self.lln)
        t1 = ad.power(self.wdrn, self.lwn)
        self.dl = self.lint + (self.ll / t0 + self.lw / t1
                               + self.lwl / (t0 * t1))
        t2 = ad.power(self.ldrn, self.wln)
        t3 = ad.power(self.wdrn, self.wwn)
        self.dw = self.wint + (self.wl / t2 + self.ww / t3
                               + self.wwl / (t2 * t3))
        self.leff = self.l - 2.0 * self.dl
        self.weff = self.w - 2.0 * self.dw
        self.AbulkCVfactor = (1. + ad.power(self.clc / self.leff, self.cle))
        self.cox = const.epOx / self.tox
        self.phi = 2. * self._Vtn * np.log(self.nch / ni)
        self.sqrtPhi = np.sqrt(self.phi)
        self.phis3 = self.sqrtPhi * self.phi
        self.Xdep0 = np.sqrt(2. * const.epSi
                             / (const.q * self.nch * 1e6)) * self.sqrtPhi
        self.litl = np.sqrt(3. * self.xj * self.tox)
        self.vbi = self._Vtn * np.log(1.0e20 * self.nch / (ni**2))
        self.cdep0 = np.sqrt(const.q * const.epSi * self.nch * 1e6
                             / 2. / self.phi)
        if not self.is_set('toxm'):
            self.toxm = self.tox
        if not self.is_set('dsub'):
            self.dsub = self.drout
        self.ldeb = np.sqrt(const.epSi * self._Vtn
                            / (const.q * self.nch * 1e6)) / 3.
        if (self.k1enable != 0.) and \
                not (self.is_set('k1') or self.is_set('k2')):
            vbx = -abs(self.phi - 7.7348e-4 * self.nch * self.xt**2)
            Vbm = -abs(self.vbm)
            gamma1 = 5.753e-12 * np.sqrt(self.nch) / self.cox
            gamma2 = 5.753e-12 * np.sqrt(self.nsub) / self.cox
            T0 = gamma1 - gamma2
            T1 = np.sqrt(self.phi - vbx) - self.sqrtPhi
            T2 = np.sqrt(self.phi * (self.phi - Vbm)) - self.phi
            self._k2 = T0 * T1 / (2. * T2 + Vbm)
            self._k1 = gamma2 - 2. * self._k2 * np.sqrt(self.phi - Vbm)
        else:
            self._k1 = self.k1
            self._k2 = self.k2
        if not self.is_set('vth0'):
            self._vth0 = self.vfb + self.phi + self._k1 * self.sqrtPhi
        else:
            self._vth0 = abs(self.vth0)
        self.k1ox = self._k1 * self.tox / self.toxm
        self.k2ox = self._k2 * self.tox / self.toxm
        t1 = np.sqrt(const.epSi / const.epOx * self.tox * self.Xdep0)
        t0 = _safe_exp(-0.5 * self.dsub * self.leff / t1)
        self.theta0vb0 = (t0 + 2.0 * t0**2)
        self._Tox = 1e8 * self.tox
        # (ad.sel / ad.uniform: scalar ``if`` or, in batch mode, np.where)
        if ad.uniform(self._k2 < 0., 'sign of k2'):
            vbsc = .9 * (self.phi - (.5 * self._k1 / self._k2)**2)
            self.vbsc = ad.sel(vbsc > -3., -3., ad.sel(vbsc < -30., -30.,
                                                       vbsc))
        else:
            self.vbsc = -30.
        self._u0 = ad.sel(self.u0 > 1., self.u0 * 1e-4, self.u0)
        if not thermal:
            self.set_temp_vars(self.temp)

    def set_temp_vars(self, temp):
        ad.delete_tape(self)
        absTemp = const.T0 + temp
        self._Vt = const.k * absTemp / const.q
        self._ToTnm1 = absTemp / self._Tn - 1.
        self.vsattemp = self.vsat - self.at * self._ToTnm1
        self.rds0 = (self.rdsw + self.prt * self._ToTnm1) \
            / ad.power(self.weff * 1e6, self.wr)
        self.V0 = self.vbi - self.phi
        self._ua = self.ua + self.ua1 * self._ToTnm1
        self._ub = self.ub + self.ub1 * self._ToTnm1
        self._uc = self.uc + self.uc1 * self._ToTnm1
        self.u0temp = self._u0 * ad.power(absTemp / self._Tn, self.ute)

    def cuda_pack(self):
        """Kernel record: attributes + folded parameter-only expressions.

        Each folded value is written exactly as the reference's eval_cqs
        computes it (line numbers of mosBSIM3v3.py in the comments).
        """
        s = self
        epSi, epOx = const.epSi, const.epOx
        v = dict(
            tf=s._tf, vbsc=s.vbsc, phi=s.phi, phis3=s.phis3, Xdep0=s.Xdep0,
            sqrtPhi=s.sqrtPhi, dvt2=s.dvt2, dvt2w=s.dvt2w,
            factor1=s.factor1, dvt0=s.dvt0, V0=s.V0, dvt0w=s.dvt0w,
            kt2=s.kt2, ToTnm1=s._ToTnm1, eta0=s.eta0, etab=s.etab,
            theta0vb0=s.theta0vb0, k1ox=s.k1ox, k2ox=s.k2ox, k3=s.k3,
            k3b=s.k3b, cdsc=s.cdsc, cdscb=s.cdscb, cdscd=s.cdscd, cit=s.cit,
            cox=s.cox, Vt=s._Vt, cdep0=s.cdep0, dwg=s.dwg, dwb=s.dwb,
            weff=s.weff, prwg=s.prwg, prwb=s.prwb, rds0=s.rds0, xj=s.xj,
            leff=s.leff, a0=s.a0, keta=s.keta, ua=s._ua, ub=s._ub, uc=s._uc,
            tox=s.tox, u0temp=s.u0temp, vsattemp=s.vsattemp, a2=s.a2,
            delta=s.delta, pclm=s.pclm, litl=s.litl, pdiblb=s.pdiblb,
            pvag=s.pvag, pscbe2=s.pscbe2, noff=s.noff, voffcv=s.voffcv,
            Tox=s._Tox, acde=s.acde, ldeb=s.ldeb, epSi=epSi,
            AbulkCVfactor=s.AbulkCVfactor)
        v['c_dvt1'] = -.5 * s.dvt1 * s.leff                       # :465
        v['c_dvt1w'] = -.5 * s.dvt1w * s.weff * s.leff            # :478
        T0 = np.sqrt(1. + s.nlx / s.leff)                         # :488
        v['sqrt1nlx'] = T0
        v['t1a'] = s.k1ox * (T0 - 1.) * s.sqrtPhi                 # :489
        v['kt1eff'] = s.kt1 + s.kt1l / s.leff                     # :490
        v['tmp2'] = s.tox * s.phi / (s.weff + s.w0)               # :492
        v['vth0mk1'] = s._vth0 - s._k1 * s.sqrtPhi                # :499
        v['nfepSi'] = s.nfactor * epSi                            # :504
        v['vtcdepcox'] = s._Vt * s.cdep0 / s.cox                  # :535
        v['b0b1'] = s.b0 / (s.weff + s.b1)                        # :555
        v['agsa0'] = s.ags * s.a0                                 # :559
        # output-resistance DIBL factor (:645-649)
        T1 = np.sqrt(epSi / epOx * s.tox * s.Xdep0)
        T0 = _safe_exp(-.5 * s.drout * s.leff / T1)
        T2 = T0 + 2. * T0**2
        v['thetaRout'] = s.pdibl1 * T2 + s.pdibl2
        # the substrate-current block multiplies Idsa by this same T2
        # (the value of its own condassign is discarded, :720-725)
        v['t2sub'] = T2
        v['npscbe1litl'] = -s.pscbe1 * s.litl                     # :676
        # zero-bias flat-band voltage (:741-766; T5 uses the exponent
        # argument T0 of :751, as the reference does)
        T0 = -.5 * s.dvt1w * s.weff * s.leff / s.factor1 / np.sqrt(s.Xdep0)
        T2 = ad.sel(T0 + EXP_THRESHOLD > 0.,
                    _safe_exp(T0) * (1. + 2. * _safe_exp(T0)), MIN_1)
        T0 = s.dvt0w * T2
        T2 = T0 * (s.vbi - s.phi)
        T0 = -.5 * s.dvt1 * s.leff / (s.factor1 * np.sqrt(s.Xdep0))
        T3 = ad.sel(T0 + EXP_THRESHOLD > 0.,
                    _safe_exp(T0) * (1. + 2. * _safe_exp(T0)), MIN_1)
        T3 = s.dvt0 * T3 * (s.vbi - s.phi)
        T4 = s.tox * s.phi / (s.weff + s.w0)
        T5 = s.k1ox * (T0 - 1.) * s.sqrtPhi \
            + (s.kt1 + s.kt1l / s.leff) * s._ToTnm1
        T6 = s._vth0 - T2 - T3 + s.k3 * T4 + T5
        vfbzb = T6 - s.phi - s._k1 * s.sqrtPhi
        v['vfbzb'] = vfbzb
        v['avfbzb08'] = .08 * abs(vfbzb)                          # :786
        v['c1tox'] = 1e-3 * s.tox                                 # :798
        v['c2toxldeb'] = 4e-3 * s.tox * s.ldeb                    # :799
        # surface-potential constants (:831-837)
        v['t2dp'] = ad.sel(s.k1ox > 0., s.moin * s._Vt * s.k1ox**2,
                           .25 * s.moin * s._Vt)
        v['t0dp'] = ad.sel(s.k1ox > 0., s.k1ox * np.sqrt(s.phi),
                           .5 * np.sqrt(s.phi))
        flags = 0
        if ad.uniform(s.pscbe1 != 0., 'pscbe1'):
            flags |= _F['PSCBE1']
        if ad.uniform((s.alpha0 + s.alpha1 * s.leff > 0.) & (s.beta0 > 0.),
                      'substrate current'):                        # :725
            flags |= _F['ISUB']
        names = LAYOUTS['bsim3'][0]
        nint = len(names) - 1 - 4 * len(EMPTY)
        rec = [v[n] for n in names[:nint]]
        # intrinsic device: multiplier 1, no junctions
        return flags, ad.stack_record(rec, [1.], *(4 * [EMPTY]))

    def power(self, vPort, currV):
        vds = vPort[0] - vPort[2]
        return vds * currV[0] + vPort[0] * currV[1] + vPort[2] * currV[2]

    def get_OP(self, vPort):
        outV = self.eval(vPort)
        return dict(temp=self.temp, VD=vPort[0], VG=vPort[1], VS=vPort[2],
                    IDS=outV[0])


ExtBSIM3 = mosExt.extrinsic_mos(BSIM3)

devList = [BSIM3, ExtBSIM3]
