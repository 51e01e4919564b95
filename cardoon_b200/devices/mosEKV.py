"""
EPFL EKV 2.6 MOSFET: intrinsic ``ekv_i`` and extrinsic ``mosekv``.

Plugin interface and parameter processing follow the reference's
devices/mosEKV.py (IntEKV :63, process_params :236-325, set_temp_vars
:327-389).  eval_cqs (:391-555) with the interpolation functions f_simple
(:19) / f_accurate (:35) is implemented by csrc/models/ekv.cuh; parameter-only
sub-expressions are folded in ``cuda_pack`` in the reference's operation
order.

Terminal order: 0 drain, 1 gate, 2 source, 3 bulk.
"""
import numpy as np

from ..globalVars import const
from .. import circuit as cir
from . import cudadev as ad
from . import mosExt
from .junction import EMPTY
from .param_layout import flag_bits, LAYOUTS

_F = flag_bits('ekv')


class IntEKV(ad.CudaModel, cir.Element):
    """
    Intrinsic EPFL EKV 2.6 MOSFET
    -----------------------------

    Currents ``[ids, idb, isb]``, charges ``[qd, qg, qs]``, controls
    ``[vdb, vgb, vsb]``.  ``ekvint = 0`` selects the accurate interpolation
    function (three Newton refinements), any other value the simple one.

    Netlist examples::

        ekv_i:m1 2 3 4 gnd w=30e-6 l=1e-6 type = n ekvint=0
        .model ekvn ekv_i (type = n kp = 200u theta = 0.6)
    """
    category = "Semiconductor devices"
    devType = "ekv_i"
    cudaModel = 'ekv'
    makeAutoThermal = True
    numTerms = 4
    isNonlinear = True

    paramDict = dict(
        cir.Element.tempItem,
        type=('N- or P-channel MOS (n or p)', '', str, 'n'),
        l=('Gate length', 'm', float, 1e-06),
        w=('Gate width', 'm', float, 1e-06),
        np=('Parallel multiple device number', '', float, 1.),
        ns=('Serial multiple device number', '', float, 1.),
        cox=('Gate oxide capacitance per area', 'F/m^2', float, 0.7e-3),
        xj=('Junction depth', 'm', float, 1e-07),
        dw=('Channel width correction', 'm', float, 0.),
        dl=('Channel length correction', 'm', float, 0.),
        vt0=('Long_channel threshold voltage', 'V', float, 0.5),
        gamma=('Body effect parameter', 'V^1/2', float, 1.),
        phi=('Bulk Fermi potential', 'V', float, 0.7),
        kp=('Transconductance parameter', 'A/V^2', float, 50e-6),
        e0=('Mobility reduction coefficient', 'V/m', float, 1e12),
        ucrit=('Longitudinal critical field', 'V/m', float, 2e6),
        tox=('Oxide thickness', 'm', float, None),
        nsub=('Channel doping', '1/cm^3', float, None),
        vfb=('Flat-band voltage', 'V', float, None),
        u0=('Low-field mobility', 'cm^2/(V.s)', float, None),
        vmax=('Saturation velocity', 'm/s', float, None),
        theta=('Mobility recuction coefficient', '1/V', float, 0.),
        Lambda=('Channel-length modulation', '', float, 0.5),
        weta=('Narrow-channel effect coefficient', '', float, 0.25),
        leta=('Short-channel effect coefficient', '', float, 0.1),
        q0=('Reverse short channel effect peak charge density', 'A.s/m^2',
            float, 0.),
        lk=('Reverse short channel effect characteristic length', 'm',
            float, 2.9e-07),
        iba=('First impact ionization coefficient', '1/m', float, 0.),
        ibb=('Second impact ionization coefficient', 'V/m', float, 3e+08),
        ibn=('Saturation voltage factor for impact ionization', '', float,
             1.),
        tcv=('Threshold voltage temperature coefficient', 'V/K', float,
             0.001),
        bex=('Mobility temperature exponent', '', float, -1.5),
        ucex=('Longitudinal critical field temperature exponent', '', float,
              0.8),
        ibbt=('Temperature coefficient for IBB', '1/K', float, 0.0009),
        avto=('Area related threshold voltage mismatch parameter', 'Vm',
              float, 0.),
        akp=('Area related gain mismatch parameter', 'm', float, 0.),
        agamma=('Area related body effect mismatch parameter', 'V^(1/2)m',
                float, 0.),
        kf=('Flicker noise coefficient', '', float, 0.),
        af=('Flicker noise exponent', '', float, 1.),
        satlim=('Ratio defining the saturation limit if/ir', '', float,
                54.5982),
        tnom=('Nominal temperature of model parameters', 'C', float, 27.),
        ekvint=('Interpolation function (0: accurate, 1: simple)', '', int,
                0),
    )
    csOutPorts = [(0, 2), (0, 3), (2, 3)]
    controlPorts = [(0, 3), (1, 3), (2, 3)]
    vPortGuess = np.array([0., 0., 0.])
    qsOutPorts = [(0, 3), (1, 3), (2, 3)]
    noisePorts = [(0, 2)]

    def __init__(self, instanceName):
        cir.Element.__init__(self, instanceName)
        self.qox = 0.

    def process_params(self, thermal=False):
        ad.delete_tape(self)
        self.Tref = 300.15
        self._Tn = self.tnom + const.T0
        vtTref = (const.k * self.Tref) / const.q
        self.egTref = 1.16 - 0.000702 * self.Tref * self.Tref \
            / (self.Tref + 1108.)
        vtTnom = (const.k * self._Tn) / const.q
        self.egTnom = 1.16 - .000702 * self._Tn * self._Tn \
            / (self._Tn + 1108.)
        niTnom = 1.45e10 * (self._Tn / self.Tref) \
            * np.exp(self.egTref / (2. * vtTref)
                     - self.egTnom / (2. * vtTnom))
        self._tcv = np.abs(self.tcv)
        if self.type == 'n':
            self.eta = 0.5
            self._tf = 1.
        elif self.type == 'p':
            self.eta = 1. / 3.
            self._tf = -1.
        else:
            raise cir.CircuitError(
                '{0}: unrecognized type: {1}. Valid types are "n" or "p"'
                .format(self.instanceName, self.type))
        # parameters derived from process data when not given explicitly
        if (not self.is_set('cox')) and self.tox is not None:
            self.cox = const.epOx / self.tox
        if (not self.is_set('gamma')) and self.nsub is not None:
            self.gamma = np.sqrt(2. * const.q * const.epSi
                                 * self.nsub * 1.e6) / self.cox
        if (not self.is_set('phi')) and self.nsub is not None:
            self.phi = 2. * vtTnom * np.log(self.nsub / niTnom)
        if not self.is_set('vt0'):
            if self.vfb is not None:
                self.vt0 = self._tf * (self.vfb + self.phi
                                       + self.gamma * np.sqrt(self.phi))
            else:
                self.vt0 *= self._tf
        self._vt0 = abs(self.vt0)
        if (not self.is_set('kp')) and self.u0 is not None:
            self.kp = self.u0 * 1.e-4 * self.cox
        if (not self.is_set('ucrit')) and (self.vmax is not None) and \
                (self.u0 is not None):
            self.ucrit = self.vmax / (self.u0 * 1.e-4)
        self._weff = self.w + self.dw
        self._leff = self.l + self.dl
        self._sga = np.sqrt(self.np * self._weff * self.ns * self._leff)
        self._gammaa = self.gamma + self.agamma / self._sga
        if self._gammaa < 0.:
            self._gammaa = 0.
        cEps = 0.001936
        cA = 0.028
        xi = cA * (10. * self._leff / self.lk - 1.)
        self._deltavRSCE = 2. * self.q0 / \
            (self.cox * pow(1. + 0.5 * (xi + np.sqrt(xi * xi + cEps)), 2))
        self._lc = np.sqrt(const.epSi * self.xj / self.cox)
        self._lmin = self.ns * self._leff / 10.
        self._Cox = self.cox * self.np * self._weff * self.ns * self._leff
        if not thermal:
            self.set_temp_vars(self.temp)

    def set_temp_vars(self, temp):
        ad.delete_tape(self)
        T = const.T0 + temp
        self._Vt = const.k * T / const.q
        vt0T = self._vt0 - self._tcv * (T - self._Tn)
        self._vt0a = vt0T + self.avto / self._sga
        kpT = self.kp * pow(T / self._Tn, self.bex)
        self.kpa = kpT * (1 + self.akp / self._sga)
        if not self.kpa > 0.:
            self.kpa = 0.
        self._t_ucrit = self.ucrit * pow(T / self._Tn, self.ucex)
        egT = 1.16 - 0.000702 * T * T / (T + 1108.)
        self.phiT = self.phi * T / self.Tref \
            - 3. * self._Vt * np.log(T / self.Tref) \
            - self.egTref * T / self.Tref + egT
        self._t_ibb = self.ibb * (1. + self.ibbt * (T - self.Tref))
        self._vc = self._t_ucrit * self._leff * self.ns
        self._qb0 = self._gammaa * np.sqrt(self.phiT)
        self._kSt = 4. * const.k * T

    def cuda_pack(self):
        """Kernel record (line numbers of the reference's mosEKV.py)."""
        s = self
        if s.theta != 0.:
            # the reference's simple mobility model reads an undefined name
            # (mosEKV.py:516) and cannot run either
            raise cir.CircuitError(
                s.instanceName + ': theta != 0 is not supported')
        epSi = const.epSi
        Vt = s._Vt
        v = dict(tf=s._tf, vt0a=s._vt0a, deltavRSCE=s._deltavRSCE,
                 phiT=s.phiT, gammaa=s._gammaa, weff=s._weff, Vt=Vt,
                 vc=s._vc, Lambda=s.Lambda, lc=s._lc, t_ucrit=s._t_ucrit,
                 e0=s.e0, epSi=epSi, eta=s.eta, iba=s.iba, t_ibb=s._t_ibb,
                 qox=s.qox)
        v['gsqphi'] = s._gammaa * np.sqrt(s.phiT)                  # :412
        v['g2o4'] = s._gammaa * s._gammaa / 4.                     # :417
        v['go2'] = s._gammaa / 2.                                  # :418
        v['vt16'] = 16. * Vt * Vt                                  # :421
        v['letaleff'] = s.leta / s._leff                           # :426
        v['weta3'] = 3. * s.weta                                   # :427
        v['epscox'] = epSi / s.cox                                 # :428
        v['vt01'] = 0.1 * Vt                                       # :429
        v['vt4'] = 4. * Vt                                         # :436,:452
        v['vdsslog'] = Vt * (np.log(s._vc / (2. * Vt)) - 0.6)      # :448
        v['lamlc'] = s.Lambda * s._lc                              # :458
        v['nsleff'] = s.ns * s._leff                               # :461
        v['lmin2'] = s._lmin * s._lmin                             # :462
        v['kpanpw'] = s.kpa * s.np * s._weff                       # :509
        v['bop'] = 1. + s.cox * s._qb0 / s.e0 / epSi               # :518
        v['coxvt'] = s.cox * Vt                                    # :521
        v['ibn2'] = 2. * s.ibn                                     # :530
        v['ntibblc'] = -s._t_ibb * s._lc                           # :532
        v['qscale'] = s._Cox * Vt * s._tf                          # :540
        flags = _F['SIMPLE'] if s.ekvint else 0
        names = LAYOUTS['ekv'][0]
        nint = len(names) - 1 - 4 * len(EMPTY)
        rec = [v[n] for n in names[:nint]]
        return flags, np.concatenate([rec, [1.]] + 4 * [EMPTY])

    def thermal_raw(self):
        """Raw inputs of set_temp_vars for the mosekv_t record
        (param_layout.EKV_T; reference mosEKV.py:329-357)."""
        from ..globalVars import glVar
        return [glVar.temp, self._vt0, self._tcv, self._Tn, self.avto,
                self._sga, self.kp, self.bex, self.akp, self.ucrit,
                self.ucex, self.phi, self.Tref, self.egTref, self.ibb,
                self.ibbt, self._leff, self.ns, self.np, self.cox, self._Cox]

    def power(self, vPort, currV):
        vds = vPort[0] - vPort[2]
        return vds * currV[0] + vPort[0] * currV[1] + vPort[2] * currV[2]

    def get_OP(self, vPort):
        (outV, jac) = self.eval_and_deriv(vPort)
        reversed_ = (self._tf * (vPort[0] - vPort[2])) <= 0.
        return dict(VD=vPort[0], VG=vPort[1], VS=vPort[2], IDS=outV[0],
                    IDB=outV[1], ISB=outV[2], QD=outV[-3], QG=outV[-2],
                    QS=outV[-1], Power=self.power(vPort, outV),
                    Temp=self.temp, gm=jac[0, 1],
                    gmbs=-jac[0, 2] - jac[0, 1] - jac[0, 0],
                    gds=jac[0, 2] if reversed_ else jac[0, 0],
                    Reversed=reversed_)


EKV = mosExt.extrinsic_mos(IntEKV)

devList = [IntEKV, EKV]
