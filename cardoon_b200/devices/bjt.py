"""
Gummel-Poon BJT (netlist name ``bjt``) and the extrinsic wrapper shared with
the state-variable version (``svbjt``).

Plugin interface, topology and parameter processing follow the reference's
devices/bjt.py (extrinsic_bjt :16-228, BJTi :230-607) and devices/svbjt.py
(SVBJTi :19-435).  The reference keeps two near-identical intrinsic classes;
here both derive from ``_GPBase`` and differ only in topology and junction
type.  The current/charge equations (bjt.py:485-607, svbjt.py:308-435) live
in csrc/models/bjt.cuh.

Terminal order: 0 collector, 1 base, 2 emitter, [3 bulk].
"""
import numpy as np

from ..globalVars import const, glVar
from .. import circuit as cir
from . import cudadev as ad
from .junction import Junction, SVJunction, EMPTY
from .param_layout import flag_bits

_F = flag_bits('bjt')

_gpParams = dict(
    cir.Element.tempItem,
    type=('Type (npn or pnp)', '', str, 'npn'),
    isat=('Transport saturation current', 'A', float, 1e-16),
    bf=('Ideal maximum forward beta', '', float, 100.),
    nf=('Forward current emission coefficient', '', float, 1.),
    vaf=('Forward early voltage', 'V', float, 0.),
    ikf=('Forward-beta high current roll-off knee current', 'A', float, 0.),
    ise=('Base-emitter leakage saturation current', 'A', float, 0.),
    ne=('Base-emitter leakage emission coefficient', '', float, 1.5),
    br=('Ideal maximum reverse beta', '', float, 1.),
    nr=('Reverse current emission coefficient', '', float, 1.),
    var=('Reverse early voltage', 'V', float, 0.),
    ikr=('Corner for reverse-beta high current roll off', 'A', float, 0.),
    isc=('Base collector leakage saturation current', 'A', float, 0.),
    nc=('Base-collector leakage emission coefficient', '', float, 2.),
    rb=('Zero bias base resistance', 'Ohm', float, 0.),
    rbm=('Minimum base resistance', 'Ohm', float, 0.),
    irb=('Current at which rb falls to half of rbm', 'A', float, 0.),
    eg=('Badgap voltage', 'eV', float, 1.11),
    cje=('Base emitter zero bias p-n capacitance', 'F', float, 0.),
    vje=('Base emitter built in potential', 'V', float, 0.75),
    mje=('Base emitter p-n grading factor', '', float, 0.33),
    cjc=('Base collector zero bias p-n capacitance', 'F', float, 0.),
    vjc=('Base collector built in potential', 'V', float, 0.75),
    mjc=('Base collector p-n grading factor', '', float, 0.33),
    xcjc=('Fraction of cbc connected internal to rb', '', float, 1.),
    fc=('Forward bias depletion capacitor coefficient', '', float, 0.5),
    tf=('Ideal forward transit time', 's', float, 0.),
    xtf=('Transit time bias dependence coefficient', '', float, 0.),
    vtf=('Transit time dependency on vbc', 'V', float, 0.),
    itf=('Transit time dependency on ic', 'A', float, 0.),
    tr=('Ideal reverse transit time', 's', float, 0.),
    xtb=('Forward and reverse beta temperature coefficient', '', float, 0.),
    xti=('IS temperature effect exponent', '', float, 3.),
    tnom=('Nominal temperature', 'C', float, 27.),
    area=('Current multiplier', '', float, 1.),
)


class _GPBase(ad.CudaModel, cir.Element):
    """Parts common to the plain and the state-variable intrinsic models."""
    paramDict = _gpParams
    junctionClass = Junction

    def __init__(self, instanceName):
        cir.Element.__init__(self, instanceName)
        self.jif = self.junctionClass()
        self.jir = self.junctionClass()
        self.jile = Junction()
        self.jilc = Junction()
        # collector / emitter terminal numbers (re-mapped by the extrinsic
        # wrapper when rc / re are given)
        self._ct = 0
        self._et = 2

    def _trim_charges(self):
        """Drop charge sources that are identically zero."""
        keep = []
        if self.cje + self.tf != 0.:
            keep.append(self.qsOutPorts[0])
        if self.cjc + self.tr != 0.:
            if self._qbx:
                keep += self.qsOutPorts[-2:]
            else:
                keep.append(self.qsOutPorts[-1])
        self.qsOutPorts = keep
        self.ncurrents = len(self.csOutPorts)
        self.ncharges = len(self.qsOutPorts)

    def _process_common(self):
        self._typef = -1. if self.type == 'pnp' else 1.
        self.Tnomabs = self.tnom + const.T0
        self.egapn = self.eg - .000702 * (self.Tnomabs**2) \
            / (self.Tnomabs + 1108.)
        self.jif.process_params(self.isat, self.nf, self.fc, self.cje,
                                self.vje, self.mje, self.xti, self.eg,
                                self.Tnomabs)
        self.jir.process_params(self.isat, self.nr, self.fc, self.cjc,
                                self.vjc, self.mjc, self.xti, self.eg,
                                self.Tnomabs)
        if self.ise != 0.:
            self.jile.process_params(self.ise, self.ne, 0, 0, 0, 0,
                                     self.xti, self.eg, self.Tnomabs)
        if self.isc != 0.:
            self.jilc.process_params(self.isc, self.nc, 0, 0, 0, 0,
                                     self.xti, self.eg, self.Tnomabs)
        if self.irb != 0.:
            self._ck1 = 144. / self.irb / self.area / np.pi / np.pi
            self._ck2 = np.pi * np.pi * np.sqrt(self.irb * self.area) / 24.

    def set_temp_vars(self, temp):
        self.Tabs = const.T0 + temp
        self.tnratio = self.Tabs / self.Tnomabs
        tnXTB = pow(self.tnratio, self.xtb)
        self.vt = const.k * self.Tabs / const.q
        self.egap_t = self.eg - .000702 * (self.Tabs**2) \
            / (self.Tabs + 1108.)
        args = (self.Tabs, self.Tnomabs, self.vt, self.egapn, self.egap_t)
        self.jif.set_temp_vars(*args)
        self.jir.set_temp_vars(*args)
        # leakage currents have a different temperature law
        if self.ise != 0.:
            self.jile.set_temp_vars(*args)
            self.jile._t_is /= tnXTB
        if self.isc != 0.:
            self.jilc.set_temp_vars(*args)
            self.jilc._t_is /= tnXTB
        self._bf_t = self.bf * tnXTB
        self._br_t = self.br * tnXTB

    def _pack_intrinsic(self):
        f = 0
        for name, cond in (('ISE', self.ise != 0.), ('ISC', self.isc != 0.),
                           ('VAR', self.var != 0.), ('VAF', self.vaf != 0.),
                           ('IKF', self.ikf != 0.), ('IKR', self.ikr != 0.),
                           ('RB', self.rb != 0.), ('IRB', self.irb != 0.),
                           ('TF', self.tf != 0.), ('VTF', self.vtf != 0.),
                           ('CJE', self.cje != 0.), ('CJC', self.cjc != 0.),
                           ('TR', self.tr != 0.), ('QBX', self._qbx),
                           ('QBE', self.cje + self.tf != 0.),
                           ('QBC', self.cjc + self.tr != 0.)):
            if cond:
                f |= _F[name]
        if (f & _F['IKF'] or f & _F['IKR']) and self.ikf + self.ikr == 0.:
            # reference tests ``ikf + ikr != 0`` (bjt.py:531)
            f &= ~(_F['IKF'] | _F['IKR'])
        ck1 = self._ck1 if self.irb != 0. else 0.
        ck2 = self._ck2 if self.irb != 0. else 0.
        head = [self._typef, self.area, self._bf_t, self._br_t, self.ikf,
                self.ikr, self.var, self.vaf, self.rb, self.rbm, ck1, ck2,
                self.tf, self.xtf, self.vtf, self.itf, self.tr, self.xcjc,
                glVar.gyr]
        recs = [self.jif.pack(), self.jir.pack(),
                self.jile.pack() if self.ise != 0. else EMPTY,
                self.jilc.pack() if self.isc != 0. else EMPTY]
        return f, np.concatenate([head] + recs)

    def cuda_pack(self):
        f, p = self._pack_intrinsic()
        return f, np.concatenate((p, EMPTY))


class BJTi(_GPBase):
    """
    Gummel-Poon intrinsic BJT model

    Controls ``[vbe, vbc]`` (``[vbie, vbic, v_ib]`` if ``rb != 0``), currents
    ``[ibe, ibc, ice(, gyr*ib*Rb)]``, charges ``[qbe, qbc(, qbx)]``.  With
    ``rb != 0`` two internal terminals (``Bi``, ``ib``) and a local reference
    implement the current-dependent base resistance with a gyrator.
    """
    devType = "bjt"
    cudaModel = 'bjt'

    def process_params(self):
        ct, et = self._ct, self._et
        self._qbx = False
        if self.rb != 0.:
            tBi = self.add_internal_term('Bi', 'V')
            tib = self.add_internal_term('ib', '{0} A'.format(glVar.gyr))
            tref = self.add_reference_term()
            self.linearVCCS = [((1, tBi), (tib, tref), glVar.gyr),
                               ((tib, tref), (1, tBi), glVar.gyr)]
            self.csOutPorts = [(tBi, et), (tBi, ct), (ct, et), (tref, tib)]
            self.controlPorts = [(tBi, et), (tBi, ct), (tib, tref)]
            self.vPortGuess = np.array([0., 0., 0.])
            self.qsOutPorts = [(tBi, et), (tBi, ct)]
            if (self.cjc != 0.) and (self.xcjc < 1.):
                self.qsOutPorts.append((1, ct))
                self._qbx = True
        else:
            self.linearVCCS = []
            self.csOutPorts = [(1, et), (1, ct), (ct, et)]
            self.controlPorts = [(1, et), (1, ct)]
            self.vPortGuess = np.array([0., 0.])
            self.qsOutPorts = [(1, et), (1, ct)]
        self._trim_charges()
        self._process_common()

    def power(self, vPort, currV):
        vce = vPort[0] - vPort[1]
        pRb = currV[3] * vPort[2] if self.rb != 0. else 0.
        return currV[0] * vPort[0] + currV[1] * vPort[1] \
            + currV[2] * vce + pRb

    def get_OP(self, vPort):
        (outV, jac) = self.eval_and_deriv(vPort)
        return dict(VBE=vPort[0], VCE=vPort[0] - vPort[1],
                    IB=outV[0] + outV[1], IC=outV[2] - outV[1],
                    IE=-outV[2] - outV[0], Temp=self.temp,
                    Power=self.power(vPort, outV),
                    gm=jac[2, 0] - jac[1, 0],
                    rpi=1. / (jac[0, 0] + jac[1, 0]))


class SVBJTi(_GPBase):
    """
    State-variable-based Gummel-Poon intrinsic BJT model

    The BE and BC diodes are parametrised by state variables ``x1``, ``x2``
    (two extra internal terminals), which removes large positive
    exponentials.  Controls ``[x1, x2(, v_ib)]``, currents
    ``[ibe, gyr*vbe, ibc, gyr*vbc, ice(, gyr*ib*Rb)]``.
    """
    devType = "svbjt"
    cudaModel = 'svbjt'
    junctionClass = SVJunction

    def process_params(self):
        ct, et = self._ct, self._et
        x2 = self.add_internal_term('x2', 's.v.')
        x1 = self.add_internal_term('x1', 's.v.')
        tref = self.add_reference_term()
        self._qbx = False
        if self.rb != 0.:
            tBi = self.add_internal_term('Bi', 'V')
            tib = self.add_internal_term('ib', '{0} A'.format(glVar.gyr))
            # second VCCS senses (Bi, terminal 0) as in the reference
            # (svbjt.py:204), not (Bi, ct)
            self.linearVCCS = [((tBi, et), (x1, tref), glVar.gyr),
                               ((tBi, 0), (x2, tref), glVar.gyr),
                               ((1, tBi), (tib, tref), glVar.gyr),
                               ((tib, tref), (1, tBi), glVar.gyr)]
            self.csOutPorts = [(tBi, et), (tref, x1), (tBi, ct), (tref, x2),
                               (ct, et), (tref, tib)]
            self.controlPorts = [(x1, tref), (x2, tref), (tib, tref)]
            self.qsOutPorts = [(tBi, et), (tBi, ct)]
            if (self.cjc != 0.) and (self.xcjc < 1.):
                self.qsOutPorts.append((1, ct))
                self._qbx = True
        else:
            self.linearVCCS = [((1, et), (x1, tref), glVar.gyr),
                               ((1, ct), (x2, tref), glVar.gyr)]
            self.csOutPorts = [(1, et), (tref, x1), (1, ct), (tref, x2),
                               (ct, et)]
            self.controlPorts = [(x1, tref), (x2, tref)]
            self.qsOutPorts = [(1, et), (1, ct)]
        self.vPortGuess = np.zeros(len(self.controlPorts))
        self._trim_charges()
        self._process_common()

    def power(self, vPort, currV):
        gyrvce = currV[1] - currV[3]
        pRb = currV[5] * vPort[2] if self.rb != 0. else 0.
        return (currV[0] * currV[1] + currV[2] * currV[3]
                + currV[4] * gyrvce) / glVar.gyr + pRb

    def get_OP(self, vPort):
        (outV, jac) = self.eval_and_deriv(vPort)
        return dict(VBE=outV[1] / glVar.gyr,
                    VCE=(outV[1] - outV[3]) / glVar.gyr,
                    IB=outV[0] + outV[2], IC=outV[4] - outV[2],
                    IE=-outV[4] - outV[0], Temp=self.temp,
                    Power=self.power(vPort, outV))


def extrinsic_bjt(IBJT):
    """Class factory adding RC, RE and the collector-bulk junction."""

    class BJT(IBJT):
        """
    Bipolar Junction Transistor
    ---------------------------

    Accepts 3 or 4 terminal connections (C, B, E[, bulk]).  RC and RE are
    added as linear conductances on internal terminals ``ct``/``et``; with 4
    terminals a collector-bulk junction is connected.

    Netlist examples::

        bjt:q1 2 3 4 1 model = mypnp isat=4e-17 bf=147 iss=10fA
        svbjt:q3 2 3 4 1 model = mypnp vaf=80 ikf=4m iss=15fA
        .model mypnp bjt (type=pnp isat=5e-17 cje=60fF vje=0.83 mje=0.35)
        """
        extraDoc = IBJT.__doc__
        makeAutoThermal = True
        isNonlinear = True
        category = "Semiconductor devices"
        devType = IBJT.devType
        paramDict = dict(
            IBJT.paramDict.items(),
            re=('Emitter ohmic resistance', 'Ohm', float, None),
            rc=('Collector ohmic resistance', 'Ohm', float, None),
            cjs=('Collector substrate capacitance', 'F', float, 0.),
            mjs=('substrate junction exponential factor', '', float, 0.),
            vjs=('substrate junction built in potential', 'V', float, 0.75),
            ns=('substrate p-n coefficient', '', float, 1.),
            iss=('Substrate saturation current', 'A', float, 1e-14),
        )

        def __init__(self, instanceName):
            IBJT.__init__(self, instanceName)
            self.cbjtn = Junction()

        def process_params(self, thermal=False):
            ad.delete_tape(self)
            extraTerms = 2 if thermal else 0
            if self.numTerms == 3 + extraTerms:
                self._addCBjtn = False
            elif self.numTerms == 4 + extraTerms:
                self._addCBjtn = True
            else:
                raise cir.CircuitError(
                    '{0}: Wrong number of connections. Can only be {1} or '
                    '{2}, {3} found.'.format(self.instanceName,
                                             3 + extraTerms, 4 + extraTerms,
                                             self.numTerms))
            self.clean_internal_terms()
            self.__addThermalPorts = True
            self._ct, self._et = 0, 2
            extraVCCS = []
            if self.re is not None:
                self._et = self.add_internal_term('et', 'V')
                extraVCCS.append(((2, self._et), (2, self._et),
                                  self.area / self.re))
            if self.rc is not None:
                self._ct = self.add_internal_term('ct', 'V')
                extraVCCS.append(((0, self._ct), (0, self._ct),
                                  self.area / self.rc))
            IBJT.process_params(self)
            self.cbjtn.process_params(self.iss, self.ns, self.fc, self.cjs,
                                      self.vjs, self.mjs, self.xti, self.eg,
                                      self.Tnomabs)
            self.linearVCCS = self.linearVCCS + extraVCCS
            if self._addCBjtn:
                self._bccn = len(self.controlPorts)
                self._bcon = len(self.csOutPorts)
                self.controlPorts = self.controlPorts + [(3, self._ct)]
                self.csOutPorts = self.csOutPorts + [(3, self._ct)]
                if self.cjs != 0.:
                    self.qsOutPorts = self.qsOutPorts + [(3, self._ct)]
                if len(self.vPortGuess) < len(self.controlPorts):
                    self.vPortGuess = np.concatenate((self.vPortGuess, [1]))
            self.set_temp_vars(self.temp)

        def set_temp_vars(self, temp):
            ad.delete_tape(self)
            IBJT.set_temp_vars(self, temp)
            self.cbjtn.set_temp_vars(self.Tabs, self.Tnomabs, self.vt,
                                     self.egapn, self.egap_t)

        def cuda_pack(self):
            f, p = self._pack_intrinsic()
            if self._addCBjtn:
                f |= _F['SUB']
                if self.cjs != 0.:
                    f |= _F['CJS']
                return f, np.concatenate((p, self.cbjtn.pack()))
            return f, np.concatenate((p, EMPTY))

        def cuda_pack_thermal(self):
            """bjt_t / svbjt_t record (param_layout.BJT_T): the BJT record
            (temperature-dependent entries are placeholders at the device's
            own temp, recomputed by the kernel from the thermal port) plus
            the raw inputs of set_temp_vars (reference bjt.py:453-482,
            :170-180)."""
            f, base = BJT.cuda_pack(self)
            raw_empty = EMPTY
            extra = [glVar.temp, self.Tnomabs, self.egapn, self.eg, self.xtb,
                     self.bf, self.br]
            recs = [self.jif.pack_raw(), self.jir.pack_raw(),
                    self.jile.pack_raw() if self.ise != 0. else raw_empty,
                    self.jilc.pack_raw() if self.isc != 0. else raw_empty,
                    self.cbjtn.pack_raw() if self._addCBjtn else raw_empty]
            return f, np.concatenate([base, extra] + recs)

        def power(self, vPort, currV):
            pout = IBJT.power(self, vPort, currV)
            if self._addCBjtn:
                pout += vPort[self._bccn] * currV[self._bcon]
            return pout

    BJT.__name__ = 'BJT_' + IBJT.devType
    return BJT


Device = extrinsic_bjt(BJTi)
