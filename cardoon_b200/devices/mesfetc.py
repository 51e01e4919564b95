"""
Curtice cubic MESFET (netlist name ``mesfetc``) with time-delayed control
ports.

Plugin interface and parameter processing follow the reference's
devices/mesfetc.py (Device :17, process_params :144-165, set_temp_vars
:167-190); eval_cqs (:192-226) is csrc/models/mesfetc.cuh.  The two delayed
copies of (vgs, vgd) are produced by the nodal engine from the stored
history (reference nodal.delay_interp :82-108).

Terminal order: 0 drain, 1 gate, 2 source.
"""
import numpy as np

from ..globalVars import const
from .. import circuit as cir
from . import cudadev as ad
from .junction import Junction
from .param_layout import flag_bits

_F = flag_bits('mesfetc')


class Device(ad.CudaModel, cir.Element):
    """
    Cubic Curtice-Ettemberg Intrinsic MESFET Model
    ----------------------------------------------

    Controls ``[vgs(t), vgd(t), vgs(t-tau), vgd(t-tau)]``, currents
    ``[igs, igd, ids]``, charges ``[qgs, qgd]``.

    Netlist example::

        mesfetc:m1 2 3 4 a0=0.09910 a1=0.08541 a2=-0.02030 a3=-0.01543
    """
    category = "Semiconductor devices"
    devType = "mesfetc"
    cudaModel = 'mesfetc'
    numTerms = 3
    makeAutoThermal = True
    isNonlinear = True
    nDelays = 2
    csOutPorts = [(1, 2), (1, 0), (0, 2)]
    controlPorts = [(1, 2), (1, 0)]
    vPortGuess = np.array([0., 0., 0., 0.])
    qsOutPorts = [(1, 2), (1, 0)]

    paramDict = dict(
        cir.Element.tempItem,
        a0=('Drain saturation current for Vgs=0', 'A', float, 0.1),
        a1=('Coefficient for V1', 'A/V', float, 0.05),
        a2=('Coefficient for V1^2', 'A/V^2', float, 0.),
        a3=('Coefficient for V1^3', 'A/V^3', float, 0.),
        beta=('V1 dependance on Vds', '1/V', float, 0.),
        vds0=('Vds at which BETA was measured', 'V', float, 4.),
        gama=('Slope of drain characteristic in the linear region', '1/V',
              float, 1.5),
        vt0=('Voltage at which the channel current is forced to be zero '
             'for Vgs<=Vto', 'V', float, -np.inf),
        cgs0=('Gate-source Schottky barrier capacitance for Vgs=0', 'F',
              float, 0.),
        cgd0=('Gate-drain Schottky barrier capacitance for Vgd=0', 'F',
              float, 0.),
        isat=('Diode saturation current', 'A', float, 0.),
        n=('Diode ideality factor', '', float, 1.),
        ib0=('Breakdown current parameter', 'A', float, 0.),
        nr=('Breakdown ideality factor', '', float, 10.),
        vbd=('Breakdown voltage', 'V', float, np.inf),
        tau=('Channel transit time', 's', float, 0.),
        vbi=('Built-in potential of the Schottky junctions', 'V', float,
             0.8),
        fcc=('Forward-bias depletion capacitance coefficient', 'V', float,
             0.5),
        tnom=('Nominal temperature', 'C', float, 27.),
        avt0=('Pinch-off voltage (VP0 or VT0) linear temp. coefficient',
              '1/K', float, 0.),
        bvt0=('Pinch-off voltage (VP0 or VT0) quadratic temp. coefficient',
              '1/K^2', float, 0.),
        tbet=('BETA power law temperature coefficient', '1/K', float, 0),
        tm=('Ids linear temp. coeff.', '1/K', float, 0.),
        tme=('Ids power law temp. coeff.', '1/K^2', float, 0.),
        eg0=('Barrier height at 0 K', 'eV', float, 0.8),
        mgs=('Gate-source grading coefficient', '', float, 0.5),
        mgd=('Gate-drain grading coefficient', '', float, 0.5),
        xti=('Diode saturation current temperature exponent', '', float,
             2.),
        area=('Area multiplier', '', float, 1.),
    )

    def __init__(self, instanceName):
        cir.Element.__init__(self, instanceName)
        self.diogs = Junction()
        self.diogd = Junction()

    def process_params(self, thermal=False):
        ad.delete_tape(self)
        self.delayedContPorts = [(1, 2, self.tau), (1, 0, self.tau)]
        self.Tnomabs = self.tnom + const.T0
        self.egapn = self.eg0 - .000702 * (self.Tnomabs**2) \
            / (self.Tnomabs + 1108.)
        self.diogs.process_params(self.isat, self.n, self.fcc, self.cgs0,
                                  self.vbi, self.mgs, self.xti, self.eg0,
                                  self.Tnomabs)
        self.diogd.process_params(self.isat, self.n, self.fcc, self.cgd0,
                                  self.vbi, self.mgd, self.xti, self.eg0,
                                  self.Tnomabs)
        if not thermal:
            self.set_temp_vars(self.temp)

    def set_temp_vars(self, temp):
        ad.delete_tape(self)
        self.Tabs = const.T0 + temp
        self.vt = const.k * self.Tabs / const.q
        self.egap_t = self.eg0 - .000702 * (self.Tabs**2) \
            / (self.Tabs + 1108.)
        args = (self.Tabs, self.Tnomabs, self.vt, self.egapn, self.egap_t)
        self.diogs.set_temp_vars(*args)
        self.diogd.set_temp_vars(*args)
        delta_T = temp - self.tnom
        self._k6 = self.nr * self.vt
        self._Vt0 = self.vt0 * (1. + (delta_T * self.avt0)
                                + (delta_T**2 * self.bvt0))
        if self.tbet:
            self._Beta = self.beta * pow(1.01, (delta_T * self.tbet))
        else:
            self._Beta = self.beta
        if self.tme * self.tm != 0.:
            self._idsFac = pow((1. + delta_T * self.tm), self.tme)
        else:
            self._idsFac = 1.

    def cuda_pack(self):
        flags = 0
        if self.cgs0 != 0.:
            flags |= _F['CGS0']
        if self.cgd0 != 0.:
            flags |= _F['CGD0']
        head = [self.area, self.ib0, self.vbd, self._k6, self._Beta,
                self.vds0, self.a0, self.a1, self.a2, self.a3, self.gama,
                self._idsFac, self._Vt0]
        return flags, np.concatenate((head, self.diogs.pack(),
                                      self.diogd.pack()))

    def cuda_pack_thermal(self):
        """mesfetc_t record (param_layout.MESFETC_T): base record with
        placeholders at ambient + raw inputs of set_temp_vars
        (reference mesfetc.py:160-189)."""
        from ..globalVars import glVar
        self.set_temp_vars(glVar.temp)
        flags, base = Device.cuda_pack(self)
        extra = [glVar.temp, self.Tnomabs, self.egapn, self.eg0, self.tnom,
                 self.nr, self.vt0, self.avt0, self.bvt0, self.beta,
                 float(self.tbet), self.tm, self.tme]
        return flags, np.concatenate((base, extra, self.diogs.pack_raw(),
                                      self.diogd.pack_raw()))

    def power(self, vPort, currV):
        vds = vPort[0] - vPort[1]
        return vds * currV[2] + vPort[0] * currV[0] + vPort[1] * currV[1]

    def get_OP(self, vPort):
        (outV, jac) = self.eval_and_deriv(vPort)
        return dict(VGS=vPort[0], VDS=vPort[0] - vPort[1], IDS=outV[2],
                    IGS=outV[0], IGD=outV[1])
