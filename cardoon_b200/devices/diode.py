"""
Junction diode (netlist name ``diode``).

Plugin interface and parameter processing follow the reference's
devices/diode.py (Device :103, process_params :209, set_temp_vars :258); the
current/charge law of eval_cqs (:275-307) is implemented by the CUDA kernel
csrc/models/diode.cuh.

Terminals: 0 anode, 1 cathode.  With ``rs != 0`` an internal terminal ``t2``
carries the series resistor as a linear VCCS ``area/rs``.
"""
import numpy as np

from ..globalVars import const
from .. import circuit as cir
from . import cudadev as ad
from .junction import Junction
from .param_layout import flag_bits

_F = flag_bits('diode')


class Device(ad.CudaModel, cir.Element):
    """
    Junction Diode
    --------------

    Based on the Spice model: junction current with series resistance,
    depletion + diffusion charge, reverse breakdown.

    Netlist examples::

        diode:d1 1 0 isat=10fA cj0=20fF
        .model dmodel1 diode (cj0 = 10pF tt=1ps)
    """
    category = "Semiconductor devices"
    devType = "diode"
    cudaModel = 'diode'
    numTerms = 2
    makeAutoThermal = True
    isNonlinear = True
    vPortGuess = np.array([0.])

    paramDict = dict(
        cir.Element.tempItem,
        isat=('Saturation current', 'A', float, 1e-14),
        n=('Emission coefficient', ' ', float, 1.),
        fc=('Coefficient for forward-bias depletion capacitance', ' ',
            float, .5),
        cj0=('Zero-bias depletion capacitance', 'F', float, 0.),
        vj=('Built-in junction potential', 'V', float, 1.),
        m=('PN junction grading coefficient', ' ', float, .5),
        tt=('Transit time', 's', float, 0.),
        xti=('Is temperature exponent', ' ', float, 3.),
        eg0=('Energy bandgap', 'eV', float, 1.11),
        tnom=('Nominal temperature', 'C', float, 27.),
        ibv=('Current at reverse breakdown voltage', 'A', float, 1e-10),
        bv=('Breakdown voltage', 'V', float, np.inf),
        area=('Area multiplier', ' ', float, 1.),
        rs=('Series resistance', 'Ohms', float, 0.),
        kf=('Flicker noise coefficient', '', float, 0.),
        af=('Flicker noise exponent', '', float, 1.),
    )

    def __init__(self, instanceName):
        cir.Element.__init__(self, instanceName)
        self.jtn = Junction()

    def process_params(self, thermal=False):
        ad.delete_tape(self)
        self.clean_internal_terms()
        self.__addThermalPorts = True
        if self.rs != 0.:
            t2 = self.add_internal_term('t2', 'V')
            self.linearVCCS = [((0, t2), (0, t2), self.area / self.rs)]
            self.csOutPorts = [(t2, 1)]
            self.noisePorts = [(t2, 1), (0, t2)]
            self.controlPorts = [(t2, 1)]
        else:
            self.linearVCCS = []
            self.csOutPorts = [(0, 1)]
            self.noisePorts = [(0, 1)]
            self.controlPorts = [(0, 1)]
        # charge source only if there is any charge to model
        self._qd = (self.tt + self.cj0 != 0.)
        self.qsOutPorts = list(self.csOutPorts) if self._qd else []
        self.Tnomabs = self.tnom + const.T0
        self.egapn = self.eg0 - .000702 * (self.Tnomabs**2) \
            / (self.Tnomabs + 1108.)
        self.jtn.process_params(self.isat, self.n, self.fc, self.cj0,
                                self.vj, self.m, self.xti, self.eg0,
                                self.Tnomabs)
        if not thermal:
            self.set_temp_vars(self.temp)

    def set_temp_vars(self, temp):
        ad.delete_tape(self)
        Tabs = temp + const.T0
        self.vt = const.k * Tabs / const.q
        egap_t = self.eg0 - .000702 * (Tabs**2) / (Tabs + 1108.)
        self._Sthermal = 4. * const.k * Tabs * self.rs
        self.jtn.set_temp_vars(Tabs, self.Tnomabs, self.vt, self.egapn,
                               egap_t)

    def cuda_pack(self):
        flags = 0
        if self.cj0 != 0.:
            flags |= _F['CJ0']
        if self.tt != 0.:
            flags |= _F['TT']
        if self.bv < np.inf:
            flags |= _F['BV']
        if self._qd:
            flags |= _F['QD']
        bv = self.bv if self.bv < np.inf else 0.
        head = [self.area, bv, self.ibv, self.n, self.vt, self.tt]
        return flags, np.concatenate((head, self.jtn.pack()))

    def cuda_pack_thermal(self):
        """diode_t record (param_layout.DIODE_T): the diode record -- its
        temperature-dependent entries (vt, jtn_t_is, jtn_kexp, jtn_t_vj,
        jtn_k5..k8) are placeholders evaluated at ambient, the kernel
        recomputes them from the port temperature -- plus the raw inputs of
        set_temp_vars (reference diode.py:254-272, :53-72)."""
        from ..globalVars import glVar
        self.set_temp_vars(glVar.temp)
        flags, base = Device.cuda_pack(self)
        extra = [glVar.temp, self.Tnomabs, self.egapn, self.eg0]
        return flags, np.concatenate((base, extra, self.jtn.pack_raw()))

    def power(self, vPort, ioutV):
        return vPort[0] * ioutV[0]

    def get_OP(self, vPort):
        (outV, jac) = self.eval_and_deriv(vPort)
        self._Sshot = 2. * const.q * outV[0]
        self._kFliker = self.kf * outV[0]
        opDict = dict(VD=vPort[0], ID=outV[0], gd=jac[0, 0],
                      Sthermal=self._Sthermal, Sshot=self._Sshot,
                      kFliker=self._kFliker)
        if self._qd:
            opDict['Cd'] = jac[-1, 0]
        return opDict
