"""
State-variable-based junction diode (netlist name ``svdiode``).

Device plugin with the reference's interface (devices/svdiode.py:113-400):
the junction is parametrised by a state variable ``x`` on an internal
terminal (Rizzoli's formulation), so that the exponential never overflows;
externally it behaves like ``diode``.  Control port ``[x]``, current outputs
``[id, gyr*vd]`` (the second drives the gyrator that enforces ``vd``), charge
``[qd]`` if ``tt + cj0 != 0``.  ``svdiode_t`` is the electro-thermal variant
(autoThermal).  The equations run in the CUDA kernels
(csrc/models/diode.cuh ``svdiode_core``, csrc/models/thermal.cuh).
"""
import numpy as np

from ..globalVars import const, glVar
from .. import circuit as cir
from . import cudadev as ad
from .junction import SVJunction
from .param_layout import flag_bits

_F = flag_bits('svdiode')


class Device(ad.CudaModel, cir.Element):
    """
    State-Variable-Based Diode
    --------------------------

    Netlist examples::

        svdiode:d1 1 0 isat=10fA cj0=20fF
        svdiode:d2 1 0 model=mydiode
        .model mydiode svdiode (cj0=10fF tt=1ps)
    """
    category = "Semiconductor devices"
    devType = "svdiode"
    cudaModel = 'svdiode'
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
        self.jtn = SVJunction()

    def process_params(self, thermal=False):
        ad.delete_tape(self)
        self.clean_internal_terms()
        self.__addThermalPorts = True
        tx = self.add_internal_term('x', 's.v.')
        tref = self.add_reference_term()
        self.linearVCCS = [((0, 1), (tx, tref), glVar.gyr)]
        self.csOutPorts = [(0, 1), (tref, tx)]
        self.noisePorts = [(0, 1)]
        self.controlPorts = [(tx, tref)]
        if self.rs != 0.:
            t2 = self.add_internal_term('t2', 'V')
            g = self.area / self.rs
            self.linearVCCS = [((t2, 1), (tx, tref), glVar.gyr),
                               ((0, t2), (0, t2), g)]
            self.csOutPorts = [(t2, 1), (tref, tx)]
            self.noisePorts = [(t2, 1), (0, t2)]
        self.qsOutPorts = []
        self._qd = False
        if self.tt + self.cj0 != 0.:
            self.qsOutPorts.append(self.csOutPorts[0])
            self._qd = True
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
        self.Tabs = temp + const.T0
        self.vt = const.k * self.Tabs / const.q
        self.egap_t = self.eg0 - .000702 * (self.Tabs**2) \
            / (self.Tabs + 1108.)
        self._Sthermal = 4. * const.k * self.Tabs * self.rs
        self.jtn.set_temp_vars(self.Tabs, self.Tnomabs, self.vt, self.egapn,
                               self.egap_t)

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
        head = [self.area, bv, self.ibv, self.n, self.vt, self.tt, glVar.gyr]
        return flags, np.concatenate((head, self.jtn.pack()))

    def cuda_pack_thermal(self):
        """svdiode_t record (param_layout.SVDIODE_T): base record with
        placeholders at ambient + raw inputs of set_temp_vars
        (reference svdiode.py:280-295, :53-78)."""
        self.set_temp_vars(glVar.temp)
        flags, base = Device.cuda_pack(self)
        extra = [glVar.temp, self.Tnomabs, self.egapn, self.eg0]
        return flags, np.concatenate((base, extra, self.jtn.pack_raw()))

    def power(self, vPort, ioutV):
        # ioutV[0]: current, ioutV[1]: gyr * voltage
        return ioutV[0] * ioutV[1] / glVar.gyr

    def get_OP(self, vPort):
        (outV, jac) = self.eval_and_deriv(vPort)
        self._Sshot = 2. * const.q * outV[0]
        self._kFliker = self.kf * outV[0]
        opDict = dict(x=vPort[0], VD=outV[1] / glVar.gyr, ID=outV[0],
                      gd=glVar.gyr * jac[0, 0] / jac[1, 0],
                      Sthermal=self._Sthermal, Sshot=self._Sshot,
                      kFliker=self._kFliker)
        if self._qd:
            opDict['Cd'] = jac[-1, 0] / jac[1, 0]
        return opDict

    def get_noise(self, f):
        rs = self._Sthermal if self.rs != 0. else None
        inoise = self._Sshot + self._kFliker / pow(f, self.af)
        return np.array([inoise]) if rs is None else np.array([inoise, rs])
