"""
Linear elements and independent sources.

They only publish data: ``linearVCCS``/``linearVCQS`` quads
``((c1, c2), (o1, o2), g)`` stamped once into G/C, and source values
(``get_DCsource``/``get_TDsource``).  Topologies and formulas follow the
reference's devices/resistor.py, capacitor.py, inductor.py, gyrator.py,
vccs.py, idc.py, vdc.py, isin.py, vsin.py, ipulse.py, vpulse.py (voltage
sources and inductors are built with gyrators, vdc.py:112-123,
inductor.py:82-91).
"""
import cmath as cm

import numpy as np

from ..globalVars import const, glVar
from .. import circuit as cir
from . import cudadev as ad


class Resistor(ad.CudaModel, cir.Element):
    """
    Resistor
    --------

    Netlist example::

        res:r1 1 2 r=1e3 tc1=10e-3

    ``res_t`` (electro-thermal, autoThermal) is nonlinear: its conductance
    follows the thermal port (kernel ``res_t_eval``, csrc/models/thermal.cuh).
    """
    category = "Basic components"
    devType = "res"
    cudaModel = 'res'
    makeAutoThermal = True
    numTerms = 2
    csOutPorts = [(0, 1)]
    noisePorts = [(0, 1)]
    controlPorts = [(0, 1)]
    qsOutPorts = []
    vPortGuess = np.array([0.])
    paramDict = dict(
        cir.Element.tempItem,
        r=('Resistance', 'Ohms', float, 0.),
        rsh=('Sheet resistance', 'Ohms', float, 0.),
        l=('Length', 'm', float, 0.),
        w=('Width', 'm', float, 0.),
        narrow=('Narrowing due to side etching', 'm', float, 0.),
        tnom=('Nominal temperature', 'C', float, 27.),
        tc1=('Temperature coefficient 1', '1/C', float, 0.),
        tc2=('Temperature coefficient 2', '1/C^2', float, 0.),
    )

    def process_params(self, thermal=False):
        if not self.r and not (self.rsh and self.l and self.w):
            raise cir.CircuitError(self.instanceName
                                   + ': Resistance can not be zero')
        if self.r != 0.:
            self.g = 1. / self.r
        else:
            self.g = (self.w - self.narrow) / (self.l - self.narrow) \
                / self.rsh
        if not thermal:
            self.set_temp_vars(self.temp)

    def set_temp_vars(self, temp):
        deltaT = temp - self.tnom
        self.g /= (1. + (self.tc1 + self.tc2 * deltaT) * deltaT)
        self._St = 4. * const.k * (temp + const.T0) * self.g
        self.linearVCCS = [((0, 1), (0, 1), self.g)]

    def cuda_pack_thermal(self):
        """res_t record (param_layout.RES_T); reference resistor.py:107-131:
        process_params(thermal=True) leaves ``g`` uncorrected and the taped
        eval_cqs divides it once by the temperature polynomial."""
        from ..globalVars import glVar
        return 0, np.array([self.g, self.tnom, self.tc1, self.tc2,
                            glVar.temp])

    def power(self, vPort, ioutV):
        return vPort[0] * ioutV[0]

    def get_noise(self, f):
        return np.array([self._St])


class Capacitor(cir.Element):
    """
    Linear Capacitor
    ----------------

    Netlist example::

        cap:c1 1 2 c=10uF
    """
    category = "Basic components"
    devType = "cap"
    numTerms = 2
    paramDict = dict(c=('Capacitance', 'F', float, 0.))

    def process_params(self):
        if not self.c:
            raise cir.CircuitError(self.instanceName
                                   + ': Capacitance can not be zero')
        self.linearVCQS = [((0, 1), (0, 1), self.c)]


class Inductor(cir.Element):
    """
    Inductor
    --------

    Implemented with a gyrator and a capacitor ``gyr^2 L`` on the internal
    terminal ``il`` (voltage = current / gyr).

    Netlist example::

        ind:l1 1 0 l=3uH
    """
    category = "Basic components"
    devType = "ind"
    numTerms = 2
    paramDict = dict(l=('Inductance', 'H', float, 0.))

    def process_params(self):
        self.clean_internal_terms()
        if not self.l:
            raise cir.CircuitError(self.instanceName
                                   + ': Inductance can not be zero')
        til = self.add_internal_term('il', '{0} A'.format(glVar.gyr))
        tref = self.add_reference_term()
        self.linearVCCS = [((0, 1), (tref, til), glVar.gyr),
                           ((til, tref), (0, 1), glVar.gyr)]
        self.linearVCQS = [((til, tref), (til, tref),
                            self.l * glVar.gyr**2)]


class Gyrator(cir.Element):
    """
    Gyrator
    -------

    Netlist example::

        gyr:gg 1 0 2 0 g=1m
    """
    category = "Controlled Sources"
    devType = "gyr"
    numTerms = 4
    paramDict = dict(g=('Gyrator gain', 'Ohms', float, 1e-3))

    def process_params(self):
        self.clean_internal_terms()
        self.linearVCCS = [((0, 1), (3, 2), self.g),
                           ((2, 3), (0, 1), self.g)]


class VCCS(cir.Element):
    """
    Voltage-controlled current source (linear)
    ------------------------------------------

    Netlist example::

        vccs:g1 gnd 4 3 gnd g=2mS
    """
    category = "Controlled Sources"
    devType = "vccs"
    numTerms = 4
    paramDict = dict(
        g=('Linear transconductance', 'S', float, 1e-3),
        f=('Nonlinear function i(vc)', 'A', str, ''),
    )

    def process_params(self):
        if self.is_set('f'):
            raise cir.CircuitError(
                self.instanceName + ': nonlinear vccs (f=...) is not built '
                'in this back-end yet')
        self.linearVCCS = [((2, 3), (0, 1), self.g)]


def _tc(dev, temp):
    deltaT = temp - dev.tnom
    return 1. + (dev.tc1 + dev.tc2 * deltaT) * deltaT


class IDC(cir.Element):
    """
    DC current source
    -----------------

    Netlist example::

        idc:is1 gnd 4 idc=2mA
    """
    category = "Sources"
    devType = "idc"
    numTerms = 2
    isDCSource = True
    sourceOutput = (0, 1)
    paramDict = dict(
        cir.Element.tempItem,
        idc=('DC current', 'A', float, 0.),
        tnom=('Nominal temperature', 'C', float, 27.),
        tc1=('Current temperature coefficient 1', '1/C', float, 0.),
        tc2=('Current temperature coefficient 2', '1/C^2', float, 0.),
    )

    def process_params(self):
        self.set_temp_vars(self.temp)

    def set_temp_vars(self, temp):
        self._adjIdc = self.idc * _tc(self, temp)

    def get_DCsource(self):
        return self._adjIdc


def _voltage_source_topology(dev):
    """Norton equivalent if rint != 0, else gyrator + current source.

    Returns the factor that converts volts to the injected current.
    """
    dev.clean_internal_terms()
    if dev.rint != 0.:
        dev.sourceOutput = (1, 0)
        dev.linearVCCS = [((0, 1), (0, 1), 1. / dev.rint)]
        return None
    ti = dev.add_internal_term('i', '{0} A'.format(glVar.gyr))
    tref = dev.add_reference_term()
    dev.sourceOutput = (tref, ti)
    dev.linearVCCS = [((0, 1), (ti, tref), glVar.gyr),
                      ((ti, tref), (0, 1), glVar.gyr)]
    return glVar.gyr


class VDC(cir.Element):
    """
    DC voltage source
    -----------------

    Netlist example::

        vdc:vdd vddnode gnd vdc=3V
    """
    category = "Sources"
    devType = "vdc"
    numTerms = 2
    isDCSource = True
    paramDict = dict(
        cir.Element.tempItem,
        vdc=('DC voltage', 'V', float, 0.),
        rint=('Internal resistance', 'Ohms', float, 0.),
        tnom=('Nominal temperature', 'C', float, 27.),
        tc1=('Voltage temperature coefficient 1', '1/C', float, 0.),
        tc2=('Voltage temperature coefficient 2', '1/C^2', float, 0.),
    )

    def process_params(self):
        gyr = _voltage_source_topology(self)
        if gyr is None:
            self._idc = self.vdc / self.rint
        else:
            self._idc = gyr * self.vdc
        self.set_temp_vars(self.temp)

    def set_temp_vars(self, temp):
        self._adjIdc = self._idc * _tc(self, temp)

    def get_DCsource(self):
        return self._adjIdc


class ISin(cir.Element):
    """
    (Co-)Sinusoidal current source
    ------------------------------

    Netlist example::

        isin:i1 gnd 4 idc=2mA mag=2mA freq=1GHz phase=90
    """
    category = "Sources"
    devType = "isin"
    numTerms = 2
    isDCSource = True
    isTDSource = True
    isFDSource = True
    sourceOutput = (0, 1)
    paramDict = dict(
        idc=('DC current', 'A', float, 0.),
        mag=('Amplitude', 'A', float, 0.),
        acmag=('Amplitude for AC analysis only', 'A', float, None),
        phase=('Phase', 'degrees', float, 0.),
        freq=('Frequency', 'Hz', float, 1e3),
    )

    def process_params(self):
        self._acmag = self.mag if self.acmag is None else self.acmag
        self._omega = 2. * np.pi * self.freq
        self._phase = self.phase * np.pi / 180.

    def get_DCsource(self):
        return self.idc

    def get_TDsource(self, time):
        return self.mag * np.cos(self._omega * time + self._phase)

    def get_FDsource(self):
        return (np.array([self.freq]),
                np.array([cm.rect(self.mag, self._phase)]))

    def get_AC(self):
        return cm.rect(self._acmag, self._phase)


class VSin(cir.Element):
    """
    (Co-)Sinusoidal voltage source
    ------------------------------

    Netlist example::

        vsin:vin 1 0 vdc=1 mag=1 freq=100MEGHz phase=120 rint=50
    """
    category = "Sources"
    devType = "vsin"
    numTerms = 2
    isDCSource = True
    isTDSource = True
    isFDSource = True
    paramDict = dict(
        vdc=('DC voltage', 'V', float, 0.),
        rint=('Internal resistance', 'Ohms', float, 0.),
        mag=('Amplitude', 'V', float, 0.),
        acmag=('Amplitude for AC analysis only', 'V', float, None),
        phase=('Phase', 'degrees', float, 0.),
        freq=('Frequency', 'Hz', float, 1e3),
    )

    def process_params(self):
        self._acmag = self.mag if self.acmag is None else self.acmag
        self._omega = 2. * np.pi * self.freq
        self._phase = self.phase * np.pi / 180.
        gyr = _voltage_source_topology(self)
        if gyr is None:
            self._idc = self.vdc / self.rint
            self._imag = self.mag / self.rint
            self._acimag = self._acmag / self.rint
        else:
            self._idc = gyr * self.vdc
            self._imag = gyr * self.mag
            self._acimag = gyr * self._acmag

    def get_DCsource(self):
        return self._idc

    def get_TDsource(self, time):
        return self._imag * np.cos(self._omega * time + self._phase)

    def get_FDsource(self):
        return (np.array([self.freq]),
                np.array([cm.rect(self._imag, self._phase)]))

    def get_AC(self):
        return cm.rect(self._acimag, self._phase)


def pulse_value(dev, time):
    """Piece-wise linear pulse (reference ipulse.py:15-51)."""
    t = time % dev.per if time > dev.per else time
    if t <= dev.td:
        return dev._i1
    t -= dev.td
    if t < dev.tr:
        return dev._i1 + dev._deltai * t / dev.tr
    t -= dev.tr
    if t < dev.pw:
        return dev._i2
    t -= dev.pw
    if t < dev.tf:
        return dev._i2 - dev._deltai * t / dev.tf
    return dev._i1


_pulseTiming = dict(
    td=('Delay time', 's', float, 0.),
    tr=('Rise time', 's', float, 0.),
    tf=('Fall time', 's', float, 0.),
    pw=('Pulse width', 's', float, np.inf),
    per=('Period', 's', float, np.inf),
)


class IPulse(cir.Element):
    """
    Pulse current source
    --------------------

    Netlist example::

        ipulse:i1 gnd 4 i1=-1mA i2=1mA td=1ms pw=10ms per=20ms
    """
    category = "Sources"
    devType = "ipulse"
    numTerms = 2
    isTDSource = True
    sourceOutput = (0, 1)
    paramDict = dict(
        _pulseTiming,
        i1=('Initial value', 'A', float, 0.),
        i2=('Pulsed value', 'A', float, 0.),
    )

    def process_params(self):
        self._i1 = self.i1
        self._i2 = self.i2
        self._deltai = self._i2 - self._i1
        self._rem = (self.per - self.tr - self.pw - self.tf
                     if self.pw < np.inf else 0.)

    get_TDsource = pulse_value


class VPulse(cir.Element):
    """
    Pulse voltage source
    --------------------

    Netlist example::

        vpulse:vin 1 0 v1=-1V v2=1V td=1ms pw=10ms per=20ms
    """
    category = "Sources"
    devType = "vpulse"
    numTerms = 2
    isTDSource = True
    paramDict = dict(
        _pulseTiming,
        v1=('Initial value', 'V', float, 0.),
        v2=('Pulsed value', 'V', float, 0.),
        rint=('Internal resistance', 'Ohms', float, 0.),
    )

    def process_params(self):
        gyr = _voltage_source_topology(self)
        if gyr is None:
            self._i1 = self.v1 / self.rint
            self._i2 = self.v2 / self.rint
        else:
            self._i1 = gyr * self.v1
            self._i2 = gyr * self.v2
        self._deltai = self._i2 - self._i1
        self._rem = (self.per - self.tr - self.pw - self.tf
                     if self.pw < np.inf else 0.)

    get_TDsource = pulse_value


devList = [Resistor, Capacitor, Inductor, Gyrator, VCCS, IDC, VDC, ISin,
           VSin, IPulse, VPulse]
