"""
Electro-thermal wrapping of nonlinear devices (``<dev>_t``).

Class factory with the interface of the reference's devices/autoThermal.py
(thermal_device :25-138): adds a thermal port pair after the regular
terminals, appends the thermal port to ``controlPorts`` and a power output to
``csOutPorts``.  The kernels need a second instantiation per model in which
``set_temp_vars`` runs on the device in dual arithmetic with the temperature
as an extra input; that instantiation is scheduled after the isothermal path
(SURVEY section 8f, rank 3), so for now ``cuda_pack`` of a ``_t`` device
raises.  Netlists and catalogs that only *name* ``_t`` devices keep parsing.
"""
import numpy as np

from ..globalVars import glVar
from .. import circuit as cir


def thermal_device(nle):
    class ThermalDevice(nle):
        """
        Generic electrothermal nonlinear element: the base device plus one
        thermal port (temperature rise over ambient in C); a current source
        equal to the instantaneous dissipated power drives that port.
        """
        devType = nle.devType + '_t'
        isNonlinear = True

        def __init__(self, instanceName):
            nle.__init__(self, instanceName)
            self._addThermalPorts = True
            if nle.numTerms:
                self.numTerms = nle.numTerms + 2

        def process_params(self):
            nle.process_params(self, thermal=True)
            if self._addThermalPorts:
                unit = '+{0} C'.format(glVar.temp)
                self.connection[self.numTerms - 1].unit = unit
                self.connection[self.numTerms - 2].unit = unit
                self.csOutPorts = list(self.csOutPorts) \
                    + [(self.numTerms - 1, self.numTerms - 2)]
                self.controlPorts = list(self.controlPorts) \
                    + [(self.numTerms - 2, self.numTerms - 1)]
                self._ton = len(self.csOutPorts) - 1
                self._tpn = len(self.controlPorts) - 1
                # reference quirk (autoThermal.py:51,71,85): base classes
                # re-arm a differently mangled flag, so ports are only added
                # on the first call
                self._addThermalPorts = False
            guess = getattr(self, 'vPortGuess', None)
            if guess is not None and \
                    len(guess) < len(self.controlPorts) + self.nDelays:
                self.vPortGuess = np.insert(guess, self._tpn, 0.)

        def cuda_pack(self):
            raise cir.CircuitError(
                '{0}: electro-thermal kernels ({1}) are not built yet'
                .format(self.instanceName, self.devType))

        def power(self, vPort, ioutV):
            return ioutV[self._ton]

    ThermalDevice.__name__ = 'Thermal_' + nle.devType
    return ThermalDevice
