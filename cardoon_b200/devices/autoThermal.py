"""
Electro-thermal wrapping of nonlinear devices (``<dev>_t``).

Class factory with the interface of the reference's devices/autoThermal.py
(thermal_device :25-138): adds a thermal port pair after the regular
terminals, appends the thermal port to ``controlPorts`` and a power output to
``csOutPorts``.  The reference's ``eval_cqs`` (:97-124) calls the base
class's ``set_temp_vars`` with the AD-typed port temperature before every
evaluation; here that happens inside the model's electro-thermal kernel
(csrc/models/thermal.cuh): the base class provides ``cuda_pack_thermal()``,
which returns the base record followed by the raw, temperature-independent
inputs of ``set_temp_vars``, and the kernel recomputes the temperature
dependent entries as duals (one more derivative lane for the temperature).
A base class without ``cuda_pack_thermal`` still parses and builds its
ports, but raises when a kernel record is requested.
"""
import numpy as np

from ..globalVars import glVar
from .. import circuit as cir
from . import cudadev as ad
from .param_layout import MODEL_IDS


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

            # ambient temperature the port temperature is added to
            # (autoThermal.py:110: vPort[tpn] + glVar.temp)
            self._tamb = glVar.temp
            ad.delete_tape(self)

        cudaModel = (nle.cudaModel + '_t') \
            if getattr(nle, 'cudaModel', None) else None

        def cuda_pack(self):
            if not hasattr(nle, 'cuda_pack_thermal') \
                    or self.cudaModel not in MODEL_IDS:
                raise cir.CircuitError(
                    '{0}: no electro-thermal kernel for {1} yet'
                    .format(self.instanceName, self.devType))
            return nle.cuda_pack_thermal(self)

        def power(self, vPort, ioutV):
            return ioutV[self._ton]

    ThermalDevice.__name__ = 'Thermal_' + nle.devType
    return ThermalDevice
