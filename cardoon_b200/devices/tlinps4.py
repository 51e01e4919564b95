"""
4-terminal physical transmission line with scattering parameters (netlist
name ``tlinps4``), reference devices/tlinps4.py.

Frequency-defined only: two extra internal ports keep track of the incident
waves, which keeps the description well behaved for lossless lines and at DC.
Used by the AC sweep (``get_Y_matrix``) and by the operating point
(``get_G_matrix``); like the reference, the transient analysis ignores
frequency-defined elements.
"""
import numpy as np

from ..globalVars import const
from .. import circuit as cir


class Device(cir.Element):
    """
    4-Terminal Physical Transmission Line (S-parameter formulation)
    ---------------------------------------------------------------

    Netlist examples::

        tlinps4:tl1 in gnd out gnd z0mag=100. length=0.3m
        .model c_line tlinps4 (z0mag=75.00 k=7 fscale=1.e10 alpha = 59.9)

    Internal terminals: ``v1p`` (4), ``v2m`` (5) and a local reference (6);
    the 4-port "Y" matrix over the ports (0,1), (2,3), (4,6), (5,6) is::

            |              1/Z0    -s12/Z0 |
        Y = |             -s21/Z0    1/Z0  |
            | -1            1        s12   |
            |        -1    s21        1    |
    """
    category = "Distributed components"
    devType = "tlinps4"
    numTerms = 4
    isFreqDefined = True
    fPortsDefinition = [(0, 1), (2, 3), (4, 6), (5, 6)]
    paramDict = dict(
        k=('Effective relative dielectric constant', '', float, 1.),
        alpha=('Attenuation', 'dB/m', float, 0.1),
        z0mag=('Magnitude of characteristic impedance', 'Ohms', float, 50.),
        fscale=('Scaling frequency for attenuation', 'Hz', float, 0.),
        tand=('Loss tangent', '', float, 0.),
        length=('Line length', 'm', float, 0.1),
    )

    def process_params(self):
        self.clean_internal_terms()
        self.c = np.sqrt(self.k) / (self.z0mag * const.c0)
        self.l = (self.z0mag * self.z0mag) * self.c
        self.alpha_nepers = self.alpha / const.Np2dB
        self.add_internal_term('v1p', 'V')      # 4
        self.add_internal_term('v2m', 'V')      # 5
        self.add_reference_term()               # 6

    @staticmethod
    def _fill(y, invZ0, s12):
        y[0, 2] = invZ0
        y[0, 3] = -invZ0 * s12
        y[1, 2] = -invZ0 * s12
        y[1, 3] = invZ0
        y[2, 0] = -1.
        y[2, 2] = 1.
        y[2, 3] = s12
        y[3, 1] = -1.
        y[3, 2] = s12
        y[3, 3] = 1.
        return y

    def get_Y_matrix(self, freq):
        """4 x 4 (x nfreq) matrix at ``freq`` (never zero);
        reference tlinps4.py get_Y_matrix."""
        omega = 2. * np.pi * freq
        if self.fscale != 0.:
            alphaf = self.alpha_nepers * np.sqrt(freq / self.fscale)
        else:
            alphaf = self.alpha_nepers
        r = 2.0 * alphaf * self.z0mag
        g = self.tand * omega * self.c
        z1 = r + 1j * self.l * omega
        y1 = g + 1j * self.c * omega
        invZ0 = np.sqrt(y1 / z1)
        gamma = np.sqrt(z1 * y1)
        s12 = np.exp(-gamma * self.length)
        shape = (4, 4) + np.shape(freq)
        return self._fill(np.zeros(shape, dtype=complex), invZ0, s12)

    def get_G_matrix(self):
        """DC parameters (reference tlinps4.py get_G_matrix)."""
        if self.fscale != 0.:
            s12 = 1.
        else:
            s12 = np.exp(-self.alpha_nepers * self.length)
        return self._fill(np.zeros((4, 4)), 1. / self.z0mag, s12)

    def get_OP(self, vPort):
        iout = np.dot(self.get_G_matrix()[:2, :2], vPort)
        return dict(V1=vPort[0], V2=vPort[1], I1=iout[0], I2=iout[1])
