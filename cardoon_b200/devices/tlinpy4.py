"""
4-terminal physical transmission line (netlist name ``tlinpy4``).

With ``nsect > 0`` the line expands into ``nsect`` R-L-C-G sections made of
gyrators, capacitors and conductances, exactly as the reference's
devices/tlinpy4.py:91-192 does -- including the input-node index rule
``inn = nps*i + tref`` (:150), which for lossy lines (R != 0, nps = 3) starts
section i at the *resistor* node of section i-1; results on soliton.net
depend on it, so it is kept.  The frequency-domain Y-matrix mode
(``nsect = 0``) belongs to AC analysis and is out of scope here.
"""
import numpy as np

from ..globalVars import const, glVar
from .. import circuit as cir


class Device(cir.Element):
    """
    4-Terminal Physical Transmission Line
    -------------------------------------

    Netlist example::

        tlinpy4:tl1 in gnd out gnd z0mag=100. length=0.3m nsect=20
        .model c_line tlinpy4 (z0mag=75.00 k=7 fscale=10.e9 alpha = 59.9)
    """
    category = "Distributed components"
    devType = "tlinpy4"
    numTerms = 4
    isFreqDefined = True
    fPortsDefinition = [(0, 1), (2, 3)]
    paramDict = dict(
        k=('Effective relative dielectric constant', '', float, 1.),
        alpha=('Attenuation', 'dB/m', float, 0.1),
        z0mag=('Magnitude of characteristic impedance', 'Ohms', float, 50.),
        fscale=('Scaling frequency for attenuation', 'Hz', float, 0.),
        tand=('Loss tangent', '', float, 0.),
        length=('Line length', 'm', float, 0.1),
        nsect=('Enable discrete approximation with n sections', '', int, 0),
        fopt=('Optimum frequency for discrete approximation', 'Hz', float,
              0),
    )

    def process_params(self):
        self.clean_internal_terms()
        # per-unit-length capacitance / inductance; k = epsilon_eff
        self.c = np.sqrt(self.k) / (self.z0mag * const.c0)
        self.l = (self.z0mag * self.z0mag) * self.c
        self.alpha_nepers = self.alpha / const.Np2dB
        if not self.nsect:
            # frequency-defined model (reference tlinpy4.py:114-118): Y
            # parameters in AC (get_Y_matrix), DC conductances in OP
            # (get_G_matrix); ignored in transient like in the reference
            self.isFreqDefined = True
            self.linearVCCS = []
            self.linearVCQS = []
            return
        self.isFreqDefined = False
        if self.connection[1] != self.connection[3]:
            raise cir.CircuitError(
                '{0}: nsect > 0 but sectional model requires port '
                'references to be the same node'.format(self.instanceName))
        delta_x = self.length / self.nsect
        if self.fscale * self.fopt != 0.:
            alphaf = self.alpha_nepers * np.sqrt(self.fopt / self.fscale)
        else:
            alphaf = self.alpha_nepers
        R = 2.0 * alphaf * self.z0mag * delta_x
        if self.fopt * self.tand != 0.:
            G = self.tand * 2. * np.pi * self.fopt * self.c * delta_x
        else:
            G = 0.
        L = self.l * delta_x
        C = self.c * delta_x
        indcap = L * glVar.gyr * glVar.gyr
        tref = self.add_reference_term()
        nps = 3 if R else 2
        gyr = glVar.gyr
        vccs = []
        vcqs = []
        for i in range(self.nsect):
            inn = nps * i + tref if i else 0
            gnn = self.add_internal_term('il{0}'.format(i),
                                         '{0} A'.format(gyr))
            if i < self.nsect - 1:
                cnn = self.add_internal_term('c{0}'.format(i), 'V')
            else:
                cnn = 2
            if R:
                rnn = self.add_internal_term('r{0}'.format(i), 'V')
                vccs += [((inn, rnn), (tref, gnn), gyr),
                         ((gnn, tref), (inn, rnn), gyr)]
                vccs.append(((rnn, cnn), (rnn, cnn), 1. / R))
            else:
                vccs += [((inn, cnn), (tref, gnn), gyr),
                         ((gnn, tref), (inn, cnn), gyr)]
            vcqs.append(((gnn, tref), (gnn, tref), indcap))
            vcqs.append(((cnn, 1), (cnn, 1), C))
            if G:
                vccs.append(((cnn, 1), (cnn, 1), G))
        self.linearVCCS = vccs
        self.linearVCQS = vcqs

    # ---- linear frequency-defined device methods (nsect = 0) ---------------
    def get_Y_matrix(self, freq):
        """Y parameters [[y11, y12], [y12, y11]] at ``freq`` (scalar or
        vector, never zero); reference tlinpy4.py:235-271."""
        omega = 2. * np.pi * freq
        if self.fscale != 0.:
            alphaf = self.alpha_nepers * np.sqrt(freq / self.fscale)
        else:
            alphaf = self.alpha_nepers
        r = 2.0 * alphaf * self.z0mag
        g = self.tand * omega * self.c
        z1 = r + 1j * self.l * omega
        y1 = g + 1j * self.c * omega
        Z0 = np.sqrt(z1 / y1)
        gamma = np.sqrt(z1 * y1)
        ctemp = gamma * self.length
        y11 = np.cosh(ctemp) / Z0 / np.sinh(ctemp)
        y12 = -1. / Z0 / np.sinh(ctemp)
        return np.array([[y11, y12], [y12, y11]])

    def get_G_matrix(self):
        """DC Y parameters (reference tlinpy4.py:274-293): attenuation
        alpha * length, floored at 1e-5."""
        ctemp = max(self.alpha_nepers * self.length, 1e-5)
        y11 = np.cosh(ctemp) / self.z0mag / np.sinh(ctemp)
        y12 = -1. / self.z0mag / np.sinh(ctemp)
        return np.array([[y11, y12], [y12, y11]])
