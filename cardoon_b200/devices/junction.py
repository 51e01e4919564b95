"""
P-N junction helper objects (parameter processing only).

Host-side mirror of the reference's ``Junction`` (devices/diode.py:20-98) and
``SVJunction`` (devices/svdiode.py:20-111): ``process_params`` and
``set_temp_vars`` compute the temperature-adjusted constants; the current and
charge laws themselves (``get_id``/``get_qd``/``get_idvd``) are evaluated by
the CUDA kernels (csrc/models/junction.cuh) from the packed record returned
by ``pack()``.
"""
import numpy as np

from ..globalVars import const
from .param_layout import NJ


class Junction:
    stateVariable = False

    def process_params(self, isat, n, fc, cj0, vj, m, xti, eg0, Tnomabs):
        self.isat = isat
        self.n = n
        self.fc = fc
        self.cj0 = cj0
        self.vj = vj
        self.m = m
        self._k1 = xti / self.n
        self._k2 = const.q * eg0 / self.n / const.k / Tnomabs
        self._k3 = const.q * eg0 / self.n / const.k
        self._k4 = 1. - self.m

    def _set_is(self, Tabs, Tnomabs):
        tnratio = Tabs / Tnomabs
        self._t_is = self.isat * pow(tnratio, self._k1) \
            * np.exp(self._k2 - self._k3 / Tabs)
        return tnratio

    def _set_cap(self, Tabs, Tnomabs, tnratio, vt, egapn, egap_t):
        self._t_vj = self.vj * tnratio - 3. * vt * np.log(tnratio) \
            - tnratio * egapn + egap_t
        self._t_cj0 = self.cj0 * (1. + self.m * (.0004 * (Tabs - Tnomabs)
                                                 + 1. - self._t_vj / self.vj))
        self._k5 = self._t_vj * self._t_cj0 / self._k4
        self._k6 = self._t_cj0 * pow(1. - self.fc, - self.m - 1.)
        self._k7 = ((1. - self.fc * (1. + self.m)) * self._t_vj * self.fc
                    + .5 * self.m * self._t_vj * self.fc * self.fc)

    def set_temp_vars(self, Tabs, Tnomabs, vt, egapn, egap_t):
        tnratio = self._set_is(Tabs, Tnomabs)
        self._kexp = vt * self.n
        if self.cj0 != 0.:
            self._set_cap(Tabs, Tnomabs, tnratio, vt, egapn, egap_t)

    def pack(self):
        """Record in param_layout.JUNCTION order."""
        rec = np.zeros(NJ)
        rec[0] = self._t_is
        rec[1] = self._kexp
        if self.cj0 != 0.:
            rec[2:7] = (self._t_vj, self._k4, self._k5, self._k6, self._k7)
            rec[7] = self.fc
            rec[8] = self.m
            # constant term of the forward-bias branch of get_qd (reference
            # diode.py:93-96), folded with the reference's own arithmetic
            rec[11] = self._k5 * (1. - pow(1. - self.fc, self._k4))
        if self.stateVariable:
            rec[9] = self._alpha
            rec[10] = self._svth
        return rec


def _pack_raw(self):
    """Raw record of the electro-thermal kernels (param_layout.JRAW): the
    temperature-independent inputs of set_temp_vars."""
    c6 = c8 = 0.
    if self.cj0 != 0.:
        c6 = pow(1. - self.fc, - self.m - 1.)
        c8 = 1. - pow(1. - self.fc, self._k4)
    return np.array([self.isat, self.n, self._k1, self._k2, self._k3,
                     self._k4, self.vj, self.cj0, self.m, self.fc, c6, c8])


Junction.pack_raw = _pack_raw


class SVJunction(Junction):
    """State-variable junction: i and v are both functions of x."""
    stateVariable = True

    def set_temp_vars(self, Tabs, Tnomabs, vt, egapn, egap_t):
        tnratio = self._set_is(Tabs, Tnomabs)
        # largest exponential argument allowed before the linear extension
        self._alpha = 1. / self.n / vt
        max_exp_arg = 5e10
        self._svth = np.log(max_exp_arg / self._alpha) / self._alpha
        self._kexp = self.n * vt * max_exp_arg
        if self.cj0 != 0.:
            self._set_cap(Tabs, Tnomabs, tnratio, vt, egapn, egap_t)


EMPTY = np.zeros(NJ)
