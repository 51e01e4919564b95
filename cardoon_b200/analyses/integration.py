"""
Fixed-step integration formulas (reference cardoon/analyses/integration.py:
BEuler :13, Trapezoidal :52).

    dq/dt at t_{n+1}  =  a0 * q_{n+1}  -  f_n1()

Host copies used by the ``nd`` interface (``set_IC``/``accept``/``get_rhs``
called from Python).  The device-resident time loop implements the same two
recurrences inside the kernels (csrc/engine_common.cuh, ``integ_accept``).

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import numpy as np


class BEuler:
    """q' = (q_{n+1} - q_n)/h"""

    def init(self, h, q):
        self.h = h
        self.a0 = 1. / h
        self.qn1 = np.zeros_like(q)

    def set_h(self, h):
        self.h = h
        self.a0 = 1. / h

    def accept(self, q):
        self.qn1[:] = q

    def f_n1(self):
        return self.a0 * self.qn1


class Trapezoidal:
    """q' = (2/h)(q_{n+1} - q_n) - q'_n"""

    def init(self, h, q, dq=None):
        self.h = h
        self.a0 = 2. / h
        self.qnm1 = np.copy(q)
        self.dqnm1 = np.zeros_like(q) if dq is None else np.copy(dq)

    def set_h(self, h):
        # reference integration.py:81 sets a0 = 1/h here (unused by the
        # fixed-step driver); kept as is
        self.h = h
        self.a0 = 1. / h

    def accept(self, q):
        self.dqnm1 *= -1.
        self.dqnm1 += self.a0 * (q - self.qnm1)
        self.qnm1[:] = q

    def f_n1(self):
        return self.a0 * self.qnm1 + self.dqnm1
