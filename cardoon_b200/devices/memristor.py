"""
Gyrator-based memristor (netlist name ``memr``).

Topology and parameter processing follow the reference's
devices/memristor.py (process_params :124-176); the nonlinear current
``gyr^2 * M(q) * v_im`` with ``q = c * v_vc`` (eval_cqs :178-197) is evaluated
on the GPU from the compiled form of the ``m`` expression (exprcode.py).

Terminals 0, 1; internal ``im`` (current/gyr), ``vc`` (q/C) and a local
reference.

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import numpy as np

from ..globalVars import glVar
from .. import circuit as cir
from . import cudadev as ad
from .exprcode import compile_expr, ExprError


class Device(ad.CudaModel, cir.Element):
    """
    Memristor
    ---------

    ``m`` is the memristance as a function of the charge ``q``.

    Netlist example::

        memr:m1 1 0 m = '1e3 * (np.cosh(1e6 * q)-1.)'
    """
    category = "Basic components"
    devType = "memr"
    cudaModel = 'memr'
    numTerms = 2
    isNonlinear = True

    paramDict = dict(
        m=('Memristance function M(q)', 'Ohms', str, 'abs(5e9*q)'),
        q0=('Initial charge', 'As', float, 0.),
        c=('Auxiliary capacitance', 'F', float, 10e-6),
        rleak=('Leackage resistance', 'Ohms', float, 1e9),
    )

    def __init__(self, instanceName):
        cir.Element.__init__(self, instanceName)

    def process_params(self):
        ad.delete_tape(self)
        self.clean_internal_terms()
        if not self.rleak:
            raise cir.CircuitError(self.instanceName
                                   + ': leackage resistance can not be zero')
        if not self.c:
            raise cir.CircuitError(self.instanceName
                                   + ': capacitance can not be zero')
        try:
            self._mcode = compile_expr(self.m, 'q')
        except ExprError as e:
            raise cir.CircuitError('{0}: Invalid expression: {1} ({2})'
                                   .format(self.instanceName, self.m, e))
        tim = self.add_internal_term('im', '{0} A'.format(glVar.gyr))
        tvc = self.add_internal_term('vc', 'V')
        tref = self.add_reference_term()
        if self.q0 != 0.:
            self.isDCSource = True
            self.sourceOutput = (tref, tvc)
            self._i0 = self.q0 / self.c / self.rleak
        self.linearVCCS = [((0, 1), (tref, tim), glVar.gyr),
                           ((tim, tref), (0, 1), glVar.gyr),
                           ((tim, tref), (tref, tvc), glVar.gyr),
                           ((tvc, tref), (tvc, tref), 1. / self.rleak)]
        self.linearVCQS = [((tvc, tref), (tvc, tref), self.c)]
        self.controlPorts = [(tim, tref), (tvc, tref)]
        self.csOutPorts = [(tim, tref)]
        self.qsOutPorts = []

    def cuda_pack(self):
        return 0, np.array([self.c, glVar.gyr] + self._mcode)

    def get_OP(self, vPort):
        out = self.eval(vPort)
        q = self.c * vPort[1]
        return {'v': out[0] / glVar.gyr, 'i': vPort[0] * glVar.gyr, 'q': q}

    def get_DCsource(self):
        return self._i0
