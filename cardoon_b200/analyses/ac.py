"""
AC sweep analysis (``.analysis ac``).

Driver with the behaviour and printed report of the reference's
cardoon/analyses/ac.py (:22-146): operating point first (device Newton), then
``nd.run_AC`` -- here one launch that factors and solves every frequency of
the sweep (csrc/ac.cu) -- then the output requests ``ac``, ``ac_mag``,
``ac_phase``, ``ac_dB``.

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import sys

import numpy as np

from ..paramset import ParamSet
from . import analysis
from . import nodalCUDA as nd
from .fsolve import solve, NoConvergenceError

reqTypes = ['ac', 'ac_mag', 'ac_phase', 'ac_dB']


class ACSweep(ParamSet):
    r"""
    AC Sweep
    --------

    Calculates an AC sweep of a circuit using the nodal approach.  After the
    analysis is complete, nodal voltages are saved in the terminals with the
    ``aC_`` prefix.  An OP analysis is performed first to obtain the Jacobian
    of the nonlinear devices.

    Example::

        .analysis ac start=100. stop=1MEG num=100 log=True
        .plot ac_dB 153 151 23
        .plot ac_phase 23
    """
    anType = "ac"

    paramDict = dict(
        start=('Frequency sweep start value', 'Hz', float, 1.),
        stop=('Frequency sweep stop value', 'Hz', float, 10.),
        log=('Use logarithmic scale', '', bool, False),
        num=('Number of points in sweep', '', int, 50),
        shell=('Drop to ipython shell after calculation', '', bool, False),
    )

    def __init__(self):
        ParamSet.__init__(self, self.paramDict)

    def run(self, circuit):
        print('******************************************************')
        print('                 AC sweep analysis')
        print('******************************************************')
        if hasattr(circuit, 'title'):
            print('\n', circuit.title, '\n')
        if not circuit._flattened:
            circuit.flatten()
            circuit.init()
        nd.make_nodal_circuit(circuit)
        dc = nd.DCNodal(circuit)
        x0 = dc.get_guess()
        sV = dc.get_source()
        print('System dimension: {0}'.format(circuit.nD_dimension))
        try:
            print('Calculating DC operating point ... ', end='')
            sys.stdout.flush()
            (x, res, iterations) = solve(x0, sV, dc.convergence_helpers)
            print('Succeded.\n')
        except NoConvergenceError as ce:
            print('Failed.\n')
            print(ce)
            return
        dc.save_OP(x)
        if self.log:
            fvec = np.logspace(start=np.log10(self.start),
                               stop=np.log10(self.stop), num=self.num)
        else:
            fvec = np.linspace(start=self.start, stop=self.stop, num=self.num)
        self.fvec = fvec
        self.xVec = nd.run_AC(circuit, fvec)
        analysis.process_requests(circuit, 'ac', fvec, 'Frequency [Hz]',
                                  'aC_V', log=self.log)
        analysis.process_requests(circuit, 'ac_mag', fvec, 'Frequency [Hz]',
                                  'aC_V', lambda x: abs(x), log=self.log)
        analysis.process_requests(circuit, 'ac_phase', fvec,
                                  'Frequency [Hz]', 'aC_V',
                                  lambda v: 180. / np.pi * np.angle(v),
                                  'degrees', self.log)
        analysis.process_requests(circuit, 'ac_dB', fvec, 'Frequency [Hz]',
                                  'aC_V', lambda v: 20. * np.log10(v), 'dB',
                                  self.log)


aClass = ACSweep
