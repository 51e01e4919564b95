"""
Transient analysis (netlist: ``.analysis tran``).

Driver with the interface, printed report and result attributes
(``term.tran_v``, ``circuit.tran_timevec``) of the reference's
cardoon/analyses/tran.py (Transient :27-231).  The operating point is one
device Newton launch per convergence helper; the fixed-step loop of
tran.py:171-206 (accept, right-hand side, chord predictor, Newton, result
capture) runs as ONE kernel launch for all samples
(``nodalCUDA.TransientNodal.run``), so a Newton iteration never returns to
the host.  Source waveforms are evaluated on the host for all samples
beforehand (a few scalar function calls per sample) and uploaded once.

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
from ..globalVars import glVar
from . import analysis
from .fsolve import solve, NoConvergenceError
from .integration import BEuler, Trapezoidal
from . import nodalCUDA as nd


class Transient(ParamSet):
    """
    Transient Analysis
    ------------------

    Solves nodal equations starting from ``t=0`` to ``tstop`` with a
    fixed time step equal to ``tstep``.  Two integration methods are
    available: Backwards Euler (``im = BE``) and trapezoidal (``im = trap``).
    Convergence parameters for the Newton method are controlled using the
    global variables in ``.options``.

    One plot window is generated for each ``.plot`` statement.  Use ``tran``
    request type for this analysis.

    Example::

        .analysis tran tstop=1ms tstep=.01ms im=BE

        .plot tran vin vout
    """
    anType = 'tran'

    paramDict = dict(
        tstop=('Simulation stop time', 's', float, 1e-3),
        tstep=('Time step size', 's', float, 1e-5),
        im=('Integration method', '', str, 'trap'),
        verbose=('Show iterations for each point', '', bool, False),
        saveall=('Save all nodal voltages', '', bool, False),
        shell=('Drop to ipython shell after calculation', '', bool, False),
    )

    def __init__(self):
        ParamSet.__init__(self, self.paramDict)

    def run(self, circuit):
        print('******************************************************')
        print('                 Transient analysis')
        print('******************************************************')
        if hasattr(circuit, 'title'):
            print('\n', circuit.title, '\n')
        if not glVar.sparse:
            print('Using dense matrices\n')
        if not circuit._flattened:
            circuit.flatten()
            circuit.init()
        if self.im == 'BE':
            imo = BEuler()
        elif self.im == 'trap':
            imo = Trapezoidal()
        else:
            raise analysis.AnalysisError(
                'Unknown integration method: {0}'.format(self.im))

        nd.make_nodal_circuit(circuit)
        dc = nd.DCNodal(circuit)
        tran = nd.TransientNodal(circuit, imo)
        x = dc.get_guess()
        sV = tran.get_source(0.)
        try:
            print('Calculating DC operating point ... ', end='')
            sys.stdout.flush()
            (x, res, iterations) = solve(x, sV, dc.convergence_helpers)
            print('Succeded.\n')
        except NoConvergenceError as ce:
            print('Failed.\n')
            print(ce)
            return
        dc.save_OP(x)
        self.op_iterations = iterations
        del dc

        timeVec = np.arange(start=0., stop=self.tstop, step=self.tstep,
                            dtype=float)
        nsamples = len(timeVec)
        circuit.tran_timevec = timeVec
        termSet = circuit.get_requested_terms('tran')
        termSet1 = [t for t in termSet if t is not circuit.nD_ref]
        if self.saveall:
            saveTerms = list(circuit.nD_termList)
        else:
            saveTerms = termSet1
        save_rows = [t.nD_namRC for t in saveTerms]

        print('System dimension: {0}'.format(circuit.nD_dimension))
        print('Number of samples: {0}'.format(nsamples))
        print('Integration method: {0}'.format(self.im))

        # ---- the whole time loop: one kernel launch ------------------------
        wave, iters, resv, status = tran.run(x, timeVec, save_rows)
        if status:
            print('No convergence at sample {0} (t = {1}): iter = {2} '
                  'res = {3}'.format(status, timeVec[status], iters[status],
                                     resv[status]))
            return

        if circuit.nD_ref in termSet or self.saveall:
            circuit.nD_ref.tran_v = np.zeros(nsamples)
        for k, term in enumerate(saveTerms):
            term.tran_v = np.ascontiguousarray(wave[:, k])
        if self.verbose:
            print('-------------------------------------------------')
            print(' Step    | Time (s)     | Iter.    | Residual    ')
            print('-------------------------------------------------')
            for i in range(1, nsamples):
                print('{0:8} | {1:12} | {2:8} | {3:12}'.format(
                    i, timeVec[i], iters[i], resv[i]))
        self.iterations = iters
        self.residuals = resv
        avei = float(iters.sum()) / nsamples
        aver = float(resv.sum()) / nsamples
        print('\nAverage iterations: {0}'.format(avei))
        print('Average residual: {0}\n'.format(aver))

        analysis.process_requests(circuit, 'tran', timeVec, 'Time [s]',
                                  'tran_v')

        def getvec(termname):
            return circuit.find_term(termname).tran_v

        if self.shell:
            analysis.ipython_drop("""
Available commands:
    timeVec: time vector
    getvec(<terminal>) to retrieve results (if result saved)
""", globals(), locals())


aClass = Transient
