"""
DC sweep analysis (netlist: ``.analysis dc``).

Driver with the interface, printed report and result attributes
(``term.dC_v``, ``circuit.dC_sweep``) of the reference's
cardoon/analyses/dc.py (DCSweep :24-255).  The sweep loop keeps the
reference's mechanism (:167-205): ``setattr`` the parameter (or
``set_temp_vars`` for a global temperature sweep), ``process_params``,
``nd.restore_RCnumbers``, ``nd.process_nodal_element``, ``dc.refresh()``,
then ``solve`` starting from the previous point's solution.  Every solve is a
device Newton launch; ``refresh`` keeps the symbolic plan when only device
records or source values changed.

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import numpy as np

from ..paramset import ParamSet
from ..globalVars import glVar
from . import analysis
from .fsolve import solve, NoConvergenceError
from . import nodalCUDA as nd


class DCSweep(ParamSet):
    """
    DC Sweep
    --------

    Calculates a DC sweep of a circuit using the nodal approach.  After the
    analysis is complete, nodal voltages are saved in circuit and terminals
    with the ``dC_`` prefix.

    The following parameters can be swept:

      * Any device parameter of ``float`` type (device name must be
        specified in this case)

      * Global temperature (no device specified): sweep temperature of all
        devices that do not explicitly have ``temp`` set.

    Examples::

        .analysis dc device=vsin:v1 param=vdc start=-2. stop=2. num=50
        .analysis dc param=temp start=-20C stop=80C
        .plot dc 153 151 23
    """
    anType = "dc"

    paramDict = dict(
        device=('Instance name of device to sweep variable', '', str, ''),
        param=('Parameter to sweep', '', str, ''),
        start=('Sweep start value', '(variable)', float, 0.),
        stop=('Sweep stop value', '(variable)', float, 0.),
        num=('Number of points in sweep', '', int, 50),
        verbose=('Show iterations for each point', '', bool, False),
        shell=('Drop to ipython shell after calculation', '', bool, False)
    )

    def __init__(self):
        ParamSet.__init__(self, self.paramDict)

    def run(self, circuit):
        print('******************************************************')
        print('                 DC sweep analysis')
        print('******************************************************')
        if hasattr(circuit, 'title'):
            print('\n', circuit.title, '\n')
        if not glVar.sparse:
            print('Using dense matrices\n')
        if not circuit._flattened:
            circuit.flatten()
            circuit.init()

        tempFlag = False
        if self.device:
            try:
                dev = circuit.elemDict[self.device]
            except KeyError:
                raise analysis.AnalysisError(
                    'Could not find: {0}'.format(self.device))
            if self.param:
                try:
                    pinfo = dev.paramDict[self.param]
                except KeyError:
                    raise analysis.AnalysisError('Unrecognized parameter: '
                                                 + self.param)
                if not pinfo[2] == float:
                    raise analysis.AnalysisError(
                        'Parameter must be float: ' + self.param)
            else:
                raise analysis.AnalysisError(
                    "Don't know what parameter to sweep!")
            paramunit = pinfo[1]
        else:
            if self.param != 'temp':
                raise analysis.AnalysisError(
                    'Only temperature sweep supported if no device specified')
            paramunit = 'C'
            tempFlag = True

        nd.make_nodal_circuit(circuit)
        dc = nd.DCNodal(circuit)
        x = dc.get_guess()

        sweepvar = np.linspace(start=self.start, stop=self.stop, num=self.num)
        circuit.dC_sweep = sweepvar
        if tempFlag:
            circuit.dC_var = 'Global temperature sweep: temp'
        else:
            circuit.dC_var = 'Device: ' + dev.instanceName \
                + '  Parameter: ' + self.param
        circuit.dC_unit = paramunit
        print('System dimension: {0}'.format(circuit.nD_dimension))
        print('Sweep: ', circuit.dC_var)

        xVec = np.zeros((circuit.nD_dimension, self.num))
        self.iterations = np.zeros(self.num, dtype=int)
        tIter = 0
        tRes = 0.
        for i, value in enumerate(sweepvar):
            if tempFlag:
                for elem in circuit.nD_elemList:
                    if not elem.is_set('temp'):
                        try:
                            elem.set_temp_vars(value)
                        except AttributeError:
                            pass
            else:
                setattr(dev, self.param, value)
                dev.process_params()
                nd.restore_RCnumbers(dev)
                nd.process_nodal_element(dev, circuit)
            dc.refresh()
            sV = dc.get_source()
            try:
                (x, res, iterations) = solve(x, sV, dc.convergence_helpers)
            except NoConvergenceError as ce:
                print(ce)
                return
            xVec[:, i] = x
            self.iterations[i] = iterations
            tIter += iterations
            tRes += res
            if self.verbose:
                print('{0} = {1}'.format(self.param, value))
                print('Number of iterations = ', iterations)
                print('Residual = ', res)

        avei = tIter / len(sweepvar)
        aver = tRes / len(sweepvar)
        print('Average iterations: {0}'.format(avei))
        print('Average residual: {0}\n'.format(aver))

        circuit.nD_ref.dC_v = np.zeros(self.num)
        for i, term in enumerate(circuit.nD_termList):
            term.dC_v = xVec[i, :]

        # restore original attribute values
        if tempFlag:
            for elem in circuit.nD_elemList:
                if not elem.is_set('temp'):
                    elem.set_attributes()
                    try:
                        elem.set_temp_vars(value)
                    except AttributeError:
                        pass
        else:
            dev.set_attributes()
            dev.process_params()
            nd.restore_RCnumbers(dev)
            nd.process_nodal_element(dev, circuit)

        analysis.process_requests(circuit, 'dc', sweepvar,
                                  '{0} [{1}]'.format(circuit.dC_var,
                                                     circuit.dC_unit),
                                  'dC_v')

        def getvec(termname):
            return circuit.find_term(termname).dC_v

        if self.shell:
            analysis.ipython_drop("""
Available commands:
    sweepvar: vector with swept parameter
    getvec(<terminal>) to retrieve results
""", globals(), locals())


aClass = DCSweep
