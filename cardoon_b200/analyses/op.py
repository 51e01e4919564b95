"""
Operating-point analysis (netlist: ``.analysis op``).

Driver with the interface and printed report of the reference's
cardoon/analyses/op.py (DCOP :21-155).  The solve itself runs on the GPU:
``nodalCUDA.DCNodal`` evaluates devices, assembles, factors and iterates
Newton inside one kernel launch per convergence helper.

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
from ..paramset import ParamSet
from ..globalVars import glVar
from .analysis import ipython_drop
from .fsolve import solve, NoConvergenceError
from . import nodalCUDA as nd


class DCOP(ParamSet):
    """
    DC Operating Point
    ------------------

    Calculates the DC operating point of a circuit using the nodal approach.
    After the analysis is complete, nodal voltages are saved in circuit and
    terminals with the ``nD_`` prefix.  After this the analysis drops to an
    interactive shell if the ``shell`` global variable is set to ``True``.

    Example::

        .analysis op intvars=1 shell=1
    """
    anType = "op"

    paramDict = dict(
        intvars=('Print internal element nodal variables', '', bool, False),
        elemop=('Print element operating points', '', bool, False),
        shell=('Drop to ipython shell after calculation', '', bool, False),
    )

    def __init__(self):
        ParamSet.__init__(self, self.paramDict)

    def run(self, circuit):
        print('******************************************************')
        print('             Operating point analysis')
        print('******************************************************')
        if hasattr(circuit, 'title'):
            print('\n', circuit.title, '\n')
        if not glVar.sparse:
            print('Using dense matrices\n')
        if not circuit._flattened:
            circuit.flatten()
            circuit.init()
        nd.make_nodal_circuit(circuit)
        dc = nd.DCNodal(circuit)
        x0 = dc.get_guess()
        sV = dc.get_source()
        try:
            (x, res, iterations) = solve(x0, sV, dc.convergence_helpers)
        except NoConvergenceError as ce:
            print(ce)
            return
        dc.save_OP(x)
        self.x, self.res, self.iterations = x, res, iterations

        print('Number of iterations = ', iterations)
        print('Residual = ', res)
        print('\n Node      |  Value               | Unit ')
        print('----------------------------------------')
        for key in sorted(circuit.termDict):
            term = circuit.termDict[key]
            print('{0:10} | {1:20} | {2}'.format(key, term.nD_vOP,
                                                 term.unit))
        if self.intvars or self.elemop:
            for elem in circuit.nD_nlinElem:
                print('\nElement: ', elem.instanceName)
                if self.intvars:
                    print('\n    Internal nodal variables:\n')
                    for term in elem.get_internal_terms():
                        print('    {0:10} : {1:20} {2}'.format(
                            term.instanceName, term.nD_vOP, term.unit))
                if self.elemop:
                    print('\n    Operating point info:\n')
                    try:
                        opDict = elem.get_OP(elem.nD_xOP)
                        for key in sorted(opDict):
                            print('    {0:10} : {1}'.format(key,
                                                            opDict[key]))
                    except AttributeError:
                        print('    (none)')
        print('\n')

        def getvar(termname):
            return circuit.find_term(termname).nD_vOP

        def getterm(termname):
            return circuit.find_term(termname)

        def getdev(elemname):
            return circuit.elemDict[elemname]

        if self.shell:
            ipython_drop("""
Available commands:
    getvar(<terminal name>) returns variable at <terminal name>
    getterm(<terminal name>) returns terminal reference
    getdev(<device name>) returns device reference
""", globals(), locals())


aClass = DCOP
