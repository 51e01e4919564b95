"""
Nonlinear-solve driver (reference cardoon/analyses/fsolve.py).

``solve`` walks the list of convergence helpers exactly like the reference
(:23-56).  ``fsolve_Newton`` (:59-128) is kept for callers that bring their
own ``get_deltax`` (e.g. homotopy helpers working through the ``nd``
interface); the CUDA back-end's own helpers run the identical iteration on
the device (csrc/engine_common.cuh ``newton_update``) and only report
``(x, res, iterations, success)``.

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import numpy as np

from ..globalVars import glVar


class NoConvergenceError(Exception):
    pass


def solve(x0, sV, convergence_helpers):
    for algorithm in convergence_helpers:
        if algorithm is None:
            raise NoConvergenceError(
                'Giving up. No convergence with any method')
        if algorithm.__doc__:
            print('\nTrying ' + algorithm.__doc__)
        try:
            (x, res, iterations) = algorithm(x0, sV)
        except NoConvergenceError as ce:
            print(str(ce))
        else:
            break
    return (x, res, iterations)


def fsolve_Newton(x0, get_deltax, f_eval):
    """Damped Newton: returns ``(x, res, niter, success)``.

    Whole-vector damping to ``maxdelta``; converged when every
    ``|dx_j| < reltol*max(|x_j|,|xnew_j|) + abstol`` (tested on the damped
    step); ``res = max|dx|``; optional error-function test (``errfunc``).
    """
    success = False
    x = x0
    n2 = True
    for i in range(glVar.maxiter):
        deltax = get_deltax(x)
        maxDelta = max(abs(deltax))
        if maxDelta > glVar.maxdelta:
            deltax *= glVar.maxdelta / maxDelta
        xnew = x + deltax
        n1 = np.all(abs(deltax) < (glVar.reltol
                                   * np.maximum(abs(x), abs(xnew))
                                   + glVar.abstol))
        res = max(abs(deltax))
        if n1:
            if glVar.errfunc:
                errFunc = max(abs(f_eval(xnew)))
                n2 = errFunc < glVar.abstol
                res = max(errFunc, res)
            else:
                n2 = True
        x = xnew
        if n1 and n2:
            success = True
            break
    return (x, res, i + 1, success)
