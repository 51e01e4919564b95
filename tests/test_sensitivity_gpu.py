"""
Parameter sensitivities (reference analyses/sensitivity.py:16-172,
nodal.py:661-672) on the CUDA back-end: device-level derivatives against
golden vectors produced by AD through the reference's own ``process_params``
+ ``eval_cqs`` (tests/golden/make_sens_golden.py), and the adjoint
sensitivity of an operating point against a re-solved finite difference.
"""
import os

import numpy as np
import pytest

from tests import support as S
from tests import ref_cases as RC
from tests.golden.make_sens_golden import CASES
from cardoon_b200.analyses import sensitivity as sens
from cardoon_b200.analyses import nodalCUDA as nd
from cardoon_b200.analyses.fsolve import solve

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'sens.npz')


@pytest.mark.parametrize('case', sorted(CASES))
def test_device_derivatives_vs_reference_ad(lib, case):
    g = np.load(GOLD)
    elem, xin = RC.build(case)
    names = [str(n) for n in g[case + '/names']]
    col = int(g[case + '/col'])
    assert np.array_equal(xin[:, col], g[case + '/xin'])
    di, dg, dsrc = sens.device_derivatives(elem, names, xin[:, col], lib)
    ref = g[case + '/di']
    assert di.shape == ref.shape
    # fourth-order differences with a 1e-3 relative step resolve a derivative
    # to ~1e-9 of (current / parameter): relative 1e-6, plus that floor
    p = np.array([abs(float(getattr(elem, n))) for n in names])
    cur = np.abs(g[case + '/i'])
    floor = 1e-8 * (cur[:, None] + 1e-6 * cur.max()) / (p[None, :] + 1e-300)
    err = np.abs(di - ref)
    assert np.all(err <= 1e-6 * np.abs(ref) + floor), \
        (names, (err / (1e-6 * np.abs(ref) + floor)).max(axis=0))
    assert np.abs(ref).max() > 0.
    if dg.size:
        rg = g[case + '/dg']
        assert np.allclose(dg, rg, rtol=1e-7,
                           atol=1e-9 * (np.abs(rg).max() + 1e-300))
    # the element is left as it was
    out0, _ = S.cuda_eval(lib, elem, xin[:, col:col + 1])
    elem2, _ = RC.build(case)
    out1, _ = S.cuda_eval(lib, elem2, xin[:, col:col + 1])
    assert np.array_equal(out0, out1)


def test_adjoint_sensitivity_of_operating_point(lib):
    """d V(node) / d p = -lambda^T (d i / d p) with Jac^T lambda = e_node,
    for a resistor and for the transistor of bias_npn, against central
    differences of re-solved operating points."""
    ckt, _ = S.load_circuit(S.net('bias_npn.net'))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    sV = dc.get_source().copy()
    x, res, it = solve(dc.get_guess(), sV, dc.convergence_helpers)
    node = ckt.termDict['1'].nD_namRC           # collector
    e_out = np.zeros(len(x))
    e_out[node] = 1.
    dc.get_i_Jac(x)
    lam = dc.get_adjoint_voltages(e_out)
    targets = []
    for elem in ckt.nD_elemList:
        if elem.devType == 'res':
            targets.append((elem, 'r'))
        if elem.isNonlinear:
            targets.append((elem, 'bf'))
            targets.append((elem, 'isat'))
    assert len(targets) >= 4
    for elem, name in targets[:5]:
        Ji = sens.get_i_derivatives(elem, [(name, getattr(elem, name))], x,
                                    lib=lib)
        dv = -float(lam @ Ji[:, 0])
        # re-solve at p (1 +- 1e-4)
        p = float(getattr(elem, name))
        vals = []
        for s in (-1., 1.):
            setattr(elem, name, p * (1. + s * 1e-4))
            elem.process_params()
            nd.restore_RCnumbers(elem)
            nd.process_nodal_element(elem, ckt)
            dc.refresh()
            xs, _, _ = solve(x.copy(), dc.get_source().copy(),
                             dc.convergence_helpers)
            vals.append(xs[node])
        setattr(elem, name, p)
        elem.process_params()
        nd.restore_RCnumbers(elem)
        nd.process_nodal_element(elem, ckt)
        dc.refresh()
        fd = (vals[1] - vals[0]) / (2e-4 * p)
        assert abs(dv - fd) < 2e-4 * abs(fd) + 1e-12, (elem.instanceName,
                                                       name, dv, fd)
