"""
GPU parity of the batched Monte-Carlo transient (BASELINE config 5, SURVEY
8d "M" / 8e): every instance of the batch must reproduce the oracle run of a
circuit carrying that instance's parameter variation.
"""
import numpy as np
import pytest

from tests import support as S
from cardoon_b200.analyses import mc
from cardoon_b200.globalVars import glVar
from oracle import nodal_ref as R

pytestmark = pytest.mark.gpu


def oracle_instance(j, nsamples):
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = queue[0]
    elems = S.nonlinear_elements(ckt)
    mc.bsim3_ring_variation()(j, elems)
    for e in elems:
        e.process_params()
    return R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im,
                      max_samples=nsamples)


def test_mc_ring_bsim3_instances(lib):
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = queue[0]
    nsamples = 90
    bt = mc.BatchTransient(ckt, (nsamples - .5) * an.tstep, an.tstep, an.im)
    assert len(bt.timeVec) == nsamples
    B = 37                       # ragged: not a multiple of the tiles per CTA
    params = bt.pack_instances(B, mc.bsim3_ring_variation())
    terms = ['1', '3', '4']
    wave, iters, status, opit = bt.run(params, terms)
    assert wave.shape == (B, nsamples, 3)
    assert np.all(status == 0)
    # instances differ from one another
    assert np.abs(wave[0] - wave[1]).max() > 1e-4
    rows = bt.save_rows(terms)
    for j in (0, 5, 36):
        ref = oracle_instance(j, nsamples)
        X = ref['X'][:, rows]
        err = np.abs(wave[j] - X)
        assert np.all(err < glVar.reltol * np.abs(X) + glVar.abstol), \
            (j, err.max())
        assert np.max(np.abs(iters[j] - ref['iters'])) <= 1
        assert abs(opit[j] - ref['op_iterations']) <= 1


def test_mc_matches_single_circuit(lib):
    """A batch whose instances all carry the nominal parameters equals the
    plain transient analysis of the circuit."""
    from cardoon_b200.analyses import nodalCUDA as nd
    from cardoon_b200.analyses.fsolve import solve
    from cardoon_b200.analyses.integration import Trapezoidal
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = queue[0]
    bt = mc.BatchTransient(ckt, 60.5 * an.tstep, an.tstep, an.im)
    params = bt.pack_instances(3, lambda j, elems: None)
    wave, iters, status, opit = bt.run(params, ['1', '3', '4'])
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    tran = nd.TransientNodal(ckt, Trapezoidal())
    x, res, it = solve(dc.get_guess(), tran.get_source(0.).copy(),
                       dc.convergence_helpers)
    w1, it1, r1, st1 = tran.run(x, bt.timeVec, bt.save_rows(['1', '3', '4']))
    for b in range(3):
        assert np.array_equal(wave[b], w1)
        assert np.array_equal(iters[b], it1)
