"""
Pins the CPU oracle to the reference's own printed known-answer values
(tests/golden/reference_kat.json, transcribed from doc/usage.rst and
doc/design.rst of cechrist/cardoon).  No GPU needed.

KAT 1 anchors svbjt + SVJunction + the dense Newton loop (iteration count
and 12 printed digits); KAT 2 anchors EKV 2.6 + mosExt + the transient loop
(trapezoidal integration, chord predictor with the stale residual, damping
and convergence test): its average iteration count and 12-digit average
residual only come out right if every one of those is restated exactly.
"""
import json
import os

import numpy as np

from tests import support as S
from cardoon_b200.globalVars import glVar
from oracle import nodal_ref as R

with open(os.path.join(S.HERE, 'golden', 'reference_kat.json')) as f:
    KAT = json.load(f)


def digits(a, b, nd=11):
    return abs(a - b) <= 0.6 * 10.0 ** (-nd) * max(abs(b), 1e-300) * 10


def test_bias_npn_operating_point():
    k = KAT['bias_npn_op']
    ckt, _ = S.load_circuit(S.net('bias_npn.net'))
    r = R.run_op(ckt, glVar)
    assert r['iterations'] == k['iterations']
    assert digits(r['res'], k['residual'], 8)
    rows = r['idx'].name_to_row()
    x = r['x']
    for name, v in k['nodes'].items():
        assert digits(x[rows[name]], v), (name, x[rows[name]], v)
    byname = {t.instanceName: r['idx'].row(t) for t in r['idx'].termList}
    for name, v in k['internal'].items():
        assert digits(x[byname[name]], v), (name, x[byname[name]], v)
    # derived operating-point values (doc/usage.rst:70-79)
    assert digits((x[rows['10']] - x[rows['1']]) / 1e3, k['op']['IC'])
    assert digits((10. - x[rows['2']]) / 96e3, k['op']['IB'])


def test_ring_osc_ahkab_transient_statistics():
    k = KAT['ring_osc_ahkab_tran']
    ckt, queue = S.load_circuit(S.net('ring_osc_ahkab.net'))
    an = [a for a in queue if a.anType == 'tran'][0]
    r = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im)
    n = len(r['t'])
    assert r['idx'].dimension == k['dimension']
    assert n == k['samples']
    assert abs(r['iters'].sum() / float(n) - k['average_iterations']) < 1e-12
    assert digits(r['res'].sum() / n, k['average_residual'], 10)


def test_soliton_shape():
    """One Newton solve per step thanks to the chord predictor
    (doc/design.rst:243-256: 505 solves for 501 steps)."""
    k = KAT['soliton_tran']
    ckt, queue = S.load_circuit(S.net('soliton.net'))
    an = [a for a in queue if a.anType == 'tran'][0]
    r = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im, max_samples=120)
    assert r['idx'].dimension == k['dimension']
    # the reference's profile: 501 samples = 500 steps; the first step takes
    # 2 iterations, every later one 1 (chord predictor), plus the operating
    # point: 4 + 2 + 499 = 505 solves
    steps = 119
    assert r['nfactor'] == steps + 1
    assert list(r['iters'][:3]) == [0, 2, 1]
    assert r['op_iterations'] + (k['steps'] - 1) + 1 == k['newton_solves']


def test_colpitts_dimension():
    ckt, _ = S.load_circuit(S.net('oscillator.net'))
    idx = R.prepare(ckt)
    assert idx.dimension == KAT['colpitts_dimension']['dimension']
