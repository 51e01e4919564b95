"""
GPU parity of the AC sweep (cdn_ac_solve through nodalCUDA.run_AC and the
`ac` analysis driver) against the oracle restatement of run_AC
(nodal.py:286-381): complex node voltages within 1e-9 relative at every
frequency.
"""
import numpy as np
import pytest

from tests import support as S
from tests.test_ac_cpu import load_text, RC
from cardoon_b200.analyses import nodalCUDA as nd
from cardoon_b200.analyses.fsolve import solve
from cardoon_b200.globalVars import glVar
from oracle import nodal_ref as R

pytestmark = pytest.mark.gpu


def ac_both(ckt, f):
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    x, res, it = solve(dc.get_guess(), dc.get_source(), dc.convergence_helpers)
    dc.save_OP(x)
    X = nd.run_AC(ckt, f)
    ref = R.run_ac(ckt, glVar, f)
    return X, ref


def check_ac(X, ref):
    Xr = ref['X']
    assert X.shape == Xr.shape
    # (nodes pinned by DC sources carry only rounding noise: absolute floor
    # relative to the largest signal in the circuit)
    scale = np.max(np.abs(Xr))
    err = np.abs(X - Xr)
    assert np.all(err <= 1e-9 * np.abs(Xr) + 1e-12 * scale), \
        float(np.max(err / (np.abs(Xr) + 1e-3 * scale)))


def test_ac_rc_and_diode(lib):
    ckt, _ = load_text(RC)
    X, ref = ac_both(ckt, np.logspace(0., 6., 40))
    check_ac(X, ref)
    assert ckt.termDict['2'].aC_V.shape == (40,)


def test_ac_amplifier_driver(lib, capsys):
    """`.analysis ac` of tests/netlists/ac_amp.net through the analysis class
    (report, log sweep, aC_V attributes)."""
    ckt, queue = S.load_circuit(S.net('ac_amp.net'))
    an = [a for a in queue if a.anType == 'ac'][0]
    ckt.saveReqList = []
    ckt.plotReqList = []
    an.run(ckt)
    out = capsys.readouterr().out
    assert 'AC sweep analysis' in out and 'System dimension: 13' in out
    ckt2, _ = S.load_circuit(S.net('ac_amp.net'))
    ref = R.run_ac(ckt2, glVar, an.fvec)
    check_ac(an.xVec, ref)
    gain = np.abs(ckt.termDict['7'].aC_V) / 0.01
    assert gain.max() > 50.


def test_ac_mesfet_delayed_ports(lib):
    ckt, _ = S.load_circuit(S.net('mesfetc_plain.net'))
    X, ref = ac_both(ckt, np.linspace(1e8, 2e10, 64))
    check_ac(X, ref)


def test_ac_large_dense_system(lib):
    """n > 56: the matrices live in the global workspace (soliton's diode
    line truncated to a 150-node ladder)."""
    lines = ['vsin:vs 1 gnd mag=1. acmag=1. freq=1GHz rint=50.']
    for k in range(1, 150):
        lines.append('ind:l{0} {0} {1} l=2n'.format(k, k + 1))
        lines.append('diode:d{0} {1} gnd isat=1e-12 cj0=1p'.format(k, k + 1))
        lines.append('res:r{0} {1} gnd r=5k'.format(k, k + 1))
    lines.append('res:rl 150 gnd r=50.')
    lines.append('.end')
    ckt, _ = load_text('\n'.join(lines) + '\n')
    X, ref = ac_both(ckt, np.logspace(7., 10., 24))
    assert ckt.nD_dimension > 200
    check_ac(X, ref)


def test_ac_panel_smem_opt_in(lib):
    """More than 600 unknowns: the panel's multipliers need more than 48 KB of
    shared memory (opt-in attribute); more frequencies than workspace slots."""
    lines = ['vsin:vs 1 gnd mag=1. acmag=1. freq=1GHz rint=50.']
    for k in range(1, 330):
        lines.append('ind:l{0} {0} {1} l=2n'.format(k, k + 1))
        lines.append('cap:c{0} {1} gnd c=1p'.format(k, k + 1))
        lines.append('res:r{0} {1} gnd r=5k'.format(k, k + 1))
    lines.append('res:rl 330 gnd r=50.')
    lines.append('.end')
    ckt, _ = load_text('\n'.join(lines) + '\n')
    X, ref = ac_both(ckt, np.logspace(7., 9.5, 12))
    assert ckt.nD_dimension > 600
    check_ac(X, ref)


def test_ac_frequency_defined_line(lib):
    """tlinpy4 nsect=0: DC Y parameters in the operating point, get_Y_matrix
    per frequency in the sweep (dense and sparse OP formulations)."""
    from tests.test_ac_cpu import TLINE
    for opts in ('', '.options sparse=1\n'):
        ckt, _ = load_text(opts + TLINE)
        X, ref = ac_both(ckt, np.linspace(1e8, 5e9, 48))
        check_ac(X, ref)
        x_op = np.array([t.nD_vOP for t in ckt.nD_termList])
        assert np.allclose(x_op, ref['x_op'], rtol=1e-9, atol=1e-12)
