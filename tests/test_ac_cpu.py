"""
AC sweep (reference analyses/ac.py + nodal.py run_AC):
the oracle restatement against closed-form answers -- no GPU.
"""
import os
import tempfile

import numpy as np

from tests import support as S
from cardoon_b200.globalVars import glVar
from oracle import nodal_ref as R

RC = """# RC low-pass and a diode small-signal conductance
vsin:vs 1 gnd mag=1. acmag=2. phase=30. freq=1kHz
res:r1 1 2 r=1k
cap:c1 2 gnd c=1uF
vdc:vb 3 gnd vdc=0.65
res:r2 3 4 r=10.
diode:d1 4 gnd isat=1e-14 cj0=2e-12
ind:l1 2 5 l=1mH
res:r3 5 gnd r=50.
.end
"""


def load_text(text):
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'deck.net')
        with open(path, 'w') as f:
            f.write(text)
        return S.load_circuit(path)


def test_rc_lowpass_closed_form():
    ckt, _ = load_text(RC)
    f = np.logspace(0., 6., 25)
    r = R.run_ac(ckt, glVar, f)
    rows = r['idx'].name_to_row()
    w = 2j * np.pi * f
    # node 2: source 2 exp(j30deg) behind 1k, loaded by C || (jwL + 50)
    zl = 1. / (w * 1e-6 + 1. / (w * 1e-3 + 50.))
    vs = 2. * np.exp(1j * np.pi / 6.)
    v2 = vs * zl / (1e3 + zl)
    assert np.allclose(r['X'][rows['2']], v2, rtol=1e-9)
    assert np.allclose(r['X'][rows['1']], vs, rtol=1e-9)
    # the DC branch carries no AC signal
    assert np.all(np.abs(r['X'][rows['4']]) < 1e-12)


def test_delay_phase_factor():
    """mesfetc: delayed control ports multiply the Jacobian columns by
    exp(-j w tau) (nodal.py:354-357): check the oracle against an explicit
    evaluation at one frequency."""
    ckt, _ = S.load_circuit(S.net('mesfetc_plain.net'))
    f = np.array([1e9, 5.1e9, 2e10])
    r = R.run_ac(ckt, glVar, f)
    assert np.all(np.isfinite(r['X']))
    rows = r['idx'].name_to_row()
    # the drain node responds to the 0.2 V gate drive with some gain
    g = np.abs(r['X'][rows['4']]) / 0.2
    assert np.all(g > 1e-3) and np.all(g < 1e3)
    # without the phase factors the answer is different at tau * w ~ 0.6 rad
    elem = [e for e in S.nonlinear_elements(ckt)][0]
    tau = elem.tau
    elem.tau = 0.
    ckt2, _ = S.load_circuit(S.net('mesfetc_plain.net'))
    for e in S.nonlinear_elements(ckt2):
        e.tau = 1e-15
        e.process_params()
    r0 = R.run_ac(ckt2, glVar, f)
    assert np.max(np.abs(r0['X'][rows['4']] - r['X'][rows['4']])
                  / np.abs(r['X'][rows['4']])) > 1e-3
    assert tau == 5e-12


TLINE = """# matched, lossless frequency-defined line (nsect = 0)
vsin:vs 1 gnd mag=1. acmag=2. freq=1GHz rint=50.
tlinpy4:tl1 1 gnd 2 gnd z0mag=50. k=4. alpha=0. length=0.03
res:rl 2 gnd r=50.
res:rdc 2 3 r=1k
vdc:vb 3 gnd vdc=1.
.end
"""


def test_frequency_defined_line_closed_form():
    """tlinpy4 with nsect=0 (tlinpy4.py:235-271): a matched lossless line
    delays the wave by length*sqrt(k)/c0 and leaves its magnitude alone; at
    DC it is a near-short (get_G_matrix)."""
    ckt, _ = load_text(TLINE)
    tl = [e for e in ckt.elemDict.values() if e.devType == 'tlinpy4'][0]
    assert tl.isFreqDefined and not tl.linearVCCS
    f = np.linspace(1e8, 5e9, 30)
    r = R.run_ac(ckt, glVar, f)
    rows = r['idx'].name_to_row()
    # load = 50 || 1k (the DC source is an AC short)
    zl = 1. / (1. / 50. + 1. / 1e3)
    gam = (zl - 50.) / (zl + 50.)
    td = 0.03 * 2. / 299792458.
    ph = np.exp(-2j * np.pi * f * td)
    # incident wave of amplitude acmag/2 at the input of a line matched at
    # the source end: V2 = V+ (1 + gamma) exp(-j w td)
    v2 = 1. * (1. + gam) * ph
    assert np.allclose(r['X'][rows['2']], v2, rtol=1e-6)
    # operating point: 1 V behind 1k into (50 + line) || 50
    assert abs(r['x_op'][rows['2']] - r['x_op'][rows['1']]) < 1e-3


def test_sparameter_line_equals_y_parameter_line():
    """tlinps4 (waves on internal ports) and tlinpy4 with nsect=0 (plain Y
    parameters) describe the same line: identical AC node voltages."""
    deck = """
isin:i1 gnd 4 idc=2m acmag=1.
res:r1 4 2 r=100
tlin{0}:t1 2 gnd 3 gnd tand=1e-3 z0mag=50 alpha=2. k=3. length=0.07
res:r2 3 gnd r=75
.end
"""
    f = np.linspace(1e9, 5e9, 40)
    res = []
    for kind in ('py4', 'ps4'):
        ckt, _ = load_text(deck.format(kind))
        r = R.run_ac(ckt, glVar, f)
        rows = r['idx'].name_to_row()
        res.append((r['X'][rows['2']], r['X'][rows['3']]))
    assert np.allclose(res[0][0], res[1][0], rtol=1e-9)
    assert np.allclose(res[0][1], res[1][1], rtol=1e-9)
