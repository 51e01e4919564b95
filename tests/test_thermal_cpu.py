"""
Electro-thermal devices (reference devices/autoThermal.py):
CPU checks of the oracle restatement and of the host packer -- no GPU.

The reference holds no golden data for `_t` devices, so the thermal oracle is
pinned against the isothermal oracle (itself anchored by the KATs): a thermal
device whose port temperature is dT must behave exactly like the isothermal
device at temp = ambient + dT, its extra output must be the dissipated power,
and the derivative lane must agree with a finite difference in temperature.
"""
import numpy as np

from tests import support as S
from cardoon_b200.globalVars import glVar
from cardoon_b200.devices.param_layout import LAYOUTS
from oracle.models import device_eval
from oracle import nodal_ref as R


def make_pair(dT, **kw):
    dt = S.make_element('diode_t', [1, 0, 2, 0], **kw)
    d = S.make_element('diode', [1, 0], temp=glVar.temp + dT, **kw)
    return dt, d


def test_ports_and_record():
    dt, d = make_pair(0., cj0=1e-12, tt=1e-9, rs=5.)
    assert dt.devType == 'diode_t' and dt.numTerms == 4
    assert len(dt.controlPorts) == len(d.controlPorts) + 1
    assert len(dt.csOutPorts) == len(d.csOutPorts) + 1
    assert len(dt.qsOutPorts) == len(d.qsOutPorts) == 1
    # thermal port: control (n-2, n-1), output (n-1, n-2) (autoThermal.py:76-79)
    assert tuple(dt.controlPorts[-1]) == (2, 3)
    assert tuple(dt.csOutPorts[-1]) == (3, 2)
    spec = dt.cuda_spec()
    assert spec['params'].shape == (len(LAYOUTS['diode_t'][0]),)
    assert spec['nxin'] == 2 and spec['ncs'] == 2 and spec['nqs'] == 1
    # the base part of the record is the isothermal record at ambient
    base = S.make_element('diode', [1, 0], cj0=1e-12, tt=1e-9, rs=5.)
    nb = len(LAYOUTS['diode'][0])
    assert np.array_equal(spec['params'][:nb], base.cuda_spec()['params'])


def test_thermal_equals_isothermal_at_port_temperature():
    rng = np.random.default_rng(7)
    vd = rng.uniform(-3., 0.85, 512)
    for dT in (0., 13.5, 73., -40.):
        for kw in (dict(), dict(cj0=2e-12, tt=1e-9), dict(bv=4., n=1.3),
                   dict(cj0=1e-12, m=0.33, fc=0.6, area=2.)):
            dt, d = make_pair(dT, **kw)
            xin = np.vstack([vd, np.full_like(vd, dT)])
            ot, jt = device_eval(dt, xin, True)
            oi, ji = device_eval(d, xin[:1], True)
            assert np.allclose(ot[0], oi[0], rtol=1e-12, atol=0.)
            assert np.allclose(ot[1], vd * oi[0], rtol=1e-12, atol=0.)   # power
            if oi.shape[0] > 1:
                assert np.allclose(ot[2], oi[1], rtol=1e-12, atol=1e-300)
            # d/dvd of the thermal device = isothermal derivative
            assert np.allclose(jt[0, 0], ji[0, 0], rtol=1e-12, atol=0.)


def test_temperature_derivative_matches_finite_difference():
    dt, _ = make_pair(0., cj0=2e-12, tt=1e-9, bv=6.)
    vd = np.linspace(-2., 0.8, 57)
    T0, h = 20., 1e-3
    xin = np.vstack([vd, np.full_like(vd, T0)])
    o, j = device_eval(dt, xin, True)
    op, _ = device_eval(dt, xin + np.array([[0.], [h]]), False)
    om, _ = device_eval(dt, xin - np.array([[0.], [h]]), False)
    fd = (op - om) / (2. * h)
    scale = np.abs(j[:, 1]).max(axis=1, keepdims=True)
    assert np.all(np.abs(fd - j[:, 1]) <= 1e-6 * np.abs(j[:, 1]) + 1e-9 * scale)


def test_diode_thermal_operating_point_self_consistent():
    """test/diode_thermal.net: the thermal node sits at rth * (vd * id), and an
    isothermal diode at the resulting temperature gives the same voltages."""
    ckt, _ = S.load_circuit(S.net('diode_thermal.net'))
    r = R.run_op(ckt, glVar)
    rows = r['idx'].name_to_row()
    x = r['x']
    v3, dT = x[rows['3']], x[rows['1000']]
    assert 0.01 < dT < 50.
    d1 = [e for e in S.nonlinear_elements(ckt)][0]
    byname = {t.instanceName: r['idx'].row(t) for t in r['idx'].termList}
    t2 = [v for k, v in byname.items() if k.endswith('t2')][0]
    vd = x[t2]
    out, _ = device_eval(d1, np.array([[vd], [dT]]), False)
    # (Newton stops at reltol = 1e-4 of the last step: consistency to ~1e-6)
    assert abs(out[1, 0] * 100. - dT) < 1e-5 * dT                 # rth = 100
    assert abs((v3 - vd) / 50. - out[0, 0]) < 1e-5 * abs(out[0, 0])   # rs = 50
    # isothermal replica at ambient + dT
    import tempfile, os
    deck = open(S.net('diode_thermal.net')).read()
    deck = deck.replace('diode_t:d1 3 gnd 1000 gnd rs=50.',
                        'diode:d1 3 gnd rs=50. temp={0:.17g}'.format(
                            float(glVar.temp + dT)))
    deck = deck.replace('.save dc 3 1000', '.save dc 3')
    deck = deck.replace('res:rth 1000 gnd r=100.', 'res:rth 1000 gnd r=100.\n'
                        'res:rth2 1000 3 r=1e12')
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'iso.net')
        open(path, 'w').write(deck)
        ckt2, _ = S.load_circuit(path)
        r2 = R.run_op(ckt2, glVar)
    assert abs(r2['x'][r2['idx'].name_to_row()['3']] - v3) < 1e-6
