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


# ---- Gummel-Poon: bjt_t / svbjt_t -------------------------------------------
BJT_KW = [dict(bf=200., vaf=100., ikf=10e-3, rb=1e3),
          dict(isat=5e-17, bf=147., vaf=80., ikf=4.3e-3, ise=8e-18, ne=1.233,
               br=1.9, var=11., ikr=6e-4, isc=5e-16, nc=1.08, re=12.,
               rb=1200., rbm=200., rc=25., cje=58e-15, vje=0.83, mje=0.35,
               cjc=133e-15, vjc=0.6, mjc=0.44, xcjc=0.8, fc=0.85, tf=60e-12,
               xtf=48., itf=3e-2, vtf=2., tr=10e-9, eg=1.16, xti=3.,
               xtb=1.6),
          dict(type='pnp', bf=49., xtb=1.5, cjs=1e-13, iss=2e-14,
               cje=54e-15)]


def bjt_pair(dev, dT, nterm, **kw):
    nodes = [1, 2, 3, 4][:nterm]
    dt = S.make_element(dev + '_t', nodes + [9, 0], **kw)
    d = S.make_element(dev, nodes, temp=glVar.temp + dT, **kw)
    return dt, d


def bjt_xin(elem, n, rng, sv):
    nx = len(elem.controlPorts)
    x = rng.uniform(-1., 1., size=(nx, n))
    if sv:
        x[0] = rng.uniform(-50., 400., n)
        x[1] = rng.uniform(-50., 50., n)
    else:
        x[0] = rng.uniform(-1., 0.85, n)
        x[1] = rng.uniform(-5., 0.85, n)
    return x


def test_bjt_thermal_equals_isothermal_at_port_temperature():
    rng = np.random.default_rng(5)
    for dev in ('bjt', 'svbjt'):
        for kw in BJT_KW:
            for nterm in (3, 4):
                for dT in (0., 41.):
                    dt, d = bjt_pair(dev, dT, nterm, **kw)
                    assert len(dt.controlPorts) == len(d.controlPorts) + 1
                    x = bjt_xin(d, 256, rng, dev == 'svbjt')
                    xt = np.vstack([x, np.full((1, x.shape[1]), dT)])
                    ot, jt = device_eval(dt, xt, True)
                    oi, ji = device_eval(d, x, True)
                    ncs = len(d.csOutPorts)
                    assert ot.shape[0] == oi.shape[0] + 1
                    assert np.allclose(ot[:ncs], oi[:ncs], rtol=1e-11,
                                       atol=1e-300)
                    assert np.allclose(ot[ncs + 1:], oi[ncs:], rtol=1e-11,
                                       atol=1e-300)
                    # power output = the device's own power() of the currents
                    pw = np.array([d.power(x[:, k], oi[:ncs, k])
                                   for k in range(x.shape[1])])
                    assert np.allclose(ot[ncs], pw, rtol=1e-10, atol=1e-300)
                    assert np.allclose(jt[:ncs, :-1], ji[:ncs], rtol=1e-10,
                                       atol=1e-300)


def test_bjt_temperature_derivative_matches_finite_difference():
    rng = np.random.default_rng(6)
    for dev in ('bjt', 'svbjt'):
        dt, d = bjt_pair(dev, 0., 4, **BJT_KW[1])
        x = bjt_xin(d, 64, rng, dev == 'svbjt')
        if dev == 'svbjt':
            x[0] = rng.uniform(-5., 40., 64)
            x[1] = rng.uniform(-50., 20., 64)
        T0, h = 30., 1e-3
        xt = np.vstack([x, np.full((1, 64), T0)])
        e = np.zeros_like(xt)
        e[-1] = h
        o, j = device_eval(dt, xt, True)
        op, _ = device_eval(dt, xt + e, False)
        om, _ = device_eval(dt, xt - e, False)
        fd = (op - om) / (2. * h)
        scale = np.abs(j[:, -1]).max(axis=1, keepdims=True)
        assert np.all(np.abs(fd - j[:, -1])
                      <= 2e-6 * np.abs(j[:, -1]) + 1e-8 * scale)


def test_npn_thermal_dc_sweep_kat():
    """doc/examples/examples.rst:122-165: electro-thermal svbjt_t DC sweep --
    system dimension 11, "Average iterations: 5" (dc.py:208 divides two ints
    under Python 2: 295 // 50), "Average residual: 0.0114478106429".  The
    oracle reproduces the residual to 8 digits (release 0.4.1 of the docs vs.
    the current sources), which pins svbjt_t: SVJunction.set_temp_vars in
    dual arithmetic, the power output and the thermal port stamps."""
    import json
    import os
    import tempfile
    with open(os.path.join(S.HERE, 'golden', 'reference_kat.json')) as f:
        k = json.load(f)['npn_thermal_dc']
    deck = open(S.net('npn_thermal.net')).read()
    assert '.options temp=25' in deck
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'npn_thermal_doc.net')
        open(path, 'w').write(deck.replace('.options temp=25', ''))
        ckt, queue = S.load_circuit(path)
        an = [a for a in queue if a.anType == 'dc'][0]
        r = R.run_dc(ckt, glVar, an.device, an.param, an.start, an.stop,
                     an.num)
    assert an.num == k['points']
    assert r['idx'].dimension == k['dimension']
    assert int(r['iters'].sum()) // an.num \
        == k['average_iterations_py2_integer_division']
    assert abs(r['res'].mean() - k['average_residual']) \
        < 1e-7 * k['average_residual']
    rows = r['idx'].name_to_row()
    X = r['X']
    # self-heating: T rise = 100 K/W * dissipated power, grows with the bias
    t1 = X[:, rows['t1']]
    assert t1[-1] > t1[0] and 0.5 < t1[-1] < 20.
    ic = (X[:, rows['10']] - X[:, rows['1']]) / 1e3
    assert np.all(np.diff(ic) > -1e-9) and 5e-3 < ic[-1] < 12e-3


# ---- Curtice MESFET: mesfetc_t ----------------------------------------------
MES_KW = dict(a0=.016542, a1=.0500214, a2=.02012, a3=-.00806592,
              gama=2.16505, tau=5e-12, beta=-.0394707, isat=1e-9, vbd=15.,
              nr=10., ib0=1e-9, cgs0=.52785e-12, cgd0=.087e-12, avt0=1e-3,
              bvt0=1e-6, tbet=0.3, tm=2e-3, tme=1.5)


def test_mesfetc_thermal_equals_isothermal_at_port_temperature():
    rng = np.random.default_rng(8)
    for dT in (0., 37.):
        dt = S.make_element('mesfetc_t', [1, 2, 3, 9, 0], **MES_KW)
        d = S.make_element('mesfetc', [1, 2, 3], temp=glVar.temp + dT,
                           **MES_KW)
        assert dt.nDelays == 2 and len(dt.controlPorts) == 3
        x = rng.uniform(-1.2, 0.6, size=(4, 256))
        x[1] = rng.uniform(-4., 0.6, 256)
        x[3] = x[1] + rng.uniform(-.1, .1, 256)
        # thermal port sits before the delayed copies (autoThermal.py:114)
        xt = np.vstack([x[:2], np.full((1, 256), dT), x[2:]])
        ot, jt = device_eval(dt, xt, True)
        oi, ji = device_eval(d, x, True)
        assert np.allclose(ot[:3], oi[:3], rtol=1e-11, atol=1e-300)
        assert np.allclose(ot[4:], oi[3:], rtol=1e-11, atol=1e-300)
        pw = np.array([d.power(x[:, k], oi[:3, k]) for k in range(256)])
        assert np.allclose(ot[3], pw, rtol=1e-10, atol=1e-300)
    # d/dT against finite differences
    h = 1e-3
    e = np.zeros_like(xt)
    e[2] = h
    o, j = device_eval(dt, xt, True)
    fd = (device_eval(dt, xt + e, False)[0]
          - device_eval(dt, xt - e, False)[0]) / (2. * h)
    scale = np.abs(j[:, 2]).max(axis=1, keepdims=True)
    assert np.all(np.abs(fd - j[:, 2]) <= 2e-6 * np.abs(j[:, 2])
                  + 1e-8 * scale)


# ---- svdiode, svdiode_t, res_t -----------------------------------------------
def test_svdiode_behaves_like_diode():
    """svdiode.py docstring: 'Externally it behaves exactly like the regular
    diode device' -- same current, voltage and charge as `diode` once the
    state variable is mapped to the junction voltage."""
    kw = dict(isat=2e-14, n=1.2, cj0=3e-12, tt=1e-9, m=0.4, area=2.)
    sv = S.make_element('svdiode', [1, 0], **kw)
    d = S.make_element('diode', [1, 0], **kw)
    x = np.linspace(-3., 60., 300)[None, :]
    osv, jsv = device_eval(sv, x, True)
    vd = osv[1] / 1e-3                                  # gyr = 1e-3
    od, jd = device_eval(d, vd[None, :], True)
    assert np.allclose(osv[0], od[0], rtol=1e-9, atol=1e-22)
    assert np.allclose(osv[2], od[1], rtol=1e-9, atol=1e-30)
    # chain rule: di/dvd = (di/dx) / (dvd/dx)
    assert np.allclose(jsv[0, 0] / (jsv[1, 0] / 1e-3), jd[0, 0], rtol=1e-8)


def test_svdiode_and_resistor_thermal_equal_isothermal():
    rng = np.random.default_rng(9)
    x = rng.uniform(-5., 45., (1, 256))
    for dT in (0., 55.):
        kw = dict(cj0=2e-12, tt=1e-9, bv=7., rs=4.)
        dt = S.make_element('svdiode_t', [1, 0, 9, 0], **kw)
        d = S.make_element('svdiode', [1, 0], temp=glVar.temp + dT, **kw)
        xt = np.vstack([x, np.full((1, 256), dT)])
        ot, jt = device_eval(dt, xt, True)
        oi, ji = device_eval(d, x, True)
        assert np.allclose(ot[:2], oi[:2], rtol=1e-11, atol=1e-300)
        assert np.allclose(ot[3], oi[2], rtol=1e-11, atol=1e-300)
        assert np.allclose(ot[2], oi[0] * oi[1] / 1e-3, rtol=1e-11)
        # resistor: i = v / (r (1 + (tc1 + tc2 dT') dT')), dT' = temp - tnom
        rt = S.make_element('res_t', [1, 0, 9, 0], r=330., tc1=4e-3,
                            tc2=1e-5, tnom=20.)
        assert rt.isNonlinear and len(rt.controlPorts) == 2
        v = rng.uniform(-5., 5., (1, 256))
        o, j = device_eval(rt, np.vstack([v, np.full((1, 256), dT)]), True)
        d_ = glVar.temp + dT - 20.
        g = 1. / 330. / (1. + (4e-3 + 1e-5 * d_) * d_)
        assert np.allclose(o[0], g * v[0], rtol=1e-13)
        assert np.allclose(o[1], g * v[0] * v[0], rtol=1e-13)
        dg = -g * (4e-3 + 2e-5 * d_) / (1. + (4e-3 + 1e-5 * d_) * d_)
        assert np.allclose(j[0, 1], dg * v[0], rtol=1e-10)


# ---- EKV: mosekv_t -----------------------------------------------------------
EKV_KW = dict(w=30e-6, l=1e-6, ad=5e-11, asrc=5e-11, pd=4e-5, ps=4e-5,
              cj=3e-4, cjsw=2e-10, js=1e-3, tcv=1.5e-3, bex=-1.4, ucex=0.9,
              ibbt=8e-4, iba=2e8, ibb=4e8)


def test_mosekv_thermal_equals_isothermal_at_port_temperature():
    rng = np.random.default_rng(12)
    for typ in ('n', 'p'):
        sgn = 1. if typ == 'n' else -1.
        for dT in (0., 45.):
            dt = S.make_element('mosekv_t', [1, 2, 3, 4, 9, 0], type=typ,
                                **EKV_KW)
            d = S.make_element('mosekv', [1, 2, 3, 4], type=typ,
                               temp=glVar.temp + dT, **EKV_KW)
            assert len(dt.controlPorts) == 4 and len(dt.csOutPorts) == 4
            x = sgn * rng.uniform(-0.3, 3.3, size=(3, 256))
            xt = np.vstack([x, np.full((1, 256), dT)])
            ot, jt = device_eval(dt, xt, True)
            oi, ji = device_eval(d, x, True)
            assert np.allclose(ot[:3], oi[:3], rtol=1e-10, atol=1e-300)
            assert np.allclose(ot[4:], oi[3:], rtol=1e-10, atol=1e-300)
            pw = np.array([d.power(x[:, k], oi[:3, k]) for k in range(256)])
            assert np.allclose(ot[3], pw, rtol=1e-10, atol=1e-300)
            assert np.allclose(jt[:3, :3], ji[:3], rtol=1e-9, atol=1e-300)
    h = 1e-3
    e = np.zeros_like(xt)
    e[3] = h
    o, j = device_eval(dt, xt, True)
    fd = (device_eval(dt, xt + e, False)[0]
          - device_eval(dt, xt - e, False)[0]) / (2. * h)
    scale = np.abs(j[:, 3]).max(axis=1, keepdims=True)
    assert np.all(np.abs(fd - j[:, 3]) <= 5e-6 * np.abs(j[:, 3])
                  + 1e-8 * scale)
