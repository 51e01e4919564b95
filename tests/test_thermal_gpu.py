"""
GPU parity of the electro-thermal diode (`diode_t`, reference
devices/autoThermal.py + diode.py) against the oracle: the evaluation kernel
through cdn_dev_eval, and operating point / DC sweep / transient through the
device-resident Newton engine.
"""
import numpy as np
import pytest

from tests import support as S
from tests.test_dev_eval_gpu import check
from tests.test_newton_gpu import tran_both, check_tran, check_op
from cardoon_b200.analyses import nodalCUDA as nd
from cardoon_b200.analyses.fsolve import solve
from cardoon_b200.globalVars import glVar
from oracle import nodal_ref as R

pytestmark = pytest.mark.gpu


def test_diode_t_kernel(lib):
    rng = np.random.default_rng(11)
    n = 2048
    xin = np.vstack([rng.uniform(-5., 0.9, n), rng.uniform(-60., 150., n)])
    for kw in (dict(), dict(cj0=1e-12), dict(tt=1e-9),
               dict(cj0=2e-12, tt=1e-9, bv=4., rs=10.),
               dict(isat=1e-15, n=1.5, area=3., cj0=1e-12, m=0.33, fc=0.6)):
        elem = S.make_element('diode_t', [1, 0, 2, 0], **kw)
        check(lib, elem, xin)


def test_diode_t_plugin_batch_of_one(lib):
    """Device-plugin interface (eval_and_deriv / eval_cqs / power) of a `_t`
    device: batch of one through the same kernel."""
    from oracle.models import device_eval
    elem = S.make_element('diode_t', [1, 0, 2, 0], cj0=1e-12, rs=3.)
    v = np.array([0.71, 35.])
    out, jac = elem.eval_and_deriv(v)
    ro, rj = device_eval(elem, v[:, None], True)
    assert np.allclose(out, ro[:, 0], rtol=1e-10)
    assert np.allclose(jac, rj[:, :, 0], rtol=1e-8)
    iV, qV = elem.eval_cqs(v)
    assert len(iV) == 2 and len(qV) == 1
    assert elem.power(v[:1], iV) == iV[1]
    assert abs(iV[1] - v[0] * iV[0]) < 1e-12 * abs(iV[1])


def test_op_and_dc_sweep_diode_thermal(lib):
    ckt, queue = S.load_circuit(S.net('diode_thermal.net'))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    x, res, it = solve(dc.get_guess(), dc.get_source(), dc.convergence_helpers)
    ref = R.run_op(ckt, glVar)
    assert abs(it - ref['iterations']) <= 1
    check_op(x, ref['x'])
    # `.analysis dc device=res:rth param=r` of the deck (dc.py:89-242)
    ckt, queue = S.load_circuit(S.net('diode_thermal.net'))
    an = [a for a in queue if a.anType == 'dc'][0]
    ckt.saveReqList = []
    ckt.plotReqList = []
    an.run(ckt)
    got = np.array([t.dC_v for t in ckt.nD_termList]).T
    ckt2, _ = S.load_circuit(S.net('diode_thermal.net'))
    r2 = R.run_dc(ckt2, glVar, an.device, an.param, an.start, an.stop, an.num)
    err = np.abs(got - r2['X'])
    assert np.all(err < 1e-9 * np.abs(r2['X']) + 1e-12), err.max()
    assert np.max(np.abs(an.iterations - r2['iters'])) <= 1


def test_tran_self_heating(lib):
    wave, iters, status, ref, tran = tran_both('diode_thermal_tran.net')
    check_tran(wave, iters, status, ref)
    rows = ref['idx'].name_to_row()
    # the junction really heats up (tens of degrees) and cools down again
    dT = wave[:, rows['1000']]
    assert dT.max() > 5. and dT[-1] < dT.max()


# ---- Gummel-Poon: bjt_t / svbjt_t -------------------------------------------
from tests.test_thermal_cpu import BJT_KW, bjt_pair, bjt_xin   # noqa: E402


@pytest.mark.parametrize('dev', ['bjt', 'svbjt'])
def test_bjt_t_kernel(lib, dev):
    rng = np.random.default_rng(21)
    for kw in BJT_KW:
        for nterm in (3, 4):
            dt, d = bjt_pair(dev, 0., nterm, **kw)
            x = bjt_xin(d, 1024, rng, dev == 'svbjt')
            if dev == 'svbjt':
                x[0] = rng.uniform(-20., 45., 1024)
                x[1] = rng.uniform(-50., 30., 1024)
            xt = np.vstack([x, rng.uniform(-40., 120., (1, 1024))])
            if kw.get('rb'):
                xt[2, :8] = 0.          # ib = 0: |x| derivative convention
            check(lib, dt, xt, jtol=1e-7)


def test_npn_thermal_dc_sweep(lib):
    """`.analysis dc` of test/npn_thermal.net on the device engine against
    the oracle (which reproduces the documented run, test_thermal_cpu)."""
    ckt, queue = S.load_circuit(S.net('npn_thermal.net'))
    an = [a for a in queue if a.anType == 'dc'][0]
    ckt.saveReqList = []
    ckt.plotReqList = []
    an.run(ckt)
    assert ckt.nD_dimension == 11
    got = np.array([t.dC_v for t in ckt.nD_termList]).T
    ckt2, _ = S.load_circuit(S.net('npn_thermal.net'))
    ref = R.run_dc(ckt2, glVar, an.device, an.param, an.start, an.stop,
                   an.num)
    err = np.abs(got - ref['X'])
    assert np.all(err < 1e-9 * np.abs(ref['X']) + 1e-12), err.max()
    assert np.max(np.abs(an.iterations - ref['iters'])) <= 1


# ---- Curtice MESFET: mesfetc_t ----------------------------------------------
from tests.test_thermal_cpu import MES_KW                      # noqa: E402


def test_mesfetc_t_kernel(lib):
    rng = np.random.default_rng(31)
    n = 1024
    dt = S.make_element('mesfetc_t', [1, 2, 3, 9, 0], **MES_KW)
    x = rng.uniform(-1.2, 0.6, size=(5, n))
    x[1] = rng.uniform(-4., 0.6, n)
    x[2] = rng.uniform(-30., 90., n)
    x[4] = x[1] + rng.uniform(-.1, .1, n)
    check(lib, dt, x, jtol=1e-7)


def test_tran_mesfetc_thermal(lib):
    """test/mesfetc_thermal.net: delayed control ports next to the thermal
    port, thermal RC sub-circuit (first 150 samples)."""
    wave, iters, status, ref, tran = tran_both('mesfetc_thermal.net',
                                               max_samples=150)
    check_tran(wave, iters, status, ref)


# ---- svdiode, svdiode_t, res_t ------------------------------------------------
def test_svdiode_and_res_t_kernels(lib):
    rng = np.random.default_rng(41)
    n = 1024
    x = rng.uniform(-8., 60., (1, n))
    for kw in (dict(), dict(cj0=2e-12, tt=1e-9, bv=7., rs=4.),
               dict(isat=1e-15, n=1.5, area=3., cj0=1e-12, m=0.33, fc=0.6)):
        check(lib, S.make_element('svdiode', [1, 0], **kw), x)
        xt = np.vstack([x, rng.uniform(-40., 120., (1, n))])
        check(lib, S.make_element('svdiode_t', [1, 0, 9, 0], **kw), xt)
    rt = S.make_element('res_t', [1, 0, 9, 0], r=330., tc1=4e-3, tc2=1e-5,
                        tnom=20.)
    check(lib, rt, np.vstack([rng.uniform(-5., 5., (1, n)),
                              rng.uniform(-40., 120., (1, n))]))


def test_op_thermal_mix(lib):
    ckt, queue = S.load_circuit(S.net('thermal_mix.net'))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    x, res, it = solve(dc.get_guess(), dc.get_source(), dc.convergence_helpers)
    ref = R.run_op(ckt, glVar)
    assert abs(it - ref['iterations']) <= 1
    check_op(x, ref['x'])
    rows = ref['idx'].name_to_row()
    assert 5. < x[rows['th']] < 80.          # tens of degrees of self-heating


# ---- EKV: mosekv_t -----------------------------------------------------------
from tests.test_thermal_cpu import EKV_KW                      # noqa: E402

EKV_T_DECK = """# self-heating EKV inverter stage
.options temp=30
vdc:vdd 1 gnd vdc=3.3
vdc:vg 2 gnd vdc=1.4
res:rd 1 3 r=2k
mosekv_t:m1 3 2 gnd gnd th gnd w=30e-6 l=1e-6 ad=5e-11 asrc=5e-11 cj=3e-4 js=1e-3 tcv=1.5e-3 bex=-1.4
res:rth th gnd r=2k
.end
"""


def test_mosekv_t_kernel_and_op(lib):
    from tests.test_ac_cpu import load_text
    rng = np.random.default_rng(51)
    n = 1024
    for typ in ('n', 'p'):
        sgn = 1. if typ == 'n' else -1.
        dt = S.make_element('mosekv_t', [1, 2, 3, 4, 9, 0], type=typ,
                            **EKV_KW)
        x = np.vstack([sgn * rng.uniform(-0.3, 3.3, size=(3, n)),
                       rng.uniform(-40., 120., (1, n))])
        check(lib, dt, x, jtol=1e-7)
    ckt, _ = load_text(EKV_T_DECK)
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    x, res, it = solve(dc.get_guess(), dc.get_source(), dc.convergence_helpers)
    ref = R.run_op(ckt, glVar)
    assert abs(it - ref['iterations']) <= 1
    check_op(x, ref['x'])
    assert x[ref['idx'].name_to_row()['th']] > 0.5
