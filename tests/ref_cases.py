"""
Device cases shared by the reference-pinning tests: every nonlinear model of
the hot path on the BASELINE decks and on the bias sets of the GPU kernel
tests.  A case builds an initialised *product* element and a bias matrix
``xin[nxin, n]``; ``oracle.build_ref.twin`` turns the element into an
instance of the reference's own device class.
"""
import numpy as np

from tests import support as S

N = 96          # biases per case (kept small: the golden file is committed)


def _deck(deck, pick=0, init=True):
    def make():
        ckt, _ = S.load_circuit(S.net(deck), init=init)
        return S.nonlinear_elements(ckt)[pick]
    return make


def _elem(dev, nodes, **kw):
    def make():
        from cardoon_b200 import simulator as sim
        sim.reset_all()         # global options (temp ...) back to defaults
        return S.make_element(dev, nodes, **kw)
    return make


def _mc_instance(j, pick):
    def make():
        from cardoon_b200.analyses import mc
        ckt, _ = S.load_circuit(S.net('ring_bsim3.net'))
        elems = S.nonlinear_elements(ckt)
        mc.bsim3_ring_variation()(j, elems)
        for e in elems:
            e.process_params()
        return elems[pick]
    return make


def _u(seed, *ranges, n=N):
    rng = np.random.default_rng(seed)
    return np.vstack([rng.uniform(lo, hi, n) for lo, hi in ranges])


def _mos(seed, lo=-0.5, hi=3.5):
    def xin(elem):
        return elem._tf * _u(seed, (lo, hi), (lo, hi), (lo, hi))
    return xin


MES = dict(a0=0.09910, a1=0.08541, a2=-0.02030, a3=-0.01543, beta=0.01865,
           gama=5., vt0=-2., isat=1e-13, cgs0=.5e-12, cgd0=.1e-12,
           tau=5e-12, ib0=1e-10, vbd=12.)
T = (-40., 120.)      # temperature-rise port of the _t devices

# name -> (element factory, xin(elem))
CASES = {
    # ---- diode / svdiode -------------------------------------------------
    'diode_soliton': (_deck('soliton.net'),
                      lambda e: _u(1, (-18., 1.5))),
    'diode_default': (_elem('diode', [1, 0]), lambda e: _u(2, (-5., .9))),
    'diode_cj_tt_bv_rs': (_elem('diode', [1, 0], cj0=2e-12, tt=1e-9, bv=4.,
                                rs=10.), lambda e: _u(2, (-5., .9))),
    'diode_area': (_elem('diode', [1, 0], isat=1e-15, n=1.5, area=3.,
                         cj0=1e-12, m=.33, fc=.6),
                   lambda e: _u(2, (-5., .9))),
    'svdiode': (_elem('svdiode', [1, 0], cj0=1e-12, tt=1e-10, rs=5.),
                lambda e: _u(3, (-30., 60.))),
    # ---- Gummel-Poon -----------------------------------------------------
    'svbjt_bias_npn': (_deck('bias_npn.net'),
                       lambda e: _u(4, (-50., 400.), (-50., 50.), (-1., 1.),
                                    (-5., .7))),
    'bjt_oscillator': (_deck('oscillator.net'),
                       lambda e: _u(5, (-1., .9), (-5., .9), (-1., 1.))),
    'bjt_default': (_elem('bjt', [1, 2, 3]),
                    lambda e: _u(6, *[(-1., .8)] * S.nxin_of(e))),
    'bjt_pnp': (_elem('bjt', [1, 2, 3], type='pnp', vaf=50., ikf=1e-3,
                      cje=1e-12, tf=1e-11),
                lambda e: _u(6, *[(-1., .8)] * S.nxin_of(e))),
    'bjt_rb': (_elem('bjt', [1, 2, 3], rb=100., cjc=1e-12, xcjc=.5, tr=1e-9),
               lambda e: _u(6, *[(-1., .8)] * S.nxin_of(e))),
    # ---- BSIM3 (configs 3, 4, 5) -----------------------------------------
    'bsim3_ring_n': (_deck('ring_bsim3.net', 0), _mos(7)),
    'bsim3_ring_p': (_deck('ring_bsim3.net', 1), _mos(8)),
    'bsim3_chain_n': (_deck('inverter_chain_bsim.net', 0), _mos(9)),
    'bsim3_chain_p': (_deck('inverter_chain_bsim.net', 1), _mos(10)),
    'bsim3_mc31_n': (_mc_instance(31, 0), _mos(11, -9.5, 11.)),
    'bsim3_mc31_p': (_mc_instance(31, 1), _mos(12, -9.5, 11.)),
    'bsim3_mc7_p': (_mc_instance(7, 3), _mos(13)),
    # ---- EKV ---------------------------------------------------------------
    'ekv_ring_n': (_deck('ring_osc_ahkab.net', 0), _mos(14, -.5, 3.)),
    'ekv_ring_p': (_deck('ring_osc_ahkab.net', 1), _mos(15, -.5, 3.)),
    # ---- MESFET, memristor -------------------------------------------------
    'mesfetc': (_elem('mesfetc', [1, 2, 3], **MES),
                lambda e: _u(16, *[(-3., .6)] * 4)),
    'mesfetc_deck': (_deck('mesfetc_plain.net'),
                     lambda e: _u(17, *[(-3., .6)] * S.nxin_of(e))),
    'memristor': (_deck('memristor.net'),
                  lambda e: _u(18, (-5., 5.), (-2., 2.))),
    # ---- electro-thermal (autoThermal.py:97-124) ---------------------------
    'diode_t': (_elem('diode_t', [1, 0, 2, 0], cj0=2e-12, tt=1e-9, bv=4.,
                      rs=10.), lambda e: _u(19, (-5., .9), T)),
    'svdiode_t': (_elem('svdiode_t', [1, 0, 2, 0], cj0=1e-12, rs=5.),
                  lambda e: _u(20, (-30., 60.), T)),
    'svbjt_t_npn_thermal': (_deck('npn_thermal.net'),
                            lambda e: _u(21, *([(-20., 45.), (-50., 30.)]
                                               + [(-1., .8)]
                                               * (S.nxin_of(e) - 3) + [T]))),
    'bjt_t': (_elem('bjt_t', [1, 2, 3, 4, 0], vaf=50., ikf=1e-3, cje=1e-12,
                    tf=1e-11, cjc=1e-12, tr=1e-9),
              lambda e: _u(22, *([(-1., .8)] * (S.nxin_of(e) - 1) + [T]))),
    'mesfetc_t': (_elem('mesfetc_t', [1, 2, 3, 4, 0], **MES),
                  lambda e: _u(23, *([(-3., .6)] * (S.nxin_of(e) - 1)
                                     + [T]))),
    'mosekv_t': (_elem('mosekv_t', [1, 2, 3, 4, 5, 0], w=30e-6, l=1e-6,
                       vt0=.5, gamma=.7, phi=.5, kp=8e-5),
                 lambda e: _u(24, (-.5, 3.), (-.5, 3.), (-.5, 3.), T)),
    'res_t': (_elem('res_t', [1, 2, 3, 0], r=50., tc1=1e-3, tc2=1e-5),
              lambda e: _u(25, (-5., 5.), T)),
}


def build(name):
    make, xin = CASES[name]
    elem = make()
    return elem, np.ascontiguousarray(xin(elem))
