#!/usr/bin/env python
"""
Golden vectors for cardoon_b200.analyses.sensitivity: derivatives of device
currents with respect to device parameters by forward-mode AD through the
REFERENCE's own ``process_params`` + ``eval_cqs`` (loaded verbatim from
/root/reference by oracle/build_ref.py) -- what the reference's
``create_sensitivity_tape`` records with CppAD (analyses/sensitivity.py:99-172).

  python tests/golden/make_sens_golden.py     -> tests/golden/sens.npz
  <case>/xin[nxin]  <case>/names  <case>/di[ncs, npar]
Linear devices: <case>/dg[nvccs, npar] (derivative of the VCCS conductances).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, 'sens.npz')

# case of tests/ref_cases.py -> parameters; the bias is the column of the
# case's bias set with the largest first output current (device conducting)
CASES = {
    'diode_cj_tt_bv_rs': ['isat', 'n', 'rs', 'bv', 'area', 'ibv'],
    'bsim3_ring_n': ['vfb', 'u0', 'tox', 'l', 'w', 'nch', 'vsat'],
    'bsim3_chain_p': ['vth0', 'u0', 'tox', 'l', 'w', 'k1', 'pclm'],
    'bjt_oscillator': ['bf', 'isat', 'vaf', 'ikf', 'br', 'ise'],
    'ekv_ring_n': ['vt0', 'kp', 'gamma', 'w', 'l', 'phi'],
    'mesfetc': ['a0', 'a1', 'beta', 'vt0', 'isat', 'gama'],
    'svbjt_bias_npn': ['bf', 'isat', 'vaf', 'ikf'],
}


def bias_column(case):
    g = np.load(os.path.join(HERE, 'ref_devices.npz'))
    out = g[case + '/out']
    ok = np.all(np.isfinite(out), axis=0)
    return int(np.argmax(np.where(ok, np.abs(out[0]), -1.)))


def ref_sens(elem, xin_col, names):
    from oracle import build_ref as B, dual as D
    r = B.twin(elem)
    vals = [float(getattr(r, n)) for n in names]
    P = len(names)
    for k, (n, v) in enumerate(zip(names, vals)):
        d = np.zeros((P, 1))
        d[k] = 1.
        setattr(r, n, D.Dual(np.array([v]), d))
    with np.errstate(all='ignore'):
        r.process_params()
        v = np.empty(len(xin_col), dtype=object)
        for k, x in enumerate(xin_col):
            v[k] = D.Dual(np.array([x]), np.zeros((P, 1)))
        iv, qv = r.eval_cqs(v)
    out = np.zeros((len(iv), P))
    for a, o in enumerate(iv):
        if isinstance(o, D.Dual):
            out[a] = o.d[:, 0]
    dg = np.zeros((len(r.linearVCCS), P))
    for a, vccs in enumerate(r.linearVCCS):
        if isinstance(vccs[2], D.Dual):
            dg[a] = vccs[2].d[:, 0]
    return out, dg


def main():
    from tests import ref_cases as RC
    data = {}
    ref = np.load(os.path.join(HERE, 'ref_devices.npz'))
    for case, names in CASES.items():
        elem, xin = RC.build(case)
        col = bias_column(case)
        di, dg = ref_sens(elem, xin[:, col], names)
        data[case + '/col'] = np.array(col)
        data[case + '/i'] = ref[case + '/out'][:di.shape[0], col]
        data[case + '/xin'] = xin[:, col]
        data[case + '/names'] = np.array(names)
        data[case + '/di'] = di
        data[case + '/dg'] = dg
        print(case, di.shape, np.abs(di).max())
    np.savez_compressed(OUT, **data)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
