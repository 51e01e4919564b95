#!/usr/bin/env python
"""
Generates tests/golden/mc_ring_bsim3.npz with the CPU oracle (test
infrastructure): the BASELINE config-5 Monte-Carlo workload (ring_bsim3 with
the per-instance BSIM3 variation of cardoon_b200.analyses.mc).

  python tests/golden/make_mc_golden.py scan      # operating points of all
                                                  # 4096 instances -> helper
                                                  # index and iterations
  python tests/golden/make_mc_golden.py waves     # full 400-sample transients
                                                  # of the chosen instances

The chosen instances are 0..47 plus every instance whose operating point
needs a convergence helper beyond plain Newton (up to a cap), so that the GPU
test covers the in-kernel homotopy.  Arrays written:
  op_helper[4096]   0 plain Newton, 1 gmin2, 2 source, 3 gmin stepping,
                    -1 none converged
  op_iters[4096]    iterations reported by the helper that succeeded
  op_x_all[4096,7]  the operating points
  inst[K]           instance numbers with full transients
  wave[K,400,3]     nodes 1, 3, 4; iters[K,400]; op_x[K,7]
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, 'mc_ring_bsim3.npz')
TERMS = ['1', '3', '4']
TOTAL = 4096


def _instance(j):
    from tests import support as S
    from cardoon_b200.analyses import mc
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = [a for a in queue if a.anType == 'tran'][0]
    elems = S.nonlinear_elements(ckt)
    mc.bsim3_ring_variation()(j, elems)
    for e in elems:
        e.process_params()
    return ckt, an


def op_of(j):
    """(helper index, iterations, x) of the operating point of instance j."""
    from cardoon_b200.globalVars import glVar
    from oracle import nodal_ref as R
    ckt, an = _instance(j)
    idx = R.prepare(ckt)
    opt = R.Options(glVar)
    dc = R.RefDC(ckt, idx, opt)
    tr = R.RefTran(ckt, idx, opt, R.Trapezoidal())
    x0 = dc.get_guess()
    sV = tr.get_source(0.).copy()
    for k, h in enumerate(dc.helpers):
        if h is None:
            return j, -1, 0, np.full(idx.dimension, np.nan)
        try:
            x, res, it = h(x0, sV)
            return j, k, it, x
        except R.NoConvergence:
            continue


def wave_of(j):
    from cardoon_b200.globalVars import glVar
    from oracle import nodal_ref as R
    ckt, an = _instance(j)
    r = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im)
    rows = [r['idx'].row(ckt.find_term(t)) for t in TERMS]
    return j, r['X'][:, rows], r['iters'], r['X'][0], r['op_iterations']


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'scan'
    old = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    t0 = time.time()
    with mp.get_context('spawn').Pool(os.cpu_count()) as pool:
        if what == 'scan':
            out = pool.map(op_of, range(TOTAL), chunksize=16)
            old['op_helper'] = np.array([o[1] for o in out], dtype=np.int8)
            old['op_iters'] = np.array([o[2] for o in out], dtype=np.int32)
            old['op_x_all'] = np.array([o[3] for o in out])
            print('helpers used:', np.bincount(old['op_helper'] + 1))
        else:
            hard = np.nonzero(old['op_helper'] != 0)[0]
            inst = sorted(set(range(48)) | set(hard[:32].tolist()))
            out = pool.map(wave_of, inst, chunksize=1)
            old['inst'] = np.array(inst, dtype=np.int32)
            old['wave'] = np.array([o[1] for o in out])
            old['iters'] = np.array([o[2] for o in out], dtype=np.int32)
            old['op_x'] = np.array([o[3] for o in out])
            old['tran_op_iters'] = np.array([o[4] for o in out],
                                            dtype=np.int32)
    np.savez_compressed(OUT, **old)
    print('wrote', OUT, 'in {0:.0f} s'.format(time.time() - t0))


if __name__ == '__main__':
    main()
