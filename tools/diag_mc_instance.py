#!/usr/bin/env python
"""Diagnostic (GPU box): plain-Newton operating-point iterates of Monte-Carlo
instance j on the device (one iteration per launch) next to the oracle's, and
the device outputs / Jacobians at the first iterate where they part.

  python tools/diag_mc_instance.py 27 [46 ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(j):
    from tests import support as S
    from tests.golden import make_mc_golden as G
    from cardoon_b200.analyses import mc, nodalCUDA as nd
    from cardoon_b200.globalVars import glVar
    from oracle import nodal_ref as R
    from oracle.models import device_eval
    np.set_printoptions(linewidth=220, precision=17)
    ckt, an = G._instance(j)
    idx = R.prepare(ckt)
    opt = R.Options(glVar)
    rdc = R.RefDC(ckt, idx, opt)
    sV = R.RefTran(ckt, idx, opt, R.Trapezoidal()).get_source(0.).copy()
    # device: same instance through the batch packer
    ckt2, queue = S.load_circuit(S.net('ring_bsim3.net'))
    bt = mc.BatchTransient(ckt2, an.tstop, an.tstep, an.im)
    z = mc.bsim3_ring_samples([j])
    params = bt.pack_batch(mc.bsim3_ring_overrides(z, ckt2.nD_nlinElem))
    dc, tr = bt.engines(1)
    dc.set_params(params)
    x_o = rdc.get_guess()
    x_d = x_o.copy()
    old = glVar.maxiter
    glVar.maxiter = 1
    try:
        for it in range(12):
            # oracle step
            try:
                iVec, Jac = rdc.get_i_Jac(x_o)
                dx = rdc.factor_and_solve(sV - iVec, Jac)
                m = np.abs(dx).max()
                if m > opt.maxdelta:
                    dx = dx * (opt.maxdelta / m)
                xo_new = x_o + dx
                fin_o = True
            except R.NonFiniteSystem:
                xo_new, fin_o = x_o, False
            xd_new, res, its, ok = dc.newton(x_d, sV)
            xd_new = xd_new[0]
            print('it', it, 'oracle finite', fin_o, 'device res', res[0])
            print('  x_o', xo_new)
            print('  x_d', xd_new)
            d = np.abs(xo_new - xd_new).max()
            print('  max diff', d)
            if d > 1e-6 or not fin_o or not np.isfinite(res[0]):
                print('  --- device outputs at the previous iterates')
                for e in idx.nlin:
                    inf = idx.info[id(e)]
                    for nm, xx in (('o', x_o), ('d', x_d)):
                        xin = R._set_xin(inf, xx)
                        out, jac = device_eval(e, xin[:, None], True)
                        e2 = [q for q in ckt2.nD_nlinElem
                              if q.instanceName == e.instanceName][0]
                        k = ckt2.nD_nlinElem.index(e2)
                        g = bt.topo.groups[0]
                        col = params[0][:, g['elems'].index(e2)]
                        og, jg = dc.lib.dev_eval_host(
                            g['model'], g['flags'], col[:, None],
                            xin[:, None], 6, True)
                        print('  ', e.instanceName, nm, 'xin', xin)
                        print('     oracle out', out[:, 0])
                        print('     device out', og[:, 0])
                        print('     oracle jac', jac[:, :, 0].ravel())
                        print('     device jac', jg[:, 0])
                break
            x_o, x_d = xo_new, xd_new
    finally:
        glVar.maxiter = old


if __name__ == '__main__':
    for a in sys.argv[1:] or ['27']:
        print('=========== instance', a)
        main(int(a))
