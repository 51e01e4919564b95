#!/usr/bin/env python
"""
Freezes outputs of the REFERENCE's own device code (loaded verbatim from
/root/reference by oracle/build_ref.py; only possible in the build container)
into tests/golden/ref_devices.npz:

  <case>/xin[nxin, n]      bias set (tests/ref_cases.py)
  <case>/out[nout, n]      currents then charges returned by the reference's
                           eval_cqs (what cppaddev.eval_and_deriv returns)
  <case>/jac[nout,nxin,n]  its Jacobian (forward-mode duals in place of CppAD)
  <case>/attr_names, <case>/attr_values
                           every numeric attribute of the reference element
                           after its own process_params / set_temp_vars

  python tests/golden/make_ref_device_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, 'ref_devices.npz')


def main():
    from oracle import build_ref as B
    from tests import ref_cases as RC
    data = {}
    for name in RC.CASES:
        elem, xin = RC.build(name)
        ref = B.twin(elem)
        attrs = B.numeric_attributes(ref)      # before eval: the _t wrappers
        out, jac = B.eval_cqs_dual(ref, xin)   # overwrite them with duals
        keys = sorted(attrs)
        data[name + '/xin'] = xin
        data[name + '/out'] = out
        data[name + '/jac'] = jac
        data[name + '/attr_names'] = np.array(keys)
        data[name + '/attr_values'] = np.array([attrs[k] for k in keys])
        print('{0:24s} {1} -> out {2} jac {3}, {4} attributes'.format(
            name, xin.shape, out.shape, jac.shape, len(keys)))
    np.savez_compressed(OUT, **data)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
