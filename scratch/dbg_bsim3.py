import sys; sys.path.insert(0,'.')
import numpy as np
from tests import support as S
from oracle.models import device_eval
from cardoon_b200._lib import get_lib
lib = get_lib(0)
for deck in ['ring_bsim3.net','inverter_chain_bsim.net']:
    ckt,_ = S.load_circuit(S.net(deck))
    rng = np.random.default_rng(6)
    for elem in S.nonlinear_elements(ckt)[:2]:
        xin = elem._tf*rng.uniform(-0.5,3.5,size=(3,8192))
        out,jac = S.cuda_eval(lib, elem, xin)
        ro,rj = device_eval(elem, xin, True)
        with np.errstate(all='ignore'):
            e = np.abs(jac-rj)/(np.abs(rj)+1e-16)
        e = np.where(np.isnan(jac)&np.isnan(rj),0,e)
        e = np.where(np.isnan(e), 1e99, e)
        idx = np.argsort(e.ravel())[::-1][:12]
        print(deck, elem.instanceName, 'n bad', (e>1e-7).sum())
        for k in idx:
            r,c,i = np.unravel_index(k, e.shape)
            print(r,c,i, 'x', xin[:,i]*elem._tf, 'gpu', jac[r,c,i], 'ref', rj[r,c,i], 'val', out[r,i], ro[r,i])
