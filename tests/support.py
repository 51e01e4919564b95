"""Helpers shared by the tests, smoke() and bench.py's checker legs."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NETS = os.path.join(HERE, 'netlists')


def net(name):
    """Path of a deck: tests/netlists/, else the reference's own regression
    decks under tests/netlists/reference_tests/ (verbatim copies, kept once)."""
    p = os.path.join(NETS, name)
    if not os.path.exists(p):
        q = os.path.join(NETS, 'reference_tests', name)
        if os.path.exists(q):
            return q
    return p


def load_circuit(path, init=True):
    """Parse a deck into a fresh main circuit; returns (ckt, analyses)."""
    from cardoon_b200 import simulator as sim
    from cardoon_b200 import circuit as cir
    sim.reset_all()
    ckt = cir.get_mainckt()
    queue = sim.parse_net(path, ckt)
    if init:
        ckt.flatten()
        ckt.init()
    return ckt, queue


def nonlinear_elements(ckt):
    return [e for e in ckt.elemDict.values() if e.isNonlinear]


def make_element(devType, nodes, model=None, **params):
    """One initialised element in a scratch circuit."""
    from cardoon_b200 import simulator as sim
    from cardoon_b200 import circuit as cir
    name = 'scratch{0}'.format(len(cir.Circuit.cktDict))
    ckt = cir.Circuit(name)
    elem = sim.new_elem(devType, 'e1', **params)
    ckt.add_elem(elem, model)
    ckt.connect(elem, [str(n) for n in nodes])
    elem.init()
    return elem


def nxin_of(elem):
    return len(elem.controlPorts) + elem.nDelays


def cuda_eval(lib, elem, xin, want_jac=True, full=True):
    """out[nout, n], jac[nout, nxin, n] of one element at xin[nxin, n]."""
    spec = elem.cuda_spec()
    nout = spec['ncs'] + spec['nqs']
    xin = np.ascontiguousarray(xin, dtype=float)
    out, jac = lib.dev_eval_host(spec['model'], spec['flags'],
                                 spec['params'][:, None].repeat(
                                     xin.shape[1] if full else 1, axis=1),
                                 xin, nout, want_jac)
    if jac is not None:
        jac = jac.reshape(nout, xin.shape[0], xin.shape[1])
    return out, jac


def rel_err(a, b, floor):
    """max |a-b| / (|b| + floor)"""
    a = np.asarray(a)
    b = np.asarray(b)
    with np.errstate(invalid='ignore'):
        e = np.abs(a - b) / (np.abs(b) + floor)
    e = np.where(np.isnan(a) & np.isnan(b), 0., e)
    return float(np.max(e)) if e.size else 0.


def smoke_check():
    """One small invocation of the hot path on cuda:0 against the oracle:
    BSIM3 device evaluation, the bias_npn operating point (the reference's
    printed known answer: 17 iterations) and 40 samples of the ring_bsim3
    transient, all through the C ABI."""
    from cardoon_b200._lib import get_lib
    from cardoon_b200.analyses import nodalCUDA as nd
    from cardoon_b200.analyses.fsolve import solve
    from cardoon_b200.analyses.integration import Trapezoidal
    from cardoon_b200.globalVars import glVar
    from oracle.models import device_eval
    from oracle import nodal_ref as R
    lib = get_lib(0)
    ckt, _ = load_circuit(net('ring_bsim3.net'))
    rng = np.random.default_rng(7)
    for elem in nonlinear_elements(ckt)[:2]:
        xin = rng.uniform(-0.5, 3.0, size=(3, 64))
        out, jac = cuda_eval(lib, elem, xin)
        ro, rj = device_eval(elem, xin, True)
        assert rel_err(out, ro, 1e-18) < 1e-9, 'BSIM3 value mismatch'
        assert rel_err(jac, rj, 1e-12) < 1e-7, 'BSIM3 Jacobian mismatch'
    # operating point of bias_npn.net
    ckt, _ = load_circuit(net('bias_npn.net'))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    x, res, it = solve(dc.get_guess(), dc.get_source(),
                       dc.convergence_helpers)
    ref = R.run_op(ckt, glVar)
    assert abs(it - 17) <= 1 and abs(it - ref['iterations']) <= 1
    assert np.all(np.abs(x - ref['x']) < 1e-9 * np.abs(ref['x']) + 1e-12)
    # device-resident transient
    ckt, queue = load_circuit(net('ring_bsim3.net'))
    an = queue[0]
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    tran = nd.TransientNodal(ckt, Trapezoidal())
    x, res, it = solve(dc.get_guess(), tran.get_source(0.).copy(),
                       dc.convergence_helpers)
    timeVec = np.arange(0., an.tstop, an.tstep)[:40]
    wave, iters, resv, status = tran.run(x, timeVec,
                                         list(range(ckt.nD_dimension)))
    ref = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im, max_samples=40)
    assert status == 0
    assert np.all(np.abs(wave - ref['X'])
                  < glVar.reltol * np.abs(ref['X']) + glVar.abstol)
    assert np.max(np.abs(iters - ref['iters'])) <= 1


def chain_netlist(K, path=None):
    """BASELINE config 4: K `inverter` sub-circuits of the reference's
    workspace/inverter_chain_bsim.net (:17-22) in a chain, same model cards,
    same vpulse:vin and vdc:vdd, cap:cload on the last node, tstep 0.1 ns,
    tstop 20 ns, `.options maxdelta=1 maxiter=30 sparse=1`.
    Returns the path of the generated deck."""
    import tempfile
    src = open(net('inverter_chain_bsim.net')).read()
    models = src[src.index('.model nch'):src.index('#.plot dc')]
    lines = ['# synthetic {0}-stage BSIM3 inverter chain'.format(K),
             '.options maxdelta=1. maxiter=30 sparse=1',
             '.analysis tran tstop=20n tstep=.1n im=trap verbose=0',
             models,
             'vdc:vdd 1 gnd vdc=3V',
             'vpulse:vin o0 gnd v1=0 v2=3 pw=2n per=10n tr=3n tf=3n']
    for k in range(1, K + 1):
        lines.append('x{0} o{1} o{0} 1 gnd inverter'.format(k, k - 1))
    lines += ['cap:cload o{0} gnd c=20fF'.format(K),
              '.subckt inverter in out 1 2',
              'mosbsim3:m1 1 in out 1 model=pch ad=8p asrc=8p cj=2e-4 '
              'cjsw=1n ps=8u m=3',
              'mosbsim3:m2 out in 2 2 model=nch ad=8p asrc=8p cj=2e-4 '
              'cjsw=1n ps=8u',
              '.ends', '.end', '']
    if path is None:
        fd, path = tempfile.mkstemp(suffix='.net', prefix='chain{0}_'.format(K))
        os.close(fd)
    with open(path, 'w') as f:
        f.write('\n'.join(lines))
    return path


def chain_nodeset(ckt, K, vdd=3.):
    """Initial guess for the operating point of the synthetic chain with the
    input low: inverter outputs alternate vdd / 0.  (From the all-zero guess
    the reference's Newton needs O(K) iterations to propagate through the
    chain and gives up at maxiter; its SuperLU path then meets exactly
    singular matrices during source stepping.)"""
    import numpy as np
    x0 = np.zeros(ckt.nD_dimension)
    x0[ckt.termDict['1'].nD_namRC] = vdd
    for k in range(1, K + 1):
        if k % 2:
            x0[ckt.termDict['o{0}'.format(k)].nD_namRC] = vdd
    return x0
