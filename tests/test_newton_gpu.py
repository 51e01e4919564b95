"""
GPU parity of the device-resident Newton engine
against the oracle, through the reference-facing ``nd`` interface
(cardoon_b200.analyses.nodalCUDA) and therefore the C ABI
(cdn_plan_create / cdn_engine_run).

Bars (BASELINE.json north_star): operating-point unknowns within 1e-9
relative / 1e-12 absolute, Newton iteration counts within +-1, transient
waveforms within the reference's own reltol/abstol at every accepted sample.
"""
import os

import numpy as np
import pytest

from tests import support as S
from cardoon_b200.analyses import nodalCUDA as nd
from cardoon_b200.analyses.fsolve import solve
from cardoon_b200.analyses.integration import Trapezoidal, BEuler
from cardoon_b200.globalVars import glVar
from oracle import nodal_ref as R

pytestmark = pytest.mark.gpu


def run_op(ckt):
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    x0 = dc.get_guess()
    sV = dc.get_source()
    return solve(x0, sV, dc.convergence_helpers), dc


def check_op(x, ref_x):
    err = np.abs(x - ref_x)
    assert np.all(err < 1e-9 * np.abs(ref_x) + 1e-12), \
        'max err {0:g}'.format(err.max())


def test_op_bias_npn_kat(lib):
    """doc/usage.rst:48-79: 17 iterations, 12-digit node voltages."""
    ckt, _ = S.load_circuit(S.net('bias_npn.net'))
    (x, res, it), dc = run_op(ckt)
    ref = R.run_op(ckt, glVar)
    assert abs(it - ref['iterations']) <= 1 and abs(it - 17) <= 1
    check_op(x, ref['x'])
    dc.save_OP(x)
    kat = {'1': 3.49950994522, '10': 11.9158367321, '2': 0.800179840214,
           '3': 10.0}
    for name, v in kat.items():
        assert abs(ckt.termDict[name].nD_vOP - v) < 2e-11 * abs(v)
    assert abs(res - ref['res']) < 1e-6 * abs(ref['res']) + 1e-12


@pytest.mark.parametrize('deck', ['ring_bsim3.net', 'oscillator.net',
                                  'ring_osc_ahkab.net', 'memristor.net',
                                  'rctransient.net', 'soliton.net',
                                  'inverter_chain_bsim.net'])
def test_op_matches_oracle(lib, deck):
    ckt, queue = S.load_circuit(S.net(deck))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    tran = nd.TransientNodal(ckt, Trapezoidal())
    x0 = dc.get_guess()
    sV = tran.get_source(0.).copy()
    x, res, it = solve(x0, sV, dc.convergence_helpers)
    idx = R.prepare(ckt)
    rdc = R.RefDC(ckt, idx, R.Options(glVar))
    rtr = R.RefTran(ckt, idx, R.Options(glVar), R.Trapezoidal())
    xr, resr, itr = R.solve(rdc.get_guess(), rtr.get_source(0.), rdc.helpers)
    assert abs(it - itr) <= 1, (it, itr)
    check_op(x, xr)


def tran_both(deck, max_samples=None, tstep=None):
    ckt, queue = S.load_circuit(S.net(deck))
    an = [a for a in queue if a.anType == 'tran'][0]
    if tstep is not None:
        an.tstep = tstep
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    imo = Trapezoidal() if an.im == 'trap' else BEuler()
    tran = nd.TransientNodal(ckt, imo)
    x, res, it = solve(dc.get_guess(), tran.get_source(0.).copy(),
                       dc.convergence_helpers)
    timeVec = np.arange(start=0., stop=an.tstop, step=an.tstep, dtype=float)
    if max_samples:
        timeVec = timeVec[:max_samples]
    rows = list(range(ckt.nD_dimension))
    wave, iters, resv, status = tran.run(x, timeVec, rows)
    ref = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im,
                     max_samples=len(timeVec))
    return wave, iters, status, ref, tran


def check_tran(wave, iters, status, ref):
    assert status == 0
    X = ref['X']
    assert wave.shape == X.shape
    err = np.abs(wave - X)
    tol = glVar.reltol * np.abs(X) + glVar.abstol
    bad = err > tol
    assert not bad.any(), 'worst sample {0}, err {1:g}'.format(
        np.argwhere(bad)[0], (err / tol).max())
    assert np.max(np.abs(iters - ref['iters'])) <= 1, \
        (iters[:20], ref['iters'][:20])
    return float((err / tol).max())


def test_tran_ring_bsim3(lib):
    """BASELINE config 3: dense 7x7, 6 mosbsim3, 401 samples (tile team)."""
    wave, iters, status, ref, tran = tran_both('ring_bsim3.net')
    # 400, not 401: the reference parser turns `.01n` into 1.0000000000000001e-11
    # (netlistparser.py:113-116) and np.arange(0, 4e-9, that) stops one short
    assert wave.shape[0] == 400
    check_tran(wave, iters, status, ref)
    assert tran.engine.launches == 1      # the whole time loop


def test_tran_soliton(lib):
    """BASELINE config 2: N = 3022 sparse, 47 diodes, 500 samples
    (thread-block cluster, boundary/interior kernel of csrc/schur.cuh)."""
    wave, iters, status, ref, tran = tran_both('soliton.net')
    assert wave.shape == (501, 3022)
    check_tran(wave, iters, status, ref)
    assert tran.engine.launches == 1
    assert not tran.engine.dense
    assert tran.engine.last_path == 'schur'
    # ~1 Newton iteration per step thanks to the chord predictor
    # (reference doc/design.rst:243-256: 505 solves for 501 steps)
    assert iters.sum() < 1.2 * len(iters)


def test_tran_soliton_general_kernel(lib, monkeypatch):
    """Same run on the general sparse path (block team, level-scheduled
    static-pivot LU), which the boundary/interior kernel falls back to."""
    monkeypatch.setattr(nd, 'SCHUR_ENABLED', False)
    wave, iters, status, ref, tran = tran_both('soliton.net')
    check_tran(wave, iters, status, ref)
    assert tran.engine.launches == 1
    assert tran.engine.last_path == 'engine'
    # ~1 Newton iteration per step thanks to the chord predictor
    # (reference doc/design.rst:243-256: 505 solves for 501 steps)
    assert iters.sum() < 1.2 * len(iters)


def ladder_deck(nsec=150, every=25, im='trap', tstep='2p', clamp='diode'):
    """Synthetic deck for the boundary/interior kernel: an RC ladder with a
    series inductor now and then and diode clamps every ``every`` sections
    (the diodes have series resistance and junction capacitance, so each adds
    an internal node)."""
    import tempfile
    lines = ['# RC ladder with diode clamps (synthetic)',
             '.options sparse=1',
             '.analysis tran tstop=0.6n tstep={0} im={1} verbose=0'.format(
                 tstep, im),
             'ipulse:iin gnd n0 i1=0 i2=40mA pw=.2n per=1n tr=50p tf=50p '
             'td=20p',
             'res:rin n0 gnd r=50.']
    for k in range(nsec):
        if k % 10 == 5:
            lines.append('ind:l{0} n{0} n{1} l=20p'.format(k, k + 1))
        else:
            lines.append('res:r{0} n{0} n{1} r=2.'.format(k, k + 1))
        lines.append('cap:c{0} n{1} gnd c=20f'.format(k, k + 1))
        if k % 7 == 3:
            # capacitive coupling across a section (C[B, I] != 0 next to a clamp)
            lines.append('cap:cc{0} n{0} n{1} c=5f'.format(k, k + 1))
        if k % every == every - 1:
            kind = 'diode' if clamp == 'diode' else \
                ('bjt', 'mos', 'diode')[(k // every) % 3]
            if kind == 'diode':
                lines.append('diode:d{0} n{1} gnd isat=1e-14 rs=10. cj0=30f '
                             'tt=1p'.format(k, k + 1))
            elif kind == 'bjt':          # diode-connected Gummel-Poon npn
                lines.append('bjt:q{0} n{1} n{1} gnd model=t122'.format(
                    k, k + 1))
            else:                        # diode-connected BSIM3 n-channel
                lines.append('mosbsim3:m{0} n{1} n{1} gnd gnd model=nch '
                             'w=20u l=1u'.format(k, k + 1))
    if clamp != 'diode':
        lines += ['.model t122 bjt (type=npn isat=0.480e-15 nf=1.008 '
                  'bf=99.655 vaf=90.000 ikf=0.190 ise=7.490e-15 ne=1.762 '
                  'nr=1.010 br=38.400 var=7.000 ikr=93.200e-3 isc=0.200e-15 '
                  'nc=1.042 cje=1.e-12 cjc=1.e-12 tf=10p)',
                  '.model nch mosbsim3 (type=n)']
    lines += ['res:rl n{0} gnd r=50.'.format(nsec), '.end', '']
    fd, path = tempfile.mkstemp(suffix='.net', prefix='ladder_')
    os.close(fd)
    with open(path, 'w') as f:
        f.write('\n'.join(lines))
    return path


@pytest.mark.parametrize('im', ['trap', 'BE'])
def test_tran_ladder_cluster_kernel(lib, im):
    """A second circuit on the boundary/interior kernel (both integration
    methods, a band wider than soliton's, inductor branches in the interior)
    against the oracle's refactor-per-iteration loop."""
    path = ladder_deck(im=im)
    try:
        wave, iters, status, ref, tran = tran_both(path)
    finally:
        os.remove(path)
    assert tran.engine.last_path == 'schur'
    assert wave.shape[1] > 150 and wave.shape[0] >= 300
    check_tran(wave, iters, status, ref)
    assert iters.max() >= 2          # the clamps do conduct


def test_tran_ladder_cluster_kernel_mixed_models(lib):
    """The all-models instantiation of the cluster kernel (256 threads, device
    sweeps of Gummel-Poon, BSIM3 and diode clamps on one warp)."""
    path = ladder_deck(nsec=120, every=20, clamp='mixed')
    try:
        wave, iters, status, ref, tran = tran_both(path)
    finally:
        os.remove(path)
    assert tran.engine.last_path == 'schur'
    check_tran(wave, iters, status, ref)


def test_tran_cluster_kernel_falls_back(lib, monkeypatch):
    """A run the boundary/interior kernel gives up on (status / pivot flag) is
    repeated on the general kernels from the same operating point."""
    orig = nd.Engine._schur_tran
    seen = []

    def refuse(self, *a):
        seen.append(orig(self, *a))
        return False
    monkeypatch.setattr(nd.Engine, '_schur_tran', refuse)
    path = ladder_deck(nsec=90)
    try:
        wave, iters, status, ref, tran = tran_both(path)
    finally:
        os.remove(path)
    assert seen == [True] and tran.engine.last_path == 'engine'
    check_tran(wave, iters, status, ref)


@pytest.mark.parametrize('deck', ['rctransient.net', 'oscillator.net',
                                  'ring_osc_ahkab.net', 'memristor.net',
                                  'inverter_chain_bsim.net'])
def test_tran_matches_oracle(lib, deck):
    wave, iters, status, ref, tran = tran_both(deck, max_samples=150)
    check_tran(wave, iters, status, ref)


@pytest.mark.parametrize('tstep', [None, 2e-12, 1.7e-12])
def test_tran_mesfetc_delayed_ports(lib, tstep):
    """Curtice MESFET: two delayed control ports (tau = 5 ps) interpolated
    from the port-voltage history (reference nodal.py:82-108).  tstep = 5 ps
    is the deck's own (tau <= h branch: derivative coefficient (h-tau)/h);
    the smaller steps take the history-walk branch."""
    wave, iters, status, ref, tran = tran_both('mesfetc_plain.net',
                                               max_samples=300, tstep=tstep)
    assert tran.engine.nhist >= 3
    check_tran(wave, iters, status, ref)


def test_sparse_engine_on_small_circuit(lib):
    """Force the sparse static-pivot path and the block team on ring_bsim3;
    results must agree with the dense tile-team path."""
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = queue[0]
    nd.make_nodal_circuit(ckt)
    topo = nd.get_topology(ckt)
    dc = nd.DCNodal(ckt)
    tran = nd.TransientNodal(ckt, Trapezoidal())
    x, res, it = solve(dc.get_guess(), tran.get_source(0.).copy(),
                       dc.convergence_helpers)
    timeVec = np.arange(0., an.tstop, an.tstep)[:100]
    src = tran.source_table(timeVec)
    rows = list(range(topo.n))
    out = []
    for dense in (True, False):
        e = nd.Engine(topo, use_q=True, a0=2. / an.tstep, x_ref=x,
                      dense=dense)
        w, iters, resv, status = e.tran(x, src, rows)
        assert status[0] == 0
        out.append((w[0], iters[0]))
    assert np.allclose(out[0][0], out[1][0], rtol=1e-7, atol=1e-9)
    assert np.max(np.abs(out[0][1] - out[1][1])) <= 1


def test_nd_interface_stepwise(lib):
    """The host-driven nd interface (get_i_Jac / factor_and_solve / accept /
    get_rhs / get_chord_deltax) reproduces the device-resident loop."""
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = queue[0]
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    imo = Trapezoidal()
    tran = nd.TransientNodal(ckt, imo)
    x, res, it = solve(dc.get_guess(), tran.get_source(0.).copy(),
                       dc.convergence_helpers)
    dc.save_OP(x)
    # Jacobian handle materialises to the oracle's matrix
    iVec, Jac = dc.get_i_Jac(x)
    idx = R.prepare(ckt)
    rdc = R.RefDC(ckt, idx, R.Options(glVar))
    iv_r, J_r = rdc.get_i_Jac(x)
    assert np.allclose(iVec, iv_r, rtol=1e-9, atol=1e-15)
    assert np.allclose(Jac.toarray(), J_r, rtol=1e-8, atol=1e-14)
    dx = dc.factor_and_solve(dc.get_source() - iVec, Jac)
    assert np.allclose(dx, np.linalg.solve(J_r, rdc.get_source() - iv_r),
                       rtol=1e-6, atol=1e-10)
    # further solves with the same factors (nodal.py:629-646): vector and
    # matrix right-hand sides, zero columns skipped, residual vector kept
    rng = np.random.default_rng(3)
    B = rng.normal(size=(len(x), 3))
    B[:, 1] = 0.
    X = dc.solve_linear_system(B)
    assert np.allclose(X, np.linalg.solve(J_r, B), rtol=1e-6, atol=1e-10)
    xb = dc.solve_linear_system(B[:, 0])
    assert np.allclose(xb, X[:, 0], rtol=0, atol=0)
    assert np.array_equal(dc.engine.view('ivec').cpu().numpy(), iVec)
    assert np.array_equal(nd.get_submatrix(J_r, 2), J_r[:, :2])
    assert np.allclose(nd.add_to_diagonal(np.zeros((3, 3)), [1., 2.]),
                       np.diag([1., 2., 0.]))
    lam = dc.get_adjoint_voltages(B[:, 0])          # nodal.py:661-672
    assert np.allclose(lam, np.linalg.solve(J_r.T, B[:, 0]), rtol=1e-6,
                       atol=1e-10)
    dc.get_i_Jac(x)          # (the adjoint solve left transposed factors)
    # stepwise transient, host loop as in tran.py:171-206
    timeVec = np.arange(0., an.tstop, an.tstep)[:30]
    tran.set_IC(an.tstep)
    xOld = x.copy()
    waves = [x.copy()]
    for i in range(1, len(timeVec)):
        tran.accept(xOld)
        sV = tran.get_rhs(timeVec[i]).copy()
        if i > 1:
            xOld += tran.get_chord_deltax(sV)
        xn, res, it = solve(xOld, sV, tran.convergence_helpers)
        xOld[:] = xn
        waves.append(xn.copy())
    ref = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im, max_samples=30)
    err = np.abs(np.array(waves) - ref['X'])
    assert np.all(err < glVar.reltol * np.abs(ref['X']) + glVar.abstol)


@pytest.mark.parametrize('team', ['block', 'grid'])
def test_tran_inverter_chain_synthetic(lib, team, monkeypatch):
    """BASELINE config 4 at reduced length (K = 40 stages): sparse plan with
    the supply rail as a hub -- its gather and Crout lists are reduced by the
    whole team (threshold lowered so that this length reaches those paths)
    -- run by one CTA and by the cooperative grid.  The operating point
    starts from the alternating nodeset (see support.chain_nodeset)."""
    from cardoon_b200 import _lib as L
    K, nsamp = 40, 25
    path = S.chain_netlist(K)
    try:
        ckt, queue = S.load_circuit(path)
        an = [a for a in queue if a.anType == 'tran'][0]
        nd.make_nodal_circuit(ckt)
        topo = nd.get_topology(ckt)
        monkeypatch.setattr(nd, 'HEAVY', 16)
        monkeypatch.setattr(nd, 'DENSE_MAX', 8)
        x0 = S.chain_nodeset(ckt, K)
        dc = nd.DCNodal(ckt)
        assert not dc.engine.dense
        tran = nd.TransientNodal(ckt, Trapezoidal())
        x, res, it = solve(x0.copy(), tran.get_source(0.).copy(),
                           dc.convergence_helpers)
        timeVec = np.arange(0., an.tstop, an.tstep)[:nsamp]
        src = tran.source_table(timeVec)
        e = nd.Engine(topo, use_q=True, a0=2. / an.tstep, x_ref=x,
                      team=L.TEAM_BLOCK if team == 'block' else L.TEAM_GRID)
        assert not e.dense
        assert e.info['max_con'] > 16 and e.info['max_terms'] > 16
        wave, iters, resv, status = e.tran(x, src, list(range(topo.n)))
        ref = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im,
                         max_samples=nsamp, x0=x0)
        assert abs(it - ref['op_iterations']) <= 1
        check_tran(wave[0], iters[0], int(status[0]), ref)
    finally:
        os.remove(path)


def test_dc_sweep_bias_npn(lib, capsys):
    """`.analysis dc` of the reference's test/bias_npn.net (dc.py:89-242):
    50-point sweep of a source value, each point starting from the previous
    solution."""
    ckt, queue = S.load_circuit(S.net('bias_npn.net'))
    an = [a for a in queue if a.anType == 'dc'][0]
    ckt.saveReqList = []
    ckt.plotReqList = []
    an.run(ckt)
    got = np.array([t.dC_v for t in ckt.nD_termList]).T
    assert got.shape == (an.num, ckt.nD_dimension)
    ckt2, _ = S.load_circuit(S.net('bias_npn.net'))
    ref = R.run_dc(ckt2, glVar, an.device, an.param, an.start, an.stop,
                   an.num)
    err = np.abs(got - ref['X'])
    assert np.all(err < 1e-9 * np.abs(ref['X']) + 1e-12), err.max()
    assert np.max(np.abs(an.iterations - ref['iters'])) <= 1
    # the swept element is restored to its netlist value
    assert ckt.elemDict[an.device].vdc == 10.
    assert 'Average iterations' in capsys.readouterr().out


def test_dc_sweep_device_parameter(lib):
    """Sweep of a MOSFET parameter (vth0 of one transistor of the BSIM3
    inverter chain): only device records change, the plan is kept."""
    from cardoon_b200.analyses.dc import DCSweep
    ckt, _ = S.load_circuit(S.net('inverter_chain_bsim.net'))
    ckt.saveReqList = []
    ckt.plotReqList = []
    an = DCSweep()
    an.set_param('device', 'x2:mosbsim3:m2')
    an.set_param('param', 'w')
    an.set_param('start', 1e-6)
    an.set_param('stop', 2e-5)
    an.set_param('num', 12)
    an.set_attributes()
    an.run(ckt)
    got = np.array([t.dC_v for t in ckt.nD_termList]).T
    ckt2, _ = S.load_circuit(S.net('inverter_chain_bsim.net'))
    ref = R.run_dc(ckt2, glVar, 'x2:mosbsim3:m2', 'w', 1e-6, 2e-5, 12)
    err = np.abs(got - ref['X'])
    assert np.all(err < 1e-9 * np.abs(ref['X']) + 1e-12), err.max()
    assert np.max(np.abs(an.iterations - ref['iters'])) <= 1


def test_dc_global_temperature_sweep(lib):
    """`.analysis dc param=temp` without a device (dc.py:137-140, 167-177):
    every element whose temp is not set gets set_temp_vars(value); only the
    device records change between points."""
    from cardoon_b200.analyses.dc import DCSweep
    for deck in ('ac_amp.net', 'ring_osc_ahkab.net'):
        ckt, _ = S.load_circuit(S.net(deck))
        ckt.saveReqList = []
        ckt.plotReqList = []
        an = DCSweep()
        an.set_param('param', 'temp')
        an.set_param('start', -20.)
        an.set_param('stop', 90.)
        an.set_param('num', 12)
        an.set_attributes()
        an.run(ckt)
        got = np.array([t.dC_v for t in ckt.nD_termList]).T
        ckt2, _ = S.load_circuit(S.net(deck))
        ref = R.run_dc(ckt2, glVar, None, 'temp', -20., 90., 12)
        err = np.abs(got - ref['X'])
        assert np.all(err < 1e-9 * np.abs(ref['X']) + 1e-12), err.max()
        assert np.max(np.abs(an.iterations - ref['iters'])) <= 1
        # the operating point really moves with temperature
        assert np.max(np.abs(ref['X'][0] - ref['X'][-1])) > 1e-6


def test_no_convergence_is_raised(lib, capsys):
    """Failure is reported the reference's way: the helpers raise
    NoConvergenceError (fsolve.py:42-54), the kernels only return a status."""
    from cardoon_b200.analyses.fsolve import NoConvergenceError
    ckt, _ = S.load_circuit(S.net('bias_npn.net'))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    old = glVar.maxiter
    glVar.maxiter = 3
    try:
        with pytest.raises(NoConvergenceError):
            dc.solve_simple(dc.get_guess(), dc.get_source())
        # the device-side helper chain reports it as a status
        x, res, it, ok = dc.engine.newton(dc.get_guess(), dc.get_source())
        assert not ok[0] and it[0] == 3
    finally:
        glVar.maxiter = old
    x, res, it = solve(dc.get_guess(), dc.get_source(),
                       dc.convergence_helpers)
    assert abs(it - 17) <= 1


def test_homotopy_helpers_on_device(lib):
    """gmin and source stepping (nodal.py:491-604) run inside one launch and
    reach the same operating point as the host-driven helper chain."""
    ckt, _ = S.load_circuit(S.net('bias_npn.net'))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    old = glVar.maxiter
    glVar.maxiter = 8          # plain Newton needs 17: force the helpers
    try:
        x0, sV = dc.get_guess(), dc.get_source()
        xh, resh, ith = dc.solve_homotopy_gmin2(x0.copy(), sV)
        e = dc.engine
        r = e._run_struct(nd.L.OP_NEWTON, helpers=1)
        e.x.copy_(e.lib.torch.as_tensor(x0[None, :]))
        e.sv.copy_(e.lib.torch.as_tensor(sV))
        e._launch(r)
        assert int(e.status[0].item()) == 0
        xd = e.x[0].cpu().numpy()
        assert np.allclose(xd, xh, rtol=1e-6, atol=1e-9)
        assert int(e.iters[0].item()) == ith
        idx = R.prepare(ckt)
        rdc = R.RefDC(ckt, idx, R.Options(glVar))
        xr, resr, itr = rdc.solve_homotopy_gmin2(rdc.get_guess(),
                                                 rdc.get_source())
        assert abs(ith - itr) <= 2
        assert np.allclose(xh, xr, rtol=1e-6, atol=1e-9)
    finally:
        glVar.maxiter = old


def test_sparse_nd_interface_soliton(lib):
    """get_i_Jac / factor_and_solve of the sparse engine against the
    oracle's assembled matrix and SuperLU solve (N = 3022)."""
    ckt, queue = S.load_circuit(S.net('soliton.net'))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    assert not dc.engine.dense
    x = np.random.default_rng(11).uniform(-0.5, 0.3, ckt.nD_dimension)
    iVec, Jac = dc.get_i_Jac(x)
    idx = R.prepare(ckt)
    rdc = R.RefDC(ckt, idx, R.Options(glVar))
    iv_r, J_r = rdc.get_i_Jac(x)
    J_r = J_r.tocsc()
    assert np.allclose(iVec, iv_r, rtol=1e-9, atol=1e-14)
    d = abs(Jac.tocsc() - J_r)
    assert d.max() < 1e-9 * abs(J_r).max()
    rhs = dc.get_source() - iVec
    dx = dc.factor_and_solve(rhs, Jac)
    import scipy.sparse.linalg as spla
    ref = spla.splu(J_r).solve(rhs)
    assert np.allclose(dx, ref, rtol=1e-6, atol=1e-9 * np.abs(ref).max())


def test_tran_backward_euler_and_errfunc(lib):
    """Backward Euler (integration.py:13-49) and the optional error-function
    test of the Newton loop (fsolve.py:113-118, `.options errfunc=1`; it also
    changes which residual the chord predictor re-uses, tran.py:181)."""
    for im, errfunc in (('BE', False), ('trap', True), ('BE', True)):
        ckt, queue = S.load_circuit(S.net('ring_osc_ahkab.net'))
        an = [a for a in queue if a.anType == 'tran'][0]
        an.im = im
        old = glVar.errfunc
        glVar.errfunc = errfunc
        try:
            nd.make_nodal_circuit(ckt)
            dc = nd.DCNodal(ckt)
            imo = Trapezoidal() if im == 'trap' else BEuler()
            tran = nd.TransientNodal(ckt, imo)
            x, res, it = solve(dc.get_guess(), tran.get_source(0.).copy(),
                               dc.convergence_helpers)
            timeVec = np.arange(0., an.tstop, an.tstep)[:120]
            wave, iters, resv, status = tran.run(
                x, timeVec, list(range(ckt.nD_dimension)))
            ref = R.run_tran(ckt, glVar, an.tstop, an.tstep, im,
                             max_samples=120)
            check_tran(wave, iters, status, ref)
        finally:
            glVar.errfunc = old


# ---- dense failure semantics of the reference (nodal.py:385-402, 443-489, 606-627) ----
def test_homotopy_gmin_external_nodes(lib):
    """The dense engine's fourth helper, solve_homotopy_gmin (gmin from
    ground to every external node, nodal.py:469-489): host-driven through the
    nd interface, in-kernel (homotopy mode 2) and the oracle agree."""
    ckt, _ = S.load_circuit(S.net('bias_npn.net'))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    assert dc.engine.dense and dc.engine.n_ext == ckt.nD_nterms
    helpers = dc.convergence_helpers
    assert [h.__name__ if h else None for h in helpers] == [
        'solve_simple', 'solve_homotopy_gmin2', 'solve_homotopy_source',
        'solve_homotopy_gmin', None]            # nodal.py:443-447
    old = glVar.maxiter
    glVar.maxiter = 8          # plain Newton needs 17
    try:
        x0, sV = dc.get_guess(), dc.get_source()
        xh, resh, ith = dc.solve_homotopy_gmin(x0.copy(), sV)
        idx = R.prepare(ckt)
        rdc = R.RefDC(ckt, idx, R.Options(glVar))
        xr, resr, itr = rdc.solve_homotopy_gmin(rdc.get_guess(),
                                                rdc.get_source())
        assert abs(ith - itr) <= 2
        assert np.allclose(xh, xr, rtol=1e-6, atol=1e-9)
        # in-kernel: skip helpers 1-3 (bits 1..3), so that the chain reaches
        # the fourth one
        e = dc.engine
        r = e._run_struct(nd.L.OP_NEWTON, helpers=1 | 2 | 4 | 8)
        e.x.copy_(e.lib.torch.as_tensor(x0[None, :]))
        e.sv.copy_(e.lib.torch.as_tensor(sV))
        e._launch(r)
        assert int(e.status[0].item()) == 0
        assert int(e.hinfo[0, 0].item()) == 3
        assert int(e.iters[0].item()) == ith
        assert np.allclose(e.x[0].cpu().numpy(), xh, rtol=1e-9, atol=1e-12)
    finally:
        glVar.maxiter = old
    # a sparse engine keeps the three-helper list (nodalSP.py:232-235)
    ckt2, _ = S.load_circuit(S.net('soliton.net'))
    nd.make_nodal_circuit(ckt2)
    dc2 = nd.DCNodal(ckt2)
    assert not dc2.engine.dense and len(dc2.convergence_helpers) == 4


FLOATING = """
* node 3 floats at DC (between two capacitors): singular dense Jacobian
vdc:v1 1 gnd vdc=2
res:r1 1 2 r=1k
diode:d1 2 gnd isat=1e-14
cap:c1 2 3 c=1p
cap:c2 3 gnd c=1p
.options sparse=0
.end
"""


def test_singular_matrix_condition_array(lib, tmp_path, capsys):
    """factor_and_solve on an exactly singular dense matrix: LAPACK goes on
    past the zero pivot, the back substitution produces inf/nan, and
    condition_array (nodal.py:385-402) scrubs the step; the oracle (scipy +
    the same scrub) and the device agree on every entry that is finite, and
    on which ones were scrubbed."""
    from cardoon_b200.analyses.fsolve import NoConvergenceError
    deck = tmp_path / 'floating.net'
    deck.write_text(FLOATING)
    ckt, _ = S.load_circuit(str(deck))
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    assert dc.engine.dense
    idx = R.prepare(ckt)
    opt = R.Options(glVar)
    rdc = R.RefDC(ckt, idx, opt)
    names = idx.name_to_row()
    x = dc.get_guess()
    sV = dc.get_source()
    iVec, Jac = dc.get_i_Jac(x)
    dx = dc.factor_and_solve(sV - iVec, Jac).copy()
    # oracle on the same vector (unknown order by terminal name)
    order = [names[nd_name(t)] for t in ckt.nD_termList]
    xr = np.zeros(len(x))
    xr[order] = x
    sr = rdc.get_source()
    ir, Jr = rdc.get_i_Jac(xr)
    dr = rdc.factor_and_solve(sr - ir, Jr)[order]
    scrub = (10. * glVar.abstol, .1 * glVar.maxdelta, -.1 * glVar.maxdelta)
    was_scrubbed = np.isin(dr, scrub)
    assert was_scrubbed.any()
    assert np.array_equal(np.isin(dx, scrub), was_scrubbed)
    assert np.allclose(dx[~was_scrubbed], dr[~was_scrubbed], rtol=1e-9)
    assert np.all(np.isfinite(dx))
    # no helper can make the floating node converge, here or there
    with pytest.raises(NoConvergenceError):
        solve(dc.get_guess(), sV, dc.convergence_helpers)
    with pytest.raises(R.NoConvergence):
        R.run_op(ckt, glVar)


def nd_name(term):
    """Terminal label as oracle.nodal_ref.NodalIndex.name_to_row keys it."""
    if hasattr(term, 'get_label') and not term.get_label().startswith('T:'):
        return term.connection[0].instanceName + ':' + term.instanceName
    return term.instanceName


def test_stateless_engine_rejects_stateful_ops(lib):
    """With ws == NULL the state lives for one launch only: operations that
    continue from kept state are refused (CDN_ERR_STATE), not run on
    uninitialised shared memory."""
    from cardoon_b200._lib import CdnError
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    nd.make_nodal_circuit(ckt)
    topo = nd.get_topology(ckt)
    e = nd.Engine(topo, use_q=True, a0=2e11, keep_state=False, h=1e-11)
    assert e.ws is None
    for op in (nd.L.OP_CHORD, nd.L.OP_ACCEPT, nd.L.OP_RHS, nd.L.OP_SOLVE,
               nd.L.OP_UPDATE_Q):
        with pytest.raises(CdnError):
            e._launch(e._run_struct(op))
    r = e._run_struct(nd.L.OP_TRAN)
    r.nsamples, r.first, r.init_ic = 4, 2, 0
    with pytest.raises(CdnError):
        e._launch(r)
    # plain Newton is fine
    e._launch(e._run_struct(nd.L.OP_NEWTON))


def test_two_contexts_select_their_device(lib):
    """Every launching entry point selects the context's GPU (one box may
    hold several contexts); with one GPU this only checks the call path."""
    import torch
    from cardoon_b200._lib import Lib
    other = Lib(0)
    elem = S.make_element('diode', [1, 0])
    out1, _ = S.cuda_eval(other, elem, np.array([[0.6, 0.7]]))
    out2, _ = S.cuda_eval(lib, elem, np.array([[0.6, 0.7]]))
    assert np.array_equal(out1, out2)
    assert torch.cuda.current_device() == 0


def test_adjoint_voltages_dense(lib):
    """get_adjoint_voltages (nodal.py:661-672): Jac^T lambda = d at the
    operating point, against numpy on the device's own Jacobian."""
    ckt, _ = S.load_circuit(S.net('bias_npn.net'))
    (x, res, it), dc = run_op(ckt)
    iVec, Jac = dc.get_i_Jac(x)
    J = Jac.toarray()
    d = np.random.default_rng(3).standard_normal(len(x))
    lam = dc.get_adjoint_voltages(d)
    ref = np.linalg.solve(J.T, d)
    assert np.allclose(lam, ref, rtol=1e-8, atol=1e-10 * np.abs(ref).max())
