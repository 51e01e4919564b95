"""
Host-side symbolic analysis (csrc/symbolic.cpp) checked on the CPU: the
gather lists must reproduce the oracle's assembled residual/Jacobian, and the
level-scheduled static-pivot LU schedule must solve the system to rounding.
Runs without a GPU through a host-only library context.
"""
import numpy as np
import pytest

from tests import support as S
from tests.plan_emul import HostPlan
from cardoon_b200.analyses import nodalCUDA as nd
from cardoon_b200.globalVars import glVar
from oracle import nodal_ref as R
from oracle.models import device_eval


def oracle_device_buffers(topo, x):
    outs, jacs = [], []
    for g in topo.groups:
        n, nxin = g['n'], g['nxin']
        o = np.zeros((g['ncs'] + g['nqs'], n))
        j = np.zeros((g['ncs'] + g['nqs'], nxin, n))
        for i, e in enumerate(g['elems']):
            xin = np.zeros(nxin)
            nd.set_xin(xin, e.nD_vpos, e.nD_vneg, x)
            oo, jj = device_eval(e, xin[:, None], True, gyr=glVar.gyr)
            o[:, i] = oo[:, 0]
            j[:, :, i] = jj[:, :, 0]
        outs.append(o)
        jacs.append(j)
    return outs, jacs


def oracle_reference_matrix(topo, x, use_q, a0):
    """Same quantity the Engine builds from GPU Jacobians, from the oracle."""
    _, jacs = oracle_device_buffers(topo, x)
    return nd.reference_matrix(topo, use_q, a0, topo.device_stamps(jacs))


def setup(deck):
    ckt, queue = S.load_circuit(S.net(deck))
    nd.make_nodal_circuit(ckt)
    topo = nd.get_topology(ckt)
    idx = R.NodalIndex(ckt)
    return ckt, queue, topo, idx


def check_plan(topo, idx, ckt, use_q, a0, dense, x, ordering=None):
    opt = R.Options(glVar)
    opt.sparse = False
    if use_q:
        im = R.Trapezoidal()
        ref = R.RefTran(ckt, idx, opt, im)
        im.a0 = a0
        ref.update_Gp()
    else:
        ref = R.RefDC(ckt, idx, opt)
    iv_ref, J_ref = ref.get_i_Jac(x)
    A_ref = None
    if not dense:
        A_ref = oracle_reference_matrix(topo, x, use_q, a0)
        assert abs(A_ref - J_ref).max() < 1e-12 * np.abs(J_ref).max()
    hp = HostPlan(topo, use_q, dense, A_ref, ordering)
    out, jac = hp.pack_device(*oracle_device_buffers(topo, x))
    assert len(out) == hp.info['dev_out'] and len(jac) == hp.info['dev_jac']
    lu = hp.assemble(jac, a0, use_q)
    iv = hp.residual(x, out, a0, use_q)
    scale = np.abs(J_ref).max()
    assert np.allclose(hp.matrix(lu), J_ref, rtol=1e-12, atol=1e-13 * scale)
    assert np.allclose(iv, iv_ref, rtol=1e-11,
                       atol=1e-12 * np.abs(iv_ref).max())
    if not dense:
        rhs = np.random.default_rng(0).standard_normal(hp.n)
        dx = hp.solve(hp.factor(lu), rhs)
        ref_dx = np.linalg.solve(J_ref, rhs)
        assert np.allclose(dx, ref_dx, rtol=1e-8,
                           atol=1e-9 * np.abs(ref_dx).max())
    info = hp.info
    hp.close()
    return info


def test_dense_plan_bias_npn():
    ckt, _, topo, idx = setup('bias_npn.net')
    rng = np.random.default_rng(1)
    x = rng.uniform(0., 1., topo.n)
    info = check_plan(topo, idx, ckt, False, 0., True, x)
    assert info['nlu'] == topo.n ** 2


def test_dense_plan_ring_bsim3_tran():
    ckt, _, topo, idx = setup('ring_bsim3.net')
    x = np.random.default_rng(2).uniform(0., 2.5, topo.n)
    check_plan(topo, idx, ckt, True, 2. / 1e-11, True, x)


@pytest.mark.parametrize('deck', ['ring_bsim3.net', 'bias_npn.net',
                                  'oscillator.net', 'memristor.net'])
def test_sparse_plan_small(deck):
    """The sparse schedule on small circuits, DC and transient."""
    ckt, _, topo, idx = setup(deck)
    x = np.random.default_rng(3).uniform(0.1, 0.7, topo.n)
    check_plan(topo, idx, ckt, False, 0., False, x)
    check_plan(topo, idx, ckt, True, 2. / 1e-10, False, x)
    check_plan(topo, idx, ckt, True, 2. / 1e-10, False, x, ordering=1)


def test_sparse_plan_soliton():
    """N = 3022 chain: nested dissection keeps the level count small
    (hard part 1)."""
    ckt, queue, topo, idx = setup('soliton.net')
    assert topo.n == 3022
    x = np.random.default_rng(4).uniform(-1., 0.3, topo.n)
    h = queue[0].tstep
    info = check_plan(topo, idx, ckt, True, 2. / h, False, x)
    assert info['nlev_f'] < 200 and info['nlev_l'] < 80
    assert info['nlu'] < 40000
    info = check_plan(topo, idx, ckt, False, 0., False, x)
    assert info["nlev_l"] < 400      # DC pivot order follows SuperLU row swaps


def test_sparse_plan_chain_with_hub(monkeypatch):
    """Synthetic inverter chain: the supply rail is a hub whose gather and
    Crout lists exceed the cooperative-reduction threshold and are ordered to
    the head of their levels."""
    import os
    monkeypatch.setattr(nd, 'HEAVY', 16)
    K = 40
    path = S.chain_netlist(K)
    try:
        ckt, queue, topo, idx = setup(path)
        x = S.chain_nodeset(ckt, K) \
            + np.random.default_rng(5).uniform(-.2, .2, topo.n)
        info = check_plan(topo, idx, ckt, True, 2. / 1e-10, False, x)
        assert info['max_con'] > 16 and info['max_terms'] > 16
    finally:
        os.remove(path)


@pytest.mark.parametrize('deck,use_q', [('ring_bsim3.net', True),
                                        ('ring_bsim3.net', False),
                                        ('ring_osc_ahkab.net', True),
                                        ('memristor.net', True)])
def test_small_dense_tile_programs(deck, use_q):
    """Dense plans with n <= 8 carry the programs of the register-resident
    8-lane solver (engine.cuh sd8_*): the flat per-lane term lists rebuild
    the assembled Jacobian, and the lane-parallel elimination without row
    moves solves like LAPACK, also for a second right-hand side (chord)."""
    from tests import plan_emul as E
    ckt, _, topo, idx = setup(deck)
    assert topo.n <= 8
    a0 = 2. / 1e-11 if use_q else 0.
    x = np.random.default_rng(7).uniform(0., 2.5, topo.n)
    hp = HostPlan(topo, use_q, True)
    sd = E.sd8_program(hp)
    assert sd is not None
    ptr, prog = sd
    lens = np.diff(ptr)
    assert lens.max() == lens.min()          # equal pieces (padded)
    outs, jacs = oracle_device_buffers(topo, x)
    out, jac = hp.pack_device(outs, jacs)
    lu_ref = hp.assemble(jac, a0, use_q)
    # charge rows of the device Jacobians times a0 (WsSink)
    scaled = []
    for g, j in zip(topo.groups, jacs):
        j = j.copy()
        j[g['ncs']:] *= (a0 if use_q else 1.)
        scaled.append(j)
    _, jac_s = hp.pack_device(outs, scaled)
    base = hp.a['e_g'] + (a0 * hp.a['e_c'] if use_q else 0.)
    lu = base + E.sd8_gather(hp, jac_s)
    assert np.allclose(lu, lu_ref, rtol=1e-13,
                       atol=1e-15 * np.abs(lu_ref).max())
    n = topo.n
    J = hp.matrix(lu)
    assert np.array_equal(J, lu.reshape(n, n))      # slot e = row * n + col
    rng = np.random.default_rng(8)
    y = rng.standard_normal(n)
    xs, fact = E.sd8_solve(J, y)
    ref = np.linalg.solve(J, y)
    assert np.allclose(xs, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())
    y2 = rng.standard_normal(n)
    ref2 = np.linalg.solve(J, y2)
    assert np.allclose(E.sd8_resolve(fact, y2), ref2, rtol=1e-9,
                       atol=1e-12 * np.abs(ref2).max())
    # a singular system goes through (zero pivot: column left unscaled)
    Js = J.copy()
    Js[:, 2] = 0.
    Js[2, :] = 0.
    with np.errstate(all='ignore'):
        xz, _ = E.sd8_solve(Js, y)
    assert not np.all(np.isfinite(xz))
    hp.close()


def test_larger_dense_plan_has_no_tile_program():
    from tests import plan_emul as E
    ckt, _, topo, idx = setup('bias_npn.net')
    hp = HostPlan(topo, False, True)
    assert topo.n > 8 and E.sd8_program(hp) is None
    hp.close()
