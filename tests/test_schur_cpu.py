"""
Boundary/interior analysis of the transient Jacobian
(cardoon_b200/analyses/schur.py) on the CPU: the operators it builds,
applied in the order the kernel applies them (tests/schur_emul.py), must
reproduce the oracle's time loop (oracle/nodal_ref.py run_tran, which
refactors the whole matrix with SuperLU on every iteration as the reference
does, nodalSP.py:281-303).
"""
import numpy as np
import pytest

from tests import support as S
from tests.schur_emul import (SchurEmul, SchurEmulSlaved, band_factor,
                              band_solve)
from tests.test_plan_cpu import setup, oracle_device_buffers
from cardoon_b200.analyses import schur
from cardoon_b200.globalVars import glVar
from oracle import nodal_ref as R


def test_band_lu():
    rng = np.random.default_rng(0)
    for band in (0, 1, 2, 4):
        n = 23
        T = np.zeros((n, n))
        for i in range(n):
            for j in range(max(0, i - band), min(n, i + band + 1)):
                T[i, j] = rng.standard_normal()
            T[i, i] += 6.
        Tb = np.zeros((n, 2 * band + 1))
        for i in range(n):
            for j in range(max(0, i - band), min(n, i + band + 1)):
                Tb[i, band + j - i] = T[i, j]
        r = rng.standard_normal(n)
        x = band_solve(band_factor(Tb, band), band, r)
        assert np.allclose(x, np.linalg.solve(T, r), rtol=1e-12, atol=1e-14)


def _soliton_plan(ncta=8):
    ckt, queue, topo, idx = setup('soliton.net')
    an = [a for a in queue if a.anType == 'tran'][0]
    a0 = 2. / an.tstep
    _, jacs = oracle_device_buffers(topo, np.zeros(topo.n))
    P = schur.analyse(topo, a0, ncta, topo.device_stamps(jacs))
    return ckt, an, topo, idx, P


def test_schur_operators_soliton():
    """Block elimination with the precomputed operators solves J dx = y."""
    import scipy.sparse as sp
    ckt, an, topo, idx, P = _soliton_plan()
    assert P is not None and P.band <= 2 and P.nB <= 96
    assert sum(c.m for c in P.ctas) + P.nB == topo.n
    n = topo.n
    rng = np.random.default_rng(1)
    G = sp.coo_matrix((topo.G[2], (topo.G[0], topo.G[1])), shape=(n, n))
    C = sp.coo_matrix((topo.C[2], (topo.C[0], topo.C[1])), shape=(n, n))
    A = (G + P.a0 * C).toarray()
    # a device part on the diode diagonals (what D(x) looks like)
    ports = schur.device_port_unknowns(topo)
    d = rng.uniform(0., 1e-2, len(ports))
    J = A.copy()
    J[ports, ports] += d
    y = rng.standard_normal(n)
    W = 2 * P.band + 1
    Tb = P.Tb.copy()
    for u, dv in zip(ports, d):
        Tb[P.bpos[u], P.band] += dv
    Tf = band_factor(Tb, P.band)
    tI = [c.Uinv.dot(c.Linv.dot(y[c.Ic])) for c in P.ctas]
    rB = y[P.B].copy()
    for c, t in zip(P.ctas, tI):
        rB -= c.Asi.dot(t)
    dB = band_solve(Tf, P.band, rB)
    dx = np.zeros(n)
    dx[P.B] = dB
    for c, t in zip(P.ctas, tI):
        dx[c.Ic] = t - c.X.dot(dB)
    ref = np.linalg.solve(J, y)
    assert np.allclose(dx, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize('variant', [SchurEmul, SchurEmulSlaved])
def test_schur_time_loop_soliton_vs_oracle(variant):
    """The kernel's phase order (SchurEmul) and the variant with the interior
    eliminated analytically (SchurEmulSlaved: one interior solve and one
    exchange per time step; not built on the device -- after the operators moved
    to sliced ELLPACK both interior phases are hidden behind the device sweeps)
    against the oracle's refactor-per-iteration loop."""
    ckt, an, topo, idx, P = _soliton_plan()
    nsamp = 40 if variant is SchurEmul else 25
    r = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im, idx=idx,
                   max_samples=nsamp)
    opt = R.Options(glVar)
    tran = R.RefTran(ckt, idx, opt, R.Trapezoidal())
    src = np.array([tran.get_source(t).copy() for t in r["t"]])

    def dev_eval(x, deriv):
        outs, jacs = oracle_device_buffers(topo, x)
        return (np.concatenate([o.ravel() for o in outs]),
                np.concatenate([j.ravel() for j in jacs]))

    def qscale(jac, a0):
        jac = jac.copy()
        off = 0
        for g in topo.groups:
            nout, nx, n = g['ncs'] + g['nqs'], g['nxin'], g['n']
            blk = jac[off:off + nout * nx * n].reshape(nout, nx, n)
            blk[g['ncs']:] *= a0
            off += nout * nx * n
        return jac

    em = variant(P, dev_eval, glVar.abstol, glVar.reltol, glVar.maxdelta,
                 int(glVar.maxiter), an.im == 'trap')
    X, iters, res = em.run(r['X'][0], src, qscale)
    assert len(X) == nsamp
    tol = glVar.reltol * np.abs(r['X']) + glVar.abstol
    assert np.all(np.abs(X - r['X']) <= tol)
    assert np.all(np.abs(iters - r['iters']) <= 1)


def test_ell_copies_match_csr():
    """The sliced-ELLPACK copies the kernel walks give the products of the CSR
    operators they were made from (long rows of Linv excluded: a warp sums
    those from the CSR part of the blob)."""
    ckt, an, topo, idx, P = _soliton_plan()
    rng = np.random.default_rng(3)
    for c in P.ctas[:3]:
        for nm in ('Linv', 'Uinv', 'X', 'Arow'):
            op, e = getattr(c, nm), getattr(c, 'e' + nm)
            v = rng.standard_normal(op.shape[1])
            held = e.rowid[e.rowid >= 0]
            ref, got = op.dot(v), e.dot(v)
            assert np.allclose(got[held], ref[held], rtol=1e-13,
                               atol=1e-13 * np.abs(ref).max())
            if nm != 'Linv':
                assert len(held) == op.shape[0]
        blob = schur.pack_blob(P, c)
        assert len(blob) % 16 == 0
