"""
Dense and sparse nodal engines, Newton, integration and the op/tran drivers
of the reference, restated on numpy/scipy (oracle, test infrastructure).

Follows analyses/nodal.py (make_nodal_circuit :113, process_nodal_element
:205, _NLFunction :433, DCNodal :677, TransientNodal :852), nodalSP.py
(triplets :36-85, factor_and_solve :281 = tocsc + SuperLU on *every*
iteration, chord step :345), fsolve.py (:23-128), integration.py and the
drivers op.py:86-102 / tran.py:107-206.  Line references are given per
function.  Device evaluation goes through oracle.models.device_eval.

Node numbering uses insertion order (the reference used Python-2 hash order,
i.e. arbitrary): compare results by terminal name.
"""
import warnings

import numpy as np
import scipy.linalg as linalg
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .models import device_eval


class NoConvergence(Exception):
    pass


class NonFiniteSystem(Exception):
    """The dense Jacobian or right-hand side handed to factor_and_solve holds
    a NaN/inf (see ``_NLFunction.factor_and_solve``)."""


def condition_array(x, opt):                            # nodal.py:385-402
    """Removes infs and nans from x in place: nan -> 10 abstol, +inf ->
    0.1 maxdelta, -inf -> -0.1 maxdelta."""
    if not np.all(np.isfinite(x)):
        x[np.isnan(x)] = 10. * opt.abstol
        x[np.isposinf(x)] = .1 * opt.maxdelta
        x[np.isneginf(x)] = -.1 * opt.maxdelta
    return x


class Options(object):
    """The ``.options`` values Newton reads (globalVars.py:44-57)."""

    def __init__(self, glVar):
        for k in ('abstol', 'reltol', 'maxiter', 'maxdelta', 'errfunc',
                  'gyr', 'sparse', 'temp'):
            setattr(self, k, getattr(glVar, k))


# ---------------------------------------------------------------------------
# index building (nodal.py:113-282)
# ---------------------------------------------------------------------------
class NodalIndex(object):
    def __init__(self, ckt):
        terms = list(ckt.termDict.values())
        for elem in ckt.elemDict.values():
            terms += elem.get_internal_terms()
        self.ref = ckt.termDict['gnd']
        terms.remove(self.ref)
        self.termList = terms
        self.rc = {id(self.ref): -1}
        self.elemList = list(ckt.elemDict.values())
        for elem in self.elemList:
            if elem.localReference:
                self.rc[id(elem.connection[elem.localReference])] = -1
        for i, t in enumerate(terms):
            self.rc[id(t)] = i
        self.dimension = len(terms)
        self.nterms = len(ckt.termDict) - 1
        self.nlin = [e for e in self.elemList if e.isNonlinear]
        self.srcDC = [e for e in self.elemList if e.isDCSource]
        self.srcTD = [e for e in self.elemList if e.isTDSource]
        self.maxDelay = 0.
        self.info = {}
        for elem in self.elemList:
            self.info[id(elem)] = self._process(elem)

    def row(self, term):
        return self.rc[id(term)]

    def name_to_row(self):
        return {t.instanceName if not hasattr(t, 'get_label') or
                t.get_label().startswith('T:') else
                t.connection[0].instanceName + ':' + t.instanceName:
                self.rc[id(t)] for t in self.termList}

    def _process(self, elem):
        rcList = [self.rc[id(t)] for t in elem.connection]
        inf = {}

        def quad(x):
            return [rcList[x[1][0]], rcList[x[0][0]], rcList[x[1][1]],
                    rcList[x[0][1]], x[2]]
        inf['vccs'] = [quad(x) for x in elem.linearVCCS]
        inf['vcqs'] = [quad(x) for x in elem.linearVCQS]

        def plist(ports):
            pos = [(i, rcList[p[0]]) for i, p in enumerate(ports)
                   if rcList[p[0]] > -1]
            neg = [(i, rcList[p[1]]) for i, p in enumerate(ports)
                   if rcList[p[1]] > -1]
            return pos, neg
        if elem.isNonlinear:
            ports = list(elem.controlPorts)
            if elem.nDelays:
                inf['delay'] = np.array([p[2]
                                         for p in elem.delayedContPorts])
                self.maxDelay = max(self.maxDelay, np.amax(inf['delay']))
                ports = ports + list(elem.delayedContPorts)
            inf['vpos'], inf['vneg'] = plist(ports)
            inf['nxin'] = len(ports)
            inf['cpos'], inf['cneg'] = plist(elem.csOutPorts)
            inf['qpos'], inf['qneg'] = plist(elem.qsOutPorts)
            nports = elem.numTerms - 1
            inf['epos'], inf['eneg'] = plist([(i, nports)
                                              for i in range(nports)])
            inf['ncs'] = len(elem.csOutPorts)
        if getattr(elem, 'isFreqDefined', False):
            inf['fpos'], inf['fneg'] = plist(elem.fPortsDefinition)
        if elem.isDCSource or elem.isTDSource:
            inf['srcOut'] = (rcList[elem.sourceOutput[0]],
                             rcList[elem.sourceOutput[1]])
        return inf


def _set_quad(M, row1, col1, row2, col2, g):           # nodal.py:24-41
    if col1 >= 0:
        if row1 >= 0:
            M[row1, col1] += g
        if row2 >= 0:
            M[row2, col1] -= g
    if col2 >= 0:
        if row1 >= 0:
            M[row1, col2] -= g
        if row2 >= 0:
            M[row2, col2] += g


def _set_xin(inf, x):                                   # nodal.py:44-52
    xin = np.zeros(inf['nxin'])
    for i, j in inf['vpos']:
        xin[i] += x[j]
    for i, j in inf['vneg']:
        xin[i] -= x[j]
    return xin


def _set_i(iVec, pos, neg, cur):                        # nodal.py:55-63
    for i, j in pos:
        iVec[j] += cur[i]
    for i, j in neg:
        iVec[j] -= cur[i]


def _set_Jac(M, prow, nrow, pcol, ncol, Jac):           # nodal.py:66-80
    for i1, i in prow:
        for j1, j in pcol:
            M[i, j] += Jac[i1, j1]
        for j1, j in ncol:
            M[i, j] -= Jac[i1, j1]
    for i1, i in nrow:
        for j1, j in ncol:
            M[i, j] += Jac[i1, j1]
        for j1, j in pcol:
            M[i, j] -= Jac[i1, j1]


def delay_interp(td, vp, h, tsV, vpV):                  # nodal.py:82-108
    if td <= h:
        coeff = (h - td) / h
        return vpV[-1] + (vp - vpV[-1]) * coeff, coeff
    timeacc = h
    for i in range(-1, -len(tsV), -1):
        timeacc += tsV[i]
        if td < timeacc:
            return vpV[i - 1] + (vpV[i] - vpV[i - 1]) \
                * (timeacc - td) / tsV[i], 0.
    raise RuntimeError('delay longer than stored history')


# ---- sparse triplet helpers (nodalSP.py:36-85, 111-180) --------------------
def _triplet_quad(T, row1, col1, row2, col2, g):
    def app(val, r, c):
        T[0].append(val)
        T[1].append(r)
        T[2].append(c)
    if col1 >= 0:
        if row1 >= 0:
            app(g, row1, col1)
        if row2 >= 0:
            app(-g, row2, col1)
    if col2 >= 0:
        if row1 >= 0:
            app(-g, row1, col2)
        if row2 >= 0:
            app(g, row2, col2)


def _triplet_jac(T, prow, nrow, pcol, ncol, ncols):
    """Append the structure of one device block; returns the flat indexes
    into the row-major device Jacobian of the '+' and '-' entries, in the
    insertion order of nodalSP.set_Jac_triplet (:64-85)."""
    pidx, nidx = [], []
    for rows, cols in ((prow, pcol), (nrow, ncol)):
        for i1, i in rows:
            for j1, j in cols:
                T[0].append(0.)
                T[1].append(i)
                T[2].append(j)
                pidx.append(i1 * ncols + j1)
    for rows, cols in ((prow, ncol), (nrow, pcol)):
        for i1, i in rows:
            for j1, j in cols:
                T[0].append(0.)
                T[1].append(i)
                T[2].append(j)
                nidx.append(i1 * ncols + j1)
    return np.array(pidx, dtype=int), np.array(nidx, dtype=int)


class _CooJac(object):
    """COO Jacobian with a write cursor (nodalSP._set_Jac :167-180)."""

    def __init__(self, T, n):
        self.data = np.array(T[0], dtype=float)
        self.row = np.array(T[1], dtype=int)
        self.col = np.array(T[2], dtype=int)
        self.n = n
        self.base = 0

    def put(self, jac, pidx, nidx):
        flat = jac.ravel()
        self.data[self.base:self.base + len(pidx)] = flat[pidx]
        self.base += len(pidx)
        self.data[self.base:self.base + len(nidx)] = -flat[nidx]
        self.base += len(nidx)

    def tocsc(self):
        return sp.coo_matrix((self.data, (self.row, self.col)),
                             (self.n, self.n)).tocsc()


# ---------------------------------------------------------------------------
# Newton (fsolve.py:59-128)
# ---------------------------------------------------------------------------
def fsolve_newton(x0, get_deltax, f_eval, opt):
    success = False
    x = x0
    n2 = True
    for i in range(opt.maxiter):
        try:
            deltax = get_deltax(x)
        except NonFiniteSystem:
            # see _NLFunction.factor_and_solve: this Newton solve has failed
            return x, np.inf, i + 1, False
        maxDelta = np.max(np.abs(deltax))
        if maxDelta > opt.maxdelta:
            deltax *= opt.maxdelta / maxDelta
        xnew = x + deltax
        n1 = np.all(np.abs(deltax) < (opt.reltol
                                      * np.maximum(np.abs(x), np.abs(xnew))
                                      + opt.abstol))
        res = np.max(np.abs(deltax))
        if n1:
            if opt.errfunc:
                errFunc = np.max(np.abs(f_eval(xnew)))
                n2 = errFunc < opt.abstol
                res = max(errFunc, res)
            else:
                n2 = True
        x = xnew
        if n1 and n2:
            success = True
            break
    return x, res, i + 1, success


def solve(x0, sV, helpers):                             # fsolve.py:23-56
    for algorithm in helpers:
        if algorithm is None:
            raise NoConvergence('Giving up. No convergence with any method')
        try:
            return algorithm(x0, sV)
        except NoConvergence:
            continue


# ---------------------------------------------------------------------------
# nonlinear function base (nodal.py:433-672, nodalSP.py:220-354)
# ---------------------------------------------------------------------------
class _NLFunction(object):
    def __init__(self, ckt, idx, opt):
        self.ckt = ckt
        self.idx = idx
        self.opt = opt
        self.sparse = bool(opt.sparse)
        self.n = idx.dimension
        self.iVec = np.zeros(self.n)
        self.deltaxVec = np.zeros(self.n)
        self.nfactor = 0
        if self.sparse:
            self.helpers = [self.solve_simple, self.solve_homotopy_gmin2,
                            self.solve_homotopy_source, None]
        else:
            self.helpers = [self.solve_simple, self.solve_homotopy_gmin2,
                            self.solve_homotopy_source,
                            self.solve_homotopy_gmin, None]

    # ---- device evaluation ----
    def _eval(self, elem, xin, deriv):
        out, jac = device_eval(elem, xin, deriv, gyr=self.opt.gyr)
        if deriv:
            return out[:, 0], jac[:, :, 0]
        return out[:, 0]

    # ---- linear algebra ----
    def factor_and_solve(self, rhs, Jac):
        self.nfactor += 1
        if self.sparse:                                 # nodalSP.py:281-303
            M = Jac.tocsc()
            self._lu = spla.splu(M)
            self.deltaxVec[:] = self._lu.solve(rhs)
            return self.deltaxVec
        # nodal.py:606-627: lu_factor + lu_solve, then condition_array on the
        # step (:626).  An exactly singular matrix passes through LAPACK
        # (dgetrf only reports info > 0, scipy warns) and the inf/nan of the
        # back substitution are scrubbed by condition_array.
        #
        # A Jacobian that already holds NaN/inf (BSIM3 far outside its
        # bias range, e.g. Monte-Carlo instance 31 of the config-5 workload)
        # has no reproducible meaning in the reference: the scipy of its
        # time passed the NaNs to dgetrf, whose pivot search on NaNs is
        # implementation defined, and scrubbed the all-NaN step; today's
        # scipy raises ValueError from asarray_chkfinite, which nothing in
        # the reference catches (the run aborts).  The oracle and the CUDA
        # engine define it as "this Newton solve has failed": solve_simple
        # raises NoConvergence (next helper, fsolve.py:44-56) and _homotopy
        # bisects (nodal.py:581-592).
        if not (np.all(np.isfinite(Jac)) and np.all(np.isfinite(rhs))):
            raise NonFiniteSystem()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            self._lu = linalg.lu_factor(Jac)
            deltax = linalg.lu_solve(self._lu, rhs)
        return condition_array(deltax, self.opt)

    def lu_solve(self, b):
        if self.sparse:
            return self._lu.solve(b)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return linalg.lu_solve(self._lu, b)

    def solve_linear_system(self, b):           # nodal.py:629-646
        return condition_array(self.lu_solve(b), self.opt)

    def get_chord_deltax(self, sV):             # nodal.py:648, nodalSP:345
        # uses the iVec left by the last get_i_Jac / get_i call
        return self.lu_solve(sV - self.iVec)

    def _jac(self):
        """Jacobian in the form factor_and_solve wants."""
        return self.Jac

    # ---- convergence helpers ----
    def solve_simple(self, x0, sV):                     # nodal.py:452-467
        def get_deltax(x):
            iVec, Jac = self.get_i_Jac(x)
            return self.factor_and_solve(sV - iVec, Jac)

        def f_eval(x):
            return self.get_i(x) - sV
        x, res, it, ok = fsolve_newton(x0, get_deltax, f_eval, self.opt)
        if ok:
            return x, res, it
        raise NoConvergence('No convergence. iter = {0} res = {1}'.format(
            it, res))

    def _gones(self):
        T = ([], [], [])
        for elem in self.idx.nlin:
            inf = self.idx.info[id(elem)]
            nE = elem.numTerms - 1
            pidx, nidx = _triplet_jac(T, inf['epos'], inf['eneg'],
                                      inf['epos'], inf['eneg'], nE)
            eye = np.eye(nE).ravel()
            k = len(T[0]) - len(pidx) - len(nidx)
            T[0][k:k + len(pidx)] = list(eye[pidx])
            T[0][k + len(pidx):] = list(-eye[nidx])
        G = sp.coo_matrix((T[0], (T[1], T[2])), (self.n, self.n)).tocsr()
        return G if self.sparse else G.toarray()

    def solve_homotopy_gmin2(self, x0, sV):     # nodal.py:491-515
        Gones = self._gones()

        def get_deltax(x):
            iVec, Jac = self.get_i_Jac(x)
            iVec += self.gmin * Gones.dot(x)
            if self.sparse:
                Jac1 = Jac.tocsc() + self.gmin * sp.csc_matrix(Gones)
            else:
                Jac1 = Jac + self.gmin * Gones
            return self.factor_and_solve(sV - iVec, Jac1)

        def f_eval(x):
            iVec = self.get_i(x)
            iVec += self.gmin * Gones.dot(x)
            return iVec - sV
        x, res, it, ok = self._homotopy(0.5, self._set_gmin, x0, get_deltax,
                                        f_eval)
        if ok:
            return x, res, it
        raise NoConvergence('gmin stepping did not converge')

    def solve_homotopy_gmin(self, x0, sV):      # nodal.py:469-489
        idx = np.arange(self.idx.nterms)

        def get_deltax(x):
            iVec, Jac = self.get_i_Jac(x)
            iVec[idx] += self.gmin * x[idx]
            Jac[idx, idx] += self.gmin
            return self.factor_and_solve(sV - iVec, Jac)

        def f_eval(x):
            iVec = self.get_i(x)
            iVec[idx] += self.gmin * x[idx]
            return iVec - sV
        x, res, it, ok = self._homotopy(0.5, self._set_gmin, x0, get_deltax,
                                        f_eval)
        if ok:
            return x, res, it
        raise NoConvergence('gmin stepping did not converge')

    def solve_homotopy_source(self, x0, sV):    # nodal.py:517-533
        def f(lam):
            self._lambda = lam

        def get_deltax(x):
            iVec, Jac = self.get_i_Jac(x)
            return self.factor_and_solve(self._lambda * sV - iVec, Jac)

        def f_eval(x):
            return self.get_i(x) - self._lambda * sV
        x, res, it, ok = self._homotopy(0.5, f, x0, get_deltax, f_eval)
        if ok:
            return x, res, it
        raise NoConvergence('Source stepping did not converge')

    def _set_gmin(self, lam):                   # nodal.py:535-543
        gbase = 1e-3
        self.gmin = gbase / lam - gbase

    def _homotopy(self, lam, f, x0, get_deltax, f_eval):  # nodal.py:545-604
        stack = [0., 1.]
        step = lam
        small = 1e-4
        x = np.copy(x0)
        success = True
        totIter = 0
        res = 0.
        while stack:
            f(lam)
            x, res, it, ok1 = fsolve_newton(x, get_deltax, f_eval, self.opt)
            totIter += it
            if ok1:
                x0[:] = x
                step = stack[-1] - lam
                lam = stack.pop()
            else:
                x[:] = x0
                stack.append(lam)
                step *= .5
                lam -= step
                if (lam < small) or (step < small):
                    success = False
                    break
        return x, res, totIter, success


# ---------------------------------------------------------------------------
# DC (nodal.py:677-847, nodalSP.py:360-553)
# ---------------------------------------------------------------------------
class RefDC(_NLFunction):
    def __init__(self, ckt, idx, opt):
        _NLFunction.__init__(self, ckt, idx, opt)
        self.sVec = np.zeros(self.n)
        if self.sparse:                         # nodalSP.py:389-429
            T = ([], [], [])
            for elem in idx.elemList:
                for q in idx.info[id(elem)]['vccs']:
                    _triplet_quad(T, *q)
            for elem in idx.elemList:           # nodalSP.py:403-407
                inf = idx.info[id(elem)]
                if 'fpos' in inf:
                    Gm = elem.get_G_matrix()
                    D_ = np.zeros((self.n, self.n))
                    _set_Jac(D_, inf['fpos'], inf['fneg'], inf['fpos'],
                             inf['fneg'], Gm)
                    for r_, c_ in zip(*np.nonzero(D_)):
                        T[0].append(D_[r_, c_])
                        T[1].append(r_)
                        T[2].append(c_)
            self.G = sp.coo_matrix((T[0], (T[1], T[2])),
                                   (self.n, self.n)).tocsr()
            self._mbaseLin = len(T[0])
            self._sidx = {}
            for elem in idx.nlin:
                inf = idx.info[id(elem)]
                self._sidx[id(elem)] = _triplet_jac(
                    T, inf['cpos'], inf['cneg'], inf['vpos'], inf['vneg'],
                    inf['nxin'])
            self.Jac = _CooJac(T, self.n)
        else:                                   # nodal.py:706-721
            self.G = np.zeros((self.n, self.n))
            for elem in idx.elemList:
                for q in idx.info[id(elem)]['vccs']:
                    _set_quad(self.G, *q)
                inf = idx.info[id(elem)]
                if 'fpos' in inf:               # nodal.py:719-721
                    _set_Jac(self.G, inf['fpos'], inf['fneg'], inf['fpos'],
                             inf['fneg'], elem.get_G_matrix())
            self.Jac = np.zeros((self.n, self.n))

    def get_guess(self):                        # nodal.py:723-740
        x0 = np.zeros(self.n)
        for elem in self.idx.nlin:
            guess = getattr(elem, 'vPortGuess', None)
            if guess is None:
                continue
            for i, j in self.idx.info[id(elem)]['vpos']:
                x0[j] = guess[i]
        return x0

    def get_source(self):                       # nodal.py:742-759
        self.sVec.fill(0.)
        for elem in self.idx.srcDC:
            n1, n2 = self.idx.info[id(elem)]['srcOut']
            cur = elem.get_DCsource()
            if n1 >= 0:
                self.sVec[n1] -= cur
            if n2 >= 0:
                self.sVec[n2] += cur
        return self.sVec

    def get_i(self, x):                         # nodal.py:761-785
        self.iVec[:] = self.G.dot(x)
        for elem in self.idx.nlin:
            inf = self.idx.info[id(elem)]
            out = self._eval(elem, _set_xin(inf, x), False)
            _set_i(self.iVec, inf['cpos'], inf['cneg'], out)
        return self.iVec

    def get_i_Jac(self, x):             # nodal.py:787-821, nodalSP:494-523
        self.iVec[:] = self.G.dot(x)
        if self.sparse:
            self.Jac.base = self._mbaseLin
        else:
            self.Jac[:, :] = self.G
        for elem in self.idx.nlin:
            inf = self.idx.info[id(elem)]
            out, jac = self._eval(elem, _set_xin(inf, x), True)
            _set_i(self.iVec, inf['cpos'], inf['cneg'], out)
            if self.sparse:
                self.Jac.put(jac[:inf['ncs']], *self._sidx[id(elem)])
            else:
                _set_Jac(self.Jac, inf['cpos'], inf['cneg'], inf['vpos'],
                         inf['vneg'], jac)
        return self.iVec, self.Jac


# ---------------------------------------------------------------------------
# integration (integration.py:13-97)
# ---------------------------------------------------------------------------
class BEuler(object):
    def init(self, h, q):
        self.h = h
        self.a0 = 1. / h
        self.qn1 = np.zeros_like(q)

    def accept(self, q):
        self.qn1[:] = q

    def f_n1(self):
        return self.a0 * self.qn1


class Trapezoidal(object):
    def init(self, h, q):
        self.h = h
        self.a0 = 2. / h
        self.qnm1 = np.copy(q)
        self.dqnm1 = np.zeros_like(q)

    def accept(self, q):
        self.dqnm1 *= -1.
        self.dqnm1 += self.a0 * (q - self.qnm1)
        self.qnm1[:] = q

    def f_n1(self):
        return self.a0 * self.qnm1 + self.dqnm1


# ---------------------------------------------------------------------------
# transient (nodal.py:852-1133, nodalSP.py:559-887)
# ---------------------------------------------------------------------------
class RefTran(_NLFunction):
    def __init__(self, ckt, idx, opt, im):
        _NLFunction.__init__(self, ckt, idx, opt)
        self.im = im
        self.qVec = np.zeros(self.n)
        self.sVec = np.zeros(self.n)
        self.tdVecList = []
        if self.sparse:                         # nodalSP.py:613-664
            T = ([], [], [])
            for elem in idx.elemList:
                for q in idx.info[id(elem)]['vccs']:
                    _triplet_quad(T, *q)
            self._Cbase = len(T[0])
            for elem in idx.elemList:
                for q in idx.info[id(elem)]['vcqs']:
                    _triplet_quad(T, *q)
            self._Cdata = np.array(T[0][self._Cbase:], dtype=float)
            self.C = sp.coo_matrix(
                (self._Cdata, (T[1][self._Cbase:], T[2][self._Cbase:])),
                (self.n, self.n)).tocsr()
            self._mbaseLin = len(T[0])
            self._sidx = {}
            for elem in idx.nlin:
                inf = idx.info[id(elem)]
                cs = _triplet_jac(T, inf['cpos'], inf['cneg'], inf['vpos'],
                                  inf['vneg'], inf['nxin'])
                qs = _triplet_jac(T, inf['qpos'], inf['qneg'], inf['vpos'],
                                  inf['vneg'], inf['nxin'])
                self._sidx[id(elem)] = (cs, qs)
            self.Jac = _CooJac(T, self.n)
        else:                                   # nodal.py:897-913
            self.G = np.zeros((self.n, self.n))
            self.C = np.zeros((self.n, self.n))
            for elem in idx.elemList:
                inf = idx.info[id(elem)]
                for q in inf['vccs']:
                    _set_quad(self.G, *q)
                for q in inf['vcqs']:
                    _set_quad(self.C, *q)
            self.Jac = np.zeros((self.n, self.n))

    def set_IC(self, x, h):             # nodal.py:915-946, nodalSP:666-696
        self.update_q(x)
        self.im.init(h, self.qVec)
        self.update_Gp()
        nsteps = int(np.ceil(self.idx.maxDelay / h))
        self._dpsteps = nsteps * [h]
        for elem in self.idx.nlin:
            if elem.nDelays:
                xin = _set_xin(self.idx.info[id(elem)], x)
                self.tdVecList.append(
                    [nsteps * [xin[i]] for i in range(-elem.nDelays, 0)])

    def update_Gp(self):                # nodal.py:948, nodalSP.py:699-716
        if self.sparse:
            J = self.Jac
            J.data[self._Cbase:self._mbaseLin] = self.im.a0 * self._Cdata
            m = self._mbaseLin
            self.Gp = sp.coo_matrix((J.data[:m], (J.row[:m], J.col[:m])),
                                    (self.n, self.n)).tocsr()
        else:
            self.Gp = self.G + self.im.a0 * self.C

    def update_q(self, x):              # nodal.py:955-969, nodalSP:719-733
        self.qVec[:] = self.C.dot(x)
        for elem in self.idx.nlin:
            inf = self.idx.info[id(elem)]
            out = self._eval(elem, _set_xin(inf, x), False)
            _set_i(self.qVec, inf['qpos'], inf['qneg'], out[inf['ncs']:])
        return self.qVec

    def accept(self, x):                # nodal.py:971-988, nodalSP:735-758
        self.im.accept(self.update_q(x))
        self._dpsteps.append(self.im.h)
        flag = False
        if self.sparse and sum(self._dpsteps[1:]) > self.idx.maxDelay:
            flag = True
            del self._dpsteps[0]
        delayElems = [e for e in self.idx.nlin if e.nDelays]
        for tdVec, elem in zip(self.tdVecList, delayElems):
            xin = _set_xin(self.idx.info[id(elem)], x)
            for i in range(-elem.nDelays, 0):
                tdVec[i].append(xin[i])
                if flag:
                    del tdVec[i][0]

    def get_source(self, t):                    # nodal.py:1004-1034
        self.sVec.fill(0.)
        for elems, fn in ((self.idx.srcDC, lambda e: e.get_DCsource()),
                          (self.idx.srcTD, lambda e: e.get_TDsource(t))):
            for elem in elems:
                n1, n2 = self.idx.info[id(elem)]['srcOut']
                cur = fn(elem)
                if n1 >= 0:
                    self.sVec[n1] -= cur
                if n2 >= 0:
                    self.sVec[n2] += cur
        return self.sVec

    def get_rhs(self, t):                       # nodal.py:1036-1044
        return self.im.f_n1() + self.get_source(t)

    def _xin_delayed(self, elem, inf, x, dcounter):
        xin = _set_xin(inf, x)
        dc = None
        if elem.nDelays:
            dc = np.empty_like(inf['delay'])
            for i in range(-elem.nDelays, 0):
                xin[i], dc[i] = delay_interp(inf['delay'][i], xin[i],
                                             self.im.h, self._dpsteps,
                                             self.tdVecList[dcounter][i])
        return xin, dc

    def get_i(self, x):                 # nodal.py:1046-1080, nodalSP:804
        self.iVec[:] = self.Gp.dot(x)
        dcounter = 0
        for elem in self.idx.nlin:
            inf = self.idx.info[id(elem)]
            xin, _ = self._xin_delayed(elem, inf, x, dcounter)
            if elem.nDelays:
                dcounter += 1
            out = self._eval(elem, xin, False)
            _set_i(self.iVec, inf['cpos'], inf['cneg'], out)
            _set_i(self.iVec, inf['qpos'], inf['qneg'],
                   self.im.a0 * out[inf['ncs']:])
        return self.iVec

    def get_i_Jac(self, x):             # nodal.py:1082-1133, nodalSP:838
        self.iVec[:] = self.Gp.dot(x)
        if self.sparse:
            self.Jac.base = self._mbaseLin
        else:
            self.Jac[:, :] = self.Gp
        dcounter = 0
        for elem in self.idx.nlin:
            inf = self.idx.info[id(elem)]
            xin, dc = self._xin_delayed(elem, inf, x, dcounter)
            out, jac = self._eval(elem, xin, True)
            if elem.nDelays:
                dcounter += 1
                for i in range(-elem.nDelays, 0):
                    jac[:, i] *= dc[i]
            ncs = inf['ncs']
            _set_i(self.iVec, inf['cpos'], inf['cneg'], out)
            _set_i(self.iVec, inf['qpos'], inf['qneg'],
                   self.im.a0 * out[ncs:])
            qJac = self.im.a0 * jac[ncs:, :]
            if self.sparse:
                cs, qs = self._sidx[id(elem)]
                self.Jac.put(jac[:ncs], *cs)
                self.Jac.put(qJac, *qs)
            else:
                _set_Jac(self.Jac, inf['cpos'], inf['cneg'], inf['vpos'],
                         inf['vneg'], jac)
                _set_Jac(self.Jac, inf['qpos'], inf['qneg'], inf['vpos'],
                         inf['vneg'], qJac)
        return self.iVec, self.Jac


# ---------------------------------------------------------------------------
# drivers (op.py:86-102, tran.py:93-206)
# ---------------------------------------------------------------------------
def prepare(ckt):
    if not ckt._flattened:
        ckt.flatten()
    ckt.init()
    return NodalIndex(ckt)


def run_op(ckt, glVar, idx=None):
    """Operating point: dict(x, res, iterations, idx, names)."""
    idx = idx or prepare(ckt)
    opt = Options(glVar)
    dc = RefDC(ckt, idx, opt)
    x0 = dc.get_guess()
    sV = dc.get_source()
    x, res, it = solve(x0, sV, dc.helpers)
    return dict(x=x, res=res, iterations=it, idx=idx, dc=dc)


def run_ac(ckt, glVar, fvec, idx=None):
    """AC sweep (ac.py:69-146 + run_AC nodal.py:286-381): operating point,
    then per frequency Y = G + jwC + device Jacobians (delayed columns times
    exp(-jw tau)) and a dense complex solve.  dict(f, X[nvars, nfreq], x_op,
    idx)."""
    from scipy import linalg
    op = run_op(ckt, glVar, idx)
    idx = op['idx']
    x_op = op['x']
    opt = Options(glVar)
    n = idx.dimension
    G = np.zeros((n, n))
    C = np.zeros((n, n))
    for elem in idx.elemList:
        inf = idx.info[id(elem)]
        for q in inf['vccs']:
            _set_quad(G, *q)
        for q in inf['vcqs']:
            _set_quad(C, *q)
    jacs = []
    for elem in idx.nlin:
        inf = idx.info[id(elem)]
        xin = _set_xin(inf, x_op)
        out, jac = device_eval(elem, xin[:, None], True, gyr=opt.gyr)
        jacs.append(jac[:, :, 0])
    sVec = np.zeros(n, dtype=complex)
    for elem in idx.elemList:
        if not elem.isFDSource:
            continue
        try:
            current = elem.get_AC()
        except AttributeError:
            continue                         # nodal.py:331-343
        n1 = idx.row(elem.connection[elem.sourceOutput[0]])
        n2 = idx.row(elem.connection[elem.sourceOutput[1]])
        if n1 >= 0:
            sVec[n1] -= current
        if n2 >= 0:
            sVec[n2] += current
    fvec = np.asarray(fvec, dtype=float)
    X = np.zeros((n, len(fvec)), dtype=complex)
    for k, jomega in enumerate(2j * np.pi * fvec):
        Y = G + jomega * C
        for elem, jac in zip(idx.nlin, jacs):
            inf = idx.info[id(elem)]
            work = jac.astype(complex)
            if elem.nDelays:
                for i in range(-elem.nDelays, 0):
                    work[:, i] *= np.exp(-jomega * inf['delay'][i])
            ncs = inf['ncs']
            _set_Jac(Y, inf['cpos'], inf['cneg'], inf['vpos'], inf['vneg'],
                     work)
            if len(elem.qsOutPorts):
                _set_Jac(Y, inf['qpos'], inf['qneg'], inf['vpos'],
                         inf['vneg'], jomega * work[ncs:, :])
        for elem in idx.elemList:            # nodal.py:366-370
            inf = idx.info[id(elem)]
            if 'fpos' in inf:
                _set_Jac(Y, inf['fpos'], inf['fneg'], inf['fpos'],
                         inf['fneg'], elem.get_Y_matrix(fvec[k]))
        X[:, k] = linalg.solve(Y, sVec)
    return dict(f=fvec, X=X, x_op=x_op, idx=idx)


def run_dc(ckt, glVar, device, param, start, stop, num):
    """DC sweep of one device parameter (dc.py:141-205): dict(sweep,
    X[num, N], iters, idx).  The circuit's element is left at the last
    value; callers reload the deck."""
    idx = prepare(ckt)
    opt = Options(glVar)
    # no device: global temperature sweep (dc.py:137-140, 167-177)
    dev = ckt.elemDict[device] if device else None
    sweep = np.linspace(start=start, stop=stop, num=num)
    x = RefDC(ckt, idx, opt).get_guess()
    X = np.zeros((num, idx.dimension))
    iters = np.zeros(num, dtype=int)
    resv = np.zeros(num)
    for i, value in enumerate(sweep):
        if dev is None:
            for elem in idx.elemList:
                if not elem.is_set('temp'):
                    try:
                        elem.set_temp_vars(value)
                    except AttributeError:
                        pass            # element independent of temperature
        else:
            setattr(dev, param, value)
            dev.process_params()
        idx = NodalIndex(ckt)                   # restore + re-process
        dc = RefDC(ckt, idx, opt)               # refresh()
        sV = dc.get_source()
        x, res, it = solve(x, sV, dc.helpers)
        X[i] = x
        iters[i] = it
        resv[i] = res
    return dict(sweep=sweep, X=X, iters=iters, res=resv, idx=idx)


def run_tran(ckt, glVar, tstop, tstep, im='trap', idx=None, max_samples=None,
             x0=None):
    """Fixed-step transient: dict(t, X[nsamples, N], iters, res, idx).
    ``x0`` overrides the devices' initial guess of the operating-point solve
    (a nodeset; the reference always starts from get_guess())."""
    idx = idx or prepare(ckt)
    opt = Options(glVar)
    imo = Trapezoidal() if im == 'trap' else BEuler()
    dc = RefDC(ckt, idx, opt)
    tran = RefTran(ckt, idx, opt, imo)
    x = dc.get_guess() if x0 is None else np.array(x0, dtype=float)
    sV = tran.get_source(0.)
    x, res, it = solve(x, sV, dc.helpers)
    op_iters = it
    tran.set_IC(x, tstep)
    timeVec = np.arange(start=0., stop=tstop, step=tstep, dtype=float)
    nsamples = len(timeVec)
    if max_samples:
        nsamples = min(nsamples, max_samples)
    X = np.empty((nsamples, idx.dimension))
    X[0] = x
    iters = np.zeros(nsamples, dtype=int)
    resv = np.zeros(nsamples)
    xOld = x
    for i in range(1, nsamples):
        tran.accept(xOld)
        sV = tran.get_rhs(timeVec[i])
        if i > 1:
            xOld += tran.get_chord_deltax(sV)
        x, res, it = solve(xOld, sV, tran.helpers)
        xOld[:] = x
        X[i] = x
        iters[i] = it
        resv[i] = res
    return dict(t=timeVec[:nsamples], X=X, iters=iters, res=resv, idx=idx,
                op_iterations=op_iters, nfactor=tran.nfactor)
