"""
CUDA nodal back-end: the ``nd`` module of the B200 build.

Same module interface as the reference's cardoon/analyses/nodal.py (dense)
and nodalSP.py (sparse) -- ``make_nodal_circuit``, ``process_nodal_element``,
``restore_RCnumbers``, ``DCNodal``, ``TransientNodal`` -- so that ``op`` and
``tran`` select it exactly the way they select those two
(``nd = nodalSP if glVar.sparse else nodal``, op.py:80-84).  What differs is
where the work happens: the index lists built here are handed once to
``libcardoon_b200.so`` (``cdn_plan_create``: symbolic analysis, gather lists),
and every numerical step -- device evaluation, assembly, LU, Newton control,
integration, and in ``TransientNodal.run`` the whole time loop -- is a kernel
launch (``cdn_engine_run``).  There is no CPU implementation of those steps
in this package.

Host-side pieces kept in Python, as in the reference: terminal numbering,
port-list translation, ``get_guess``/``get_source`` (a handful of scalar
source evaluations), ``save_OP`` and the homotopy bisection control
(nodal.py:545-604) around device Newton solves.
"""
import ctypes as C

import numpy as np

from ..globalVars import glVar
from .. import _lib as L
from .fsolve import NoConvergenceError


# ---------------------------------------------------------------------------
# index building (reference nodal.py:113-282)
# ---------------------------------------------------------------------------
def make_nodal_circuit(ckt, termList=None):
    """Add the ``nD_`` attributes to circuit, elements and terminals."""
    ckt.nD_termList = list(ckt.termDict.values()) + ckt.get_internal_terms()
    try:
        ckt.nD_ref = ckt.termDict['gnd']
        ckt.nD_termList.remove(ckt.nD_ref)
        ckt.nD_ref.nD_namRC = -1
    except KeyError:
        pass
    ckt.nD_elemList = list(ckt.elemDict.values())
    for elem in ckt.nD_elemList:
        if elem.localReference:
            elem.connection[elem.localReference].nD_namRC = -1
    if termList is not None:
        assert len(termList) > 0
        for term in termList:
            ckt.nD_termList.remove(term)
        ckt.nD_termList = list(termList) + ckt.nD_termList
        for i, term in enumerate(ckt.nD_termList):
            term.nD_namRC = i
        ckt.nD_extRClist = [term.nD_namRC for term in termList]
    else:
        for i, term in enumerate(ckt.nD_termList):
            term.nD_namRC = i
    for elem in ckt.nD_elemList:
        elem.nD_intRC = [term.nD_namRC
                         for term in elem.connection[elem.numTerms:]]
    ckt.nD_dimension = len(ckt.nD_termList)
    ckt.nD_nterms = len(ckt.termDict) - 1
    ckt.nD_maxDelay = 0.
    ckt.nD_nlinElem = [e for e in ckt.nD_elemList if e.isNonlinear]
    ckt.nD_delayElem = [e for e in ckt.nD_nlinElem if e.nDelays]
    ckt.nD_freqDefinedElem = [e for e in ckt.nD_elemList if e.isFreqDefined]
    ckt.nD_sourceDCElem = [e for e in ckt.nD_elemList if e.isDCSource]
    ckt.nD_sourceTDElem = [e for e in ckt.nD_elemList if e.isTDSource]
    ckt.nD_sourceFDElem = [e for e in ckt.nD_elemList if e.isFDSource]
    for elem in ckt.nD_elemList:
        process_nodal_element(elem, ckt)
    ckt.__dict__.pop('_cuda_topology', None)


def restore_RCnumbers(elem):
    """Restore RC numbers of internal terminals (nodal.py:196-203)."""
    for term, i in zip(elem.connection[elem.numTerms:], elem.nD_intRC):
        term.nD_namRC = i


def process_nodal_element(elem, ckt):
    """Translate port descriptions into row/column numbers
    (nodal.py:205-282)."""
    rcList = [x.nD_namRC for x in elem.connection]

    def convert_vcs(x):
        return [rcList[x[1][0]], rcList[x[0][0]], rcList[x[1][1]],
                rcList[x[0][1]], x[2]]
    elem.nD_linVCCS = [convert_vcs(x) for x in elem.linearVCCS]
    elem.nD_linVCQS = [convert_vcs(x) for x in elem.linearVCQS]

    def create_list(portlist):
        tmp0 = [rcList[x1[0]] for x1 in portlist]
        tmp1 = [rcList[x1[1]] for x1 in portlist]
        return ([(i, j) for i, j in enumerate(tmp0) if j > -1],
                [(i, j) for i, j in enumerate(tmp1) if j > -1])

    if elem.isNonlinear:
        if elem.nDelays:
            elem.nD_delay = np.array([x1[2] for x1 in elem.delayedContPorts])
            ckt.nD_maxDelay = max(ckt.nD_maxDelay, np.amax(elem.nD_delay))
            allPorts = list(elem.controlPorts) + list(elem.delayedContPorts)
        else:
            allPorts = list(elem.controlPorts)
        elem.nD_allPorts = [(p[0], p[1]) for p in allPorts]
        (elem.nD_vpos, elem.nD_vneg) = create_list(allPorts)
        elem.nD_nxin = len(allPorts)
        (elem.nD_cpos, elem.nD_cneg) = create_list(elem.csOutPorts)
        (elem.nD_qpos, elem.nD_qneg) = create_list(elem.qsOutPorts)
        nports = elem.numTerms - 1
        elem.nD_extPorts = [(i, nports) for i in range(nports)]
        (elem.nD_epos, elem.nD_eneg) = create_list(elem.nD_extPorts)
        elem.nD_rcList = rcList
    if elem.isFreqDefined:
        (elem.nD_fpos, elem.nD_fneg) = create_list(elem.fPortsDefinition)
    if elem.isDCSource or elem.isTDSource or elem.isFDSource:
        elem.nD_sourceOut = (rcList[elem.sourceOutput[0]],
                             rcList[elem.sourceOutput[1]])
    ckt.__dict__.pop('_cuda_topology', None)


def _set_quad(M, row1, col1, row2, col2, g):
    """nodal.py:24-41."""
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


def _jac_entries(prow, nrow, pcol, ncol, Jac):
    """(row, col, value, jacobian column) of T_out Jac T_in, in the order of
    set_Jac (nodal.py:66-80)."""
    out = []
    for i1, i in prow:
        for j1, j in pcol:
            out.append((i, j, Jac[i1, j1], j1))
        for j1, j in ncol:
            out.append((i, j, -Jac[i1, j1], j1))
    for i1, i in nrow:
        for j1, j in ncol:
            out.append((i, j, Jac[i1, j1], j1))
        for j1, j in pcol:
            out.append((i, j, -Jac[i1, j1], j1))
    return out


def ac_operands(ckt, fvec):
    """Operands of the AC sweep (see run_AC): (G, C, delayed, fvec, sVec,
    fd_rc, fd_val)."""
    return _ac_operands(ckt, fvec)


def run_AC(ckt, fvec):
    """Set up and solve the small-signal AC equations (reference
    nodal.py:286-381).

    ckt: nodal-ready circuit with the operating point saved (``save_OP``);
    fvec: frequencies, all different from zero.  Results are stored in the
    terminals as ``aC_V``; returns the matrix (nvars x nfreq).

    The reference assembles and solves ``Y = G + j w C + device Jacobians``
    once per frequency in Python.  Here the frequency-independent part (linear
    stamps, dI/dv into G, dQ/dv into C) is assembled once on the host from the
    device Jacobians at the operating point (one ``cdn_dev_eval`` per
    element), the columns of delayed control ports go to a triplet list that
    gets its ``exp(-j w tau)`` factor per frequency, and all frequencies are
    factored and solved in one launch (``cdn_ac_solve``, one CTA per
    frequency).
    """
    from .._lib import get_lib
    x = get_lib().ac_solve(*_ac_operands(ckt, fvec))
    xVec = np.ascontiguousarray(x.T)
    ckt.nD_ref.aC_V = 0.
    for k, term in enumerate(ckt.nD_termList):
        term.aC_V = xVec[k, :]
    return xVec


def _ac_operands(ckt, fvec):
    n = ckt.nD_dimension
    G = np.zeros((n, n))
    Cm = np.zeros((n, n))
    for elem in ckt.nD_elemList:
        for vccs in elem.nD_linVCCS:
            _set_quad(G, *vccs)
        for vcqs in elem.nD_linVCQS:
            _set_quad(Cm, *vcqs)
    delayed = []
    for elem in ckt.nD_nlinElem:
        (outV, outJac) = elem.eval_and_deriv(elem.nD_xOP)
        ncs = len(elem.csOutPorts)
        first_delayed = outJac.shape[1] - elem.nDelays
        for rows, M, J in (((elem.nD_cpos, elem.nD_cneg), G, outJac[:ncs]),
                           ((elem.nD_qpos, elem.nD_qneg), Cm, outJac[ncs:])):
            if J.shape[0] == 0:
                continue
            for i, j, v, j1 in _jac_entries(rows[0], rows[1], elem.nD_vpos,
                                            elem.nD_vneg, J):
                if j1 >= first_delayed:
                    tau = elem.nD_delay[j1 - first_delayed]
                    delayed.append((i, j, v, 0., tau) if M is G
                                   else (i, j, 0., v, tau))
                else:
                    M[i, j] += v
    sVec = np.zeros(n, dtype=complex)
    for source in ckt.nD_sourceFDElem:
        try:
            current = source.get_AC()
        except AttributeError:
            continue                 # AC routine not implemented: ignored
        outTerm = source.nD_sourceOut
        if outTerm[0] >= 0:
            sVec[outTerm[0]] -= current
        if outTerm[1] >= 0:
            sVec[outTerm[1]] += current
    # frequency-defined elements: Y parameters for the whole sweep at once
    # (get_Y_matrix is vectorised over frequency, nodal.py:366-370)
    fvec = np.asarray(fvec, dtype=float)
    fd_rc, fd_val = [], []
    for elem in ckt.nD_freqDefinedElem:
        Ym = elem.get_Y_matrix(fvec)
        for sa, rows in ((1., elem.nD_fpos), (-1., elem.nD_fneg)):
            for a, i in rows:
                for sb, cols in ((1., elem.nD_fpos), (-1., elem.nD_fneg)):
                    for b, j in cols:
                        fd_rc.append((i, j))
                        fd_val.append(sa * sb * Ym[a, b])
    return G, Cm, delayed, fvec, sVec, fd_rc, fd_val


def add_to_diagonal(M, vec):
    """Adds vec to the leading diagonal entries of the dense matrix M
    (nodal.py:405-416)."""
    n = len(vec)
    M[0:n, 0:n] += np.diag(vec)
    return M


def get_submatrix(M, n):
    """Copy of the n first columns of M (nodal.py:418-426)."""
    return M[:, 0:n].copy()


def get_guess(ckt):
    """Initial guess from the devices' ``vPortGuess`` (nodal.py:723-740):
    only the positive node of each control port is written."""
    x0 = np.zeros(ckt.nD_dimension)
    for elem in ckt.nD_nlinElem:
        try:
            for i, j in elem.nD_vpos:
                x0[j] = elem.vPortGuess[i]
        except AttributeError:
            pass
    return x0


def set_xin(xin, posCols, negCols, xVec):
    """xin = T_cv x (nodal.py:44-52)."""
    for i, j in posCols:
        xin[i] += xVec[j]
    for i, j in negCols:
        xin[i] -= xVec[j]


def set_i(iVec, posRows, negRows, current):
    """iVec += T_i current (nodal.py:55-63; used by the sensitivity module,
    the engines assemble on the device)."""
    for i, j in posRows:
        iVec[j] += current[i]
    for i, j in negRows:
        iVec[j] -= current[i]


# ---------------------------------------------------------------------------
# topology -> plan input
# ---------------------------------------------------------------------------
def _quads_to_coo(quads):
    """[row1, col1, row2, col2, g] stamps -> COO triplets (nodal.py:24-41),
    reference node (-1) dropped."""
    r, c, v = [], [], []
    for row1, col1, row2, col2, g in quads:
        if col1 >= 0:
            if row1 >= 0:
                r.append(row1), c.append(col1), v.append(g)
            if row2 >= 0:
                r.append(row2), c.append(col1), v.append(-g)
        if col2 >= 0:
            if row1 >= 0:
                r.append(row1), c.append(col2), v.append(-g)
            if row2 >= 0:
                r.append(row2), c.append(col2), v.append(g)
    return (np.array(r, dtype=np.int32), np.array(c, dtype=np.int32),
            np.array(v, dtype=np.float64))


def _port_rows(elem, ports):
    rc = elem.nD_rcList
    pos = np.array([rc[p[0]] for p in ports], dtype=np.int32)
    neg = np.array([rc[p[1]] for p in ports], dtype=np.int32)
    return pos, neg


class Topology(object):
    """Formulation data shared by the DC and transient engines of a circuit:
    linear stamps, homogeneous device groups and their port rows."""

    def __init__(self, ckt):
        self.ckt = ckt
        self.n = ckt.nD_dimension
        quadsG, quadsC, quadsE = [], [], []
        for elem in ckt.nD_elemList:
            quadsG += elem.nD_linVCCS
            quadsC += elem.nD_linVCQS
        for elem in ckt.nD_nlinElem:
            # gmin-stepping pattern: unit conductance across every external
            # port of every nonlinear device (nodal.py:491-515)
            rc = elem.nD_rcList
            for (p, q) in elem.nD_extPorts:
                quadsE.append([rc[p], rc[p], rc[q], rc[q], 1.])
        # frequency-defined elements enter the DC equations through their DC
        # Y parameters (nodal.py:719-721, nodalSP.py:403-407) and are absent
        # from the transient ones (nodal.py:921, commented out there)
        quadsFD = []
        for elem in ckt.nD_freqDefinedElem:
            rc = [t.nD_namRC for t in elem.connection]
            Gm = elem.get_G_matrix()
            for a, pa in enumerate(elem.fPortsDefinition):
                for b, pb in enumerate(elem.fPortsDefinition):
                    quadsFD.append([rc[pa[0]], rc[pb[0]], rc[pa[1]],
                                    rc[pb[1]], Gm[a, b]])
        self.G = _quads_to_coo(quadsG)
        self.Gdc = _quads_to_coo(quadsG + quadsFD) if quadsFD else self.G
        self.C = _quads_to_coo(quadsC)
        self.E = _quads_to_coo(quadsE)
        # device groups
        keyed = {}
        for elem in ckt.nD_nlinElem:
            if not hasattr(elem, 'cuda_spec'):
                raise NotImplementedError(
                    '{0}: device type {1} has no CUDA model'.format(
                        elem.instanceName, elem.devType))
            spec = elem.cuda_spec()
            key = (spec['model'], spec['flags'], spec['nxin'], spec['ncs'],
                   spec['nqs'], len(spec['params']), elem.nDelays)
            keyed.setdefault(key, []).append(elem)
        self.groups = []
        for key, elems in keyed.items():
            model, flags, nxin, ncs, nqs, npar, ndelay = key
            g = dict(model=model, flags=flags, nxin=nxin, ncs=ncs, nqs=nqs,
                     ndelay=ndelay, elems=elems, n=len(elems))
            vp, vn, cp, cn, qp, qn = [], [], [], [], [], []
            for e in elems:
                a, b = _port_rows(e, e.nD_allPorts)
                vp.append(a), vn.append(b)
                a, b = _port_rows(e, e.csOutPorts)
                cp.append(a), cn.append(b)
                a, b = _port_rows(e, e.qsOutPorts)
                qp.append(a), qn.append(b)

            def kmajor(lst, k):
                if k == 0:
                    return np.zeros(1, dtype=np.int32)
                return np.ascontiguousarray(
                    np.array(lst, dtype=np.int32).reshape(len(elems), k).T)
            g['vpos'], g['vneg'] = kmajor(vp, nxin), kmajor(vn, nxin)
            g['cpos'], g['cneg'] = kmajor(cp, ncs), kmajor(cn, ncs)
            g['qpos'], g['qneg'] = kmajor(qp, nqs), kmajor(qn, nqs)
            g['delay'] = None
            if ndelay:
                # [ndelay][n] delays of the delayed control ports
                g['delay'] = np.ascontiguousarray(np.array(
                    [e.nD_delay for e in elems], dtype=np.float64).T)
            self.groups.append(g)

    def same_formulation(self, other):
        """True if ``other`` has the same linear stamps, device groups and
        port rows (then a plan built for one serves the other)."""
        if self.n != other.n or len(self.groups) != len(other.groups):
            return False
        # (Gdc: the DC Y parameters of frequency-defined elements are part of
        # the DC plan's linear values, nodal.py:719-721)
        for a, b in zip(self.G + self.C + self.E + self.Gdc,
                        other.G + other.C + other.E + other.Gdc):
            if a.shape != b.shape or not np.array_equal(a, b):
                return False
        keys = ('model', 'flags', 'nxin', 'ncs', 'nqs', 'ndelay', 'n')
        arrs = ('vpos', 'vneg', 'cpos', 'cneg', 'qpos', 'qneg')
        for g, h in zip(self.groups, other.groups):
            if any(g[k] != h[k] for k in keys):
                return False
            if any(not np.array_equal(g[k], h[k]) for k in arrs):
                return False
            if (g['delay'] is None) != (h['delay'] is None) or (
                    g['delay'] is not None
                    and not np.array_equal(g['delay'], h['delay'])):
                return False
        return True

    def pack_params(self):
        """[npar, n] packed records per group (post-process_params)."""
        return [np.ascontiguousarray(
            np.array([e.cuda_spec()['params'] for e in g['elems']]).T)
            for g in self.groups]

    def device_jacobians(self, lib, x):
        """Per group: jac[nout, nxin, n] at the nodal vector x (GPU)."""
        jacs = []
        for g, par in zip(self.groups, self.pack_params()):
            n, nxin = g['n'], g['nxin']
            xin = np.zeros((nxin, n))
            for k in range(nxin):
                a, b = g['vpos'][k], g['vneg'][k]
                xin[k] = np.where(a >= 0, x[np.maximum(a, 0)], 0.) \
                    - np.where(b >= 0, x[np.maximum(b, 0)], 0.)
            _, jac = lib.dev_eval_host(g['model'], g['flags'], par, xin,
                                       g['ncs'] + g['nqs'], True)
            jacs.append(jac.reshape(g['ncs'] + g['nqs'], nxin, n))
        return jacs

    def device_stamps(self, jacs):
        """Signed nodal stamps of the device Jacobians ``jacs`` as COO:
        (rows, cols, dI/dv values, dQ/dv values), nodal.py:66-80."""
        rows, cols, vals_i, vals_q = [], [], [], []
        for g, jac in zip(self.groups, jacs):
            ncs, nqs, nxin = g['ncs'], g['nqs'], g['nxin']
            jac = np.nan_to_num(jac, nan=0., posinf=1e30, neginf=-1e30)
            for o in range(ncs + nqs):
                if o < ncs:
                    rp, rn = g['cpos'][o], g['cneg'][o]
                else:
                    rp, rn = g['qpos'][o - ncs], g['qneg'][o - ncs]
                for k in range(nxin):
                    cp, cn = g['vpos'][k], g['vneg'][k]
                    for r, c, sg in ((rp, cp, 1.), (rn, cn, 1.),
                                     (rp, cn, -1.), (rn, cp, -1.)):
                        m = (r >= 0) & (c >= 0)
                        rows.append(r[m]), cols.append(c[m])
                        z = np.zeros(m.sum())
                        if o < ncs:
                            vals_i.append(sg * jac[o, k][m])
                            vals_q.append(z)
                        else:
                            vals_i.append(z)
                            vals_q.append(sg * jac[o, k][m])
        if not rows:
            e = np.zeros(0)
            return e.astype(np.int32), e.astype(np.int32), e, e
        return (np.concatenate(rows), np.concatenate(cols),
                np.concatenate(vals_i), np.concatenate(vals_q))


def get_topology(ckt):
    topo = ckt.__dict__.get('_cuda_topology')
    if topo is None:
        topo = Topology(ckt)
        ckt._cuda_topology = topo
    return topo


def reference_matrix(topo, use_q, a0, stamps):
    """Numeric Jacobian G + a0 C + device stamps at the reference point
    (scipy CSR, duplicates summed; structural entries kept)."""
    import scipy.sparse as sp
    Gr, Gc, Gv = topo.G if use_q else topo.Gdc
    Cr, Cc, Cv = topo.C
    dr, dc_, di, dq = stamps
    r = [Gr, dr]
    c = [Gc, dc_]
    v = [Gv, di + (a0 * dq if use_q else 0.)]
    if use_q:
        r.append(Cr), c.append(Cc), v.append(a0 * Cv)
    A = sp.coo_matrix((np.concatenate(v),
                       (np.concatenate(r), np.concatenate(c))),
                      shape=(topo.n, topo.n)).tocsr()
    A.sum_duplicates()
    return A


def static_pivot_matching(A):
    """Row -> column matching that maximises the product of matched
    magnitudes of the sparse matrix A (the first static pivot choice).

    Gyrator-built sources and inductors leave structurally zero diagonals at
    DC (reference devices/vdc.py:112-123) which SuperLU survives by partial
    pivoting; the device LU has a fixed pattern, so large entries are
    permuted onto the diagonal once, on the host.
    """
    import scipy.sparse as sp
    from scipy.sparse.csgraph import min_weight_full_bipartite_matching
    n = A.shape[0]
    Wc = abs(A).tocoo()
    # every structural entry stays a candidate; prefer the diagonal among
    # comparable ones
    w = np.maximum(Wc.data, 1e-300) * np.where(Wc.row == Wc.col, 4., 1.)
    colmax = np.zeros(n)
    np.maximum.at(colmax, Wc.col, w)
    cost = 1. + np.log(colmax[Wc.col] / w)
    Bm = sp.coo_matrix((cost, (Wc.row, Wc.col)), shape=(n, n)).tocsr()
    try:
        rows, cols = min_weight_full_bipartite_matching(Bm)
    except ValueError as err:
        raise NoConvergenceError(
            'Structurally singular Jacobian: ' + str(err))
    match = np.empty(n, dtype=np.int32)
    match[rows] = cols
    return match


def refine_pivots(A, rowmap, colmap, thresh=1e-3):
    """Numerically validated static pivot sequence (KLU-style refactor
    strategy).  ``rowmap``/``colmap`` give the row and the unknown at every
    elimination position of a first plan.  The permuted reference matrix is
    factored on the host by SuperLU with threshold partial pivoting in that
    same column order; the row interchanges it needed are folded into a new
    row -> unknown matching, for which elimination *without* pivoting is, at
    the reference point, exactly SuperLU's factorisation.  Returns the new
    matching, or None if no interchange was needed.
    """
    import scipy.sparse.linalg as spla
    n = A.shape[0]
    Bm = A[rowmap][:, colmap].tocsc()
    try:
        lu = spla.splu(Bm, permc_spec='NATURAL', diag_pivot_thresh=thresh,
                       options=dict(SymmetricMode=False))
    except RuntimeError as err:
        raise NoConvergenceError('Singular Jacobian at the reference '
                                 'point: ' + str(err))
    if np.array_equal(lu.perm_r, np.arange(n)):
        return None
    match = np.empty(n, dtype=np.int32)
    match[np.asarray(rowmap)] = np.asarray(colmap)[lu.perm_r]
    return match


PLAN_INFO = ('n', 'nlu', 'nlev_f', 'nlev_l', 'nlev_u', 'nterms', 'ncon',
             'nnzA', 'max_terms', 'max_con', 'dev_out', 'dev_jac')


HEAVY = 0          # 0 = library default (tests lower it to reach the hub paths)


def create_plan(dll, ctx, topo, use_q, dense, match=None, ordering=None,
                check=None, colorder=None, nhist=0):
    """cdn_plan_create on the index lists of ``topo``: returns the plan
    handle and its size figures.  Host-only (works without a GPU)."""
    n = topo.n
    garr = (L.GroupHost * max(1, len(topo.groups)))()
    for k, g in enumerate(topo.groups):
        gh = garr[k]
        gh.model, gh.flags, gh.n = g['model'], g['flags'], g['n']
        gh.nxin, gh.ncs, gh.nqs = g['nxin'], g['ncs'], g['nqs']
        gh.ndelay = g['ndelay']
        for name in ('vpos', 'vneg', 'cpos', 'cneg', 'qpos', 'qneg'):
            setattr(gh, name, g[name].ctypes.data_as(L._vp))
        if g['delay'] is not None:
            gh.delay = g['delay'].ctypes.data_as(L._vp)
    ch = L.CircuitHost()
    ch.n = n
    ch.heavy = HEAVY
    ch.nhist = nhist
    ch.skip_q = 0 if use_q else 1
    if dense:
        ch.ordering = L.ORDER_DENSE
    else:
        ch.ordering = L.ORDER_ND if ordering is None else ordering
    empty = (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
    for nm, (r, c, v) in (('G', topo.G if use_q else topo.Gdc),
                          ('C', topo.C if use_q else empty),
                          ('E', topo.E)):
        low = nm.lower()
        setattr(ch, 'n' + nm, len(v))
        setattr(ch, low + '_row', r.ctypes.data_as(L._vp))
        setattr(ch, low + '_col', c.ctypes.data_as(L._vp))
        setattr(ch, low + '_val', v.ctypes.data_as(L._vp))
    ch.ngroups = len(topo.groups)
    ch.groups = garr
    if match is not None:
        match = np.ascontiguousarray(match, dtype=np.int32)
        ch.rowmatch = match.ctypes.data_as(L._vp)
    if colorder is not None and not dense:
        colorder = np.ascontiguousarray(colorder, dtype=np.int32)
        ch.colorder = colorder.ctypes.data_as(L._vp)
        ch.ordering = L.ORDER_GIVEN
    plan = L._vp()
    rc = dll.cdn_plan_create(ctx, C.byref(ch), C.byref(plan))
    if rc != 0:
        if check is not None:
            check(rc)
        msg = dll.cdn_last_error(ctx)
        raise L.CdnError('cdn_plan_create failed ({0}): {1}'.format(
            rc, msg.decode() if msg else ''))
    info = (C.c_long * 16)()
    dll.cdn_plan_info(plan, info)
    return plan, dict(zip(PLAN_INFO, info[:12]))


def sparse_plan(dll, ctx, topo, use_q, A_ref, ordering=None, check=None,
                nhist=0):
    """Static-pivot sparse plan validated numerically at the reference
    matrix ``A_ref``: matching + nested dissection, then the pivot sequence
    is corrected with the row interchanges SuperLU needs in that order."""
    match = static_pivot_matching(A_ref)
    plan, info = create_plan(dll, ctx, topo, use_q, False, match, ordering,
                             check, nhist=nhist)
    rowmap = plan_array(dll, plan, 'rowmap')
    colmap = plan_array(dll, plan, 'colmap')
    match2 = refine_pivots(A_ref, rowmap, colmap)
    if match2 is not None:
        dll.cdn_plan_free(plan)
        match = match2
        plan, info = create_plan(dll, ctx, topo, use_q, False, match, None,
                                 check, colorder=colmap, nhist=nhist)
    return plan, info, match


_ARRAY_TYPES = dict(e_g=np.float64, e_c=np.float64, e_gones=np.float64,
                    r_lin_g=np.float64, r_lin_c=np.float64,
                    r_lin_gones=np.float64, e_con_kind=np.int8,
                    r_con_kind=np.int8, e_con_pk=np.uint32,
                    r_con_pk=np.uint32, sd_prog=np.uint32)


def plan_array(dll, plan, name):
    """Host copy of one packed plan array (tests, tooling)."""
    ptr = L._vp()
    cnt = C.c_long()
    rc = dll.cdn_plan_array(plan, name.encode(), C.byref(ptr),
                            C.byref(cnt))
    if rc != 0:
        raise KeyError(name)
    dt = np.dtype(_ARRAY_TYPES.get(name, np.int32))
    if cnt.value == 0:
        return np.zeros(0, dtype=dt)
    buf = (C.c_char * (cnt.value * dt.itemsize)).from_address(ptr.value)
    return np.frombuffer(buf, dtype=dt).copy()


# ---------------------------------------------------------------------------
# engine wrapper
# ---------------------------------------------------------------------------
DENSE_MAX = 64
# transients of one sparse circuit try the boundary/interior kernel first
# (analyses/schur.py); tests switch it off to compare the two paths
SCHUR_ENABLED = True


class Engine(object):
    """One plan (symbolic analysis + gather lists) bound to the GPU, plus the
    per-instance state, for a batch of ``B`` circuit instances that share
    the topology."""

    def __init__(self, topo, use_q, a0=0., x_ref=None, batch=1, params=None,
                 lib=None, keep_state=True, dense=None, team=L.TEAM_AUTO,
                 ordering=None, h=0.):
        self.lib = lib or L.get_lib()
        lib = self.lib
        torch = lib.torch
        self.topo = topo
        self.n = n = topo.n
        self.B = batch
        self.use_q = bool(use_q)
        self.a0 = float(a0)
        self.h = float(h)
        # history samples kept per delayed control port (nodalSP.py:686-696)
        maxDelay = max([g['delay'].max() for g in topo.groups
                        if g['delay'] is not None] + [0.])
        self.nhist = 0
        if use_q and maxDelay > 0.:
            if not h > 0.:
                raise ValueError('delayed control ports need the time step')
            self.nhist = int(np.ceil(maxDelay / h)) + 2
        if dense is None:
            dense = uses_dense(n)
        self.dense = dense
        self.team = team
        self.match = None
        self._ref_stamps = None
        self._schur = None      # boundary/interior plan: None = not tried yet
        self.last_path = None   # 'schur' or 'engine': what the last tran() ran on
        if dense:
            self.plan, self.info = create_plan(
                lib.dll, lib.ctx, topo, use_q, True, check=lib.check,
                nhist=self.nhist)
        else:
            xr = np.zeros(n) if x_ref is None else np.asarray(x_ref)
            stamps = topo.device_stamps(topo.device_jacobians(lib, xr))
            self._ref_stamps = stamps
            A_ref = reference_matrix(topo, use_q, a0, stamps)
            self.plan, self.info, self.match = sparse_plan(
                lib.dll, lib.ctx, topo, use_q, A_ref, ordering, lib.check,
                nhist=self.nhist)
        nbytes = lib.dll.cdn_plan_bytes(self.plan)
        self._planbuf = torch.empty(nbytes, dtype=torch.uint8,
                                    device=lib.device)
        lib.check(lib.dll.cdn_plan_bind(self.plan, lib.ptr(self._planbuf),
                                        nbytes, lib.stream()))
        # ---- parameters ------------------------------------------------------------
        self.set_params(params)
        # ---- state -------------------------------------------------------------------
        self.ws_per = lib.dll.cdn_engine_ws_doubles(self.plan)
        off = (C.c_long * 15)()
        lib.check(lib.dll.cdn_engine_layout(self.plan, off))
        self.off = dict(zip(('x', 'dx', 'ivec', 'sv', 'q', 'qprev', 'dq',
                             'y', 'tmp', 'xb', 'lu', 'out', 'jac', 'hist',
                             'piv'), off))
        self.keep_state = keep_state
        self.ws = None
        if keep_state or not (dense and n <= DENSE_MAX):
            self.ws = torch.zeros((batch, self.ws_per), dtype=torch.float64,
                                  device=lib.device)
        keep = []
        self.profile = False
        self.prof = torch.zeros(128, dtype=torch.int64, device=lib.device)
        self.red = torch.zeros(4096, dtype=torch.float64, device=lib.device)
        self.x = torch.zeros((batch, n), dtype=torch.float64,
                             device=lib.device)
        self.iters = torch.zeros(batch, dtype=torch.int32, device=lib.device)
        self.res = torch.zeros(batch, dtype=torch.float64, device=lib.device)
        self.status = torch.zeros(batch, dtype=torch.int32,
                                  device=lib.device)
        # [:, 0] highest convergence helper used, [:, 1] pivot-trouble flag
        self.hinfo = torch.zeros((batch, 2), dtype=torch.int32,
                                device=lib.device)
        # external terminals come first in the unknown vector (nodal.py:472)
        self.n_ext = int(getattr(topo.ckt, 'nD_nterms', 0)) if dense else 0
        self.sv = torch.zeros(n, dtype=torch.float64, device=lib.device)
        self._keep = keep
        self.launches = 0

    def __del__(self):
        try:
            if self.plan:
                self.lib.dll.cdn_plan_free(self.plan)
                self.plan = None
        except Exception:
            pass

    # ---- parameters ----------------------------------------------------------
    def set_params(self, params=None):
        """``params``: per group, [npar, n] (shared by the batch) or
        [npar, B*n] (instance-major columns b*n + i), numpy or device tensors.
        None packs the circuit's own elements."""
        lib = self.lib
        if params is None:
            params = self.topo.pack_params()
        self.params = []
        self.pidx = []
        for k, (g, p) in enumerate(zip(self.topo.groups, params)):
            pidx = None
            if not hasattr(p, 'data_ptr') and p.shape[1] >= 64:
                # identical records (devices sharing a model card and
                # geometry) are stored once and addressed through an index:
                # a warp then reads one address per parameter
                uniq, inv = np.unique(p, axis=1, return_inverse=True)
                if uniq.shape[1] * 4 <= p.shape[1]:
                    pidx = lib.dev(np.ascontiguousarray(
                        inv.reshape(-1).astype(np.int32)))
                    ncol = p.shape[1]
                    p = np.ascontiguousarray(uniq)
            t = p if hasattr(p, 'data_ptr') else lib.dev(p)
            if pidx is None:
                ncol = t.shape[1]
            if ncol == g['n']:
                per_inst = 0
            elif ncol == g['n'] * self.B:
                per_inst = 1
            else:
                raise ValueError('group {0}: {1} parameter columns for {2} '
                                 'devices x {3} instances'.format(
                                     k, ncol, g['n'], self.B))
            self.params.append(t)
            self.pidx.append(pidx)
            lib.check(lib.dll.cdn_plan_set_group_params(
                self.plan, k, lib.ptr(t), int(t.stride(0)), lib.ptr(pidx),
                per_inst))

    # ---- launches ---------------------------------------------------------------
    def _run_struct(self, op, a0=None, gmin=0., lam=1., helpers=0,
                    gmin_ext=0.):
        r = L.RunHost()
        r.helpers = helpers
        r.n_ext = self.n_ext
        r.gmin_ext = gmin_ext
        r.info = self.lib.ptr(self.hinfo)
        r.abstol, r.reltol = glVar.abstol, glVar.reltol
        r.maxdelta, r.maxiter = glVar.maxdelta, int(glVar.maxiter)
        r.errfunc = int(bool(glVar.errfunc))
        r.a0 = self.a0 if a0 is None else a0
        r.gmin, r.lam = gmin, lam
        r.h = self.h
        r.use_q = int(self.use_q)
        r.im = 1
        r.op, r.B = op, self.B
        r.team, r.T = self.team, 0
        lib = self.lib
        r.x = lib.ptr(self.x)
        r.sv_in = lib.ptr(self.sv)
        r.ws = lib.ptr(self.ws)
        r.iters, r.res = lib.ptr(self.iters), lib.ptr(self.res)
        r.status, r.red = lib.ptr(self.status), lib.ptr(self.red)
        r.prof = lib.ptr(self.prof) if self.profile else None
        return r

    def _launch(self, r):
        lib = self.lib
        lib.check(lib.dll.cdn_engine_run(self.plan, C.byref(r),
                                         lib.stream()))
        self.launches += 1

    PHASES = ('device_sweep', 'assembly', 'factor', 'solve', 'update',
              'accept_rhs', 'chord', 'spare', 'stage_L', 'permute_rhs',
              'levels_L', 'stage_U', 'levels_U', 'scatter_dx')

    # timers of the boundary/interior kernel (csrc/schur.cuh SCH_P slots)
    SCHUR_PHASES = ('unused', 'newton_rows_interior_solve', 'boundary_lu',
                    'interior_update_max', 'update', 'accept_boundary_rows',
                    'chord_update', 'newton_exchange', 'accept_chord_interior',
                    'accept_chord_exchange', 'chord_boundary_resolve',
                    'boundary_assembly', 'convergence_exchange')

    def phase_cycles(self, reset=True):
        """SM cycles instance 0 spent per phase since the last reset (set
        ``engine.profile = True`` before launching)."""
        c = self.prof.cpu().numpy()
        if reset:
            self.prof.zero_()
        names = self.SCHUR_PHASES if self.last_path == 'schur' else self.PHASES
        return dict(zip(names, (int(v) for v in c)))

    def last_team(self):
        """Thread team of the last launch: 'tile', 'block' or 'grid'."""
        info = (C.c_long * 16)()
        self.lib.dll.cdn_plan_info(self.plan, info)
        return {0: 'tile', 1: 'block', 2: 'grid'}.get(info[12], 'none')

    def view(self, name, b=0):
        """Slice of the kept state of instance ``b`` (device tensor)."""
        names = list(self.off)
        k = names.index(name)
        lo = self.off[name]
        hi = self.off[names[k + 1]] if k + 1 < len(names) else self.ws_per
        return self.ws[b, lo:hi]

    def newton(self, x0, sV, gmin=0., lam=1., gmin_ext=0.):
        """Damped Newton on the device from ``x0`` [n] or [B, n] with source
        vector ``sV`` [n].  Returns numpy (x, res, iters, ok) per batch."""
        torch = self.lib.torch
        x0 = np.asarray(x0, dtype=np.float64)
        self.x.copy_(torch.as_tensor(np.broadcast_to(
            x0, (self.B, self.n)).copy()))
        self.sv.copy_(torch.as_tensor(np.ascontiguousarray(sV)))
        r = self._run_struct(L.OP_NEWTON, gmin=gmin, lam=lam,
                             gmin_ext=gmin_ext)
        self._launch(r)
        x = self.x.cpu().numpy()
        return (x, self.res.cpu().numpy(), self.iters.cpu().numpy(),
                self.status.cpu().numpy() == 0)

    def simple_op(self, op, x=None, sV=None, a0=None):
        torch = self.lib.torch
        if x is not None:
            self.view('x').copy_(torch.as_tensor(np.ascontiguousarray(x)))
        if sV is not None:
            self.sv.copy_(torch.as_tensor(np.ascontiguousarray(sV)))
        self._launch(self._run_struct(op, a0=a0))

    # ---- boundary / interior transient path ---------------------------------
    SCHUR_CTAS = 8          # CTAs of the cluster (8 is the portable maximum)

    def _schur_plan(self):
        """Boundary/interior plan (analyses/schur.py) of this engine, built on
        first use; None when the circuit is not eligible."""
        if self._schur is not None:
            return self._schur or None
        self._schur = False
        if (self.dense or self.B != 1 or not self.use_q or self.nhist
                or glVar.errfunc or not SCHUR_ENABLED
                or self._ref_stamps is None):
            return None
        from . import schur
        lib = self.lib
        P = schur.analyse(self.topo, self.a0, self.SCHUR_CTAS,
                          self._ref_stamps)
        if P is None:
            return None
        if (P.out_size != self.info['dev_out']
                or P.jac_size != self.info['dev_jac']):
            return None
        blob, off = schur.pack(P)
        need = max(lib.dll.cdn_schur_smem_bytes(
            int(off[k + 1] - off[k]), int(c.m), P.nB, P.band, P.out_size,
            P.jac_size, P.ncta) for k, c in enumerate(P.ctas))
        if need > 227 * 1024:
            return None
        keep = dict(P=P, schur=schur)
        keep['blob'] = lib.dev(np.frombuffer(blob, dtype=np.uint8).copy())
        keep['off'] = np.ascontiguousarray(off, dtype=np.int64)
        keep['m'] = np.array([c.m for c in P.ctas], dtype=np.int32)
        ports = [schur.remap_ports(P, g) for g in self.topo.groups]
        keep['ports'] = [(lib.dev(a), lib.dev(b)) for a, b in ports]
        ng = len(ports)
        keep['vp'] = (C.c_void_p * max(1, ng))(
            *[t[0].data_ptr() for t in keep['ports']])
        keep['vn'] = (C.c_void_p * max(1, ng))(
            *[t[1].data_ptr() for t in keep['ports']])
        self._schur = keep
        return keep

    def _schur_tran(self, r, var, save_rows, check):
        """Launch the boundary/interior kernel for the run described by ``r``;
        False when it asks for the general kernels (status / pivot flag; only
        looked at if ``check``, which synchronises)."""
        lib = self.lib
        k = self._schur
        P, schur = k['P'], k['schur']
        key = (np.asarray(var).tobytes(),
               np.asarray(save_rows, dtype=np.int32).tobytes())
        if k.get('key') != key:
            own, loc = schur.locate(P, var) if len(var) else (
                np.zeros(0, np.int32), np.zeros(0, np.int32))
            sp, sk, sl = schur.save_lists(P, save_rows)
            k['maps'] = [lib.dev(a) if len(a) else None for a in (own, loc)] \
                + [lib.dev(sp), lib.dev(sk) if len(sk) else None,
                   lib.dev(sl) if len(sl) else None]
            k['key'] = key
        dev = k['maps']
        h = L.SchurHost()
        h.ncta, h.nB, h.band = P.ncta, P.nB, P.band
        h.blob = lib.ptr(k['blob'])
        h.cta_off = k['off'].ctypes.data_as(L._vp)
        h.m = k['m'].ctypes.data_as(L._vp)
        h.vpos_b = C.cast(k['vp'], L._vp)
        h.vneg_b = C.cast(k['vn'], L._vp)
        h.src_own, h.src_loc = lib.ptr(dev[0]), lib.ptr(dev[1])
        h.save_ptr, h.save_k, h.save_loc = (lib.ptr(dev[2]), lib.ptr(dev[3]),
                                            lib.ptr(dev[4]))
        self._x_keep = self.x.clone()
        self.hinfo.zero_()
        rc = lib.dll.cdn_schur_tran(self.plan, C.byref(r), C.byref(h),
                                    lib.stream())
        if rc != 0:
            # (a launch the device refuses -- cluster size, shared memory --
            # must not cost the run: the general kernels take over for good)
            self._schur = False
            return False
        self.launches += 1
        if not check:
            return True
        st = int(self.status[0].item())
        piv = int(self.hinfo[0, 1].item())
        return st == 0 and piv == 0

    def jacobian_coords(self):
        rows = np.empty(self.info['nlu'], dtype=np.int32)
        cols = np.empty(self.info['nlu'], dtype=np.int32)
        self.lib.check(self.lib.dll.cdn_plan_slot_coords(
            self.plan, rows.ctypes.data_as(L._vp),
            cols.ctypes.data_as(L._vp)))
        return rows, cols

    def tran(self, x0, src, save_rows, im='trap', first=1, init_ic=True,
             wave=None, iters=None, res=None, sync=True, reuse_src=False,
             var_rows=None):
        """Whole fixed-step time loop on the device (tran.py:171-206).

        ``x0``: operating point [n] or [B, n], numpy or a device tensor (None:
        keep self.x);
        ``src``: source vectors [nsamples, n] (tran.get_source(t_i));
        ``save_rows``: unknowns to record.  Returns (wave[B, nsamples, nsave],
        iters[B, nsamples], res[B, nsamples], status[B]) as device tensors
        (numpy if ``sync``).  ``reuse_src``: ``src`` and ``save_rows`` are the
        arrays of the previous call and their device copies are kept;
        ``var_rows``: the rows of ``src`` that vary with time, when the caller
        knows them (TransientNodal.source_rows), else they are found by
        comparing the samples.
        """
        lib = self.lib
        torch = lib.torch
        nsamples = src.shape[0]
        if x0 is not None:
            if hasattr(x0, 'data_ptr'):          # already on the device
                self.x.copy_(x0.reshape(-1, self.n).expand(self.B, self.n))
            else:
                x0 = np.asarray(x0, dtype=np.float64)
                self.x.copy_(torch.as_tensor(np.broadcast_to(
                    x0, (self.B, self.n)).copy()))
        # constant rows go to src_dc, time-varying rows to a compact table
        if reuse_src and getattr(self, '_src_key', None) == (id(src),
                                                             src.shape):
            var = self._src_var
        else:
            if var_rows is not None:
                var = np.asarray(var_rows, dtype=np.int32)
            else:
                var = np.nonzero(np.any(src != src[0],
                                        axis=0))[0].astype(np.int32)
            dc = src[0].copy()
            dc[var] = 0.
            self._src = (lib.dev(var) if len(var) else None,
                         lib.dev(np.ascontiguousarray(src[:, var]))
                         if len(var) else None, lib.dev(dc),
                         lib.dev(np.asarray(save_rows, dtype=np.int32)))
            self._src_var = var
            self._src_key = (id(src), src.shape)
        nsave = len(save_rows)
        if wave is None:
            wave = torch.zeros((self.B, nsamples, nsave),
                               dtype=torch.float64, device=lib.device)
        if iters is None:
            iters = torch.zeros((self.B, nsamples), dtype=torch.int32,
                                device=lib.device)
        if res is None:
            res = torch.zeros((self.B, nsamples), dtype=torch.float64,
                              device=lib.device)
        if first == 1:
            wave[:, 0, :] = self.x[:, self._src[3].long()]
        r = self._run_struct(L.OP_TRAN, helpers=1)
        r.im = 1 if im == 'trap' else 0
        r.nsamples, r.first, r.init_ic = nsamples, first, int(init_ic)
        r.nsrc, r.nsave = len(var), nsave
        r.src_rows, r.src_val = lib.ptr(self._src[0]), lib.ptr(self._src[1])
        r.src_dc, r.save_rows = lib.ptr(self._src[2]), lib.ptr(self._src[3])
        r.wave, r.iters, r.res = lib.ptr(wave), lib.ptr(iters), lib.ptr(res)
        self.last_path = 'engine'
        if first == 1 and init_ic and self._schur_plan() is not None:
            # one sparse circuit with few device ports: boundary/interior
            # kernel on a thread-block cluster (analyses/schur.py); Newton
            # failures and degraded pivots fall back to the general kernels,
            # which have the helper chain
            if self._schur_tran(r, var, save_rows, sync):
                self.last_path = 'schur'
            else:
                self.x.copy_(self._x_keep)
        if self.last_path == 'engine':
            self._launch(r)
        if not sync:
            return wave, iters, res, self.status
        return (wave.cpu().numpy(), iters.cpu().numpy(), res.cpu().numpy(),
                self.status.cpu().numpy())


# ---------------------------------------------------------------------------
# nonlinear-function base: convergence helpers (nodal.py:433-604)
# ---------------------------------------------------------------------------
def uses_dense(n):
    """Storage rule of the engines: dense LU with partial pivoting for small
    systems or ``sparse=0`` (the reference's nodal.py engine), else the
    static-pivot sparse LU (nodalSP.py)."""
    return (n <= DENSE_MAX) or (not glVar.sparse and n <= 2000)


class _NLFunction(object):
    @property
    def convergence_helpers(self):
        """Helper list of the engine in use: the dense engine has a fourth
        helper (nodal.py:443-447), the sparse one does not
        (nodalSP.py:232-235)."""
        helpers = [self.solve_simple, self.solve_homotopy_gmin2,
                   self.solve_homotopy_source]
        engine = getattr(self, 'engine', None)
        dense = engine.dense if engine is not None \
            else uses_dense(self.ckt.nD_dimension)
        if dense:
            helpers.append(self.solve_homotopy_gmin)
        return helpers + [None]

    def __init__(self):
        pass

    def _newton(self, x0, sV, gmin=0., lam=1., gmin_ext=0.):
        x, res, it, ok = self.engine.newton(x0, sV, gmin=gmin, lam=lam,
                                            gmin_ext=gmin_ext)
        if not ok[0] and int(self.engine.hinfo[0, 1].item()):
            # a static pivot of the sparse plan degraded on the way (the
            # reference re-pivots on every iteration, nodalSP.py:281-303):
            # build a new plan validated at the starting point of this solve
            # and repeat it once
            if self._replan(x0):
                x, res, it, ok = self.engine.newton(
                    x0, sV, gmin=gmin, lam=lam, gmin_ext=gmin_ext)
        return x[0], float(res[0]), int(it[0]), bool(ok[0])

    def _replan(self, x_ref):
        """Hook: rebuild the sparse plan at ``x_ref``; True if done."""
        return False

    def solve_simple(self, x0, sV):
        # no docstring: the driver prints helper docstrings (fsolve.py:47)
        x, res, it, ok = self._newton(x0, sV)
        if ok:
            return (x, res, it)
        raise NoConvergenceError(
            'No convergence. iter = {0} res = {1}'.format(it, res))

    def solve_homotopy_gmin2(self, x0, sV):
        """gmin stepping (nonlinear ports)"""
        state = dict(gmin=0.)

        def f(lam):
            gbase = 1e-3
            state['gmin'] = gbase / lam - gbase
        x, res, it, ok = self._homotopy(
            0.5, f, x0, lambda x: self._newton(x, sV, gmin=state['gmin']))
        if ok:
            return (x, res, it)
        raise NoConvergenceError('gmin stepping did not converge')

    def solve_homotopy_gmin(self, x0, sV):
        """gmin stepping"""
        # gmin from ground to every external node (nodal.py:469-489)
        state = dict(gmin=0.)

        def f(lam):
            gbase = 1e-3
            state['gmin'] = gbase / lam - gbase
        x, res, it, ok = self._homotopy(
            0.5, f, x0,
            lambda x: self._newton(x, sV, gmin_ext=state['gmin']))
        if ok:
            return (x, res, it)
        raise NoConvergenceError('gmin stepping did not converge')

    def solve_homotopy_source(self, x0, sV):
        """source stepping"""
        state = dict(lam=1.)

        def f(lam):
            state['lam'] = lam
        x, res, it, ok = self._homotopy(
            0.5, f, x0, lambda x: self._newton(x, sV, lam=state['lam']))
        if ok:
            return (x, res, it)
        raise NoConvergenceError('Source stepping did not converge')

    @staticmethod
    def _homotopy(lam, f, x0, newton):
        """Bisection on the continuation parameter (nodal.py:545-604); every
        inner solve is one device Newton launch."""
        stack = [0., 1.]
        step = lam
        small = 1e-4
        x = np.copy(x0)           # (x0 itself is updated in place below,
        success = True            #  as in the reference)
        totIter = 0
        res = 0.
        while stack:
            f(lam)
            x, res, it, ok1 = newton(x)
            totIter += it
            if ok1:
                x0[:] = x
                step = stack[-1] - lam
                lam = stack.pop()
            else:
                x = np.copy(x0)
                stack.append(lam)
                step *= .5
                lam -= step
                if (lam < small) or (step < small):
                    success = False
                    break
        return x, res, totIter, success

    # ---- nd-interface linear algebra (device state kept in engine.ws) ------
    def get_i(self, xVec):
        e = self.engine
        e.simple_op(L.OP_GET_I, x=xVec)
        self.iVec = e.view('ivec').cpu().numpy()
        return self.iVec

    def get_i_Jac(self, xVec):
        """(iVec, Jac): Jac is a handle to the assembled device matrix
        (``toarray()``/``tocsc()`` materialise it on the host)."""
        e = self.engine
        e.simple_op(L.OP_GET_I_JAC, x=xVec)
        self.iVec = e.view('ivec').cpu().numpy()
        return (self.iVec, DeviceJacobian(e))

    def factor_and_solve(self, rhs, Jac):
        """Solve ``Jac dx = rhs`` with the Jacobian assembled by the last
        ``get_i_Jac`` (it lives on the device)."""
        e = self.engine
        torch = e.lib.torch
        e.view('y').copy_(torch.as_tensor(np.ascontiguousarray(rhs)))
        e.simple_op(L.OP_SOLVE)
        self.deltaxVec = e.view('dx').cpu().numpy()
        return self.deltaxVec

    def solve_linear_system(self, b):
        """Solve ``Jac x = b`` with the factors of the last
        ``factor_and_solve`` (reference nodal.py:629-646, nodalSP.py:305-332);
        ``b`` may be a matrix, solved column by column.  The residual vector
        the chord predictor relies on is left untouched."""
        e = self.engine
        torch = e.lib.torch
        b = np.asarray(b, dtype=float)
        keep = e.view('ivec').clone()
        cols = b.reshape(b.shape[0], -1)
        x = np.zeros_like(cols)
        for k in range(cols.shape[1]):
            if np.any(cols[:, k] != 0.):          # zero rhs: skipped
                e.view('ivec').zero_()
                e.view('sv').copy_(torch.as_tensor(
                    np.ascontiguousarray(cols[:, k])))
                e.simple_op(L.OP_CHORD)
                x[:, k] = e.view('dx').cpu().numpy()
        e.view('ivec').copy_(keep)
        return x.reshape(b.shape)

    def get_adjoint_voltages(self, d):
        """Adjoint voltages for sensitivity calculations: solves
        ``Jac^T lambda = d`` with the Jacobian of the last ``get_i_Jac``
        (reference nodal.py:661-672, ``lu_solve(..., trans=1)``).

        Dense engines only: the Jacobian is assembled again at the vector of
        the last ``get_i_Jac`` / ``get_i`` (the in-place factorisation has
        overwritten it), its transpose is written into the device's Jacobian
        slots and factored and solved there (CDN_OP_SOLVE); the static-pivot
        sparse plan has no schedule for the transposed pattern.  The factors
        kept for ``solve_linear_system`` / the chord step are those of the
        transpose afterwards: call ``get_i_Jac`` + ``factor_and_solve`` again
        before going on with the forward system."""
        e = self.engine
        if not e.dense:
            raise NotImplementedError(
                'get_adjoint_voltages: the sparse plan of the CUDA back-end '
                'has no transposed schedule (use sparse=0)')
        torch = e.lib.torch
        n = e.n
        keep = e.view('ivec').clone()
        e.simple_op(L.OP_GET_I_JAC)
        e.view('ivec').copy_(keep)
        lu = e.view('lu')
        J = lu.view(n, n).clone()
        lu.copy_(J.t().contiguous().view(-1))
        e.view('y').copy_(torch.as_tensor(np.ascontiguousarray(
            d, dtype=np.float64)))
        e.simple_op(L.OP_SOLVE)
        return e.view('dx').cpu().numpy()

    def get_chord_deltax(self, sV, iVec=None):
        e = self.engine
        torch = e.lib.torch
        if iVec is not None:
            e.view('ivec').copy_(torch.as_tensor(np.ascontiguousarray(iVec)))
        e.view('sv').copy_(torch.as_tensor(np.ascontiguousarray(sV)))
        e.simple_op(L.OP_CHORD)
        self.deltaxVec = e.view('dx').cpu().numpy()
        return self.deltaxVec


class DeviceJacobian(object):
    def __init__(self, engine):
        self.engine = engine
        self.shape = (engine.n, engine.n)

    def tocoo(self):
        import scipy.sparse as sp
        e = self.engine
        rows, cols = e.jacobian_coords()
        vals = e.view('lu').cpu().numpy()
        return sp.coo_matrix((vals, (rows, cols)), shape=self.shape)

    def tocsc(self):
        return self.tocoo().tocsc()

    def toarray(self):
        return self.tocoo().toarray()


# ---------------------------------------------------------------------------
# DC (nodal.py:677-847, nodalSP.py:360-553)
# ---------------------------------------------------------------------------
class DCNodal(_NLFunction):
    def __init__(self, ckt):
        _NLFunction.__init__(self)
        self.ckt = ckt
        self.sVec = np.zeros(ckt.nD_dimension)
        self.refresh()

    def refresh(self):
        """(Re)build the device-side formulation; call after parameter
        changes (dc.py:179-188 pattern)."""
        self.ckt.__dict__.pop('_cuda_topology', None)
        topo = get_topology(self.ckt)
        old = getattr(self, 'engine', None)
        if old is not None and old.topo.same_formulation(topo):
            # only device records (or source values) changed: the symbolic
            # plan stays, the records are uploaded again
            old.topo = topo
            old.set_params(None)
            return
        self.engine = Engine(topo, use_q=False, a0=0.,
                             x_ref=self.get_guess())

    def _replan(self, x_ref):
        if self.engine.dense:
            return False
        self.engine = Engine(self.engine.topo, use_q=False, a0=0.,
                             x_ref=np.array(x_ref, dtype=float))
        self.replans = getattr(self, 'replans', 0) + 1
        return True

    def get_guess(self):
        return get_guess(self.ckt)

    def get_source(self):
        self.sVec.fill(0.)
        for elem in self.ckt.nD_sourceDCElem:
            outTerm = elem.nD_sourceOut
            current = elem.get_DCsource()
            if outTerm[0] >= 0:
                self.sVec[outTerm[0]] -= current
            if outTerm[1] >= 0:
                self.sVec[outTerm[1]] += current
        return self.sVec

    def save_OP(self, xVec):
        self.ckt.nD_ref.nD_vOP = 0.
        self.xop = xVec
        for v, term in zip(xVec, self.ckt.nD_termList):
            term.nD_vOP = v
        for elem in self.ckt.nD_nlinElem:
            xin = np.zeros(elem.nD_nxin)
            set_xin(xin, elem.nD_vpos, elem.nD_vneg, xVec)
            elem.nD_xOP = xin


# ---------------------------------------------------------------------------
# transient (nodal.py:852-1133, nodalSP.py:559-887)
# ---------------------------------------------------------------------------
class TransientNodal(_NLFunction):
    def __init__(self, ckt, im):
        _NLFunction.__init__(self)
        self.ckt = ckt
        self.im = im
        self.sVec = np.zeros(ckt.nD_dimension)
        self.engine = None

    def _imcode(self):
        return 1 if type(self.im).__name__ == 'Trapezoidal' else 0

    def _make_engine(self, h, x_ref):
        a0 = (2. / h) if self._imcode() else (1. / h)
        self.h = h
        self.engine = Engine(get_topology(self.ckt), use_q=True, a0=a0,
                             x_ref=x_ref, h=h)

    def refresh(self):
        self.ckt.__dict__.pop('_cuda_topology', None)
        self.engine = None

    def set_IC(self, h, x=None):
        """Initial conditions from the saved operating point
        (nodalSP.py:666-696)."""
        if x is None:
            x = np.array([t.nD_vOP for t in self.ckt.nD_termList])
        self._make_engine(h, x)
        e = self.engine
        r = e._run_struct(L.OP_SET_IC)
        r.im = self._imcode()
        e.view('x').copy_(e.lib.torch.as_tensor(np.ascontiguousarray(x)))
        e._launch(r)
        self.im.init(h, e.view('q').cpu().numpy())

    def update_Gp(self):
        pass      # G' = G + a0 C is formed inside the kernels

    def update_q(self, xVec):
        e = self.engine
        e.simple_op(L.OP_UPDATE_Q, x=xVec)
        self.qVec = e.view('q').cpu().numpy()
        return self.qVec

    def accept(self, xVec):
        e = self.engine
        r = e._run_struct(L.OP_ACCEPT)
        r.im = self._imcode()
        e.view('x').copy_(e.lib.torch.as_tensor(np.ascontiguousarray(xVec)))
        e._launch(r)
        self.im.accept(e.view('q').cpu().numpy())

    def get_source(self, ctime):
        self.sVec.fill(0.)
        for elem in self.ckt.nD_sourceDCElem:
            outTerm = elem.nD_sourceOut
            current = elem.get_DCsource()
            if outTerm[0] >= 0:
                self.sVec[outTerm[0]] -= current
            if outTerm[1] >= 0:
                self.sVec[outTerm[1]] += current
        for elem in self.ckt.nD_sourceTDElem:
            outTerm = elem.nD_sourceOut
            current = elem.get_TDsource(ctime)
            if outTerm[0] >= 0:
                self.sVec[outTerm[0]] -= current
            if outTerm[1] >= 0:
                self.sVec[outTerm[1]] += current
        return self.sVec

    def get_rhs(self, ctime):
        e = self.engine
        s = self.get_source(ctime)
        r = e._run_struct(L.OP_RHS)
        r.im = self._imcode()
        e.sv.copy_(e.lib.torch.as_tensor(s))
        e._launch(r)
        self.rhs = e.view('sv').cpu().numpy()
        return self.rhs

    def source_rows(self):
        """Rows of the source vector that time-dependent sources stamp."""
        rows = set()
        for elem in self.ckt.nD_sourceTDElem:
            rows.update(r for r in elem.nD_sourceOut if r >= 0)
        return np.array(sorted(rows), dtype=np.int32)

    def source_table(self, timeVec):
        """get_source(t) for every sample: [nsamples, n]."""
        return np.array([self.get_source(t).copy() for t in timeVec])

    def run(self, x0, timeVec, save_rows):
        """The time loop of tran.py:171-206 in one kernel launch.

        Returns (wave[nsamples, nsave], iters[nsamples], res[nsamples],
        status) -- status 0, or the first sample whose Newton solve failed.
        """
        h = timeVec[1] - timeVec[0] if len(timeVec) > 1 else 1.
        if self.engine is None:
            self._make_engine(h, x0)
        src = self.source_table(timeVec)
        wave, iters, res, status = self.engine.tran(
            x0, src, save_rows, im='trap' if self._imcode() else 'BE',
            var_rows=self.source_rows())
        return wave[0], iters[0], res[0], int(status[0])
