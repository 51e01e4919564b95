"""
Host analysis for the *boundary / interior* transient engine (csrc/schur.cuh).

The sparse engine of the reference refactors the whole Jacobian on every
Newton iteration (SuperLU, nodalSP.py:281-303, :838-887).  With a fixed time
step only the entries stamped by nonlinear devices change: J = A + D(x) with
A = G + a0 C constant (nodalSP.py:603-612 builds it once per step size) and D
confined to the rows/columns of the device ports (nodal.py:66-80).  This
module splits the unknowns into

  * a small *boundary* set B = device port unknowns + a few cut nodes that
    separate the rest into independent pieces, and
  * the *interior* I, dealt to the CTAs of one thread-block cluster piece by
    piece,

and precomputes, per CTA c with interior rows Ic, the constant operators

  Linv_c, Uinv_c   explicit inverses of the LU factors of A[Ic, Ic]
                   (so that a solve is two sparse matrix-vector products
                   instead of level-scheduled triangular sweeps),
  X_c   = A[Ic, Ic]^-1 A[Ic, B],   Asi_c = A[B, Ic],   Csi_c = C[B, Ic],
  A_c, C_c         rows Ic of A and C (residual and charge),

and the replicated boundary data: A[B, B], C[B, B], the Schur complement
T0 = A[B, B] - sum_c Asi_c X_c in band storage, and gather lists of the device
outputs / Jacobians into the boundary rows and the band of T = T0 + D(x).

One Newton step  J dx = y  is then
  t_I = Uinv (Linv y_I);  r_B = y_B - A[B, I] t_I;  T dx_B = r_B (banded LU,
  no pivoting);  dx_I = t_I - X dx_B,
with one exchange of boundary partial sums between the CTAs.

Entries of T0 below 2**-60 of both diagonals they couple are dropped before
the bandwidth is measured: they cannot change a double-precision pivot.  A
circuit whose boundary system is not narrow-banded after that is not eligible
(`analyse` returns None) and runs on the general kernels.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.sparse.csgraph import (breadth_first_order, connected_components,
                                  reverse_cuthill_mckee)

LONG_ROW = 32              # rows of Linv longer than this are summed by a warp
MAX_BOUNDARY = 96          # unknowns of the replicated boundary system
MAX_BAND = 4               # half bandwidth of T the kernel handles
DROP = 2. ** -60


def _levels(adj, nodes, start):
    """BFS level of every node of the sub-graph ``nodes`` from ``start``."""
    sub = adj[nodes][:, nodes]
    order, pred = breadth_first_order(sub, start, directed=False)
    lev = np.full(len(nodes), -1)
    lev[start] = 0
    for v in order[1:]:
        lev[v] = lev[pred[v]] + 1
    return lev, order


def _bisect(adj, nodes):
    """Vertex separator of the sub-graph ``nodes`` (a connected piece) from
    its BFS level structure: (part1, part2, separator) as node arrays."""
    if len(nodes) < 3:
        return nodes, nodes[:0], nodes[:0]
    lev, order = _levels(adj, nodes, 0)
    lev, order = _levels(adj, nodes, order[-1])      # pseudo-peripheral start
    if lev.max() < 2:
        return nodes, nodes[:0], nodes[:0]
    cnt = np.bincount(lev[lev >= 0])
    half = np.searchsorted(np.cumsum(cnt), len(nodes) / 2.)
    half = min(max(half, 1), len(cnt) - 2)
    unreached = lev < 0
    return (nodes[(lev < half) & ~unreached],
            nodes[(lev > half) | unreached], nodes[lev == half])


def _nd_order(adj, nodes, leaf=4):
    """Nested-dissection elimination order of the sub-graph ``nodes``."""
    if len(nodes) <= leaf:
        return list(nodes)
    ncomp, lab = connected_components(adj[nodes][:, nodes], directed=False)
    if ncomp > 1:
        out = []
        for k in range(ncomp):
            out += _nd_order(adj, nodes[lab == k], leaf)
        return out
    p1, p2, s = _bisect(adj, nodes)
    if len(s) == 0:
        return list(nodes)
    return _nd_order(adj, p1, leaf) + _nd_order(adj, p2, leaf) + list(s)


def _cut_nodes(adj, interior, parts):
    """Separators that cut the interior graph into at least ``parts`` pieces
    of comparable size."""
    pieces = []
    ncomp, lab = connected_components(adj[interior][:, interior],
                                      directed=False)
    for k in range(ncomp):
        pieces.append(interior[lab == k])
    cuts = []
    target = max(8, int(1.15 * len(interior) / parts))
    for _ in range(4 * parts):
        pieces.sort(key=len)
        big = pieces[-1]
        if len(pieces) >= parts and len(big) <= target:
            break
        p1, p2, s = _bisect(adj, big)
        if len(s) == 0:
            break
        pieces.pop()
        cuts += list(s)
        for p in (p1, p2):
            if len(p):
                nc, lb = connected_components(adj[p][:, p], directed=False)
                for k in range(nc):
                    pieces.append(p[lb == k])
    return np.array(sorted(cuts), dtype=np.int64), pieces


class Csr(object):
    """CSR operator with 16-bit column indices (what the kernel stages)."""

    def __init__(self, M, drop_zeros=True):
        M = sp.csr_matrix(M)
        if drop_zeros:
            M.eliminate_zeros()
        M.sort_indices()
        self.shape = M.shape
        self.ptr = M.indptr.astype(np.int32)
        self.col = M.indices.astype(np.uint16)
        self.val = M.data.astype(np.float64)
        assert M.shape[1] < 65536

    @property
    def nnz(self):
        return len(self.val)

    def dot(self, v):
        out = np.zeros(self.shape[0])
        for r in range(self.shape[0]):
            a, b = self.ptr[r], self.ptr[r + 1]
            out[r] = np.dot(self.val[a:b], v[self.col[a:b]])
        return out

    def tosp(self):
        return sp.csr_matrix((self.val, self.col.astype(np.int32), self.ptr),
                             shape=self.shape)


class Ell(object):
    """Sliced ELLPACK copy of the short rows of a CSR operator, for the
    one-thread-per-row products of the kernel: rows sorted by length, slices
    of 32 rows, every slice stored column-major (entry j of lane l at
    sbase[s] + 32 j + l) and padded to a multiple of four entries per row with
    (column 0, value 0).  A warp then reads consecutive shared-memory words
    (no bank conflicts on the value and column streams) and needs no
    predicates.  ``rowid[slot]`` = row a slot computes (-1: padding slot);
    rows longer than ``long_row`` are left to the caller (a warp sums each)."""

    def __init__(self, op, long_row=None):
        ln = np.diff(op.ptr)
        rows = np.arange(op.shape[0])
        if long_row is not None:
            rows = rows[ln <= long_row]
        order = rows[np.argsort(-ln[rows], kind='stable')]
        nsl = (len(order) + 31) // 32
        self.nslots = 32 * nsl
        self.rowid = np.full(self.nslots, -1, dtype=np.int16)
        self.rowid[:len(order)] = order
        self.sbase = np.zeros(nsl + 1, dtype=np.int32)
        vals, cols = [], []
        for s_ in range(nsl):
            rs = order[32 * s_:32 * s_ + 32]
            w = int(ln[rs].max()) if len(rs) else 0
            w = (w + 3) // 4 * 4
            v = np.zeros((w, 32))
            c = np.zeros((w, 32), dtype=np.uint16)
            for lane, r in enumerate(rs):
                a, b = op.ptr[r], op.ptr[r + 1]
                v[:b - a, lane] = op.val[a:b]
                c[:b - a, lane] = op.col[a:b]
            vals.append(v.ravel())
            cols.append(c.ravel())
            self.sbase[s_ + 1] = self.sbase[s_] + 32 * w
        self.val = np.concatenate(vals) if vals else np.zeros(0)
        self.col = np.concatenate(cols) if cols else np.zeros(0, np.uint16)
        self.shape = op.shape

    def dot(self, v, out=None):
        """Product of the rows held here (numpy statement of the kernel's
        loop); rows not held keep the value of ``out``."""
        out = np.zeros(self.shape[0]) if out is None else out
        for slot in range(self.nslots):
            r = self.rowid[slot]
            if r < 0:
                continue
            s_, lane = slot >> 5, slot & 31
            base = self.sbase[s_]
            w = (self.sbase[s_ + 1] - base) >> 5
            idx = base + 32 * np.arange(w) + lane
            out[r] = np.dot(self.val[idx], v[self.col[idx]])
        return out


def device_port_unknowns(topo):
    parts = [np.asarray(g[k]).ravel() for g in topo.groups
             for k in ('vpos', 'vneg', 'cpos', 'cneg', 'qpos', 'qneg')]
    if not parts:
        return np.zeros(0, dtype=np.int64)
    a = np.unique(np.concatenate(parts))
    return a[a >= 0].astype(np.int64)


def device_gathers(topo, bpos, out_off, jac_off):
    """Gather lists of the device buffers into the boundary system.

    ``bpos[u]`` = position of unknown u in B (-1: interior); ``out_off`` /
    ``jac_off``: offsets of every group inside the device-output / Jacobian
    buffers (layout of csrc/engine.cuh WsSink: out[slot * n + i],
    jac[(slot * nxin + c) * n + i]).  Returns COO lists
    (row, out index, sign) for currents and for charges, and
    (row, col, jac index, sign) for the Jacobian (charge rows of the
    Jacobian buffer arrive multiplied by a0)."""
    cur, chg, jac = [], [], []
    for g, oo, jo in zip(topo.groups, out_off, jac_off):
        n, ncs, nqs, nxin = g['n'], g['ncs'], g['nqs'], g['nxin']
        for o in range(ncs + nqs):
            rp = g['cpos'][o] if o < ncs else g['qpos'][o - ncs]
            rn = g['cneg'][o] if o < ncs else g['qneg'][o - ncs]
            dst = cur if o < ncs else chg
            for i in range(n):
                for r, sg in ((rp[i], 1.), (rn[i], -1.)):
                    if r >= 0:
                        dst.append((bpos[r], oo + o * n + i, sg))
                for k in range(nxin):
                    cp, cn = g['vpos'][k][i], g['vneg'][k][i]
                    src = jo + (o * nxin + k) * n + i
                    for r, c, sg in ((rp[i], cp, 1.), (rn[i], cn, 1.),
                                     (rp[i], cn, -1.), (rn[i], cp, -1.)):
                        if r >= 0 and c >= 0:
                            jac.append((bpos[r], bpos[c], src, sg))
    return cur, chg, jac


class SchurPlan(object):
    pass


def analyse(topo, a0, ncta=8, stamps=None, out_off=None, jac_off=None,
            max_band=MAX_BAND):
    """Boundary/interior plan of the transient Jacobian of ``topo`` with
    integration coefficient ``a0``, for a cluster of ``ncta`` CTAs.
    ``stamps``: nodal stamps (rows, cols, dI/dv, dQ/dv) of the device
    Jacobians at a reference point (Topology.device_stamps); they only serve
    to scale the drop test of the boundary system.  Returns None when the
    circuit is not eligible."""
    n = topo.n
    if any(g['ndelay'] for g in topo.groups) or n >= 60000:
        return None
    if sum(g['n'] for g in topo.groups) > 4 * MAX_BOUNDARY:
        return None                    # (cannot have few port unknowns)
    G = sp.coo_matrix((topo.G[2], (topo.G[0], topo.G[1])), shape=(n, n)).tocsr()
    Cm = sp.coo_matrix((topo.C[2], (topo.C[0], topo.C[1])), shape=(n, n)).tocsr()
    A = (G + a0 * Cm).tocsr()
    adj = (abs(A) + abs(A.T)).tocsr()
    S = device_port_unknowns(topo)
    if len(S) == 0 or len(S) > MAX_BOUNDARY:
        return None
    # operators of a piece must fit the shared memory of one CTA (>= ~200 bytes
    # per interior row): do not even try pieces of thousands of rows
    if (n - len(S)) > 1200 * ncta:
        return None
    mask = np.ones(n, bool)
    mask[S] = False
    interior = np.nonzero(mask)[0]
    cuts, pieces = _cut_nodes(adj, interior, ncta)
    B = np.array(sorted(set(S) | set(cuts)), dtype=np.int64)
    if len(B) > MAX_BOUNDARY:
        return None
    # ---- deal the pieces to the CTAs (largest first, least loaded CTA) ---------
    pieces.sort(key=len, reverse=True)
    load = [0] * ncta
    owned = [[] for _ in range(ncta)]
    for p in pieces:
        c = int(np.argmin(load))
        load[c] += len(p)
        owned[c].append(p)
    # ---- boundary order: reverse Cuthill-McKee of the pattern of T0 -----------
    ctas = []
    ABB = A[B][:, B].toarray()
    T0 = ABB.copy()
    for c in range(ncta):
        Ic = []
        for p in owned[c]:
            Ic += _nd_order(adj, np.asarray(p))
        Ic = np.array(Ic, dtype=np.int64)
        m = len(Ic)
        d = dict(Ic=Ic, m=m)
        if m:
            Acc = A[Ic][:, Ic].tocsc()
            try:
                lu = spla.splu(Acc, permc_spec='NATURAL', diag_pivot_thresh=0.01,
                               options=dict(SymmetricMode=False))
            except RuntimeError:
                return None
            eye = np.eye(m)
            Linv = spla.spsolve_triangular(lu.L.tocsr(), eye, lower=True)
            Uinv = spla.spsolve_triangular(lu.U.tocsr(), eye, lower=False)
            # Pr A Pc = L U  =>  A^-1 = Pc U^-1 L^-1 Pr
            Pr = sp.csc_matrix((np.ones(m), (lu.perm_r, np.arange(m))), shape=(m, m))
            Pc = sp.csc_matrix((np.ones(m), (np.arange(m), lu.perm_c)), shape=(m, m))
            d['Linv'] = sp.csr_matrix(Linv) @ Pr
            d['Uinv'] = Pc @ sp.csr_matrix(Uinv)
            AiB = A[Ic][:, B].toarray()
            Xc = d['Uinv'] @ (d['Linv'] @ AiB)
            Xc[AiB.any(axis=0)[None, :].repeat(m, 0) == 0] = 0.
            d['X'] = Xc
            d['Asi'] = A[B][:, Ic]
            d['Csi'] = Cm[B][:, Ic]
            T0 -= d['Asi'] @ Xc
        ctas.append(d)
    # the boundary matrix at the reference point scales the drop test (T0
    # alone may have zero diagonals: gyrator equations of device branches)
    Tref = T0.copy()
    if stamps is not None:
        bp = np.full(n, -1, dtype=np.int64)
        bp[B] = np.arange(len(B))
        sr, sc, si, sq = stamps
        np.add.at(Tref, (bp[sr], bp[sc]), si + a0 * sq)
    dg = np.abs(np.diag(Tref))
    if not (np.all(np.isfinite(Tref)) and np.all(dg > 0.)):
        return None
    keep = np.abs(T0) > DROP * np.minimum(dg[:, None], dg[None, :])
    keep |= np.eye(len(B), dtype=bool)
    perm = np.asarray(reverse_cuthill_mckee(sp.csr_matrix(keep | keep.T),
                                            symmetric_mode=True))
    kp = keep[perm][:, perm]
    ii, jj = np.nonzero(kp)
    band = int(np.abs(ii - jj).max()) if len(ii) else 0
    if band > max_band:
        return None
    # ---- renumber the boundary -------------------------------------------------
    B = B[perm]
    nB = len(B)
    bpos = np.full(n, -1, dtype=np.int64)
    bpos[B] = np.arange(nB)
    T0 = T0[perm][:, perm] * kp
    P = SchurPlan()
    P.n, P.nB, P.band, P.ncta, P.a0 = n, nB, band, ncta, float(a0)
    P.B, P.bpos = B, bpos
    P.T0 = T0
    P.Abb = Csr(A[B][:, B])
    P.Cbb = Csr(Cm[B][:, B])
    P.ctas = []
    for d in ctas:
        m, Ic = d['m'], d['Ic']
        c = SchurPlan()
        c.m, c.Ic = m, Ic
        # local vector space of a CTA: [x_B (nB) | x_Ic (m)]
        loc = np.concatenate([B, Ic])
        if m:
            c.Linv, c.Uinv = Csr(d['Linv']), Csr(d['Uinv'])
            c.X = Csr(d['X'][:, perm])
            c.Asi = Csr(d['Asi'][perm])
            c.Csi = Csr(d['Csi'][perm])
            c.Arow = Csr(A[Ic][:, loc])
            c.Crow = Csr(Cm[Ic][:, loc])
        else:
            z = sp.csr_matrix((0, 0))
            c.Linv = c.Uinv = Csr(z)
            c.X = Csr(sp.csr_matrix((0, nB)))
            c.Asi = c.Csi = Csr(sp.csr_matrix((nB, 0)))
            c.Arow = c.Crow = Csr(sp.csr_matrix((0, nB)))
        # what the kernel's row products read (the CSR copies stay for the
        # emulation and the long rows of Linv)
        c.eLinv = Ell(c.Linv, LONG_ROW)
        c.eUinv, c.eX = Ell(c.Uinv), Ell(c.X)
        c.eArow = Ell(c.Arow)
        P.ctas.append(c)
    P.has_csi = any(c.Csi.nnz for c in P.ctas)
    # ---- device gathers ------------------------------------------------------------
    if out_off is None:
        out_off, jac_off, o, j = [], [], 0, 0
        for g in topo.groups:
            out_off.append(o), jac_off.append(j)
            o += (g['ncs'] + g['nqs']) * g['n']
            j += (g['ncs'] + g['nqs']) * g['nxin'] * g['n']
        P.out_size, P.jac_size = o, j
    cur, chg, jac = device_gathers(topo, bpos, out_off, jac_off)
    P.out_size = sum((g['ncs'] + g['nqs']) * g['n'] for g in topo.groups)
    P.jac_size = sum((g['ncs'] + g['nqs']) * g['nxin'] * g['n']
                     for g in topo.groups)
    if P.out_size >= 65536 or P.jac_size >= 65536:
        return None

    def coo(lst, nrow, ncol):
        if not lst:
            return Csr(sp.csr_matrix((nrow, ncol)))
        r, c, v = zip(*lst)
        return Csr(sp.coo_matrix((v, (r, c)), shape=(nrow, ncol)), False)
    P.Dcur = coo(cur, nB, P.out_size)
    P.Dchg = coo(chg, nB, P.out_size)
    # Jacobian contributions must stay inside the band
    W = 2 * band + 1
    tj = []
    for r, c_, src, sg in jac:
        if abs(r - c_) > band:
            return None
        tj.append((r * W + (c_ - r + band), src, sg))
    P.Tjac = coo(tj, nB * W, P.jac_size)
    # band storage of T0: Tb[r, band + (c - r)]
    Tb = np.zeros((nB, W))
    for r in range(nB):
        for c_ in range(max(0, r - band), min(nB, r + band + 1)):
            Tb[r, c_ - r + band] = T0[r, c_]
    P.Tb = Tb
    return P


# ---------------------------------------------------------------------------
# device packaging (layout: csrc/schur.cuh, enums SOP_* / SH_*)
# ---------------------------------------------------------------------------
NOPS = 12
NELL = 4
SH_OPS, SH_WORDS = 8, 96
SH_TB = SH_OPS + 4 * NOPS
SH_ELL = 64


def pack_blob(P, c):
    """Operator blob of CTA ``c``: 64 header words, then per operator row
    pointers (int32), columns (uint16) and values (float64), the band of T0 and
    the unknown numbers of the interior rows and of the boundary."""
    # CSR part: the long rows of Linv (by decreasing length, so that dealing them
    # to the warps round-robin balances), the boundary operators
    ln = np.diff(c.Linv.ptr)
    llong = np.nonzero(ln > LONG_ROW)[0]
    llong = llong[np.argsort(-ln[llong], kind='stable')].astype(np.int32)
    Llong = c.Linv.tosp()[llong] if len(llong) else sp.csr_matrix((0, max(1, c.m)))
    empty = Csr(sp.csr_matrix((0, 1)))
    ops = [Csr(Llong, False), empty, empty, c.Asi, c.Csi, empty, c.Crow, P.Abb,
           P.Cbb, P.Dcur, P.Dchg, P.Tjac]
    hdr = np.zeros(SH_WORDS, dtype=np.int32)
    hdr[0:6] = (c.m, P.nB, P.band, P.out_size, P.jac_size, int(P.has_csi))
    parts = [None]
    pos = [SH_WORDS * 4]

    def put(arr):
        b = np.ascontiguousarray(arr).tobytes()
        pad = (-len(b)) % 8
        off = pos[0]
        parts.append(b + b'\0' * pad)
        pos[0] += len(b) + pad
        return off
    for k, op in enumerate(ops):
        hdr[SH_OPS + 4 * k] = op.shape[0]
        hdr[SH_OPS + 4 * k + 1] = put(op.ptr.astype(np.int32))
        hdr[SH_OPS + 4 * k + 2] = put(op.col.astype(np.uint16))
        hdr[SH_OPS + 4 * k + 3] = put(op.val.astype(np.float64))
    hdr[SH_TB] = put(P.Tb.astype(np.float64).ravel())
    hdr[SH_TB + 1] = put(c.Ic.astype(np.int32))
    hdr[SH_TB + 2] = put(P.B.astype(np.int32))
    # rows of Linv a warp sums together: row k of the CSR operator SOP_LINV is
    # row llong[k] of the piece
    hdr[6] = put(llong)
    hdr[7] = len(llong)
    for k, e in enumerate((c.eLinv, c.eUinv, c.eX, c.eArow)):
        w0 = SH_ELL + 5 * k
        hdr[w0] = e.nslots
        hdr[w0 + 1] = put(e.sbase.astype(np.int32))
        hdr[w0 + 2] = put(e.col.astype(np.uint16))
        hdr[w0 + 3] = put(e.val.astype(np.float64))
        hdr[w0 + 4] = put(e.rowid.astype(np.int16))
    pad = (-pos[0]) % 16
    parts.append(b'\0' * pad)
    pos[0] += pad
    hdr[SH_TB + 3] = pos[0]
    parts[0] = hdr.tobytes()
    return b''.join(parts)


def pack(P):
    """All CTA blobs back to back: (bytes, offsets[ncta + 1])."""
    blobs = [pack_blob(P, c) for c in P.ctas]
    off = np.zeros(len(blobs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in blobs])
    return b''.join(blobs), off


def locate(P, rows):
    """Owner CTA (-1: boundary) and position in the owner's local vector
    [x_B | x_Ic] of the unknowns ``rows``."""
    own = np.full(P.n, -1, dtype=np.int32)
    loc = np.zeros(P.n, dtype=np.int32)
    loc[P.B] = np.arange(P.nB)
    for k, c in enumerate(P.ctas):
        own[c.Ic] = k
        loc[c.Ic] = P.nB + np.arange(c.m)
    rows = np.asarray(rows, dtype=np.int64)
    return own[rows], loc[rows]


def save_lists(P, save_rows):
    """Per CTA the columns of ``wave`` it writes and where it finds them
    (boundary unknowns are written by CTA 0)."""
    own, loc = locate(P, save_rows)
    own = np.where(own < 0, 0, own)
    order = np.argsort(own, kind='stable')
    ptr = np.zeros(P.ncta + 1, dtype=np.int32)
    ptr[1:] = np.cumsum(np.bincount(own, minlength=P.ncta))
    return ptr, order.astype(np.int32), loc[order].astype(np.int32)


def remap_ports(P, g):
    """Control-port rows of group ``g`` as boundary positions."""
    def m(a):
        a = np.asarray(a, dtype=np.int64)
        return np.ascontiguousarray(
            np.where(a >= 0, P.bpos[np.maximum(a, 0)], -1).astype(np.int32))
    return m(g['vpos']), m(g['vneg'])
