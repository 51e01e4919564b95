"""
Numpy emulation of the device-side gather assembly, level-scheduled Crout
factorisation and triangular solves, driven by the packed plan arrays of
``cdn_plan_create`` (csrc/symbolic.cpp).  Test infrastructure: lets the
host-side symbolic analysis be checked on the CPU against scipy, and
documents what csrc/engine.cuh does with each array.
"""
import ctypes as C

import numpy as np

from cardoon_b200 import _lib as L
from cardoon_b200.analyses import nodalCUDA as nd

_HOST = {}


def host_ctx():
    """(dll, ctx) of a GPU-less library context."""
    if 'ctx' not in _HOST:
        dll = L.load_dll()
        ctx = L._vp()
        assert dll.cdn_init_host(C.byref(ctx)) == 0
        _HOST['dll'], _HOST['ctx'] = dll, ctx
    return _HOST['dll'], _HOST['ctx']


NAMES = ['f_lvl_ptr', 'e_term_ptr', 'e_term_l', 'e_term_u', 'e_piv', 'e_g',
         'e_c', 'e_gones', 'e_con_ptr', 'e_con_src', 'e_con_kind',
         'l_lvl_ptr', 'l_rows', 'l_ptr', 'l_col', 'l_slot', 'u_lvl_ptr',
         'u_rows', 'u_ptr', 'u_col', 'u_slot', 'u_diag', 'rowmap', 'colmap',
         'r_lin_ptr', 'r_lin_col', 'r_lin_g', 'r_lin_c', 'r_lin_gones',
         'r_con_ptr', 'r_con_src', 'r_con_kind', 'dev_g', 'dev_i']


class HostPlan(object):
    def __init__(self, topo, use_q, dense, A_ref=None, ordering=None):
        dll, ctx = host_ctx()
        self.dll = dll
        self.topo = topo
        self.match = None
        if dense:
            self.plan, self.info = nd.create_plan(dll, ctx, topo, use_q, True)
        else:
            self.plan, self.info, self.match = nd.sparse_plan(
                dll, ctx, topo, use_q, A_ref, ordering)
        self.a = {k: nd.plan_array(dll, self.plan, k) for k in NAMES}
        self.n = self.info['n']
        self.nlu = self.info['nlu']
        self.dense = dense
        rows = np.empty(self.nlu, dtype=np.int32)
        cols = np.empty(self.nlu, dtype=np.int32)
        dll.cdn_plan_slot_coords(self.plan, rows.ctypes.data_as(L._vp),
                                 cols.ctypes.data_as(L._vp))
        self.slot_row, self.slot_col = rows, cols

    def close(self):
        self.dll.cdn_plan_free(self.plan)

    # ---- device buffers of one circuit instance -------------------------
    def pack_device(self, outs, jacs):
        """Flatten per-group out[nout, n] / jac[nout, nxin, n] the way the
        kernels lay them out (offsets follow group order)."""
        o = np.concatenate([x.ravel() for x in outs]) if outs else \
            np.zeros(0)
        j = np.concatenate([x.ravel() for x in jacs]) if jacs else \
            np.zeros(0)
        return o, j

    def assemble(self, jac, a0, use_q, gmin=0.):
        a = self.a
        lu = a['e_g'] + (a0 * a['e_c'] if use_q else 0.) \
            + gmin * a['e_gones']
        ptr, src, kind = a['e_con_ptr'], a['e_con_src'], a['e_con_kind']
        for e in range(self.nlu):
            for k in range(ptr[e], ptr[e + 1]):
                kd = kind[k]
                if abs(kd) == 1:
                    lu[e] += np.sign(kd) * jac[src[k]]
                elif use_q:
                    lu[e] += np.sign(kd) * a0 * jac[src[k]]
        return lu

    def residual(self, x, out, a0, use_q, gmin=0.):
        a = self.a
        iv = np.zeros(self.n)
        for r in range(self.n):
            acc = 0.
            for k in range(a['r_lin_ptr'][r], a['r_lin_ptr'][r + 1]):
                g = a['r_lin_g'][k] + gmin * a['r_lin_gones'][k]
                if use_q:
                    g += a0 * a['r_lin_c'][k]
                acc += g * x[a['r_lin_col'][k]]
            for k in range(a['r_con_ptr'][r], a['r_con_ptr'][r + 1]):
                kd = a['r_con_kind'][k]
                v = out[a['r_con_src'][k]]
                if abs(kd) == 1:
                    acc += np.sign(kd) * v
                elif use_q:
                    acc += np.sign(kd) * a0 * v
            iv[r] = acc
        return iv

    def matrix(self, lu):
        """Assembled slots as a dense matrix in original numbering."""
        M = np.zeros((self.n, self.n))
        np.add.at(M, (self.slot_row, self.slot_col), lu)
        return M

    def factor(self, lu):
        a = self.a
        lu = lu.copy()
        lp = a['f_lvl_ptr']
        for l in range(len(lp) - 1):
            new = {}
            for e in range(lp[l], lp[l + 1]):
                v = lu[e]
                for t in range(a['e_term_ptr'][e], a['e_term_ptr'][e + 1]):
                    tl, tu = a['e_term_l'][t], a['e_term_u'][t]
                    # operands must come from earlier levels
                    assert tl < lp[l] and tu < lp[l]
                    v -= lu[tl] * lu[tu]
                pv = a['e_piv'][e]
                if pv >= 0:
                    assert pv < lp[l]
                    v /= lu[pv]
                new[e] = v
            for e, v in new.items():
                lu[e] = v
        return lu

    def solve(self, lu, y):
        a = self.a
        n = self.n
        tmp = np.zeros(n)
        dx = np.zeros(n)
        lp = a['l_lvl_ptr']
        done = np.zeros(n, dtype=bool)
        for l in range(len(lp) - 1):
            rows = a['l_rows'][lp[l]:lp[l + 1]]
            for i in rows:
                v = y[a['rowmap'][i]]
                for k in range(a['l_ptr'][i], a['l_ptr'][i + 1]):
                    assert done[a['l_col'][k]]
                    v -= lu[a['l_slot'][k]] * tmp[a['l_col'][k]]
                tmp[i] = v
            done[rows] = True
        up = a['u_lvl_ptr']
        done[:] = False
        for l in range(len(up) - 1):
            rows = a['u_rows'][up[l]:up[l + 1]]
            for i in rows:
                v = tmp[i]
                for k in range(a['u_ptr'][i], a['u_ptr'][i + 1]):
                    assert done[a['u_col'][k]]
                    v -= lu[a['u_slot'][k]] * tmp[a['u_col'][k]]
                v /= lu[a['u_diag'][i]]
                tmp[i] = v
                dx[a['colmap'][i]] = v
            done[rows] = True
        return dx


# ---------------------------------------------------------------------------
# small dense systems on 8-lane tiles (csrc/engine.cuh sd8_*)
# ---------------------------------------------------------------------------
def sd8_program(hp):
    """(sd_ptr, sd_prog) of a dense plan with n <= 8, or None."""
    ptr = nd.plan_array(hp.dll, hp.plan, 'sd_ptr')
    prog = nd.plan_array(hp.dll, hp.plan, 'sd_prog')
    return (ptr, prog) if ptr[-1] > 0 else None


def sd8_gather(hp, jac_scaled):
    """w.lu after sd8_gather: per-slot sums of the device contributions
    (``jac_scaled``: charge rows already times a0, as WsSink stores them).
    Emulates the instance state the terms address: 10 n-vectors, lu, device
    outputs, device Jacobians, factor storage (engine.cuh ws_carve)."""
    ptr, prog = sd8_program(hp)
    fix = nd.plan_array(hp.dll, hp.plan, 'sd_fix')
    n, nlu = hp.n, hp.nlu
    lu_off = 10 * n
    jac_off = lu_off + nlu + hp.info['dev_out']
    fact_off = jac_off + hp.info['dev_jac']
    PART, NFIX = 77, 10
    state = np.zeros(fact_off + PART + NFIX + 1)
    state[jac_off:jac_off + len(jac_scaled)] = jac_scaled
    maxlen = ptr[1] - ptr[0]
    assert np.all(np.diff(ptr) == maxlen)
    for lane in range(8):
        acc = 0.
        for t in range(ptr[lane], ptr[lane + 1]):
            pk = int(prog[t])
            v = state[pk & 0x3fff]
            acc += -v if pk & 0x80000000 else v
            if pk & (1 << 30):
                state[(pk >> 14) & 0x3fff] = acc
                acc = 0.
        assert acc == 0.          # every piece ends with a store (or padding)
    lu = state[lu_off:lu_off + nlu].copy()
    for j, e in enumerate(fix):
        lu[e] += state[fact_off + PART + j]
    return lu


def sd8_solve(J, y):
    """sd8_factor + sd8_backward + sd8_unknown on rows J[r] / rhs y (n <= 8):
    lane-parallel elimination without row moves, emulated lane by lane.
    Returns (x, factors) where factors = (rows, pinv, step, perm)."""
    n = len(y)
    row = np.zeros((8, 8))
    row[:n, :n] = J
    yy = np.zeros(8)
    yy[:n] = y
    step = np.full(8, 8)
    pinv = np.zeros(8)
    perm = []
    for k in range(n):
        v = np.abs(row[:, k])
        v = np.where(np.isnan(v), -1., v)
        cand = (np.arange(8) < n) & (step == 8)
        v = np.where(cand, v, -2.)
        p = int(np.argmax(v))                 # first maximum: lowest lane
        perm.append(p)
        pv = row[p, k]
        with np.errstate(divide='ignore'):
            pi = 1. / pv
        m = row[:, k] * (pi if pv != 0. else 1.)
        elim = cand.copy()
        elim[p] = False
        step[p] = k
        pinv[p] = pi
        pr = row[p].copy()
        py = yy[p]
        for l in range(8):
            if elim[l]:
                row[l, k] = m[l]
                row[l, k + 1:n] = row[l, k + 1:n] - m[l] * pr[k + 1:n]
                yy[l] = yy[l] - m[l] * py
    fact = (row.copy(), pinv.copy(), step.copy(), list(perm))
    x = sd8_backward(fact, yy, n)
    return x, fact


def sd8_backward(fact, yy, n):
    row, pinv, step, perm = fact
    yy = yy.copy()
    for k in range(n - 1, -1, -1):
        p = perm[k]
        xk = yy[p] * pinv[p]
        for l in range(8):
            if step[l] < k:
                yy[l] -= row[l, k] * xk
        yy[p] = xk
    return np.array([yy[perm[l]] for l in range(n)])


def sd8_resolve(fact, y):
    """sd8_forward + sd8_backward with stored factors (the chord step)."""
    row, pinv, step, perm = fact
    n = len(y)
    yy = np.zeros(8)
    yy[:n] = y
    for k in range(n):
        yk = yy[perm[k]]
        for l in range(8):
            if k < step[l] < 8:
                yy[l] -= row[l, k] * yk
    return sd8_backward(fact, yy, n)
