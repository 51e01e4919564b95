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
