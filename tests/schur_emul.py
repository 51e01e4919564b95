"""
Numpy emulation of the boundary/interior transient kernel (csrc/schur.cuh)
driven by the host analysis of cardoon_b200/analyses/schur.py.  Test
infrastructure: checks the analysis on the CPU against the oracle's time loop
and documents, phase by phase, what the kernel does with every operator.
"""
import numpy as np


def band_factor(Tb, band):
    """In-place LU without pivoting of a band matrix stored as
    Tb[r, band + (c - r)] (multipliers below the diagonal)."""
    nB = Tb.shape[0]
    for k in range(nB):
        piv = Tb[k, band]
        for i in range(k + 1, min(nB, k + band + 1)):
            m = Tb[i, band + (k - i)] / piv
            Tb[i, band + (k - i)] = m
            for j in range(k + 1, min(nB, k + band + 1)):
                Tb[i, band + (j - i)] -= m * Tb[k, band + (j - k)]
    return Tb


def band_solve(Tb, band, r):
    nB = Tb.shape[0]
    y = r.copy()
    for i in range(nB):
        for k in range(max(0, i - band), i):
            y[i] -= Tb[i, band + (k - i)] * y[k]
    for i in range(nB - 1, -1, -1):
        for j in range(i + 1, min(nB, i + band + 1)):
            y[i] -= Tb[i, band + (j - i)] * y[j]
        y[i] /= Tb[i, band]
    return y


class SchurEmul(object):
    def __init__(self, P, dev_eval, abstol, reltol, maxdelta, maxiter,
                 trap=True):
        """``dev_eval(x_full, deriv)`` -> (out, jac) flattened in the kernels'
        layout, the charge rows of jac NOT yet multiplied by a0."""
        self.P = P
        self.dev_eval = dev_eval
        self.abstol, self.reltol = abstol, reltol
        self.maxdelta, self.maxiter = maxdelta, maxiter
        self.trap = trap

    # -- helpers ---------------------------------------------------------------
    def _full_x(self):
        P = self.P
        x = np.zeros(P.n)
        x[P.B] = self.xB
        for c, xi in zip(P.ctas, self.xI):
            x[c.Ic] = xi
        return x

    def _eval(self, deriv):
        self.out, _ = self.dev_eval(self._full_x(), deriv)

    def _charge(self):
        P = self.P
        qB = P.Cbb.dot(self.xB) + P.Dchg.dot(self.out)
        qI = []
        for c, xi in zip(P.ctas, self.xI):
            loc = np.concatenate([self.xB, xi])
            qI.append(c.Crow.dot(loc))
            qB = qB + c.Csi.dot(xi)            # (exchange of partial sums)
        return qB, qI

    def _solve(self, yB, yI, factor):
        """dx of J dx = y.  ``yB`` already holds y_B; interior parts local."""
        P = self.P
        tI = [c.Uinv.dot(c.Linv.dot(y)) for c, y in zip(P.ctas, yI)]
        rB = yB.copy()
        for c, t in zip(P.ctas, tI):
            rB -= c.Asi.dot(t)                 # (exchange of partial sums)
        if factor:
            W = 2 * P.band + 1
            Tb = P.Tb + P.Tjac.dot(self.jac_a0).reshape(P.nB, W)
            self.Tf = band_factor(Tb, P.band)
        dB = band_solve(self.Tf, P.band, rB)
        dI = [t - c.X.dot(dB) for c, t in zip(P.ctas, tI)]
        return dB, dI

    def run(self, x0, src, qscale_jac):
        """Time loop (tran.py:171-206).  ``src``: [nsamples, n] source
        vectors; ``qscale_jac(jac, a0)`` multiplies the charge rows of the
        Jacobian buffer by a0.  Returns X[nsamples, n], iters, res."""
        P = self.P
        a0 = P.a0
        ns = src.shape[0]
        self.xB = x0[P.B].copy()
        self.xI = [x0[c.Ic].copy() for c in P.ctas]
        X = np.zeros((ns, P.n))
        X[0] = x0
        iters = np.zeros(ns, dtype=int)
        resv = np.zeros(ns)
        # set_IC
        self._eval(False)
        qB, qI = self._charge()
        qpB, qpI = qB.copy(), [q.copy() for q in qI]
        dqB, dqI = np.zeros_like(qB), [np.zeros_like(q) for q in qI]
        ivB = np.zeros(P.nB)
        ivI = [np.zeros(c.m) for c in P.ctas]
        for i in range(1, ns):
            # accept
            self._eval(False)
            qB, qI = self._charge()
            if self.trap:
                dqB = -dqB + a0 * (qB - qpB)
                dqI = [-d + a0 * (q - qp) for d, q, qp in zip(dqI, qI, qpI)]
            qpB, qpI = qB, qI
            svB = a0 * qpB + (dqB if self.trap else 0.) + src[i][P.B]
            svI = [a0 * qp + (d if self.trap else 0.) + src[i][c.Ic]
                   for qp, d, c in zip(qpI, dqI, P.ctas)]
            if i > 1:
                dB, dI = self._solve(svB - ivB, [s - v for s, v in zip(svI, ivI)],
                                     False)
                self.xB += dB
                self.xI = [x + d for x, d in zip(self.xI, dI)]
            ok = False
            it = 0
            while it < self.maxiter:
                out, jac = self.dev_eval(self._full_x(), True)
                self.out = out
                self.jac_a0 = qscale_jac(jac, a0)
                ivI = []
                yI = []
                ivB = P.Abb.dot(self.xB) + P.Dcur.dot(out) + a0 * P.Dchg.dot(out)
                for c, xi, s in zip(P.ctas, self.xI, svI):
                    loc = np.concatenate([self.xB, xi])
                    v = c.Arow.dot(loc)
                    ivI.append(v)
                    yI.append(s - v)
                    ivB = ivB + c.Asi.dot(xi)          # (exchange)
                dB, dI = self._solve(svB - ivB, yI, True)
                m = max([np.abs(dB).max()] + [np.abs(d).max() for d in dI if len(d)])
                scale = self.maxdelta / m if m > self.maxdelta else 1.
                conv = True
                newI = []
                d = dB * scale
                xn = self.xB + d
                conv &= bool(np.all(np.abs(d) < self.reltol * np.maximum(
                    np.abs(self.xB), np.abs(xn)) + self.abstol))
                self.xB = xn
                for xi, di in zip(self.xI, dI):
                    d = di * scale
                    xn = xi + d
                    conv &= bool(np.all(np.abs(d) < self.reltol * np.maximum(
                        np.abs(xi), np.abs(xn)) + self.abstol))
                    newI.append(xn)
                self.xI = newI
                res = m * scale
                it += 1
                if conv:
                    ok = True
                    break
            iters[i], resv[i] = it, res
            if not ok:
                return X[:i], iters[:i], resv[:i]
            X[i] = self._full_x()
        return X, iters, resv



class SchurEmulSlaved(SchurEmul):
    """The same time loop with the interior eliminated analytically: the
    interior equations are linear, so with z_I = A_II^-1 sv_I (ONE interior solve
    per time step) every solve of that step needs only
        t_I = z_I - (x_I + X x_B)          (interior part of J^-1 y)
        r_B = (sv_B - A_BI z_I) - T0 x_B - dev(x_B)
    and the chord predictor the stored x'_B, dev(x'_B) and w'_I = x'_I + X x'_B
    of the last residual evaluation.  One exchange of boundary partial sums per
    time step instead of one per solve."""

    def run(self, x0, src, qscale_jac):
        P = self.P
        a0 = P.a0
        W = 2 * P.band + 1
        ns = src.shape[0]
        self.xB = x0[P.B].copy()
        self.xI = [x0[c.Ic].copy() for c in P.ctas]
        X = np.zeros((ns, P.n))
        X[0] = x0
        iters = np.zeros(ns, dtype=int)
        resv = np.zeros(ns)
        T0 = np.zeros((P.nB, P.nB))
        for r in range(P.nB):
            for c_ in range(max(0, r - P.band), min(P.nB, r + P.band + 1)):
                T0[r, c_] = P.Tb[r, c_ - r + P.band]
        self._eval(False)
        qB, qI = self._charge()
        qpB, qpI = qB.copy(), [q.copy() for q in qI]
        dqB, dqI = np.zeros_like(qB), [np.zeros_like(q) for q in qI]
        xBp = devp = wIp = None
        for i in range(1, ns):
            self._eval(False)
            qB, qI = self._charge()
            if self.trap:
                dqB = -dqB + a0 * (qB - qpB)
                dqI = [-d + a0 * (q - qp) for d, q, qp in zip(dqI, qI, qpI)]
            qpB, qpI = qB, qI
            svB = a0 * qpB + (dqB if self.trap else 0.) + src[i][P.B]
            svI = [a0 * qp + (d if self.trap else 0.) + src[i][c.Ic]
                   for qp, d, c in zip(qpI, dqI, P.ctas)]
            zI = [c.Uinv.dot(c.Linv.dot(s)) for c, s in zip(P.ctas, svI)]
            gB = np.zeros(P.nB)
            for c, z in zip(P.ctas, zI):
                gB += c.Asi.dot(z)                     # (the exchange of the step)
            bB = svB - gB
            if i > 1:
                rB = bB - T0 @ xBp - devp
                dB = band_solve(self.Tf, P.band, rB)
                dI = [z - w - c.X.dot(dB) for c, z, w in zip(P.ctas, zI, wIp)]
                self.xB = self.xB + dB
                self.xI = [x + d for x, d in zip(self.xI, dI)]
            ok = False
            it = 0
            while it < self.maxiter:
                out, jac = self.dev_eval(self._full_x(), True)
                jac_a0 = qscale_jac(jac, a0)
                devB = P.Dcur.dot(out) + a0 * P.Dchg.dot(out)
                wI = [xi + c.X.dot(self.xB) for c, xi in zip(P.ctas, self.xI)]
                xBp, devp, wIp = self.xB.copy(), devB, wI
                rB = bB - T0 @ self.xB - devB
                Tb = P.Tb + P.Tjac.dot(jac_a0).reshape(P.nB, W)
                self.Tf = band_factor(Tb, P.band)
                dB = band_solve(self.Tf, P.band, rB)
                dI = [z - w - c.X.dot(dB) for c, z, w in zip(P.ctas, zI, wI)]
                m = max([np.abs(dB).max()] + [np.abs(d).max() for d in dI if len(d)])
                scale = self.maxdelta / m if m > self.maxdelta else 1.
                conv = True
                d = dB * scale
                xn = self.xB + d
                conv &= bool(np.all(np.abs(d) < self.reltol * np.maximum(
                    np.abs(self.xB), np.abs(xn)) + self.abstol))
                self.xB = xn
                newI = []
                for xi, di in zip(self.xI, dI):
                    d = di * scale
                    xn = xi + d
                    conv &= bool(np.all(np.abs(d) < self.reltol * np.maximum(
                        np.abs(xi), np.abs(xn)) + self.abstol))
                    newI.append(xn)
                self.xI = newI
                res = m * scale
                it += 1
                if conv:
                    ok = True
                    break
            iters[i], resv[i] = it, res
            if not ok:
                return X[:i], iters[:i], resv[:i]
            X[i] = self._full_x()
        return X, iters, resv
