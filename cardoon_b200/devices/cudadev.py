"""
Device-plugin side of the CUDA boundary.

Replaces the reference's AD-tape module devices/cppaddev.py: there,
``eval``/``eval_and_deriv`` (:134-170) taped ``eval_cqs`` with CppAD and
replayed the tape; here they hand the port voltages and the packed
post-``process_params`` parameter record to the forward-mode dual-number
kernel of the model (``cdn_dev_eval`` in libcardoon_b200.so) with a batch of
one.  The nodal engines never call these: they evaluate whole device groups
inside the Newton kernels from the same packed records.

There is deliberately no Python/CPU implementation of the model equations in
this package: without the CUDA library and a GPU these functions raise.
"""
import numpy as np

from .param_layout import MODEL_IDS, LAYOUTS


class CudaModel(object):
    """Mixin for nonlinear device classes evaluated by a CUDA kernel.

    Sub-classes set ``cudaModel`` (a key of param_layout.MODEL_IDS) and
    implement ``cuda_pack() -> (flags, params)`` where ``params`` follows
    param_layout.LAYOUTS[cudaModel].  ``process_params``/``set_temp_vars``
    must call ``delete_tape(self)`` (same contract as the reference,
    cppaddev.py:99-111) so that the cached record is rebuilt.
    """
    cudaModel = None

    def cuda_spec(self):
        spec = self.__dict__.get('_cuda_spec')
        if spec is None:
            flags, params = self.cuda_pack()
            params = np.ascontiguousarray(params, dtype=np.float64)
            nexp = len(LAYOUTS[self.cudaModel][0])
            if self.cudaModel != 'memr':
                assert params.shape == (nexp,), \
                    '{0}: packed {1} parameters, layout has {2}'.format(
                        self.instanceName, params.shape, nexp)
            spec = dict(model=MODEL_IDS[self.cudaModel], flags=int(flags),
                        params=params,
                        nxin=len(self.controlPorts) + self.nDelays,
                        ncs=len(self.csOutPorts), nqs=len(self.qsOutPorts))
            self._cuda_spec = spec
        return spec

    # ---- reference plugin API (batch of one) --------------------------
    def eval_and_deriv(self, vPort):
        return eval_and_deriv(self, vPort)

    def eval(self, vPort):
        return eval(self, vPort)

    def eval_cqs(self, vPort, getOP=False):
        """(iVec, qVec) at the control voltages ``vPort``.

        Same meaning as the reference's ``eval_cqs``; computed on the GPU.
        """
        out = eval(self, vPort)
        ncs = len(self.csOutPorts)
        return (out[:ncs], out[ncs:])


# ---- batch mode ---------------------------------------------------------------
# A Monte-Carlo / corner batch packs the records of B parameter variations of
# one element at once: the varied attributes are numpy arrays of shape (B,)
# and ``process_params`` / ``set_temp_vars`` / ``cuda_pack`` run ONCE, with
# the same FP64 operations element-wise (numpy evaluates +, -, *, /, sqrt,
# exp, log and pow per element exactly as for a scalar), so every column is
# bit-identical to the scalar path.  Branches on varied values go through
# ``sel`` / ``uniform`` below; anything else that cannot be vectorised raises
# and the caller falls back to the per-instance loop.
def sel(cond, a, b):
    """``a if cond else b`` for scalars, ``np.where`` for batches."""
    if np.ndim(cond) == 0:
        return a if cond else b
    return np.where(cond, a, b)


def power(x, y):
    """``pow(x, y)``; in batch mode through the scalar libm ``pow`` element by
    element (numpy's vectorised power is not bit-identical to it)."""
    if np.ndim(x) == 0 and np.ndim(y) == 0:
        return pow(x, y)
    import math
    if np.ndim(y) == 0 and y == 1.:
        return np.asarray(x, dtype=np.float64) + 0.   # pow(x, 1) == x exactly
    xb, yb = np.broadcast_arrays(np.asarray(x, dtype=np.float64),
                                 np.asarray(y, dtype=np.float64))
    return np.array([math.pow(a, b) for a, b in zip(xb.ravel().tolist(),
                                                    yb.ravel().tolist())]
                    ).reshape(xb.shape)


def uniform(cond, what):
    """A condition that selects kernel structure must agree over the batch."""
    if np.ndim(cond) == 0:
        return bool(cond)
    if np.all(cond) or not np.any(cond):
        return bool(cond.flat[0])
    raise NotImplementedError(
        'the variation changes the model structure ({0})'.format(what))


class BatchRecord(object):
    """Kernel record of a batch: a list of rows, each a scalar or an array
    [B]; materialised once, straight into its destination (``write_into``)."""

    def __init__(self, rows, B):
        self.rows = rows
        self.B = B
        self.ndim = 2

    def __len__(self):
        return len(self.rows)

    @property
    def shape(self):
        return (len(self.rows), self.B)

    def __getitem__(self, key):
        if not isinstance(key, slice):
            raise TypeError('BatchRecord supports row slices only')
        return BatchRecord(self.rows[key], self.B)

    def write_into(self, dest):
        for k, r in enumerate(self.rows):
            dest[k] = r
        return dest

    def __array__(self, dtype=None, copy=None):
        return self.write_into(np.empty(self.shape))


def _rows(tails):
    out = []
    for t in tails:
        out.extend(np.atleast_1d(np.asarray(t, dtype=np.float64)).tolist())
    return out


def stack_record(entries, *tails):
    """Kernel record from a list of scalar entries (batch mode: some may be
    arrays [B]) followed by 1-D scalar records: an array [npar], or a
    ``BatchRecord`` [npar, B]."""
    B = max([np.size(e) for e in entries if np.ndim(e) == 1] + [0])
    if B == 0:
        tail = np.concatenate([np.atleast_1d(np.asarray(t, dtype=np.float64))
                               for t in tails]) if tails else np.zeros(0)
        return np.concatenate([np.array(entries, dtype=np.float64), tail])
    return BatchRecord(list(entries) + _rows(tails), B)


def append_rows(rec, *tails):
    """``np.concatenate([rec] + tails)`` for [npar] records and batches."""
    if isinstance(rec, BatchRecord):
        return BatchRecord(rec.rows + _rows(tails), rec.B)
    return np.concatenate([rec] + [np.atleast_1d(np.asarray(
        t, dtype=np.float64)) for t in tails])


def delete_tape(dev):
    """Invalidate the cached kernel record (reference cppaddev.py:99)."""
    dev.__dict__.pop('_cuda_spec', None)


def _run(dev, vPort, want_jac):
    from .._lib import get_lib
    spec = dev.cuda_spec()
    vPort = np.asarray(vPort, dtype=np.float64)
    if vPort.shape != (spec['nxin'],):
        raise ValueError('{0}: expected {1} port voltages, got {2}'.format(
            dev.instanceName, spec['nxin'], vPort.shape))
    return get_lib().dev_eval_host(spec['model'], spec['flags'],
                                   spec['params'][:, None], vPort[:, None],
                                   spec['ncs'] + spec['nqs'], want_jac)


def eval_and_deriv(dev, vPort):
    """(out, Jac): out = [currents, charges], Jac[r, c] = d out_r / d vPort_c."""
    out, jac = _run(dev, vPort, True)
    nout = out.shape[0]
    return out[:, 0].copy(), jac[:, 0].reshape(nout, -1).copy()


def eval(dev, vPort):
    out, _ = _run(dev, vPort, False)
    return out[:, 0].copy()
