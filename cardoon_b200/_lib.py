"""
ctypes binding of libcardoon_b200.so (include/cardoon_b200.h).

PyTorch only owns device buffers here (``tensor.data_ptr()``) and provides
the current stream; every computation is a call into the library.  There is
no CPU fallback: a missing library or GPU raises ``CudaUnavailable``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (CARDOON_B200_LIB: a differently tuned build of the same library, used by
# tools/build_variant.py experiments)
LIBPATH = os.environ.get('CARDOON_B200_LIB') or os.path.join(
    _HERE, 'libcardoon_b200.so')


class CudaUnavailable(RuntimeError):
    pass


class CdnError(RuntimeError):
    pass


_vp = C.c_void_p


class GroupHost(C.Structure):
    """cdn_group_host (csrc/plan_host.h)."""
    _fields_ = [('model', C.c_int), ('flags', C.c_uint), ('n', C.c_int),
                ('nxin', C.c_int), ('ncs', C.c_int), ('nqs', C.c_int),
                ('ndelay', C.c_int),
                ('vpos', _vp), ('vneg', _vp), ('cpos', _vp), ('cneg', _vp),
                ('qpos', _vp), ('qneg', _vp), ('delay', _vp)]


class CircuitHost(C.Structure):
    """cdn_circuit_host (csrc/plan_host.h)."""
    _fields_ = [('n', C.c_int), ('ordering', C.c_int),
                ('nG', C.c_long), ('g_row', _vp), ('g_col', _vp),
                ('g_val', _vp),
                ('nC', C.c_long), ('c_row', _vp), ('c_col', _vp),
                ('c_val', _vp),
                ('nE', C.c_long), ('e_row', _vp), ('e_col', _vp),
                ('e_val', _vp),
                ('ngroups', C.c_int), ('groups', C.POINTER(GroupHost)),
                ('rowmatch', _vp), ('colorder', _vp), ('heavy', C.c_int),
                ('skip_q', C.c_int), ('nhist', C.c_int)]


class RunHost(C.Structure):
    """cdn_run_host (csrc/engine.cu)."""
    _fields_ = [('abstol', C.c_double), ('reltol', C.c_double),
                ('maxdelta', C.c_double), ('a0', C.c_double),
                ('gmin', C.c_double), ('gmin_ext', C.c_double),
                ('lam', C.c_double),
                ('h', C.c_double),
                ('maxiter', C.c_int), ('errfunc', C.c_int),
                ('use_q', C.c_int), ('im', C.c_int),
                ('op', C.c_int), ('B', C.c_int),
                ('team', C.c_int), ('T', C.c_int),
                ('helpers', C.c_int), ('n_ext', C.c_int),
                ('x', _vp), ('sv_in', _vp), ('ws', _vp),
                ('nsamples', C.c_int), ('first', C.c_int),
                ('init_ic', C.c_int), ('nsrc', C.c_int), ('nsave', C.c_int),
                ('blocks', C.c_int),
                ('src_rows', _vp), ('src_val', _vp), ('src_dc', _vp),
                ('save_rows', _vp), ('wave', _vp), ('iters', _vp),
                ('res', _vp), ('status', _vp), ('red', _vp), ('prof', _vp),
                ('info', _vp)]


class SchurHost(C.Structure):
    """cdn_schur_host (csrc/schur.cu)."""
    _fields_ = [('ncta', C.c_int), ('nB', C.c_int), ('band', C.c_int),
                ('blob', _vp), ('cta_off', _vp), ('m', _vp),
                ('vpos_b', _vp), ('vneg_b', _vp),
                ('src_own', _vp), ('src_loc', _vp),
                ('save_ptr', _vp), ('save_k', _vp), ('save_loc', _vp)]


OP_NEWTON, OP_TRAN, OP_UPDATE_Q, OP_SET_IC, OP_ACCEPT, OP_RHS, OP_CHORD, \
    OP_GET_I, OP_GET_I_JAC, OP_SOLVE = range(10)
TEAM_AUTO, TEAM_TILE, TEAM_BLOCK, TEAM_GRID = -1, 0, 1, 2
ORDER_ND, ORDER_NATURAL, ORDER_DENSE, ORDER_GIVEN = 0, 1, 2, 3


def _declare(dll):
    dll.cdn_version.restype = C.c_int
    dll.cdn_init.argtypes = [C.c_int, C.POINTER(_vp)]
    dll.cdn_free.argtypes = [_vp]
    dll.cdn_last_error.argtypes = [_vp]
    dll.cdn_last_error.restype = C.c_char_p
    dll.cdn_sm_count.argtypes = [_vp]
    dll.cdn_model_nparams.argtypes = [C.c_int]
    dll.cdn_dev_eval.argtypes = [_vp, C.c_int, C.c_uint, C.c_long, C.c_int,
                                 _vp, C.c_long, _vp, _vp, _vp, _vp, _vp]
    dll.cdn_dev_eval_split.argtypes = [_vp, C.c_int, C.c_uint, C.c_long, _vp,
                                       C.c_int, _vp, _vp, C.c_long, _vp, _vp,
                                       _vp, _vp]
    dll.cdn_ac_solve.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int,
                                 _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp,
                                 _vp, _vp, _vp, _vp, C.c_long, _vp, _vp, _vp]
    dll.cdn_ac_work_matrices.argtypes = [_vp, C.c_int]
    dll.cdn_ac_work_matrices.restype = C.c_long
    dll.cdn_fp64_peak.argtypes = [_vp, C.POINTER(C.c_double)]
    dll.cdn_set_tuning.argtypes = [C.c_int, C.c_int]
    dll.cdn_plan_create.argtypes = [_vp, C.POINTER(CircuitHost),
                                    C.POINTER(_vp)]
    dll.cdn_plan_bytes.argtypes = [_vp]
    dll.cdn_plan_bytes.restype = C.c_long
    dll.cdn_plan_info.argtypes = [_vp, C.POINTER(C.c_long)]
    dll.cdn_plan_bind.argtypes = [_vp, _vp, C.c_long, _vp]
    dll.cdn_plan_set_group_params.argtypes = [_vp, C.c_int, _vp, C.c_long,
                                              _vp, C.c_int]
    dll.cdn_plan_slot_coords.argtypes = [_vp, _vp, _vp]
    dll.cdn_plan_free.argtypes = [_vp]
    dll.cdn_plan_array.argtypes = [_vp, C.c_char_p, C.POINTER(_vp),
                                   C.POINTER(C.c_long)]
    dll.cdn_init_host.argtypes = [C.POINTER(_vp)]
    dll.cdn_engine_ws_doubles.argtypes = [_vp]
    dll.cdn_engine_ws_doubles.restype = C.c_long
    dll.cdn_engine_layout.argtypes = [_vp, C.POINTER(C.c_long)]
    dll.cdn_engine_run.argtypes = [_vp, C.POINTER(RunHost), _vp]
    dll.cdn_schur_tran.argtypes = [_vp, C.POINTER(RunHost),
                                   C.POINTER(SchurHost), _vp]
    dll.cdn_schur_smem_bytes.argtypes = [C.c_long] + [C.c_int] * 6
    dll.cdn_schur_smem_bytes.restype = C.c_long


def load_dll():
    """Load the shared library (no GPU needed); raises if it is not built."""
    if not os.path.exists(LIBPATH):
        raise CudaUnavailable(
            'libcardoon_b200.so is not built: run '
            '`python -c "import __graft_entry__ as g; g.build()"`')
    dll = C.CDLL(LIBPATH)
    _declare(dll)
    return dll


class Lib(object):
    """One library context bound to one GPU."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise CudaUnavailable(
                'cardoon_b200 needs a CUDA GPU (B200, sm_100a); there is no '
                'CPU implementation of the hot path in this package')
        self.torch = torch
        self.dll = load_dll()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device('cuda', device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)      # make sure a context exists
        self.ctx = _vp()
        rc = self.dll.cdn_init(int(device), C.byref(self.ctx))
        if rc != 0:
            raise CdnError('cdn_init failed ({0})'.format(rc))

    # ---- helpers ---------------------------------------------------------
    def check(self, rc):
        if rc != 0:
            msg = self.dll.cdn_last_error(self.ctx)
            raise CdnError('libcardoon_b200 error {0}: {1}'.format(
                rc, msg.decode() if msg else ''))

    def stream(self):
        return _vp(self.torch.cuda.current_stream(self.device).cuda_stream)

    def dev(self, array, dtype=None):
        """numpy array -> contiguous device tensor."""
        t = self.torch.as_tensor(np.ascontiguousarray(array))
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.device)

    def empty(self, shape, dtype=None):
        return self.torch.empty(shape, dtype=dtype or self.torch.float64,
                                device=self.device)

    def zeros(self, shape, dtype=None):
        return self.torch.zeros(shape, dtype=dtype or self.torch.float64,
                                device=self.device)

    @staticmethod
    def ptr(t):
        return _vp(t.data_ptr()) if t is not None else None

    # ---- device evaluation -------------------------------------------------
    def dev_eval(self, model, flags, params, xin, nout, want_jac=True,
                 pidx=None, out=None, jac=None):
        """Device tensors in, device tensors out.

        params: [npar, ncols]; xin: [nxin, n]; returns out[nout, n] and
        jac[nout*nxin, n] (None if not requested).
        """
        nxin, n = xin.shape
        if pidx is None and params.shape[1] != n:
            if params.shape[1] != 1:
                raise ValueError('params must have 1 or n columns')
            # one parameter set shared by every instance
            pidx = self.zeros((n,), self.torch.int32)
        if out is None:
            out = self.empty((nout, n))
        if want_jac and jac is None:
            jac = self.empty((nout * nxin, n))
        rc = self.dll.cdn_dev_eval(
            self.ctx, int(model), int(flags), int(n), int(nxin),
            self.ptr(params), int(params.stride(0)), self.ptr(pidx),
            self.ptr(xin), self.ptr(out), self.ptr(jac) if want_jac else None,
            self.stream())
        self.check(rc)
        return out, (jac if want_jac else None)

    def split_record(self, params):
        """Split packed records [npar, n] (device tensor) of devices that
        share model cards: returns ``(card[npar], vrow[npar] int16 host,
        inst[nvar, n])`` -- the rows all instances agree on once, the others
        per instance -- for ``dev_eval_split``."""
        torch = self.torch
        uniform = params.max(dim=1).values == params.min(dim=1).values
        u = uniform.cpu().numpy()
        vrow = np.full(params.shape[0], -1, dtype=np.int16)
        vrow[~u] = np.arange(int((~u).sum()), dtype=np.int16)
        card = params[:, 0].contiguous()
        inst = params[~uniform].contiguous() if (~u).any() else None
        return card, vrow, inst

    def dev_eval_split(self, model, flags, split, xin, nout, out=None,
                       jac=None):
        """Value + Jacobian from a split record (``split_record``)."""
        card, vrow, inst = split
        nxin, n = xin.shape
        if out is None:
            out = self.empty((nout, n))
        if jac is None:
            jac = self.empty((nout * nxin, n))
        rc = self.dll.cdn_dev_eval_split(
            self.ctx, int(model), int(flags), int(n), self.ptr(card),
            int(card.shape[0]), vrow.ctypes.data_as(_vp), self.ptr(inst),
            int(inst.stride(0)) if inst is not None else 0, self.ptr(xin),
            self.ptr(out), self.ptr(jac), self.stream())
        self.check(rc)
        return out, jac

    def dev_eval_host(self, model, flags, params, xin, nout, want_jac=True):
        """numpy in / numpy out convenience (plugin batch-of-one path)."""
        params = np.ascontiguousarray(params, dtype=np.float64)
        xin = np.ascontiguousarray(xin, dtype=np.float64)
        out, jac = self.dev_eval(model, flags, self.dev(params),
                                 self.dev(xin), nout, want_jac)
        self.torch.cuda.synchronize(self.device)
        return (out.cpu().numpy(),
                jac.cpu().numpy() if jac is not None else None)

    def ac_prepare(self, G, Cm, delayed, freqs, sVec, fd_rc=(), fd_val=()):
        """Upload the operands of an AC sweep; returns the launch state
        (device tensors + sizes) for ``ac_launch`` / ``ac_fetch``."""
        torch = self.torch
        n = G.shape[0]
        freqs = np.ascontiguousarray(freqs, dtype=np.float64)
        nf = freqs.shape[0]
        ndel = len(delayed)
        d = np.array(delayed, dtype=np.float64).reshape(ndel, 5)
        s2 = np.ascontiguousarray(
            np.stack([sVec.real, sVec.imag], axis=1), dtype=np.float64)

        def ti(a):
            return torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32),
                                   device=self.device)
        nfd = len(fd_rc)
        frc = np.array(fd_rc, dtype=np.int32).reshape(nfd, 2)
        fv = np.zeros((nf, nfd, 2))
        for e, v in enumerate(fd_val):
            v = np.broadcast_to(np.asarray(v, dtype=complex), (nf,))
            fv[:, e, 0] = v.real
            fv[:, e, 1] = v.imag
        st = dict(n=int(n), nf=int(nf), ndel=int(ndel), nfd=int(nfd),
                  freqs_h=freqs,
                  G=self.dev(np.ascontiguousarray(G, dtype=np.float64)),
                  C=self.dev(np.ascontiguousarray(Cm, dtype=np.float64)),
                  drow=ti(d[:, 0]), dcol=ti(d[:, 1]),
                  dg=self.dev(d[:, 2].copy()), dc=self.dev(d[:, 3].copy()),
                  dtau=self.dev(d[:, 4].copy()), frow=ti(frc[:, 0]),
                  fcol=ti(frc[:, 1]), fval=self.dev(fv),
                  freqs=self.dev(freqs), s=self.dev(s2),
                  # one matrix per CTA in flight
                  nwork=int(self.dll.cdn_ac_work_matrices(self.ctx, int(nf))),
                  work=self.empty((int(self.dll.cdn_ac_work_matrices(
                      self.ctx, int(nf))) * n * n * 2,)) if n > 56 else None,
                  x=self.empty((nf, n, 2)),
                  status=self.zeros((1,), torch.int32))
        return st

    def ac_launch(self, st):
        """One ``cdn_ac_solve`` launch on the prepared operands (no sync)."""
        rc = self.dll.cdn_ac_solve(
            self.ctx, st['n'], st['nf'], self.ptr(st['G']), self.ptr(st['C']),
            st['ndel'], self.ptr(st['drow']), self.ptr(st['dcol']),
            self.ptr(st['dg']), self.ptr(st['dc']), self.ptr(st['dtau']),
            st['nfd'], self.ptr(st['frow']), self.ptr(st['fcol']),
            self.ptr(st['fval']), self.ptr(st['freqs']), self.ptr(st['s']),
            self.ptr(st['work']), st['nwork'], self.ptr(st['x']),
            self.ptr(st['status']),
            self.stream())
        self.check(rc)

    def ac_fetch(self, st):
        self.torch.cuda.synchronize(self.device)
        bad = int(st['status'].item())
        if bad:
            raise np.linalg.LinAlgError(
                'AC: singular matrix at frequency {0} Hz'.format(
                    st['freqs_h'][bad - 1]))
        xh = st['x'].cpu().numpy()
        return xh[:, :, 0] + 1j * xh[:, :, 1]

    def ac_solve(self, G, Cm, delayed, freqs, sVec, fd_rc=(), fd_val=()):
        """x[nf, n] (complex) with Y(w) x = s for every frequency;
        Y = G + j w C + sum (g + j w c) exp(-j w tau) over the `delayed`
        entries (row, col, g, c, tau), plus fd_val[e][k] at (row, col) =
        fd_rc[e] for frequency k (frequency-defined elements).  numpy in /
        numpy out."""
        st = self.ac_prepare(G, Cm, delayed, freqs, sVec, fd_rc, fd_val)
        self.ac_launch(st)
        return self.ac_fetch(st)

    def fp64_peak(self):
        v = C.c_double()
        self.check(self.dll.cdn_fp64_peak(self.ctx, C.byref(v)))
        return v.value


_LIB = {}


def get_lib(device=None):
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() \
            else 0
    if device not in _LIB:
        _LIB[device] = Lib(device)
    return _LIB[device]
