"""
Batched transient of many parameter variations of one circuit (Monte Carlo,
corner and parameter sweeps).

The reference has no such analysis; its nearest mechanism is the DC sweep
loop (cardoon/analyses/dc.py:167-205), which varies a parameter with
``setattr(elem, name, value)`` -> ``elem.process_params()`` ->
``nd.restore_RCnumbers(elem)`` -> ``nd.process_nodal_element(elem, ckt)`` ->
``refresh()`` and re-solves, one point after another.  Here the same
mechanism is used to *pack* the post-``process_params`` device records of
every instance (host), and then all instances are solved together: the
operating points in one batched Newton launch and the transients in one
launch of the device-resident time loop (one thread team per instance, state
in shared memory).  Instances are independent, so several GPUs simply take
consecutive ranges of instances (one process per GPU) and the waveforms are
gathered once at the end.
"""
import numpy as np

from ..globalVars import glVar
from . import nodalCUDA as nd
from .integration import BEuler, Trapezoidal


class BatchTransient(object):
    """``B`` variations of one flattened, initialised circuit.

    Typical use::

        bt = BatchTransient(ckt, tstop, tstep, im='trap')
        params = bt.pack_instances(B, vary)      # host, reference mechanism
        wave, iters, status = bt.run(params, ['1', '3', '4'])
    """

    def __init__(self, ckt, tstop, tstep, im='trap', lib=None):
        self.ckt = ckt
        self.tstop, self.tstep, self.im = tstop, tstep, im
        if not hasattr(ckt, 'nD_nlinElem'):
            nd.make_nodal_circuit(ckt)
        self.topo = nd.get_topology(ckt)
        self.lib = lib
        self.timeVec = np.arange(start=0., stop=tstop, step=tstep,
                                 dtype=float)
        helper = nd.TransientNodal(
            ckt, Trapezoidal() if im == 'trap' else BEuler())
        self.src = helper.source_table(self.timeVec)
        self.x0 = nd.get_guess(ckt)
        self._engines = {}

    # ---- host: per-instance device records ------------------------------------
    def pack_batch(self, overrides, out=None, device_major=False):
        """Records of B instances in one vectorised pass.

        ``overrides``: one dict per nonlinear element (order of
        ``ckt.nD_nlinElem``) mapping attribute names to arrays of shape (B,).
        The arrays are set on the element (the ``setattr`` of dc.py:179-188)
        and ``process_params`` / ``cuda_pack`` run once in batch mode
        (devices/cudadev.py): every column is bit-identical to what
        ``pack_instances`` computes for that instance, at a few milliseconds
        per thousand instances instead of Python per-element calls.  Raises
        ``NotImplementedError`` when a device class or the variation cannot
        be vectorised (callers fall back to ``pack_instances``).

        ``device_major=True`` returns ``[npar, n_g, B]`` arrays (contiguous
        per element: the fastest to fill on the host, e.g. straight into
        pinned memory passed as ``out``); ``run_device`` re-orders them to the
        kernels' instance-major columns on the GPU.
        """
        ckt, topo = self.ckt, self.topo
        elems = ckt.nD_nlinElem
        B = max(np.size(v) for o in overrides for v in o.values())
        saved = [dict(e.__dict__) for e in elems]
        given = out is not None
        out = list(out) if given else [None] * len(topo.groups)
        where = {}
        for gi, g in enumerate(topo.groups):
            npar = len(g['elems'][0].cuda_spec()['params'])
            shape = (npar, g['n'], B) if device_major else (npar, B, g['n'])
            if not given:
                out[gi] = np.empty(shape)
            elif out[gi].shape != shape:
                raise ValueError('out[{0}]: shape {1}, expected {2}'.format(
                    gi, out[gi].shape, shape))
            for i, e in enumerate(g['elems']):
                where[id(e)] = (gi, i)
        try:
            for e, over in zip(elems, overrides):
                lin = (list(e.linearVCCS), list(e.linearVCQS))
                nterm = len(e.connection)
                for name, val in over.items():
                    setattr(e, name, np.asarray(val, dtype=np.float64))
                try:
                    e.process_params()
                    flags, params = e.cuda_pack()
                except (ValueError, TypeError) as err:
                    # an ``if`` on a varied value that is not batch-aware
                    raise NotImplementedError(
                        '{0}: not vectorisable: {1}'.format(
                            e.instanceName, err))
                if (list(e.linearVCCS), list(e.linearVCQS)) != lin \
                        or len(e.connection) != nterm or any(
                            np.ndim(q[2]) for q in e.linearVCCS
                            + e.linearVCQS):
                    raise NotImplementedError(
                        '{0}: the variation changes linear stamps; '
                        'per-instance linear matrices are not supported'
                        .format(e.instanceName))
                gi, i = where[id(e)]
                if int(flags) != topo.groups[gi]['flags']:
                    raise NotImplementedError(
                        '{0}: the variation changes the model structure '
                        '(kernel flags)'.format(e.instanceName))
                dest = out[gi][:, i, :] if device_major \
                    else out[gi][:, :, i]
                if hasattr(params, 'write_into'):      # batch record
                    params.write_into(dest)
                else:                                  # nothing varied
                    dest[...] = np.asarray(params)[:, None]
        finally:
            for e, d in zip(elems, saved):
                e.__dict__.clear()
                e.__dict__.update(d)
        if device_major:
            return out
        # columns instance-major: b * n_g + i
        return [o.reshape(o.shape[0], -1) for o in out]

    def pack_instances(self, B, vary, first=0, instances=None):
        """Records of instances ``first .. first+B-1`` (or of the listed
        ``instances``), one Python ``process_params`` per element and
        instance: the reference mechanism, kept as the definition that
        ``pack_batch`` is tested against.

        ``vary(j, elems)`` sets attributes on the nonlinear elements for
        instance ``j`` (plain ``setattr``, as dc.py:179-188 does); each
        element is then re-initialised with ``process_params()`` and its
        kernel record read back.  Returns one ``[npar, B*n_g]`` array per
        device group (columns instance-major: ``b*n_g + i``).  The circuit's
        own element attributes are restored afterwards.
        """
        ckt, topo = self.ckt, self.topo
        elems = ckt.nD_nlinElem
        saved = [dict(e.__dict__) for e in elems]
        lin0 = [(list(e.nD_linVCCS), list(e.nD_linVCQS)) for e in elems]
        out = [np.empty((len(g['elems'][0].cuda_spec()['params']),
                         B * g['n'])) for g in topo.groups]
        where = {}
        for gi, g in enumerate(topo.groups):
            for i, e in enumerate(g['elems']):
                where[id(e)] = (gi, i)
        ids = list(instances) if instances is not None \
            else range(first, first + B)
        try:
            for b, j in enumerate(ids):
                vary(j, elems)
                for e, lin in zip(elems, lin0):
                    e.process_params()
                    nd.restore_RCnumbers(e)
                    nd.process_nodal_element(e, ckt)
                    if (list(e.nD_linVCCS), list(e.nD_linVCQS)) != lin:
                        raise NotImplementedError(
                            '{0}: the variation changes linear stamps; '
                            'per-instance linear matrices are not supported'
                            .format(e.instanceName))
                    gi, i = where[id(e)]
                    g = topo.groups[gi]
                    spec = e.cuda_spec()
                    if spec['flags'] != g['flags']:
                        raise NotImplementedError(
                            '{0}: the variation changes the model '
                            'structure (kernel flags)'.format(e.instanceName))
                    out[gi][:, b * g['n'] + i] = spec['params']
        finally:
            for e, d in zip(elems, saved):
                e.__dict__.clear()
                e.__dict__.update(d)
            ckt.__dict__.pop('_cuda_topology', None)
            ckt._cuda_topology = topo
        return out

    # ---- device ---------------------------------------------------------------------
    def engines(self, B):
        """(dc, tran) engines for a batch of B instances (cached)."""
        if B not in self._engines:
            a0 = (2. / self.tstep) if self.im == 'trap' else (1. / self.tstep)
            dc = nd.Engine(self.topo, use_q=False, batch=B, lib=self.lib,
                           keep_state=False, x_ref=self.x0)
            tr = nd.Engine(self.topo, use_q=True, a0=a0, batch=B,
                           lib=self.lib, keep_state=False, x_ref=self.x0,
                           h=self.tstep)
            self._engines[B] = (dc, tr)
        return self._engines[B]

    def save_rows(self, terms):
        rows = []
        for t in terms:
            term = self.ckt.find_term(t) if isinstance(t, str) else t
            rows.append(term.nD_namRC)
        return rows

    def operating_points(self, params, B):
        """Batched operating points (sources at t = 0, tran.py:112-117):
        plain Newton, then the convergence helpers for the instances that
        need them, all inside one launch (fsolve.py:23-56 helper chain).
        Returns the DC engine: ``x`` [B, n], ``iters``, ``status`` and
        ``hinfo[:, 0]`` (helper used) are device tensors."""
        dc, tr = self.engines(B)
        torch = dc.lib.torch
        if any(p.dim() == 3 if hasattr(p, 'dim') else p.ndim == 3
               for p in params):
            # device-major [npar, n_g, B] -> instance-major columns, on the
            # GPU, into buffers kept for the batch size
            bufs = self.__dict__.setdefault('_pbuf', {})
            cols = []
            for gi, p in enumerate(params):
                t = p if hasattr(p, 'data_ptr') else dc.lib.dev(p)
                npar, ng, nb = t.shape
                key = (gi, npar, ng, nb)
                if key not in bufs:
                    bufs[key] = torch.empty((npar, nb * ng),
                                            dtype=torch.float64,
                                            device=dc.lib.device)
                bufs[key].view(npar, nb, ng).copy_(t.permute(0, 2, 1))
                cols.append(bufs[key])
            params = cols
        dc.set_params(params)
        if getattr(self, '_x0_dev', None) is None or \
                self._x0_dev.shape[0] != B:
            self._x0_dev = torch.as_tensor(np.broadcast_to(
                self.x0, (B, self.topo.n)).copy()).to(dc.lib.device)
            self._sv0_dev = torch.as_tensor(self.src[0].copy()).to(
                dc.lib.device)
        dc.x.copy_(self._x0_dev)
        dc.sv.copy_(self._sv0_dev)
        dc._launch(dc._run_struct(nd.L.OP_NEWTON, helpers=1))
        return dc

    def run_device(self, params, rows, B):
        """Operating points + transients of B instances; everything stays on
        the device.  ``params``: per group ``[npar, B*n_g]`` device tensors
        (or numpy).  Returns device tensors (wave[B, nsamples, nsave],
        iters[B, nsamples], status[B], op_iters[B])."""
        dc, tr = self.engines(B)
        torch = dc.lib.torch
        self.operating_points(params, B)
        tr.set_params(dc.params)
        tr.x.copy_(dc.x)
        wave, iters, res, status = tr.tran(None, self.src, rows, im=self.im,
                                           sync=False)
        # an instance whose operating point failed is flagged at sample 0
        status = torch.where(dc.status != 0, torch.full_like(status, -1),
                             status)
        # highest convergence helper each instance needed (operating point,
        # transient): 0 = plain Newton only
        self.helper_used = (dc.hinfo[:, 0], tr.hinfo[:, 0])
        return wave, iters, status, dc.iters

    def run(self, params, terms):
        B = params[0].shape[1] // self.topo.groups[0]['n']
        wave, iters, status, opit = self.run_device(
            params, self.save_rows(terms), B)
        return (wave.cpu().numpy(), iters.cpu().numpy(),
                status.cpu().numpy(), opit.cpu().numpy())


def bsim3_ring_samples(instances, nelem=6, seed=20240):
    """The standard-normal draws of the config-5 variation, z[B, nelem, 5]
    clipped to +-4: instance j uses ``default_rng([seed, j])`` and draws five
    numbers per MOSFET in element order (SURVEY 8d "M")."""
    z = np.empty((len(instances), nelem, 5))
    for b, j in enumerate(instances):
        rng = np.random.default_rng([seed, int(j)])
        for k in range(nelem):
            z[b, k] = np.clip(rng.standard_normal(5), -4., 4.)
    return z


def bsim3_ring_overrides(z, elems):
    """Per-element attribute arrays of ``bsim3_ring_variation`` for the draws
    ``z`` [B, nelem, 5] (input of ``BatchTransient.pack_batch``)."""
    out = []
    for k, e in enumerate(elems):
        ntype = e._tf > 0
        zk = z[:, k, :]
        out.append(dict(
            vfb=-1. * (1. + 0.03 * zk[:, 0]),
            u0=(670. if ntype else 250.) * (1. + 0.05 * zk[:, 1]),
            tox=1.5e-8 * (1. + 0.02 * zk[:, 2]),
            l=1e-6 * (1. + 0.02 * zk[:, 3]),
            w=(1e-6 if ntype else 3e-6) * (1. + 0.02 * zk[:, 4])))
    return out


def bsim3_ring_variation(seed=20240):
    """The BASELINE config-5 variation (SURVEY 8d "M"): for each MOSFET of
    instance j, independently, z ~ N(0,1) clipped to +-4:
    vfb <- -1 (1 + 0.03 z1), u0 <- u0_nom (1 + 0.05 z2) with u0_nom 670 (n) /
    250 (p), tox <- 1.5e-8 (1 + 0.02 z3), l <- 1e-6 (1 + 0.02 z4),
    w <- w_nom (1 + 0.02 z5) with w_nom 1 um (n) / 3 um (p)."""
    def vary(j, elems):
        rng = np.random.default_rng([seed, j])
        for e in elems:
            z = np.clip(rng.standard_normal(5), -4., 4.)
            ntype = e._tf > 0
            e.vfb = -1. * (1. + 0.03 * z[0])
            e.u0 = (670. if ntype else 250.) * (1. + 0.05 * z[1])
            e.tox = 1.5e-8 * (1. + 0.02 * z[2])
            e.l = 1e-6 * (1. + 0.02 * z[3])
            e.w = (1e-6 if ntype else 3e-6) * (1. + 0.02 * z[4])
    return vary


def shard(total, rank, world):
    """Instance range [lo, hi) of ``rank``."""
    lo = rank * total // world
    hi = (rank + 1) * total // world
    return lo, hi


def gather_results(wave, iters, status, dist=None):
    """All ranks' results concatenated along the instance axis (one
    all_gather over NCCL/NVLink at the end of the run; nothing is exchanged
    inside the time loop)."""
    if dist is None or not dist.is_initialized() or \
            dist.get_world_size() == 1:
        return wave, iters, status
    import torch
    world = dist.get_world_size()
    # shards may be unequal (total % world != 0) or empty: exchange the
    # counts, pad to the largest, gather, cut back
    cnt = torch.tensor([wave.shape[0]], dtype=torch.int64,
                       device=wave.device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    bmax = max(cnts)
    outs = []
    for t in (wave, iters, status):
        pad = torch.zeros((bmax,) + tuple(t.shape[1:]), dtype=t.dtype,
                          device=t.device)
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        outs.append(torch.cat([p[:c] for p, c in zip(parts, cnts)], dim=0))
    return tuple(outs)
