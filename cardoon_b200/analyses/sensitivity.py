"""
Nodal sensitivity analysis: derivatives of the DC currents of a device with
respect to its own parameters.

Interface of the reference's cardoon/analyses/sensitivity.py
(``get_i_derivatives`` :16-97): returns the matrix ``Ji[N, npar]`` that adds
``d i / d p`` of every listed float parameter into the rows of the nodal
equations the device feeds -- linear VCCS stamps (:63-80), DC source values
(:81-87) and the nonlinear current sources (:88-90).  Together with
``nd.DCNodal.get_adjoint_voltages`` (nodal.py:661-672) this is what an
adjoint sensitivity of an output voltage needs:
``d v_out / d p = -lambda^T Ji`` with ``Jac^T lambda = e_out``.

The reference differentiates a CppAD tape of ``process_params`` followed by
``eval_cqs`` (``create_sensitivity_tape`` :99-172).  Here ``process_params``
is host code and ``eval_cqs`` a CUDA kernel whose dual numbers carry the
derivatives with respect to the control voltages only, so the parameter
derivatives are taken by sixth-order central differences *through both*:
for each parameter the element is re-processed at p + s h, s = -2, -1, -1/2,
1/2, 1, 2 (the reference's own sweep mechanism, dc.py:179-188), the packed
records of all 6 x npar variants are evaluated in ONE ``cdn_dev_eval`` launch
at the operating point, and the fourth-order differences with steps h and
h/2 are Richardson-extrapolated.  With h = 1e-3 |p| the truncation error
(h^6 f^(7) / 5040; 3e-9 relative even for a breakdown exponential in bv/vt)
and the rounding error (~eps f / h) stay below 1e-8 of the derivative's scale;
tests pin the result to the reference's AD (golden vectors generated from
the reference's own device code, tests/golden/make_sens_golden.py).  A
parameter whose value is 0 gets a zero column when moving it would change
the element's structure (the reference's tape, recorded at 0, does not take
that branch either), otherwise an absolute step.

There is no CPU implementation of the device equations in this package: the
nonlinear part needs the CUDA library and a GPU.
"""
import numpy as np

from . import nodalCUDA as nd

STENCIL = (-2., -1., -0.5, 0.5, 1., 2.)


def _stencil(f, h):
    """Sixth-order derivative from f at p + s h, s in STENCIL: the
    fourth-order central differences with steps h and h/2, Richardson
    extrapolated.  Differences are formed first so that a quantity the
    parameter does not touch gives exactly 0."""
    d1 = (8. * (f[1.] - f[-1.]) - (f[2.] - f[-2.])) / (12. * h)
    d2 = (8. * (f[.5] - f[-.5]) - (f[1.] - f[-1.])) / (6. * h)
    return (16. * d2 - d1) / 15.


def _structure(device):
    return (len(device.connection), len(device.linearVCCS),
            len(getattr(device, 'controlPorts', ())),
            len(getattr(device, 'csOutPorts', ())),
            len(getattr(device, 'qsOutPorts', ())),
            device.cuda_spec()['flags'] if device.isNonlinear else 0)


def _snapshot(device):
    """What the derivative is taken of, after ``process_params``."""
    g = [float(v[2]) for v in device.linearVCCS]
    src = float(device.get_DCsource()) if device.isDCSource else None
    rec = device.cuda_spec()['params'].copy() if device.isNonlinear else None
    return g, src, rec


def device_derivatives(device, names, xin=None, lib=None, rel_step=1e-3):
    """Derivatives of what a device contributes to the DC equations with
    respect to its parameters ``names``: ``(di[ncs, npar], dg[nvccs, npar],
    dsrc[npar])`` -- nonlinear currents at the control voltages ``xin``
    (None for linear devices), linear VCCS conductances, DC source value.
    Columns of parameters that cannot be moved are zero (module docstring).
    """
    npar = len(names)
    p0 = [float(getattr(device, name)) for name in names]
    base_struct = _structure(device)
    saved = dict(device.__dict__)
    steps = np.zeros(npar)
    snaps = {}
    try:
        for i, (name, p) in enumerate(zip(names, p0)):
            h = rel_step * abs(p) if p != 0. else rel_step * 1e-3
            ok = True
            for s in STENCIL:
                setattr(device, name, p + s * h)
                try:
                    device.process_params()
                    if _structure(device) != base_struct:
                        ok = False
                    else:
                        snaps[(i, s)] = _snapshot(device)
                except Exception:
                    ok = False
                if not ok:
                    break
            setattr(device, name, p)
            steps[i] = h if ok else 0.
    finally:
        device.__dict__.clear()
        device.__dict__.update(saved)
        device.process_params()
        if hasattr(device, 'nD_intRC'):
            nd.restore_RCnumbers(device)
    live = [i for i in range(npar) if steps[i] > 0.]
    nvccs = len(device.linearVCCS)
    dg = np.zeros((nvccs, npar))
    dsrc = np.zeros(npar)
    for i in live:
        g = {s: np.asarray(snaps[(i, s)][0]) for s in STENCIL}
        dg[:, i] = _stencil(g, steps[i])
        if device.isDCSource:
            dsrc[i] = _stencil({s: snaps[(i, s)][1] for s in STENCIL},
                               steps[i])
    di = None
    if device.isNonlinear:
        spec = device.cuda_spec()
        ncs = spec['ncs']
        di = np.zeros((ncs, npar))
        if live:
            # all stencil points of all parameters in one launch
            from .._lib import get_lib
            lib = lib or get_lib()
            cols = [(i, s) for i in live for s in STENCIL]
            params = np.array([snaps[c][2] for c in cols]).T.copy()
            xin = np.asarray(xin, dtype=float)
            out, _ = lib.dev_eval_host(
                spec['model'], spec['flags'], params,
                np.repeat(xin[:, None], len(cols), axis=1),
                spec['ncs'] + spec['nqs'], want_jac=False)
            for k, i in enumerate(live):
                f = {s: out[:ncs, len(STENCIL) * k + m]
                     for m, s in enumerate(STENCIL)}
                di[:, i] = _stencil(f, steps[i])
    return di, dg, dsrc


def get_i_derivatives(device, parList, xVec, force=False, lib=None,
                      rel_step=1e-3):
    """DC current derivatives with respect to the given parameters.

    device: element with nodal attributes (``nD_*``), linear or nonlinear;
    parList: list of (name, value) as returned by
    ``get_float_attributes()``; xVec: nodal vector.  Returns ``Ji`` of shape
    ``(len(xVec), len(parList))``.  ``force`` is accepted for compatibility
    (there is no tape to regenerate).
    """
    xVec = np.asarray(xVec, dtype=float)
    names = [name for name, val in parList]
    npar = len(names)
    Ji = np.zeros((len(xVec), npar))
    xin = None
    if device.isNonlinear:
        xin = np.zeros(device.nD_nxin)
        nd.set_xin(xin, device.nD_vpos, device.nD_vneg, xVec)
    di, dg, dsrc = device_derivatives(device, names, xin, lib, rel_step)
    for i in range(npar):
        # linear VCCS: d g / d p times the controlling voltage (:63-80)
        for k, vccs in enumerate(device.nD_linVCCS):
            vc = 0.
            if vccs[1] >= 0:
                vc = xVec[vccs[1]]
            if vccs[3] >= 0:
                vc -= xVec[vccs[3]]
            d = vc * dg[k, i]
            if vccs[0] != -1:
                Ji[vccs[0], i] += d
            if vccs[2] != -1:
                Ji[vccs[2], i] -= d
        if device.isDCSource:                       # :81-87
            row = device.nD_sourceOut
            if row[0] >= 0:
                Ji[row[0], i] -= dsrc[i]
            if row[1] >= 0:
                Ji[row[1], i] += dsrc[i]
        if di is not None:                          # :88-90
            nd.set_i(Ji[:, i], device.nD_cpos, device.nD_cneg, di[:, i])
    return Ji
