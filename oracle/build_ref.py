"""
oracle/_ref loader -- the reference's OWN device modules, executed here
(test infrastructure; never imported by the product).

The reference (cechrist/cardoon) is Python 2 and its device modules import
pycppad, so the package cannot be imported as a whole.  The device modules
themselves (cardoon/devices/{diode,svdiode,bjt,svbjt,mosExt,mosBSIM3v3,mosEKV,
mesfetc,memristor,resistor,autoThermal}.py) are, however, valid Python 3.
This module loads them VERBATIM from ``/root/reference`` (nothing is copied
into this repository) after installing three stand-ins in ``sys.modules``:

* ``cppaddev``  -> ``condassign`` / ``safe_exp`` on ``oracle.dual`` numbers
  (the semantics of devices/cppaddev.py:56-96), ``delete_tape`` a no-op.
  pycppad's ``a_float`` is replaced by forward-mode ``oracle.dual.Dual``;
  numpy ufuncs the device code applies to AD values (``np.exp(x)`` ...) reach
  the dual through ``Dual.__array_ufunc__``.
* ``cardoon.circuit`` / ``cardoon.globalVars`` / ``cardoon.paramset`` -> the
  Python-3 host layer of this repository (the part of cardoon that "stays as
  is": element base class, parameter sets, global options, constants).
* sibling device modules under their bare names (``from diode import
  Junction`` is a Python-2 implicit relative import).

What this gives: ``process_params`` / ``set_temp_vars`` / ``eval_cqs`` of the
reference run unmodified on the same parameter values as the product's
devices, so that (i) every numeric attribute the reference computes can be
compared with the product's host packers and (ii) ``eval_cqs`` values and
Jacobians can be compared with ``oracle.models`` (the restatement the GPU
kernels are tested against).  ``/root/reference`` exists only in the build
container: ``tests/golden/make_ref_device_golden.py`` freezes the outputs
into ``tests/golden/ref_devices.npz`` for the GPU box.

There is nothing to compile, so ``oracle/_ref/`` only receives a marker file
recording which reference files were loaded (``build()``).
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

from . import dual as D

REF = '/root/reference/cardoon'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')

# load order respects the implicit relative imports
MODULES = ['diode', 'svdiode', 'bjt', 'svbjt', 'mosExt', 'mosBSIM3v3',
           'mosEKV', 'mesfetc', 'memristor', 'resistor', 'autoThermal']


def available():
    return os.path.isdir(os.path.join(REF, 'devices'))


def _cppaddev_shim():
    m = types.ModuleType('cppaddev')
    m.__doc__ = 'oracle.dual stand-in for cardoon/devices/cppaddev.py'

    def condassign(b, c, d):                 # cppaddev.py:72-96
        if isinstance(b, D.Dual) or isinstance(c, D.Dual) \
                or isinstance(d, D.Dual) or np.ndim(b) > 0:
            return D.condassign(b, c, d)
        return c if b > 0. else d

    def safe_exp(x):                         # cppaddev.py:56-70
        if isinstance(x, D.Dual) or np.ndim(x) > 0:
            return D.safe_exp(x)
        threshold = 50.
        if x < threshold:
            return np.exp(x)
        return np.exp(threshold) * (x - threshold + 1.)

    def delete_tape(dev):                    # cppaddev.py:99-111
        dev.__dict__.pop('_func', None)

    def eval_and_deriv(dev, vPort):          # cppaddev.py:134-157
        return eval_cqs_dual(dev, np.asarray(vPort, dtype=float)[:, None])

    def eval(dev, vPort):                    # cppaddev.py:160-170
        out, _ = eval_and_deriv(dev, vPort)
        return out
    m.condassign, m.safe_exp, m.delete_tape = condassign, safe_exp, \
        delete_tape
    m.eval_and_deriv, m.eval = eval_and_deriv, eval
    return m


_LOADED = {}


def load():
    """dict name -> reference device module (cached)."""
    if _LOADED:
        return _LOADED
    if not available():
        raise RuntimeError('the reference tree is not present at ' + REF)
    import cardoon_b200.circuit as cir
    import cardoon_b200.globalVars as gv
    import cardoon_b200.paramset as ps
    saved = {k: sys.modules.get(k) for k in
             ['cppaddev', 'cardoon', 'cardoon.circuit', 'cardoon.globalVars',
              'cardoon.paramset'] + MODULES}
    pkg = types.ModuleType('cardoon')
    pkg.__path__ = []
    pkg.circuit, pkg.globalVars, pkg.paramset = cir, gv, ps
    sys.modules.update({'cardoon': pkg, 'cardoon.circuit': cir,
                        'cardoon.globalVars': gv, 'cardoon.paramset': ps,
                        'cppaddev': _cppaddev_shim()})
    try:
        for name in MODULES:
            path = os.path.join(REF, 'devices', name + '.py')
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')      # '\ ' in doc strings
                spec.loader.exec_module(mod)
            _LOADED[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return _LOADED


def device_classes():
    """devType -> reference device class, incl. the ``_t`` classes built by
    the reference's own autoThermal.thermal_device (autoThermal.py:25-138,
    registered as in devices/__init__.py:54-63)."""
    mods = load()
    out = {}
    for name, mod in mods.items():
        if name in ('mosExt', 'autoThermal'):
            continue
        devs = getattr(mod, 'devList', None) or [mod.Device]
        for dev in devs:
            out[dev.devType] = dev
            if dev.makeAutoThermal:
                t = mods['autoThermal'].thermal_device(dev)
                out[t.devType] = t
    return out


def twin(elem):
    """A reference-class element carrying the parameters, model card and
    terminal names of the product element ``elem``, initialised with the
    reference's own ``process_params`` in a scratch circuit."""
    import cardoon_b200.circuit as cir
    cls = device_classes()[elem.devType]
    name = elem.instanceName.split(':', 1)[-1].replace(':', '_')
    r = cls(name)
    r.valueDict = dict(elem.valueDict)
    r.varDict = dict(elem.varDict)
    r.dotModel = elem.dotModel
    ckt = cir.Circuit('_ref_twin_{0}'.format(len(cir.Circuit.cktDict)))
    ckt.add_elem(r)
    ckt.connect(r, [t.instanceName for t in elem.connection[:elem.numTerms]])
    r.init()
    # attributes set from outside after init (sweeps: dc.py:179-188)
    for key in elem.paramDict:
        if key in elem.__dict__ and getattr(r, key, None) != elem.__dict__[key]:
            setattr(r, key, elem.__dict__[key])
            r._dirty = True
    if r.__dict__.pop('_dirty', False):
        r.clean_internal_terms() if not r.isNonlinear else None
        r.process_params()
    return r


def eval_cqs_dual(dev, xin):
    """The reference's ``eval_cqs`` on dual numbers: xin[nxin, n] ->
    (out[nout, n], jac[nout, nxin, n]); what cppaddev.eval_and_deriv returns
    (currents then charges)."""
    xin = np.asarray(xin, dtype=float)
    P, n = xin.shape
    v = np.empty(P, dtype=object)
    for k, s in enumerate(D.seed(xin)):
        v[k] = s
    with np.errstate(all='ignore'):
        res = dev.eval_cqs(v)
    outs = list(res[0]) + list(res[1])
    return D.jac_rows(outs, P, n)


def numeric_attributes(dev):
    """Every float/int/bool attribute (and those of nested Junction-like
    helper objects) an element holds after ``process_params``."""
    out = {}

    def walk(prefix, obj, depth):
        for k, v in vars(obj).items():
            if k in ('connection', 'dotModel', 'valueDict', 'varDict',
                     'paramDict', 'netVar'):
                continue
            key = prefix + k
            if isinstance(v, (bool, int, float, np.floating, np.integer)):
                out[key] = float(v)
            elif isinstance(v, np.ndarray) and v.dtype.kind in 'fi' \
                    and v.size <= 64:
                for i, x in enumerate(v.ravel()):
                    out['{0}[{1}]'.format(key, i)] = float(x)
            elif depth < 1 and hasattr(v, '__dict__') \
                    and type(v).__module__ not in ('builtins',) \
                    and not isinstance(v, (types.ModuleType, type)) \
                    and not hasattr(v, 'instanceName'):
                walk(key + '.', v, depth + 1)
    walk('', dev, 0)
    return out


def build():
    """Record which reference files the loader executes (oracle/_ref/)."""
    if not available():
        return False
    load()
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, 'LOADED'), 'w') as f:
        for name in MODULES:
            f.write(os.path.join(REF, 'devices', name + '.py') + '\n')
    return True
