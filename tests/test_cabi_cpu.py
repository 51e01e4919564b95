"""
The C-ABI library loads without a GPU and exports every symbol that
include/cardoon_b200.h declares; host-only entry points (version, parameter
counts, host context, plan creation) work; GPU entry points fail loudly
without a device instead of falling back to anything.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests import support as S

ROOT = os.path.dirname(S.HERE)


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'cardoon_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cdn_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def dll():
    import __graft_entry__ as g
    g.build()
    from cardoon_b200._lib import load_dll
    return load_dll()


def test_exports_every_declared_symbol(dll):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(dll, n), 'missing export: ' + n


def test_host_only_entry_points(dll):
    from cardoon_b200.devices.param_layout import MODEL_IDS, LAYOUTS
    assert dll.cdn_version() >= 1
    for name, mid in MODEL_IDS.items():
        if name == 'svdiode':
            continue
        assert dll.cdn_model_nparams(mid) == len(LAYOUTS[name][0])
    ctx = C.c_void_p()
    assert dll.cdn_init_host(C.byref(ctx)) == 0
    assert dll.cdn_sm_count(ctx) == 0
    assert dll.cdn_free(ctx) == 0


def test_no_cpu_fallback():
    """Without a GPU the product path raises; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from cardoon_b200._lib import get_lib, CudaUnavailable
    with pytest.raises(CudaUnavailable):
        get_lib()
    elem = S.make_element('diode', [1, 0], cj0=1e-12)
    with pytest.raises(CudaUnavailable):
        elem.eval_and_deriv(np.array([0.6]))
    ckt, _ = S.load_circuit(S.net('bias_npn.net'))
    from cardoon_b200.analyses import nodalCUDA as nd
    nd.make_nodal_circuit(ckt)
    with pytest.raises(CudaUnavailable):
        nd.DCNodal(ckt)


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, 'cardoon_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text,
                                     flags=re.M), os.path.join(base, f)
