"""
CPU oracle for the cardoon_b200 hot path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's (cechrist/cardoon) nonlinear Newton
path: forward-mode dual numbers instead of pycppad/CppAD tapes
(devices/cppaddev.py), the device equations (devices/*.py ``eval_cqs``), the
dense and sparse nodal engines (analyses/nodal.py, nodalSP.py), Newton
(analyses/fsolve.py), integration (analyses/integration.py) and the op/tran
drivers (analyses/op.py, tran.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product
(``cardoon_b200``) never does.  The oracle reads circuits built by the
product's netlist/circuit/parameter layer (the part of cardoon that "stays as
is" and is not on the hot path) and re-derives everything from the raw
post-``process_params`` device attributes.

Parity pinning: the reference itself cannot run here (Python 2 + pycppad +
pyparsing, none available) and its ``*_ref.npz`` files are not in its tree.
The oracle is pinned by the reference's printed known-answer test
doc/usage.rst:48-79 (bias_npn.net: 17 iterations and 12-digit node voltages /
OP values), checked in tests/test_oracle_kat.py.  BSIM3, EKV, MESFET,
memristor and the soliton waveforms have no in-tree golden data: for those
the oracle is "parity unpinned" and is cross-checked only by
finite-difference / eval-vs-eval_and_deriv consistency tests.
"""
