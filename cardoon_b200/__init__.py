"""
cardoon_b200: B200-native (sm_100a) nonlinear Newton hot path behind
cardoon's own API (device plugins, ``op``/``tran`` analyses, netlists).

Python host code drives CUDA kernels through the C-ABI library
``libcardoon_b200.so`` (see include/cardoon_b200.h); there is no CPU
fallback: evaluating a device or running an analysis without the library
and a GPU raises.
"""
from .simulator import (run_netlist, parse_net, run_analyses, new_elem,
                        reset_all)

__all__ = ['run_netlist', 'parse_net', 'run_analyses', 'new_elem',
           'reset_all']
