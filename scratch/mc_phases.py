import sys, numpy as np
sys.path.insert(0, '.')
import bench
from cardoon_b200._lib import get_lib
import torch
lib = get_lib(0)
ckt, an, bt, mc = bench.mc_setup(lib)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
params = bt.pack_instances(B, mc.bsim3_ring_variation())
pd = [lib.dev(p) for p in params]
rows = bt.save_rows(bench.MC_TERMS)
dc, tr = bt.engines(B)
for k in range(3):
    out = bt.run_device(pd, rows, B)
torch.cuda.synchronize()
tr.profile = True
tr.phase_cycles()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); wave, iters, status, opit = bt.run_device(pd, rows, B); e1.record()
torch.cuda.synchronize()
c = tr.phase_cycles()
it0 = int(iters[0].sum().item()); ns = iters.shape[1]
print('ms', e0.elapsed_time(e1), 'iters inst0', it0, 'samples', ns)
for k, v in c.items():
    per = v / (it0 if k in ('device_sweep', 'assembly', 'factor', 'solve', 'update') else ns)
    print('%-14s total %12d  per call %9.0f cycles' % (k, v, per))
print('sum cycles', sum(c.values()), '=', sum(c.values()) / 1.965e6, 'ms at 1965 MHz')
