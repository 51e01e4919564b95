import sys, numpy as np
sys.path.insert(0, '.')
import bench, torch
ckt, an, tran, x, timeVec, op_it = bench.soliton_setup()
rows = list(range(ckt.nD_dimension))
tran._make_engine(an.tstep, x)
e = tran.engine
src = tran.source_table(timeVec)
for k in range(3):
    out = e.tran(x, src, rows, im='trap', sync=False)
torch.cuda.synchronize()
e.profile = True; e.phase_cycles()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); wave, iters, res, status = e.tran(x, src, rows, im='trap', sync=False); e1.record()
torch.cuda.synchronize()
c = e.phase_cycles()
it0 = int(iters[0].sum().item()); ns = iters.shape[1]
print('ms', e0.elapsed_time(e1), 'iters', it0, 'samples', ns, e.info)
for k, v in c.items():
    per = v / (it0 if k in ('device_sweep', 'assembly', 'factor', 'solve', 'update') else ns)
    if k in ('stage_L', 'permute_rhs', 'levels_L', 'stage_U', 'levels_U', 'scatter_dx'): per = v / (it0 + ns)
    print('%-14s total %12d  per call %9.0f cycles = %.1f us' % (k, v, per, per / 1965.))
