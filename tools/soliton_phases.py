#!/usr/bin/env python
"""GPU box: in-kernel phase timers of the soliton transient (block kernel),
including the fine-grained timers of the staged triangular sweeps."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ckt, an, tran, x, timeVec, op_it = bench.soliton_setup()
tran._make_engine(an.tstep, x)
e = tran.engine
src = tran.source_table(timeVec)
rows = list(range(ckt.nD_dimension))
e.tran(x, src, rows, im='trap', sync=True)
e.profile = True
e.phase_cycles()
import torch
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
w, it, rs, st = e.tran(x, src, rows, im='trap', sync=False)
t1.record(); torch.cuda.synchronize()
raw = e.prof.cpu().numpy().copy()
c = e.phase_cycles()
tot = sum(v for k, v in c.items() if k in e.PHASES[:7] or e.last_path == 'schur')
print('ms', t0.elapsed_time(t1), 'steps', len(timeVec), 'info', e.info)
for k, v in c.items():
    print('%-14s %12d cycles  %6.2f us/step' % (k, v, v / 1.965e3 / (len(timeVec) - 1)))
print('sum main', tot / 1.965e3 / (len(timeVec) - 1), 'us/step')
if e.last_path == 'schur':
    names = e.SCHUR_PHASES
    tot = 0
    for k, nm in enumerate(names):
        print('%-48s %8.2f us/step' % (nm, raw[k] / 1.965e3 / (len(timeVec) - 1)))
        tot += raw[k]
    print('sum', tot / 1.965e3 / (len(timeVec) - 1), 'us/step')
    sys.exit(0)
# cycles per level of the staged L / U sweeps (slots 16.., 48..), per sweep
nsweep = 2 * (len(timeVec) - 1) - 1
print('L levels:', ' '.join('%d' % (v / nsweep) for v in raw[16:16 + e.info['nlev_l']]))
print('U levels:', ' '.join('%d' % (v / nsweep) for v in raw[48:48 + e.info['nlev_u']]))
