#!/usr/bin/env python
"""
bench.py -- throughput of the cardoon_b200 Newton hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME]
  python bench.py --impl reference ...     (CPU oracle arm, see below)

One JSON line on stdout (rank 0).  Workloads (BASELINE.json `metric`):

  deveval  device evals/s: one step = one cdn_dev_eval launch (values + full
           Jacobian) over n BSIM3 (mosbsim3) instances with per-instance
           parameter columns in HBM (SURVEY 8d "E").
  mc       circuits/s: complete ring_bsim3 transients, one circuit instance
           per Monte-Carlo sample, instances sharded across ranks (8d "M").
  soliton  accepted timepoints/s of test/soliton.net (8d "T"), 1 GPU.

`value` is measured with the inputs resident in HBM; `e2e` is the same
metric through the public host-buffer API with the host<->device copies in
the timed region.  `--impl reference` times the CPU oracle restatement of
the reference's path (the reference itself is Python 2 + pycppad and cannot
run here) on the same workload definition.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get('hbm_gbs', 6650.)), 'measured'
    return 6650., 'fallback'


def measured_traffic(workload):
    """DRAM bytes of the dominant kernel from the committed ncu capture."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        with open(p) as f:
            return json.load(f).get(workload)
    except (OSError, ValueError):
        return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,'
         'clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unsampled'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for ts, line in self.lines:
            if t0 is not None and (ts < t0 - 0.05 or ts > t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unsampled'])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                    power_w_max=float(max(pw)), samples=len(sm),
                    reasons=sorted(reasons))


class Dist(object):
    """torch.distributed plumbing (one process per GPU, NCCL)."""

    def __init__(self, ngpus):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA GPU -- the hot path has no '
                             'CPU implementation (use --impl reference for '
                             'the CPU oracle arm)')
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            os.environ.setdefault('MASTER_PORT', '29511')
            dist.init_process_group('nccl', rank=self.rank,
                                    world_size=self.world,
                                    device_id=torch.device('cuda', self.local))
            self.dist = dist
        if ngpus != self.world:
            if self.rank == 0:
                print('bench.py: --gpus {0} but WORLD_SIZE={1}; using {1}'
                      .format(ngpus, self.world), file=sys.stderr)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64,
                              device='cuda')
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64,
                              device='cuda')
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def phase_shares(D, engine, step):
    """One extra, untimed step with the in-kernel phase timers on: share of
    instance 0's cycles per phase of the time loop."""
    engine.profile = True
    engine.phase_cycles()
    step()
    D.torch.cuda.synchronize()
    c = engine.phase_cycles()
    engine.profile = False
    if getattr(engine, 'last_path', None) == 'schur':
        main = {k: v for k, v in c.items() if k != 'unused'}
    else:
        main = {k: v for k, v in c.items() if k in engine.PHASES[:7]}
    tot = float(sum(main.values())) or 1.
    return {k: round(v / tot, 4) for k, v in main.items()}


def timed_steps(D, step, steps, warmup, flush=None):
    """W untimed + K timed steps; device time (CUDA events), max over ranks.

    Returns (total_ms, wall_t0, wall_t1).  `flush` (optional) runs between
    steps outside the per-step event pairs.
    """
    torch = D.torch
    for _ in range(warmup):
        step()
        if flush is not None:
            flush()
    D.barrier()
    t0 = time.time()
    if flush is None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        ms = 0.
        for _ in range(steps):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            flush()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
    D.barrier()
    t1 = time.time()
    return D.max_over_ranks(ms), t0, t1


# ---------------------------------------------------------------------------
# workload: device evaluation (SURVEY 8d "E")
# ---------------------------------------------------------------------------
def bsim3_param_columns(ncols, seed=1234):
    """Packed mosbsim3 records of the nch and pch models of
    inverter_chain_bsim.net with w ~ U(1u, 30u), l ~ U(0.5u, 2u).

    Returns one group per channel type (their kernel flag words differ:
    the nch card has a substrate-current term)."""
    from tests import support as S
    ckt, _ = S.load_circuit(S.net('inverter_chain_bsim.net'))
    elems = S.nonlinear_elements(ckt)
    rng = np.random.default_rng(seed)
    groups = []
    for tf in (1., -1.):
        e = [x for x in elems if x._tf == tf][0]
        cols, flags = [], None
        for k in range(ncols):
            e.w = float(rng.uniform(1e-6, 30e-6))
            e.l = float(rng.uniform(0.5e-6, 2e-6))
            e.process_params()
            spec = e.cuda_spec()
            flags = spec['flags'] if flags is None else flags
            assert spec['flags'] == flags
            cols.append(spec['params'].copy())
        groups.append(dict(elem=e, tf=tf, model=spec['model'], flags=flags,
                           cols=np.array(cols).T.copy()))
    return groups


def deveval_inputs(n, ncols=256, seed=1234):
    """Two groups (nch, pch) of n/2 instances each."""
    groups = bsim3_param_columns(ncols, seed)
    rng = np.random.default_rng(seed + 1)
    for g in groups:
        m = n // 2
        g['n'] = m
        g['col'] = rng.integers(0, ncols, size=m).astype(np.int64)
        g['xin'] = rng.uniform(-0.5, 3.5, size=(3, m)) * g['tf']
    return dict(groups=groups, n=2 * (n // 2), nout=6, nxin=3)


def gp_inputs(n, seed=4321):
    """Gummel-Poon `bjt` of test/oscillator.net (model t122), one shared
    record; biases vbe ~ U(-1, 0.9), vbc ~ U(-5, 0.9), other ports ~ U(-1, 1)
    (SURVEY 8d "E")."""
    from tests import support as S
    ckt, _ = S.load_circuit(S.net('oscillator.net'))
    e = [x for x in S.nonlinear_elements(ckt) if x.devType == 'bjt'][0]
    spec = e.cuda_spec()
    rng = np.random.default_rng(seed)
    nxin = spec['nxin']
    xin = rng.uniform(-1., 1., size=(nxin, n))
    xin[0] = rng.uniform(-1., 0.9, size=n)
    xin[1] = rng.uniform(-5., 0.9, size=n)
    return dict(elem=e, spec=spec, xin=xin, n=n, nxin=nxin,
                nout=spec['ncs'] + spec['nqs'])


def thermal_inputs(n, seed=987):
    """Electro-thermal svbjt_t of test/npn_thermal.net (x1, x2, v_ib, v_sub,
    dT): the `_t` kernels rebuild the temperature laws per evaluation."""
    from tests import support as S
    ckt, _ = S.load_circuit(S.net('npn_thermal.net'))
    e = S.nonlinear_elements(ckt)[0]
    spec = e.cuda_spec()
    rng = np.random.default_rng(seed)
    xin = rng.uniform(-1., 1., size=(spec['nxin'], n))
    xin[0] = rng.uniform(-20., 45., size=n)
    xin[1] = rng.uniform(-50., 30., size=n)
    xin[-1] = rng.uniform(-20., 100., size=n)
    return dict(elem=e, spec=spec, xin=xin, n=n, nxin=spec['nxin'],
                nout=spec['ncs'] + spec['nqs'],
                label='svbjt_t (electro-thermal state-variable Gummel-Poon, '
                      'npn_thermal.net)')


def run_gp_leg(lib, torch, n, steps, warmup, D, g=None):
    """Second half of BASELINE.json's device-eval metric: GP evals/s with
    the same timing rules; reported inside the deveval line's config.  Also
    used for the electro-thermal leg (g = thermal_inputs(n))."""
    g = g or gp_inputs(n)
    spec = g['spec']
    # per-instance records in HBM like the BSIM3 leg (identical columns)
    params = lib.dev(spec['params'].reshape(-1, 1)).expand(-1, n).contiguous()
    xin = lib.dev(g['xin'])
    out = lib.empty((g['nout'], n))
    jac = lib.empty((g['nout'] * g['nxin'], n))

    def step():
        lib.dev_eval(spec['model'], spec['flags'], params, xin, g['nout'],
                     want_jac=True, out=out, jac=jac)

    ms, _, _ = timed_steps(D, step, steps, warmup)
    from oracle import dual as OD
    from oracle.models import device_eval
    OD.reset_ops()
    o_ref, j_ref = device_eval(g['elem'], g['xin'][:, :2048], True)
    flop, trans = OD.OPS['flop'], OD.OPS['trans']
    torch.cuda.synchronize()
    o = out[:, :2048].cpu().numpy()
    err = float(np.max(np.abs(o - o_ref) / (np.abs(o_ref) + 1e-30)))
    return dict(model=g.get('label', 'bjt (Gummel-Poon, oscillator.net t122)'),
                instances=n,
                nxin=g['nxin'], nout=g['nout'],
                evals_per_s=n * steps / (ms * 1e-3), ms_per_step=ms / steps,
                flop_per_eval=flop, transcendental_per_eval=trans,
                max_rel_err_vs_oracle_first_2048=err)


def bsim3_flops(inp, nsample=2048):
    """Algorithmic FLOPs of one eval (oracle op counter)."""
    from oracle import dual as OD
    from oracle.models import device_eval
    OD.reset_ops()
    for g in inp['groups']:
        device_eval(g['elem'], g['xin'][:, :nsample], True)
    # the counter tallies one scalar operation per vectorised numpy call
    k = float(len(inp['groups']))
    return OD.OPS['flop'] / k, OD.OPS['trans'] / k


def run_deveval(args, D):
    torch = D.torch
    from cardoon_b200._lib import get_lib
    lib = get_lib(D.local)
    inp = deveval_inputs(args.n or (1 << 22), seed=1234 + D.rank)
    lib.dll.cdn_set_tuning(0, args.minblocks)
    lib.dll.cdn_set_tuning(1, args.npre)
    lib.dll.cdn_set_tuning(2, args.ring)
    n, nout, nxin = inp['n'], inp['nout'], inp['nxin']
    G = []
    for g in inp['groups']:
        colidx = torch.as_tensor(g['col'], device=lib.device)
        m = g['n']
        G.append(dict(model=g['model'], flags=g['flags'], n=m,
                      params=lib.dev(g['cols']).index_select(1, colidx)
                      .contiguous(),
                      xin_h=torch.as_tensor(g['xin']).pin_memory(),
                      xin=lib.dev(g['xin']), out=lib.empty((nout, m)),
                      jac=lib.empty((nout * nxin, m)),
                      out_h=torch.empty((nout, m), dtype=torch.float64)
                      .pin_memory(),
                      jac_h=torch.empty((nout * nxin, m),
                                        dtype=torch.float64).pin_memory()))
    npar = G[0]['params'].shape[0]
    pbytes = sum(g['params'].numel() for g in G) * 8
    # devices of a group share a model card: only the rows that depend on
    # w and l differ between instances (split once, outside the timed region:
    # it is part of loading the circuit)
    for g in G:
        g['split'] = lib.split_record(g['params'])
    nvar = float(np.mean([(g['split'][1] >= 0).sum() for g in G]))
    lib.dll.cdn_set_tuning(6, args.split_minb)

    def step_full():
        for g in G:
            lib.dev_eval(g['model'], g['flags'], g['params'], g['xin'], nout,
                         want_jac=True, out=g['out'], jac=g['jac'])

    def step():
        for g in G:
            lib.dev_eval_split(g['model'], g['flags'], g['split'], g['xin'],
                               nout, out=g['out'], jac=g['jac'])

    if args.full_record:
        step = step_full
    # outputs + inputs of one step (~1 GB at 4 M instances) exceed the L2
    ms_full, _, _ = timed_steps(D, step_full, max(2, args.steps // 2),
                                args.warmup)
    ms_full /= max(2, args.steps // 2)
    clk = ClockSampler(D.local)
    clk.start()
    ms, t0, t1 = timed_steps(D, step, args.steps, args.warmup)
    clocks = clk.stop(t0, t1)
    timed_launches = args.steps * len(G)
    total = D.sum_over_ranks(n) * args.steps
    value = total / (ms * 1e-3)

    # ---- end to end: host xin -> device, eval, out + jac -> host ---------
    def e2e_step():
        for g in G:
            g['xin'].copy_(g['xin_h'], non_blocking=True)
            if args.full_record:
                lib.dev_eval(g['model'], g['flags'], g['params'], g['xin'],
                             nout, want_jac=True, out=g['out'], jac=g['jac'])
            else:
                lib.dev_eval_split(g['model'], g['flags'], g['split'],
                                   g['xin'], nout, out=g['out'], jac=g['jac'])
            g['out_h'].copy_(g['out'], non_blocking=True)
            g['jac_h'].copy_(g['jac'], non_blocking=True)

    e2e_steps = max(2, min(args.steps, 5))
    ms_e, _, _ = timed_steps(D, e2e_step, e2e_steps, 1)
    e2e = dict(value=D.sum_over_ranks(n) * e2e_steps / (ms_e * 1e-3),
               unit='device evals/s',
               h2d_bytes_per_step=int(n * nxin * 8),
               d2h_bytes_per_step=int(n * (nout + nout * nxin) * 8))

    # every rank runs the Gummel-Poon leg (it holds a barrier)
    gp = run_gp_leg(lib, torch, n, args.steps, args.warmup, D)
    th = run_gp_leg(lib, torch, n // 4, args.steps, args.warmup, D,
                    thermal_inputs(n // 4))
    line = None
    if D.rank == 0:
        flop, trans = bsim3_flops(inp)
        peak = lib.fp64_peak()
        hbm, how = measured_peaks()
        # rows the evaluation reads: junction records whose mosExt flag
        # (AD, PD, ASRC, PS) is off are not touched
        from cardoon_b200.devices.param_layout import NJ, flag_bits
        fb = flag_bits('bsim3')
        nrows = np.mean([npar - NJ * sum(1 for f in ('AD', 'PD', 'ASRC', 'PS')
                                         if not g['flags'] & fb[f])
                         for g in G])
        bytes_full = 8. * (nrows + nxin + nout + nout * nxin)
        bytes_eval = bytes_full if args.full_record else \
            8. * (nvar + nxin + nout + nout * nxin)
        t_step = ms * 1e-3 / args.steps
        ach = n * flop / t_step / 1e12
        tr = measured_traffic('deveval')
        # roofline: the algorithmic intensity F / bytes_eval decides which
        # roof binds (ridge = FP64 peak / HBM peak)
        ach_bw = n * bytes_eval / t_step / 1e9
        fp64 = dict(achieved=ach, peak=peak / 1e12, unit='TFLOP/s',
                    frac=ach / (peak / 1e12),
                    peak_source='cdn_fp64_peak (DFMA microbenchmark, this '
                                'run)', flop_per_eval=flop,
                    transcendental_per_eval=trans,
                    pipe_fp64_active_pct_ncu=(tr or {}).get(
                        'pipe_fp64_active_pct'))
        hbm_r = dict(achieved=ach_bw, peak=hbm, unit='GB/s',
                     frac=ach_bw / hbm, bytes_per_eval=bytes_eval,
                     peak_source=how)
        intensity = flop / bytes_eval
        ridge = peak / (hbm * 1e9)
        bound = 'hbm' if intensity < ridge else 'fp64'
        roof = dict(hbm_r if bound == 'hbm' else fp64)
        roof.update(bound=bound, intensity_flop_per_byte=intensity,
                    ridge_flop_per_byte=ridge,
                    traffic=(tr['bytes_per_instance'] * n / len(G)
                             if tr else None),
                    traffic_source=(tr or {}).get('source'),
                    kernel=('bsim3_ring_kernel<128,3,48>' if args.full_record
                            else 'bsim3_split_kernel<256,2>')
                    + ' (one launch per channel type)', fp64=fp64, hbm=hbm_r)
        cpu = cpu_baseline_deveval(inp)
        gp['fp64_frac'] = gp['evals_per_s'] * gp['flop_per_eval'] / peak
        th['fp64_frac'] = th['evals_per_s'] * th['flop_per_eval'] / peak
        line = dict(metric='device evals/s (BSIM3v3 mosbsim3, value + '
                           'Jacobian)', value=value, unit='device evals/s',
                    n_gpus=D.world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling='weak', vs_baseline=None, dtype='f64',
                    data='synthetic',
                    config=dict(workload='deveval: {0} mosbsim3 instances per '
                                'GPU (half nch, half pch of inverter_chain_'
                                'bsim.net, w and l per instance); record of '
                                '{1} doubles: {2}'.format(
                                    n, npar,
                                    'every row per instance in HBM'
                                    if args.full_record else
                                    '{0:.0f} rows per instance in HBM, the '
                                    'rest once per model card'.format(nvar)),
                                instances_per_gpu=n,
                                full_record=dict(
                                    evals_per_s=n / (ms_full * 1e-3),
                                    bytes_per_eval=bytes_full,
                                    hbm_frac=n * bytes_full / (ms_full * 1e-3)
                                    / 1e9 / hbm,
                                    note='every record row per instance '
                                         '(bsim3_ring_kernel), the round-1 '
                                         'configuration'),
                                gummel_poon=gp, electro_thermal=th,
                                l2='inputs larger than L2 ({0:.1f} GB of '
                                   'parameters per step)'.format(pbytes
                                                                 / 1e9)),
                    clocks=clocks, e2e=e2e, gpu_launches=timed_launches,
                    roofline=roof, cpu_baseline=cpu)
    return line


# ---------------------------------------------------------------------------
# workload: AC sweep
# ---------------------------------------------------------------------------
def ac_setup(nf):
    """test/sum741_profile.net (741 op-amp summing amplifier, 26 svbjt, 189
    unknowns): operating point on the device, then the operands of a
    log sweep of nf points from 100 Hz to 1 MHz."""
    from tests import support as S
    from cardoon_b200.analyses import nodalCUDA as nd
    from cardoon_b200.analyses.fsolve import solve
    path = os.path.join(os.path.dirname(S.net('x')), 'reference_tests',
                        'sum741_profile.net')
    ckt, _ = S.load_circuit(path)
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    x, res, it = solve(dc.get_guess(), dc.get_source(), dc.convergence_helpers)
    dc.save_OP(x)
    f = np.logspace(2., 6., nf)
    return ckt, path, f, nd.ac_operands(ckt, f)


def run_ac(args, D):
    torch = D.torch
    from cardoon_b200._lib import get_lib
    lib = get_lib(D.local)
    nf = args.n or 2048
    lib.dll.cdn_set_tuning(4, args.ac_blocks)
    lib.dll.cdn_set_tuning(5, args.ac_tpb)
    ckt, path, f, ops = ac_setup(nf)
    n = ckt.nD_dimension
    st = lib.ac_prepare(*ops)

    def step():
        lib.ac_launch(st)

    clk = ClockSampler(D.local)
    clk.start()
    ms, t0, t1 = timed_steps(D, step, args.steps, args.warmup)
    clocks = clk.stop(t0, t1)
    x = lib.ac_fetch(st)
    value = D.world * nf * args.steps / (ms * 1e-3)

    def e2e_step():
        lib.ac_solve(*ops)          # host operands in, host result out

    ne = max(2, min(args.steps, 5))
    ms_e, _, _ = timed_steps(D, e2e_step, ne, 1)
    e2e = dict(value=D.world * nf * ne / (ms_e * 1e-3),
               unit='frequency points/s',
               h2d_bytes_per_step=int(2 * n * n * 8 + nf * 8 + n * 16),
               d2h_bytes_per_step=int(nf * n * 16))
    line = None
    if D.rank == 0:
        peak = lib.fp64_peak()
        # complex LU: n^3/3 multiply-adds of 8 real flops, two triangular
        # solves of n^2 / 2 each
        flop = 8. * (n ** 3 / 3. + n ** 2)
        t_step = ms * 1e-3 / args.steps
        ach = nf * flop / t_step / 1e12
        roof = dict(bound='fp64', achieved=ach, peak=peak / 1e12,
                    unit='TFLOP/s', frac=ach / (peak / 1e12), traffic=None,
                    kernel='ac_solve_kernel (one CTA per frequency, matrix '
                           'in the L2-resident workspace)',
                    peak_source='cdn_fp64_peak (DFMA microbenchmark, this '
                                'run)', flop_per_frequency=flop,
                    workspace_bytes=int(nf * n * n * 16))
        # CPU: the oracle's per-frequency loop on a sample of the sweep
        from oracle import nodal_ref as R
        from tests import support as S
        from cardoon_b200.globalVars import glVar
        ckt2, _ = S.load_circuit(path)
        ns = 64
        tc = time.time()
        ref = R.run_ac(ckt2, glVar, f[:: nf // ns][:ns])
        tc = time.time() - tc
        err = float(np.max(np.abs(x[:: nf // ns][:ns].T - ref['X']))
                    / np.max(np.abs(ref['X'])))
        cpu = dict(value=ns / tc, unit='frequency points/s', cores=1,
                   kind='port',
                   sample='{0} of the {1} frequencies incl. the operating '
                          'point (numpy oracle: per-frequency assembly + '
                          'scipy.linalg.solve as nodal.py:346-372)'.format(
                              ns, nf))
        line = dict(metric='AC frequency points/s', value=value,
                    unit='frequency points/s', n_gpus=D.world,
                    steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling='weak', vs_baseline=None, dtype='c128',
                    data='synthetic',
                    config=dict(workload='ac: test/sum741_profile.net '
                                '(N={0}), {1}-point log sweep 100 Hz - 1 MHz '
                                'per GPU; one step = one cdn_ac_solve launch'
                                .format(n, nf), unknowns=n, frequencies=nf,
                                max_rel_err_vs_oracle=err,
                                l2='per-frequency matrices ({0:.2f} GB) '
                                   'exceed L2'.format(nf * n * n * 16 / 1e9)),
                    clocks=clocks, e2e=e2e, gpu_launches=args.steps,
                    roofline=roof, cpu_baseline=cpu)
    return line


def cpu_baseline_deveval(inp, nsample=1 << 15, seconds=10.):
    from oracle.models import device_eval
    t0 = time.time()
    reps = 0
    while time.time() - t0 < seconds:
        for g in inp['groups']:
            device_eval(g['elem'], g['xin'][:, :nsample], True)
        reps += 1
    dt = time.time() - t0
    return dict(value=2 * nsample * reps / dt, unit='device evals/s', cores=1,
                kind='port',
                sample='{0} x {1} mosbsim3 evals (numpy dual-number oracle, '
                       'vectorised over instances)'.format(2 * reps, nsample))


# ---------------------------------------------------------------------------
# workload: Monte Carlo of ring_bsim3 transients (SURVEY 8d "M", config 5)
# ---------------------------------------------------------------------------
MC_TERMS = ['1', '3', '4']


def count_flops(elem, xin, deriv):
    """Oracle op counter for one eval of ``elem`` (values only: zero tangent
    lanes)."""
    from oracle import dual as OD
    from oracle import models as OM
    xin = np.asarray(xin, dtype=float)
    if xin.ndim == 1:
        xin = xin[:, None]
    P, n = xin.shape
    OD.reset_ops()
    with np.errstate(all='ignore'):
        if deriv:
            OM.eval_cqs(elem, OD.seed(xin))
        else:
            OM.eval_cqs(elem, [OD.Dual(xin[k].copy(), np.zeros((0, n)))
                               for k in range(P)])
    return float(OD.OPS['flop']), float(OD.OPS['trans'])


def mc_setup(lib=None):
    from tests import support as S
    from cardoon_b200.analyses import mc
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = [a for a in queue if a.anType == 'tran'][0]
    bt = mc.BatchTransient(ckt, an.tstop, an.tstep, an.im, lib=lib)
    return ckt, an, bt, mc


def oracle_mc_instance(args):
    """One Monte-Carlo instance on the CPU oracle (reference mechanism:
    setattr + process_params, then the reference's tran loop)."""
    j, max_samples = args
    from tests import support as S
    from cardoon_b200.analyses import mc
    from cardoon_b200.globalVars import glVar
    from oracle import nodal_ref as R
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = [a for a in queue if a.anType == 'tran'][0]
    elems = S.nonlinear_elements(ckt)
    mc.bsim3_ring_variation()(j, elems)
    for e in elems:
        e.process_params()
    t0 = time.time()
    r = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im,
                   max_samples=max_samples)
    dt = time.time() - t0
    nfull = len(np.arange(0., an.tstop, an.tstep))
    return dt, len(r['t']), nfull, int(r['iters'].sum())


def oracle_mc_instance_safe(args):
    try:
        return oracle_mc_instance(args)
    except Exception as err:                  # reported, never fatal
        return ('failed', args[0], '{0}: {1}'.format(type(err).__name__, err))


def cpu_baseline_mc(procs=1, instances=(0,), max_samples=100):
    """CPU oracle on a bounded sample: the listed instances (a fixed list,
    whatever the core count) for the first ``max_samples`` samples, on
    ``procs`` worker processes; circuits/s is EXTRAPOLATED to complete
    transients (the time loop costs the same per sample)."""
    import multiprocessing as mp
    jobs = [(int(j), max_samples) for j in instances]
    procs = max(1, min(procs, len(jobs)))
    t0 = time.time()
    if procs > 1:
        with mp.get_context('spawn').Pool(procs) as pool:
            out = pool.map(oracle_mc_instance_safe, jobs)
    else:
        out = [oracle_mc_instance_safe(j) for j in jobs]
    wall = time.time() - t0
    good = [o for o in out if o[0] != 'failed']
    bad = [o for o in out if o[0] == 'failed']
    if not good:
        return dict(value=0., unit='circuits/s', cores=procs, kind='port',
                    sample='all {0} oracle instances failed: {1}'.format(
                        len(jobs), bad[0][2]))
    busy = sum(o[0] for o in good)
    frac = good[0][1] / float(good[0][2])
    # complete-transient equivalents finished per second by `procs` workers
    value = len(good) * frac / (busy / procs)
    return dict(value=value, unit='circuits/s', cores=procs, kind='port',
                sample='instances {0}..{1} x first {2} of {3} samples on {4} '
                       'worker process(es) (numpy oracle restatement of the '
                       'reference tran loop; EXTRAPOLATED to complete '
                       'transients; wall {5:.1f} s; {6} failed)'.format(
                           jobs[0][0], jobs[-1][0], good[0][1], good[0][2],
                           procs, wall, len(bad)),
                failed_instances=[[o[1], o[2]] for o in bad])


MC_TOTAL = 4096


def mc_roofline(lib, ckt, bt, B, nsamples, iters_sum, opit_sum, t_step):
    elems = ckt.nD_nlinElem
    xs = np.array([[2.5, 1.2, 0.]]).T
    fd = np.mean([count_flops(e, xs * e._tf, True)[0] for e in elems])
    fv = np.mean([count_flops(e, xs * e._tf, False)[0] for e in elems])
    n = bt.topo.n
    ndev = len(elems)
    it_total = iters_sum + opit_sum
    lu_flop = 2. / 3. * n ** 3 + 2. * n * n
    # algorithmic work (the accept step needs values only, whatever code
    # variant computes them)
    flop = it_total * (ndev * fd + lu_flop) \
        + B * nsamples * (ndev * fv + 2. * n * n)
    peak = lib.fp64_peak()
    ach = flop / t_step / 1e12
    tr = measured_traffic('mc')
    return dict(bound='fp64', achieved=ach, peak=peak / 1e12,
                unit='TFLOP/s', frac=ach / (peak / 1e12),
                traffic=tr['bytes_per_instance'] * B if tr else None,
                traffic_source=(tr or {}).get('source'),
                pipe_fp64_active_pct_ncu=(tr or {}).get(
                    'pipe_fp64_active_pct'),
                kernel='engine_tile_kernel<BSIM3, 8> (whole time loop, '
                       'sd8 path)',
                peak_source='cdn_fp64_peak (DFMA microbenchmark, this run)',
                flop_per_eval_jac=fd, flop_per_eval_values=fv,
                newton_iterations=it_total)


def run_mc(args, D):
    """MC circuits/s.  Default = the configuration BASELINE.json names: 4096
    instances in total, sharded over the ranks (strong scaling, mc.shard);
    ``--weak`` gives every rank 4096 instances of its own."""
    torch = D.torch
    from cardoon_b200._lib import get_lib
    lib = get_lib(D.local)
    ckt, an, bt, mc = mc_setup(lib)
    lib.dll.cdn_set_tuning(3, args.tile_tpb)
    total = args.n or MC_TOTAL
    if args.weak:
        lo, hi = D.rank * total, (D.rank + 1) * total
        job = total * D.world
    else:
        lo, hi = mc.shard(total, D.rank, D.world)
        job = total
    B = hi - lo
    elems = ckt.nD_nlinElem
    # the Monte-Carlo sample (standard-normal draws of every instance) is the
    # input data set; records are packed from it on the host
    z = mc.bsim3_ring_samples(range(lo, hi))
    over = mc.bsim3_ring_overrides(z, elems)
    rows = bt.save_rows(MC_TERMS)
    nsamples = len(bt.timeVec)
    shapes = [o.shape for o in bt.pack_batch(over, device_major=True)]
    pinned = [[torch.empty(sh, dtype=torch.float64).pin_memory()
               for sh in shapes] for _ in range(2)]
    bt.pack_batch(over, out=[t.numpy() for t in pinned[0]], device_major=True)
    params_d = [t.to(lib.device) for t in pinned[0]]
    dc, tr = bt.engines(B)
    state = {}

    def step():
        state['out'] = bt.run_device(params_d, rows, B)

    clk = ClockSampler(D.local)
    clk.start()
    l0 = dc.launches + tr.launches
    ms, t0, t1 = timed_steps(D, step, args.steps, args.warmup)
    launches = (dc.launches + tr.launches - l0) \
        * args.steps // (args.steps + args.warmup)
    clocks = clk.stop(t0, t1)
    wave, iters, status, opit = state['out']
    nfail = int(D.sum_over_ranks((status != 0).sum().item()))
    phases = phase_shares(D, tr, step)
    value = job * args.steps / (ms * 1e-3)
    it_sum = D.sum_over_ranks(float(iters.sum().item()))
    op_sum = D.sum_over_ranks(float(opit.sum().item()))
    helper_op = int(D.sum_over_ranks((bt.helper_used[0] != 0).sum().item()))

    # ---- end to end through the public API --------------------------------
    # host: raw parameter samples -> BatchTransient.pack_batch (the
    # process_params arithmetic of every instance) into pinned memory ->
    # H2D -> operating points + transients -> one gather over the ranks
    # (NCCL) -> D2H of waveforms and iteration counts.  Double-buffered:
    # while the GPU runs step k the host packs step k + 1.
    gshape = (job,) + tuple(wave.shape[1:])
    wave_h = torch.empty(gshape, dtype=torch.float64).pin_memory()
    iters_h = torch.empty((job, iters.shape[1]), dtype=torch.int32).pin_memory()
    # three streams: H2D of step k + 1 and D2H of step k - 1 run beside the
    # kernels of step k (records and results are double-buffered on the device)
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    params_dd = [params_d, [t.clone() for t in params_d]]
    h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
    run_done = [torch.cuda.Event(), torch.cuda.Event()]
    keep = [None, None]

    def e2e_run(K):
        main = torch.cuda.current_stream()
        for k in range(K):
            buf, pd = pinned[k % 2], params_dd[k % 2]
            if k >= 2:
                h2d_done[k % 2].synchronize()   # pinned buffer consumed
            bt.pack_batch(over, out=[t.numpy() for t in buf],
                          device_major=True)
            with torch.cuda.stream(s_h2d):
                if k >= 2:
                    s_h2d.wait_event(run_done[k % 2])   # run k - 2 has read pd
                for d, h in zip(pd, buf):
                    d.copy_(h, non_blocking=True)
                h2d_done[k % 2].record(s_h2d)
            main.wait_event(h2d_done[k % 2])
            w, it, st, _ = bt.run_device(pd, rows, B)
            w, it, st = mc.gather_results(w, it, st, D.dist)
            run_done[k % 2].record(main)
            keep[k % 2] = (w, it, st)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(run_done[k % 2])
                wave_h.copy_(w, non_blocking=True)
                iters_h.copy_(it, non_blocking=True)
                w.record_stream(s_d2h)
                it.record_stream(s_d2h)
        torch.cuda.synchronize()

    # enough steps that filling and draining the pipeline (one un-overlapped
    # pack + H2D, one D2H) do not dominate
    e2e_steps = max(3, min(2 * args.steps, 20))
    e2e_run(2)
    D.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record()
    e2e_run(e2e_steps)
    e1.record()
    torch.cuda.synchronize()
    ms_e = D.max_over_ranks(max(e0.elapsed_time(e1),
                                (time.time() - tw0) * 1e3))
    D.barrier()
    e2e = dict(value=job * e2e_steps / (ms_e * 1e-3), unit='circuits/s',
               h2d_bytes_per_step=int(sum(t.numel() for t in pinned[0]) * 8),
               d2h_bytes_per_step=int(wave_h.numel() * 8
                                      + iters_h.numel() * 4),
               steps=e2e_steps,
               path='raw samples (host) -> BatchTransient.pack_batch -> '
                    'pinned -> H2D -> run_device -> gather_results (NCCL) '
                    '-> D2H; host packing and H2D of step k+1 and D2H of '
                    'step k-1 overlap the kernels of step k (three streams, '
                    'double buffers); time = max(CUDA events, wall clock)')
    # one un-overlapped pass: what each stage costs
    tp = time.time()
    bt.pack_batch(over, out=[t.numpy() for t in pinned[0]],
                  device_major=True)
    e2e['host_pack_ms'] = (time.time() - tp) * 1e3

    weak = None
    if D.world > 1 and not args.weak:
        # secondary: every rank runs 4096 instances of its own
        zw = mc.bsim3_ring_samples(range(D.rank * total,
                                         (D.rank + 1) * total))
        pw = [lib.dev(p) for p in bt.pack_batch(
            mc.bsim3_ring_overrides(zw, elems), device_major=True)]

        def wstep():
            bt.run_device(pw, rows, total)
        msw, _, _ = timed_steps(D, wstep, 3, 2)
        weak = dict(instances_per_gpu=total,
                    value=total * D.world * 3 / (msw * 1e-3),
                    unit='circuits/s', ms_per_step=msw / 3)

    line = None
    if D.rank == 0:
        roof = mc_roofline(lib, ckt, bt, job, nsamples, it_sum, op_sum,
                           ms * 1e-3 / args.steps * (1 if args.weak else 1))
        if D.world > 1:
            # per-GPU fraction: the job's flops over N GPUs' peak
            roof['frac'] /= D.world
            roof['peak'] *= D.world
            roof['note'] = 'peak = {0} x one GPU'.format(D.world)
        cpu = cpu_baseline_mc(procs=1, instances=(0,), max_samples=120)
        line = dict(metric='MC circuits/s (complete ring_bsim3 transients)',
                    value=value, unit='circuits/s', n_gpus=D.world,
                    steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling='weak' if args.weak else 'strong',
                    vs_baseline=None, dtype='f64', data='synthetic',
                    config=dict(workload='mc: {0}-instance Monte Carlo of '
                                'workspace/ring_bsim3.net (N=7, 6 mosbsim3, '
                                '{1} samples, trapezoidal) with per-instance '
                                'BSIM3 records, {2}; one step = operating '
                                'points + transients of all instances'
                                .format(job, nsamples,
                                        '{0} per GPU'.format(total)
                                        if args.weak else
                                        'sharded over the GPUs '
                                        '(rank r: [r*{0}/G, (r+1)*{0}/G))'
                                        .format(total)),
                                instances_total=job,
                                instances_this_rank=B, samples=nsamples,
                                failed_instances=nfail,
                                op_instances_using_helpers=helper_op,
                                phase_shares_instance0=phases,
                                avg_newton_iterations_per_sample=it_sum
                                / (job * nsamples),
                                weak_scaling=weak,
                                l2='per-instance records ({0:.1f} MB per '
                                   'rank) are L2-resident by nature of the '
                                   'workload; every step re-reads them'
                                   .format(sum(t.numel() for t in pinned[0])
                                           * 8 / 1e6)),
                    clocks=clocks, e2e=e2e, gpu_launches=launches,
                    roofline=roof, cpu_baseline=cpu)
    return line


# ---------------------------------------------------------------------------
# workload: soliton transient (SURVEY 8d "T", config 2)
# ---------------------------------------------------------------------------
def soliton_setup():
    from tests import support as S
    from cardoon_b200.analyses import nodalCUDA as nd
    from cardoon_b200.analyses.fsolve import solve
    from cardoon_b200.analyses.integration import Trapezoidal
    ckt, queue = S.load_circuit(S.net('soliton.net'))
    an = [a for a in queue if a.anType == 'tran'][0]
    nd.make_nodal_circuit(ckt)
    dc = nd.DCNodal(ckt)
    tran = nd.TransientNodal(ckt, Trapezoidal())
    x, res, it = solve(dc.get_guess(), tran.get_source(0.).copy(),
                       dc.convergence_helpers)
    timeVec = np.arange(start=0., stop=an.tstop, step=an.tstep, dtype=float)
    return ckt, an, tran, x, timeVec, it


def cpu_baseline_soliton(max_samples=150):
    from tests import support as S
    from cardoon_b200.globalVars import glVar
    from oracle import nodal_ref as R
    ckt, queue = S.load_circuit(S.net('soliton.net'))
    an = [a for a in queue if a.anType == 'tran'][0]
    idx = R.prepare(ckt)
    t0 = time.time()
    r = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im, idx=idx,
                   max_samples=max_samples)
    dt = time.time() - t0
    return dict(value=(len(r['t']) - 1) / dt, unit='timepoints/s', cores=1,
                kind='port',
                sample='first {0} samples of test/soliton.net incl. the '
                       'operating point (numpy oracle; SuperLU refactored '
                       'every iteration as nodalSP.py:281-303)'.format(
                           len(r['t'])))


def run_soliton(args, D):
    torch = D.torch
    ckt, an, tran, x, timeVec, op_it = soliton_setup()
    rows = list(range(ckt.nD_dimension))
    nsamples = len(timeVec)
    tran._make_engine(an.tstep, x)
    e = tran.engine
    lib = e.lib
    src = tran.source_table(timeVec)
    state = {}

    x_dev = lib.dev(x)      # `value`: inputs resident in HBM (no blocking copy
                            # from pageable memory between two launches)

    def step():
        state['out'] = e.tran(x_dev, src, rows, im='trap', sync=False,
                              reuse_src=True)

    # first pass allocates; reuse buffers afterwards
    clk = ClockSampler(D.local)
    clk.start()
    l0 = e.launches
    ms, t0, t1 = timed_steps(D, step, args.steps, args.warmup)
    launches = (e.launches - l0) * args.steps // (args.steps + args.warmup)
    clocks = clk.stop(t0, t1)
    wave, iters, res, status = state['out']
    assert int(status[0].item()) == 0, 'soliton transient did not converge'
    phases = phase_shares(D, e, step)
    value = D.world * (nsamples - 1) * args.steps / (ms * 1e-3)

    var_rows = tran.source_rows()

    def e2e_step():
        w, it, rs, st = e.tran(x, src, rows, im='trap', sync=True,
                               var_rows=var_rows)

    ms_e, _, _ = timed_steps(D, e2e_step, max(2, min(args.steps, 5)), 1)
    e2e = dict(value=D.world * (nsamples - 1) * max(2, min(args.steps, 5))
               / (ms_e * 1e-3), unit='timepoints/s',
               h2d_bytes_per_step=int(src.size * 8 + x.size * 8),
               d2h_bytes_per_step=int(wave.numel() * 8 + iters.numel() * 4
                                      + res.numel() * 8))
    line = None
    if D.rank == 0:
        info = e.info
        n, nlu, nnz = info['n'], info['nlu'], info['nnzA']
        it_total = float(iters.sum().item())
        b_asm = 8. * (nnz + 2 * n) + 4. * (info['ncon'] + nnz)
        b_lu = 8. * (nnz + 2 * nlu)
        b_solve = 8. * (nlu + 3 * n)
        per_iter = b_asm + b_lu + b_solve
        per_sample = b_solve + 8. * 8 * n        # chord + accept/rhs vectors
        algo = it_total * per_iter + (nsamples - 1) * per_sample
        t_step = ms * 1e-3 / args.steps
        hbm, how = measured_peaks()
        ach = algo / t_step / 1e9
        tr = measured_traffic('soliton')
        roof = dict(bound='hbm', achieved=ach, peak=hbm, unit='GB/s',
                    frac=ach / hbm,
                    traffic=tr['bytes_per_launch'] if tr else None,
                    traffic_source=(tr or {}).get('source'),
                    kernel=('engine_schur_kernel<DIODE, 512> (whole time loop '
                            'on one 8-CTA cluster: interior operators, state '
                            'and the boundary system in shared memory; the '
                            'algorithmic bytes are those of the reference\'s '
                            'refactor-per-iteration path, SURVEY 8d)'
                            if e.last_path == 'schur' else
                            'engine_{0}_kernel<DIODE> (whole time loop, '
                            'state L2-resident)'.format(e.last_team())),
                    peak_source=how, bytes_per_newton_iteration=per_iter,
                    newton_iterations=it_total, nnzJ=nnz, nnzLU=nlu,
                    levels=dict(factor=info['nlev_f'], L=info['nlev_l'],
                                U=info['nlev_u']))
        cpu = cpu_baseline_soliton()
        line = dict(metric='soliton tran timepoints/s', value=value,
                    unit='timepoints/s', n_gpus=D.world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps,
                    higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f64', data='synthetic',
                    config=dict(workload='soliton: test/soliton.net transient '
                                '(N=3022, 47 diodes, {0} samples, '
                                'trapezoidal, h=0.1 ps); one step = the whole '
                                'time loop after the operating point; with '
                                '--gpus N every rank runs a replica'
                                .format(nsamples), samples=nsamples,
                                op_iterations=op_it,
                                avg_newton_iterations=it_total / nsamples,
                                phase_shares=phases,
                                path=e.last_path,
                                l2='single-circuit state and operators '
                                   '(~1.5 MB) live in the shared memory of '
                                   'the cluster; bound by the latency of '
                                   'dependent FP64 operations, not by bytes'),
                    clocks=clocks, e2e=e2e, gpu_launches=launches,
                    roofline=roof, cpu_baseline=cpu)
    return line


# ---------------------------------------------------------------------------
# workload: synthetic BSIM3 inverter chain
# ---------------------------------------------------------------------------
def chain_setup(K):
    import io
    import contextlib
    from tests import support as S
    from cardoon_b200.analyses import nodalCUDA as nd
    path = S.chain_netlist(K)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ckt, queue = S.load_circuit(path)
    finally:
        os.remove(path)
    an = [a for a in queue if a.anType == 'tran'][0]
    nd.make_nodal_circuit(ckt)
    return ckt, an, nd, S.chain_nodeset(ckt, K)


def cpu_baseline_chain(K_full, K=150, max_samples=8):
    from tests import support as S
    from cardoon_b200.globalVars import glVar
    from oracle import nodal_ref as R
    import io
    import contextlib
    path = S.chain_netlist(K)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ckt, queue = S.load_circuit(path)
    finally:
        os.remove(path)
    an = [a for a in queue if a.anType == 'tran'][0]
    idx = R.prepare(ckt)
    x0 = np.zeros(idx.dimension)
    rows = idx.name_to_row()
    x0[rows['1']] = 3.
    for k in range(1, K + 1, 2):
        x0[rows['o{0}'.format(k)]] = 3.
    t0 = time.time()
    r = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im, idx=idx,
                   max_samples=max_samples, x0=x0)
    dt = time.time() - t0
    tps_small = (len(r['t']) - 1) / dt
    return dict(value=tps_small * K / float(K_full), unit='timepoints/s',
                cores=1, kind='port',
                sample='{0}-stage chain, first {1} samples: {2:.3f} '
                       'timepoints/s, scaled by {0}/{3} (the oracle\'s cost '
                       'is linear in the number of devices)'.format(
                           K, len(r['t']), tps_small, K_full))


def run_chain(args, D):
    torch = D.torch
    K = args.n or 100000
    t_setup = time.time()
    ckt, an, nd, x0 = chain_setup(K)
    from cardoon_b200.analyses.integration import Trapezoidal
    from cardoon_b200.analyses.fsolve import NoConvergenceError
    topo = nd.get_topology(ckt)
    helper = nd.TransientNodal(ckt, Trapezoidal())
    timeVec = np.arange(start=0., stop=an.tstop, step=an.tstep, dtype=float)
    src = helper.source_table(timeVec)
    dc = nd.Engine(topo, use_q=False, x_ref=x0)
    r = dc._run_struct(nd.L.OP_NEWTON, helpers=1)
    dc.x.copy_(torch.as_tensor(x0[None, :]))
    dc.sv.copy_(torch.as_tensor(src[0]))
    dc._launch(r)
    if int(dc.status[0].item()) != 0:
        raise NoConvergenceError('chain operating point did not converge')
    op_it = int(dc.iters[0].item())
    x = dc.x[0].cpu().numpy()
    e = nd.Engine(topo, use_q=True, a0=2. / an.tstep, x_ref=x)
    del dc
    t_setup = time.time() - t_setup
    rows = [ckt.termDict[nm].nD_namRC
            for nm in ('o1', 'o2', 'o{0}'.format(K // 2), 'o{0}'.format(K))]
    nsamples = len(timeVec)
    state = {}

    def step():
        state['out'] = e.tran(x, src, rows, im='trap', sync=False)

    clk = ClockSampler(D.local)
    clk.start()
    l0 = e.launches
    ms, t0, t1 = timed_steps(D, step, args.steps, args.warmup)
    launches = (e.launches - l0) * args.steps // (args.steps + args.warmup)
    clocks = clk.stop(t0, t1)
    wave, iters, res, status = state['out']
    assert int(status[0].item()) == 0, 'chain transient did not converge'
    phases = phase_shares(D, e, step)
    value = D.world * (nsamples - 1) * args.steps / (ms * 1e-3)

    def e2e_step():
        e.tran(x, src, rows, im='trap', sync=True)

    ne = max(1, min(args.steps, 2))
    ms_e, _, _ = timed_steps(D, e2e_step, ne, 1)
    e2e = dict(value=D.world * (nsamples - 1) * ne / (ms_e * 1e-3),
               unit='timepoints/s',
               h2d_bytes_per_step=int(src.size * 8 + x.size * 8),
               d2h_bytes_per_step=int(wave.numel() * 8 + iters.numel() * 4
                                      + res.numel() * 8))
    line = None
    if D.rank == 0:
        info = e.info
        n, nlu, nnz = info['n'], info['nlu'], info['nnzA']
        it_total = float(iters.sum().item())
        ndev = sum(g['n'] for g in topo.groups)
        npar = len(topo.groups[0]['elems'][0].cuda_spec()['params'])
        # evaluation and stamping are fused and the devices' records
        # are stored once per distinct card + geometry (two columns here,
        # addressed through an index), so device records add no per-iteration
        # HBM term; what moves is the control voltages in and the outputs and
        # Jacobians out of the state buffers
        nout = sum(g['n'] * (g['ncs'] + g['nqs']) for g in topo.groups)
        njac = sum(g['n'] * (g['ncs'] + g['nqs']) * g['nxin']
                   for g in topo.groups)
        b_eval = 8. * (nout + njac) + 4. * ndev
        b_asm = 8. * (nnz + 2 * n) + 4. * (info['ncon'] + nnz)
        b_lu = 8. * (nnz + 2 * nlu)
        b_solve = 8. * (nlu + 3 * n)
        per_iter = b_eval + b_asm + b_lu + b_solve
        per_sample = 8. * nout + 4. * ndev + b_solve + 8. * 8 * n
        algo = it_total * per_iter + (nsamples - 1) * per_sample
        t_step = ms * 1e-3 / args.steps
        hbm, how = measured_peaks()
        ach = algo / t_step / 1e9
        roof = dict(bound='hbm', achieved=ach, peak=hbm, unit='GB/s',
                    frac=ach / hbm, traffic=None,
                    kernel='engine_{0}_kernel<BSIM3, 256> (whole time loop)'
                    .format(e.last_team()),
                    peak_source=how, bytes_per_newton_iteration=per_iter,
                    newton_iterations=it_total, nnzJ=nnz, nnzLU=nlu,
                    devices=ndev,
                    levels=dict(factor=info['nlev_f'], L=info['nlev_l'],
                                U=info['nlev_u']))
        cpu = cpu_baseline_chain(K)
        line = dict(metric='inverter-chain tran timepoints/s', value=value,
                    unit='timepoints/s', n_gpus=D.world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps,
                    higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f64', data='synthetic',
                    config=dict(workload='chain: synthetic {0}-stage BSIM3v3 '
                                'inverter chain (N={1}, {2} mosbsim3, {3} '
                                'samples, trapezoidal, h=0.1 ns), operating '
                                'point from the alternating nodeset; one step '
                                '= the whole time loop'.format(
                                    K, n, ndev, nsamples),
                                stages=K, samples=nsamples,
                                op_iterations=op_it,
                                avg_newton_iterations=it_total / nsamples,
                                phase_shares=phases,
                                host_setup_s=t_setup,
                                l2='state {0:.0f} MB (L2: 126 MB); device '
                                   'records de-duplicated to {1} columns'
                                   .format(e.ws_per * 8 / 1e6,
                                           sum(int(t.shape[1])
                                               for t in e.params))),
                    clocks=clocks, e2e=e2e, gpu_launches=launches,
                    roofline=roof, cpu_baseline=cpu)
    return line


# ---------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    common = dict(impl='reference', n_gpus=args.gpus, steps=args.steps,
                  warmup=args.warmup, ms_per_step=None,
                  higher_is_better=True, scaling='weak', vs_baseline=None,
                  dtype='f64', data='synthetic')
    if args.workload == 'deveval':
        cpu = cpu_baseline_deveval(deveval_inputs(1 << 16), seconds=20.)
        metric = 'device evals/s (BSIM3v3 mosbsim3, value + Jacobian)'
        wl = 'deveval'
    elif args.workload == 'mc':
        # a fixed instance list (independent of the core count), every step
        # the same bounded sample; per-instance failures are reported, never
        # fatal
        procs = min(cores, 16)
        steps = []
        for k in range(max(0, args.warmup) + max(1, args.steps)):
            c = cpu_baseline_mc(procs=procs, instances=range(16),
                                max_samples=120)
            if k >= max(0, args.warmup):
                steps.append(c)
        cpu = dict(steps[-1])
        cpu['value'] = float(np.mean([c['value'] for c in steps]))
        cpu['sample'] += '; mean of {0} such steps'.format(len(steps))
        metric = 'MC circuits/s (complete ring_bsim3 transients)'
        wl = ('mc: {0}-instance Monte Carlo of workspace/ring_bsim3.net '
              '(N=7, 6 mosbsim3, 400 samples, trapezoidal) with '
              'per-instance BSIM3 records'.format(args.n or MC_TOTAL))
        common['scaling'] = 'strong'
    elif args.workload == 'chain':
        cpu = cpu_baseline_chain(args.n or 100000, K=300, max_samples=10)
        metric = 'inverter-chain tran timepoints/s'
        wl = 'chain: synthetic BSIM3v3 inverter chain'
    elif args.workload == 'ac':
        from oracle import nodal_ref as R
        from tests import support as S
        from cardoon_b200.globalVars import glVar
        path = os.path.join(os.path.dirname(S.net('x')), 'reference_tests',
                            'sum741_profile.net')
        ckt, _ = S.load_circuit(path)
        ns = 256
        t0 = time.time()
        R.run_ac(ckt, glVar, np.logspace(2., 6., ns))
        cpu = dict(value=ns / (time.time() - t0), unit='frequency points/s',
                   cores=1, kind='port',
                   sample='{0}-point log sweep incl. the operating point '
                          '(numpy oracle, one scipy.linalg.solve per '
                          'frequency as nodal.py:346-372)'.format(ns))
        metric = 'AC frequency points/s'
        wl = 'ac: test/sum741_profile.net'
    else:
        cpu = cpu_baseline_soliton(max_samples=300)
        metric = 'soliton tran timepoints/s'
        wl = 'soliton: test/soliton.net transient'
    v = cpu['value']
    common.update(metric=metric, value=v, unit=cpu['unit'],
                  config=dict(workload=wl), cpu_baseline=cpu,
                  e2e=dict(value=v, unit=cpu['unit'], h2d_bytes_per_step=0,
                           d2h_bytes_per_step=0))
    return common


def main():
    # stdout carries exactly ONE line (the JSON result): everything else any
    # library prints (NCCL banner, netlist warnings) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + '\n').encode())

    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native',
                    choices=['native', 'reference'])
    ap.add_argument('--workload', default=None,
                    choices=['deveval', 'mc', 'soliton', 'chain', 'ac'])
    ap.add_argument('--minblocks', type=int, default=0,
                    help='deveval tuning: CTAs/SM register limit (0 = default)')
    ap.add_argument('--npre', type=int, default=-1,
                    help='deveval tuning: record rows prefetched to L1')
    ap.add_argument('--ring', type=int, default=1,
                    help='deveval tuning: BSIM3 record through the '
                         'shared-memory ring (0 = on-demand loads)')
    ap.add_argument('--full-record', action='store_true',
                    help='deveval: every record row per instance (round-1 '
                         'configuration) instead of the split record')
    ap.add_argument('--split-minb', type=int, default=5,
                    help='deveval tuning: CTAs/SM of the split-record kernel')
    ap.add_argument('--ac-blocks', type=int, default=0,
                    help='ac tuning: CTAs (matrices) in flight, 0 = automatic')
    ap.add_argument('--ac-tpb', type=int, default=0,
                    help='ac tuning: threads per CTA (512 | 1024)')
    ap.add_argument('--tile-tpb', type=int, default=0,
                    help='mc tuning: threads per CTA of the tile kernel '
                         '(128 | 256, 0 = automatic)')
    ap.add_argument('--n', type=int, default=0,
                    help='problem size (0 = workload default; mc: instances '
                         'of the whole job)')
    ap.add_argument('--weak', action='store_true',
                    help='mc: --n instances per GPU instead of in total')
    ap.add_argument('--legs', type=int, default=-1,
                    help='mc on one GPU: also run short deveval and soliton '
                         'legs (the other two parts of BASELINE.json\'s '
                         'metric) and report them under config (-1 = when '
                         'the workload flag is not given)')
    args = ap.parse_args()
    if args.legs < 0:
        args.legs = 1 if args.workload is None else 0
    if args.workload is None:
        args.workload = 'mc'
    args.warmup = max(args.warmup, 3) if args.impl == 'native' \
        else args.warmup
    if args.impl == 'reference':
        line = run_reference(args)
        if line is not None:
            emit(line)
        return
    import __graft_entry__ as g
    if int(os.environ.get('LOCAL_RANK', '0')) == 0:
        g.build()
    D = Dist(args.gpus)
    D.barrier()
    line = dict(deveval=run_deveval, mc=run_mc, soliton=run_soliton, ac=run_ac,
                chain=run_chain)[args.workload](args, D)
    if args.workload == 'mc' and args.legs and D.world == 1:
        # the other two parts of BASELINE.json's metric, short runs
        import copy
        keep = ('metric', 'value', 'unit', 'ms_per_step', 'steps', 'warmup',
                'e2e', 'roofline', 'cpu_baseline', 'gpu_launches', 'clocks')
        for name, fn, n in (('deveval', run_deveval, 1 << 21),
                            ('soliton', run_soliton, 0)):
            a2 = copy.copy(args)
            a2.workload, a2.n = name, n
            a2.steps = max(3, min(args.steps, 5))
            sub = fn(a2, D)
            leg = {k: sub[k] for k in keep if k in sub}
            leg['workload'] = sub['config']['workload']
            if name == 'deveval':
                leg['gummel_poon'] = sub['config']['gummel_poon']
            line['config'][name] = leg
            line['gpu_launches'] += sub.get('gpu_launches', 0)
    if D.rank == 0:
        emit(line)
    D.close()


if __name__ == '__main__':
    main()
