#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics per launch as a
small CSV, plus the hottest source lines by sampled stalls.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_name \
      [mangled-kernel-substring [library.so]]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio']


def ncu(args):
    return subprocess.run(['ncu'] + args, stdout=subprocess.PIPE,
                          stderr=subprocess.DEVNULL, text=True).stdout


def main(rep, out, func=None, lib='cardoon_b200/libcardoon_b200.so'):
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw',
                                            '--csv']))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out + '_ncu_full.csv', 'w') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + ['launch%d' % i
                                         for i in range(len(data))])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in data])
    if func:
        hot_lines(rep, out, func, lib)


def line_table(lib, func):
    """offset -> 'file:line' of one kernel, from nvdisasm line info of the
    cubins embedded in the shared library."""
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)],
                   cwd=tmp, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    table = {}
    for cub in sorted(os.listdir(tmp)):
        txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)],
                             stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        inside = False
        cur = '?'
        for line in txt.splitlines():
            if line.startswith('//---') and '.text.' in line:
                inside = func in line
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = os.path.basename(m.group(1)) + ':' + m.group(2)
                continue
            m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', line)
            if m:
                table[int(m.group(1), 16)] = (cur, m.group(2).strip())
        if table:
            break
    return table


def hot_lines(rep, out, func, lib):
    table = line_table(lib, func)
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source',
                                            '--csv']))))
    # first kernel section only
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == 'Address':
            if hdr is not None:
                break
            hdr = r
        elif hdr is not None and r and r[0].startswith('0x'):
            data.append(r)
    if hdr is None or not data:
        print('no source page')
        return
    si = hdr.index('# Samples')
    stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_')
              and 'Not Issued' not in h]
    base = int(data[0][0], 16)
    per_line, per_stall, src_text = {}, {}, {}
    total = 0.
    for r in data:
        off = int(r[0], 16) - base
        n = float(r[si] or 0)
        total += n
        where = table.get(off, ('?', ''))[0]
        per_line[where] = per_line.get(where, 0.) + n
        for i in stalls:
            per_stall[hdr[i]] = per_stall.get(hdr[i], 0.) + float(r[i] or 0)
    srcs = {}
    with open(out + '_hot_lines.txt', 'w') as f:
        f.write('# {0}: warp-stall samples by source line (innermost inlined '
                'location), {1:.0f} samples\n'.format(func, total))
        st = sum(per_stall.values()) or 1.
        f.write('# stall mix: ' + ', '.join(
            '{0} {1:.1f}%'.format(k, 100. * v / st)
            for k, v in sorted(per_stall.items(), key=lambda x: -x[1])[:8])
            + '\n')
        for where, n in sorted(per_line.items(), key=lambda x: -x[1])[:45]:
            fn, _, ln = where.partition(':')
            text = ''
            for root in ('cardoon_b200/csrc', 'cardoon_b200/csrc/models'):
                pth = os.path.join(root, fn)
                if os.path.exists(pth) and ln.isdigit():
                    if pth not in srcs:
                        srcs[pth] = open(pth).read().splitlines()
                    if int(ln) <= len(srcs[pth]):
                        text = srcs[pth][int(ln) - 1].strip()[:100]
                    break
            f.write('{0:6.2f}%  {1:<24} {2}\n'.format(
                100. * n / (total or 1.), where, text))


if __name__ == '__main__':
    main(*sys.argv[1:])
