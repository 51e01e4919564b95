#!/usr/bin/env python
"""Per-region breakdown of one kernel from an .ncu-rep (read here, no GPU):
executed warp instructions and stall samples grouped by the source function
that encloses each SASS instruction's innermost line (nvdisasm line table of
the shipped cubin + a scan of `engine.cuh` for function starts).

  python tools/ncu_phases.py gpurun_out/x.ncu-rep engine_tile_kernel [out.txt]
"""
import csv
import io
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_summary as NS                                       # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def function_ranges(path):
    """[(first line, name)] of every function definition in a source file."""
    out = []
    pat = re.compile(r'^\s*(?:template.*>)?\s*(?:static\s+)?(?:__device__|__global__|CDN_HD|'
                     r'__host__)[^;{(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(')
    with open(path) as f:
        for n, line in enumerate(f, 1):
            m = pat.match(line)
            if m and not line.strip().endswith(';'):
                out.append((n, m.group(1)))
    return out


def main(rep, func, out=None, lib='cardoon_b200/libcardoon_b200.so'):
    table = NS.line_table(lib, func)
    rows = list(csv.reader(io.StringIO(NS.ncu(['-i', rep, '--page', 'source',
                                               '--csv']))))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == 'Address':
            if hdr is not None:
                break
            hdr = r
        elif hdr is not None and r and r[0].startswith('0x'):
            data.append(r)
    si = hdr.index('# Samples')
    ei = hdr.index('Warp Instructions Executed') \
        if 'Warp Instructions Executed' in hdr else hdr.index('Instructions Executed')
    ti = hdr.index('Thread Instructions Executed') \
        if 'Thread Instructions Executed' in hdr else None
    base = int(data[0][0], 16)
    franges = {}
    agg = {}
    tot = [0., 0., 0.]
    for r in data:
        off = int(r[0], 16) - base
        where, sass = table.get(off, ('?:0', ''))
        fn, _, ln = where.partition(':')
        region = fn
        if fn.endswith(('engine.cuh',)):
            pth = os.path.join(ROOT, 'cardoon_b200/csrc', fn)
            if pth not in franges:
                franges[pth] = function_ranges(pth)
            name = '?'
            for first, nm in franges[pth]:
                if first <= int(ln):
                    name = nm
                else:
                    break
            region = 'engine:' + name
        elif fn.endswith('.cuh') or fn.endswith('.hpp') or fn.endswith('.h'):
            region = fn
        a = agg.setdefault(region, [0., 0., 0., 0])
        a[0] += float(r[si] or 0)
        a[1] += float(r[ei] or 0)
        if ti is not None:
            a[2] += float(r[ti] or 0)
        a[3] += 1
        tot[0] += float(r[si] or 0)
        tot[1] += float(r[ei] or 0)
        tot[2] += float(r[ti] or 0) if ti is not None else 0.
    lines = ['# {0}: {1} SASS instructions, {2:.3g} warp instructions '
             'executed, {3:.0f} stall samples'.format(func, len(data), tot[1],
                                                     tot[0]),
             '# region                         samples%  warp-inst%  '
             'lanes/inst  static SASS']
    for region, a in sorted(agg.items(), key=lambda x: -x[1][0]):
        lines.append('{0:<34}{1:7.2f}  {2:9.2f}  {3:9.1f}  {4:8d}'.format(
            region, 100. * a[0] / tot[0], 100. * a[1] / tot[1],
            a[2] / a[1] if a[1] else 0., a[3]))
    txt = '\n'.join(lines) + '\n'
    if out:
        open(out, 'w').write(txt)
    print(txt)


if __name__ == '__main__':
    main(*sys.argv[1:])
