#!/usr/bin/env python
"""
Build a differently tuned copy of the library for an A/B measurement:

  python tools/build_variant.py TAG "-DCDN_TILE_MAXT=384" engine_k_bsim3_tile8.cu [more.cu]

recompiles the listed translation units with the extra flags, links them with
the regular objects of build/ into cardoon_b200/variants/lib_TAG.so (git-ignored,
travels with gpurun) -- run with CARDOON_B200_LIB=cardoon_b200/variants/lib_TAG.so.
"""
import os
import shlex
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g   # noqa: E402


def main():
    tag, flags, units = sys.argv[1], shlex.split(sys.argv[2]), sys.argv[3:]
    g.build()
    vdir = os.path.join(g.BUILD, 'variant_' + tag)
    os.makedirs(vdir, exist_ok=True)
    objs = {f + '.o': os.path.join(g.BUILD, f + '.o')
            for f in os.listdir(g.CSRC) if f.endswith(('.cu', '.cpp'))}
    procs = []
    for u in units:
        obj = os.path.join(vdir, u + '.o')
        cmd = ['nvcc'] + g.NVCC_FLAGS + flags + ['-c', os.path.join(g.CSRC, u), '-o', obj]
        if u.endswith('.cpp'):
            cmd[1:1] = []
        procs.append((u, obj, subprocess.Popen(cmd)))
    for u, obj, p in procs:
        if p.wait() != 0:
            raise SystemExit('nvcc failed for ' + u)
        objs[u + '.o'] = obj
    out_dir = os.path.join(g.PKG, 'variants')
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, 'lib_%s.so' % tag)
    subprocess.check_call(['nvcc', '-shared', '-o', out] + sorted(objs.values())
                          + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    print(out)


if __name__ == '__main__':
    main()
