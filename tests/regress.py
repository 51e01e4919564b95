"""
Regression harness over the reference's own test decks (SURVEY 8f rank 4;
reference test/run_tests.py, test/README).

The reference compares the ``<deck>_<type>.npz`` files its ``.save`` requests
write (analysis.py:59-116) against ``*_ref.npz`` files that are not in its
tree.  Here the reference results are regenerated from the oracle:
``oracle_results`` runs the deck's analyses on the CPU restatement and
returns what the ``.save`` requests would store; ``product_results`` runs the
deck through the CUDA back-end exactly like ``run_tests.py`` does
(``parse_net`` + ``run_analyses``) and loads the files it wrote;
``compare`` applies run_tests.py's acceptance test
``|delta| < reltol |ref| + abstol``.

Test infrastructure: lives under tests/ because it imports the oracle.
"""
import os
import shutil

import numpy as np

from tests import support as S
from cardoon_b200.globalVars import glVar
from oracle import nodal_ref as R

REF_DIR = os.path.join(S.HERE, 'netlists', 'reference_tests')


def _saved(ckt, reqtype, idx):
    rows = idx.name_to_row()
    names = []
    for req in ckt.saveReqList:
        if req.type == reqtype:
            names += [t.instanceName for t in req.termlist]
    return [(n, rows[n]) for n in names if n in rows]


def oracle_results(path):
    """{request type: {'xaxis': ..., terminal: values}} from the oracle."""
    out = {}
    ckt, queue = S.load_circuit(path)
    # one circuit for the whole queue, like run_analyses: state left behind by
    # an analysis is seen by the next ones
    for an in queue:
        if an.anType == 'tran':
            r = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im)
            d = {'xaxis': r['t']}
            for name, row in _saved(ckt, 'tran', r['idx']):
                d[name] = r['X'][:, row]
            out['tran'] = d
        elif an.anType == 'dc':
            r = R.run_dc(ckt, glVar, an.device, an.param, an.start, an.stop,
                         an.num)
            d = {'xaxis': r['sweep']}
            for name, row in _saved(ckt, 'dc', r['idx']):
                d[name] = r['X'][:, row]
            out['dc'] = d
            # dc.py:219-236 "restore original attribute values":
            # set_attributes() only touches parameters that are set in the
            # netlist or not yet attributes (paramset.py:258-280), so a
            # swept parameter the deck leaves at its default KEEPS the last
            # sweep value for the analyses that follow (sum741_profile.net:
            # vsin:v1 vdc stays at 0.5 V for the ac and tran runs)
            dev = ckt.elemDict[an.device]
            dev.set_attributes()
            dev.process_params()
        elif an.anType == 'ac':
            ckt2 = ckt
            if an.log:
                f = np.logspace(np.log10(an.start), np.log10(an.stop), an.num)
            else:
                f = np.linspace(an.start, an.stop, an.num)
            r = R.run_ac(ckt2, glVar, f)
            for rt, fn in (('ac', lambda x: x), ('ac_mag', np.abs),
                           ('ac_phase',
                            lambda v: 180. / np.pi * np.angle(v)),
                           ('ac_dB', lambda v: 20. * np.log10(v))):
                sv = _saved(ckt2, rt, r['idx'])
                if sv:
                    d = {'xaxis': f}
                    for name, row in sv:
                        d[name] = fn(r['X'][row])
                    out[rt] = d
    return out


def product_results(path, workdir):
    """Run the deck like test/run_tests.py and load the .npz files."""
    from cardoon_b200 import simulator as sim
    from cardoon_b200 import circuit as cir
    deck = os.path.join(workdir, os.path.basename(path))
    shutil.copy(path, deck)
    sim.reset_all()
    ckt = cir.get_mainckt()
    queue = sim.parse_net(deck, ckt)
    ckt.plotReqList = []
    sim.run_analyses(queue, ckt)
    base = deck.split('.net')[0]
    out = {}
    for req in ckt.saveReqList:
        fn = base + '_' + req.type + '.npz'
        if os.path.exists(fn):
            with np.load(fn) as z:
                out[req.type] = {k: z[k] for k in z.files}
    sim.reset_all()
    return out


def compare(ref, got):
    """run_tests.py:78-108: every variable of every reference file within
    reltol |ref| + abstol.  Returns (ok, sum of average residuals, worst)."""
    residual = 0.
    worst = ('', 0.)
    ok = True
    for rt, rd in ref.items():
        if rt not in got:
            return False, np.inf, (rt, np.inf)
        for var, rv in rd.items():
            if var not in got[rt]:
                return False, np.inf, (rt + ':' + var, np.inf)
            delta = np.abs(rv - got[rt][var])
            residual += float(np.average(delta))
            tol = np.abs(glVar.reltol * rv) + glVar.abstol
            ratio = float(np.max(delta / tol)) if delta.size else 0.
            if ratio > worst[1]:
                worst = (rt + ':' + var, ratio)
            if not np.all(delta < tol):
                ok = False
    return ok, residual, worst
