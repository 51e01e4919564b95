import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import regress as G
from cardoon_b200.globalVars import glVar
path = os.path.join(G.REF_DIR, 'sum741_profile.net')
ref = G.oracle_results(path)
for variant in ('asis', 'dense', 'tranonly'):
    deck = open(path).read()
    if variant == 'dense':
        deck = deck.replace('.options', '.options sparse=0') if '.options' in deck else '.options sparse=0\n' + deck
    if variant == 'tranonly':
        deck = '\n'.join(l for l in deck.split('\n') if not l.startswith('.analysis dc') and not l.startswith('.analysis ac'))
    with tempfile.TemporaryDirectory() as td:
        p2 = os.path.join(td, 'sum741_profile.net')
        open(p2, 'w').write(deck)
        with tempfile.TemporaryDirectory() as td2:
            got = G.product_results(p2, td2)
    for rt, rd in ref.items():
        if rt not in got:
            print(variant, rt, 'missing'); continue
        for var, rv in rd.items():
            d = np.abs(rv - got[rt][var])
            tol = np.abs(glVar.reltol * rv) + glVar.abstol
            print(variant, rt, var, 'max ratio %.3g' % np.max(d / tol), 'first bad', int(np.argmax(d > tol)) if np.any(d > tol) else -1,
                  'ref', np.round(np.real(rv[:3]), 6), 'got', np.round(np.real(got[rt][var][:3]), 6))
