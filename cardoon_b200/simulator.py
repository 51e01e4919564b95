"""
Library façade: parse netlists and run the analyses they request.

Mirror of the reference's cardoon/simulator.py (reset_all :43, parse_net :54,
run_analyses :66, new_elem :90, run_netlist :109).

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import time

from . import circuit as cir
from . import analyses
from . import devices
from .globalVars import glVar
from .paramset import ParamSet
from .netlistparser import parse_file, ParseError   # noqa: F401

version = '0.1'


def reset_all():
    """Forget all circuits/models/variables and restore default options."""
    glVar.reset()
    glVar.set_attributes()
    ParamSet.netVar.clear()
    cir.reset_allckt()


def parse_net(filename, ckt=None):
    if not ckt:
        ckt = cir.get_mainckt()
    return parse_file(filename, ckt)


def run_analyses(analysisQueue, ckt=None):
    if not ckt:
        ckt = cir.get_mainckt()
    if not analysisQueue:
        print('Nothing to do. Printing circuit:\n')
        print(ckt.netlist_string())
        print(ckt.globals_to_str())
        return
    for an in analysisQueue:
        try:
            start = time.perf_counter()
            an.run(ckt)
            elapsed = time.perf_counter() - start
            print('{0} analysis time: {1} s\n'.format(an.anType, elapsed))
        except analyses.AnalysisError as ae:
            print(ae)


def new_elem(elemType, name, **kwargs):
    """``new_elem('diode', 'd4', isat=2.1e-15, cj0=1e-12)``"""
    elem = devices.devClass[elemType](name)
    elem.set_params(**kwargs)
    return elem


def run_netlist(fileName, circuitName=None, reset=True):
    if reset:
        reset_all()
    ckt = cir.Circuit(circuitName) if circuitName else cir.get_mainckt()
    queue = parse_net(fileName, ckt)
    run_analyses(queue, ckt)
    return ckt
