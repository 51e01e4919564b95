"""
Shared analysis helpers (reference cardoon/analyses/analysis.py:20-116).

Plots need matplotlib, which is optional here: ``.plot`` requests are
skipped with a note when it is missing; ``.save`` requests write the same
``<netlist>_<reqtype>.npz`` files (keys ``xaxis`` + terminal names).

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import numpy as np


class AnalysisError(Exception):
    pass


def ipython_drop(msg, glo, loc):
    try:
        from IPython import embed
    except ImportError:
        print('(IPython not available: interactive shell skipped)')
        return
    print(msg)
    embed(user_ns=dict(glo, **loc))


def process_requests(circuit, reqtype, xaxis, xlabel, attribute,
                     f1=lambda x: x, unitOverride='', log=False):
    plots = [r for r in circuit.plotReqList if r.type == reqtype]
    if plots:
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            plt = None
            print('(matplotlib not available: .plot requests skipped)')
        if plt is not None:
            pltfunc = plt.semilogx if log else plt.plot
            for outreq in plots:
                plt.figure()
                plt.grid(True)
                plt.title(reqtype)
                plt.xlabel(xlabel)
                for term in outreq.termlist:
                    unit = unitOverride if unitOverride else term.unit
                    pltfunc(xaxis, f1(getattr(term, attribute)),
                            label='{0} [{1}]'.format(term.get_label(), unit))
                plt.legend()
            plt.draw()
    saveDict = {'xaxis': xaxis}
    flag = False
    for outreq in circuit.saveReqList:
        if outreq.type == reqtype:
            flag = True
            for term in outreq.termlist:
                saveDict[term.instanceName] = f1(getattr(term, attribute))
    if flag:
        outfile = circuit.filename.split('.net')[0] + '_' + reqtype
        np.savez(outfile, **saveDict)
