from ..paramset import ParamSet
class Transient(ParamSet):
    anType = 'tran'
    paramDict = dict(
        tstop=('Simulation stop time', 's', float, 1e-3),
        tstep=('Time step size', 's', float, 1e-5),
        im=('Integration method', '', str, 'trap'),
        verbose=('Show iterations for each point', '', bool, False),
        saveall=('Save all nodal voltages', '', bool, False),
        shell=('Drop to ipython shell after calculation', '', bool, False),
    )
    def __init__(self):
        ParamSet.__init__(self, self.paramDict)
aClass = Transient
