"""
Analysis registry (``anClass``) and output requests.

Mirror of the reference's cardoon/analyses/__init__.py (:21-70).  This
back-end provides the two analyses on the Newton hot path, ``op`` and
``tran``, the ``dc`` sweep and the ``ac`` sweep; ``testdev`` is registered as
a placeholder that reports it is not available, so netlists that request it
still load and run the rest.
"""
from .analysis import AnalysisError
from ..paramset import ParamSet


class OutRequest:
    """``.plot``/``.save`` request: type + list of variables."""

    def __init__(self, reqtype, varlist):
        if reqtype not in validReqTypes:
            raise AnalysisError(
                'Not a valid output request type: {0}\nValid types: {1}'
                .format(reqtype, validReqTypes))
        self.type = reqtype
        self.varlist = varlist


def _placeholder(name, params):
    class NotProvided(ParamSet):
        anType = name
        paramDict = params

        def __init__(self):
            ParamSet.__init__(self, self.paramDict)

        def set_param(self, paramName, value):
            # accept and ignore any parameter
            self.valueDict[paramName] = value

        def set_attributes(self, useDefaults=True):
            pass

        def run(self, circuit):
            print('\n"{0}" analysis is not provided by the B200 back-end '
                  '(op, dc, ac and tran are); skipped.\n'.format(name))

    NotProvided.__name__ = 'NotProvided_' + name
    return NotProvided


class _AnyParams(dict):
    """paramDict that accepts every name (all typed as strings)."""

    def __contains__(self, key):
        return True

    def __getitem__(self, key):
        return ('', '', str, '')


anClass = {}
validReqTypes = ['tran', 'dc', 'ac', 'ac_mag', 'ac_phase', 'ac_dB']

from . import op, tran, dc, ac      # noqa: E402

for _module in (op, tran, dc, ac):
    anClass[_module.aClass.anType] = _module.aClass
for _name in ('testdev',):
    anClass[_name] = _placeholder(_name, _AnyParams())
