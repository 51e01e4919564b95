"""
Parameter sets for elements, models, analyses and global options.

Host-side mirror of the reference's cardoon/paramset.py (ParamSet :27,
Model :285).  A parameter table is ``name -> (doc, unit, type, default)``;
values given in the netlist live in ``valueDict``, references to ``.vars``
variables in ``varDict``; ``set_attributes`` turns them into instance
attributes with the precedence element > model > default
(reference circuit.py:318-336).

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""


class ParamError(Exception):
    pass


class ParamSet:
    # netlist variables (.vars), shared by every parameter set
    netVar = dict()

    def __init__(self, paramDict):
        self.paramDict = paramDict
        self.valueDict = dict()
        self.varDict = dict()

    # ---- description helpers ------------------------------------------
    def __str__(self):
        rows = [' Name    |   Value    | Unit  ',
                '------------------------------']
        for key in sorted(self.paramDict):
            info = self.paramDict[key]
            value = getattr(self, key,
                            self.valueDict.get(key, info[3]))
            rows.append('{0:^8} | {1!s:^10} | {2:^5}'.format(key, value,
                                                            info[1]))
        return '\n'.join(rows) + '\n'

    def netlist_string(self):
        items = list(self.varDict.items()) + list(self.valueDict.items())
        return ''.join('{0} = {1} '.format(k, v) for k, v in items)

    def format(self, paramName):
        info = self.paramDict[paramName]
        return ' {0:<10}   {1!s:<18}   {2:<10}   {3:<52} \n'.format(
            paramName, info[3], info[1], info[0])

    def describe_parameters(self, paramName=None):
        bar = (' =========== ==================== ============ '
               '===================================================== \n')
        if paramName:
            return self.format(paramName) + bar
        out = bar + (' Name         Default              Unit         '
                     'Description                                          \n')
        out += bar
        for key in sorted(self.paramDict):
            out += self.format(key)
        return out + bar

    def list_parameters(self):
        return ''.join(' ' + key for key in self.paramDict)

    # ---- setting values ------------------------------------------------
    def set_param(self, paramName, value):
        """Store a value (or the name of a netlist variable) for a parameter.

        Attributes are only updated by set_attributes().
        """
        if paramName not in self.paramDict:
            raise ParamError(
                '{0}: not a valid parameter name'.format(paramName))
        ptype = self.paramDict[paramName][2]
        if type(value) == ptype:
            self.valueDict[paramName] = value
        elif ptype == float and type(value) == int:
            self.valueDict[paramName] = float(value)
        elif type(value) == str:
            self.varDict[paramName] = value
        else:
            raise ParamError(
                '{0}: not a valid value or variable name'.format(paramName))

    def set_params(self, **kwargs):
        for name, value in kwargs.items():
            self.set_param(name, value)

    def get_type(self, paramName):
        try:
            return self.paramDict[paramName][2]
        except KeyError:
            raise ParamError(
                '{0}: not a valid parameter name'.format(paramName))

    def is_set(self, paramName):
        return paramName in self.valueDict or paramName in self.varDict

    def get_float_attributes(self):
        return [(key, getattr(self, key))
                for key, info in self.paramDict.items() if info[2] == float]

    # ---- attributes ----------------------------------------------------
    def reset(self):
        self.valueDict = dict()
        self.varDict = dict()
        self.clean_attributes()

    def clean_attributes(self):
        for key in self.paramDict:
            if key in self.__dict__:
                delattr(self, key)

    def _get_variable_value(self, paramName, variable):
        try:
            var = self.netVar[variable]
        except KeyError:
            raise ParamError(
                'Undefined netlist variable: {0} (parameter: {1})'.format(
                    variable, paramName))
        ptype = self.paramDict[paramName][2]
        try:
            if ptype == float:
                return float(var) if type(var) == int else var
            if ptype == int:
                return int(var)
            if ptype == bool:
                return bool(var)
            if type(var) == ptype:
                return var
            raise TypeError()
        except (TypeError, ValueError):
            raise ParamError(
                '{0}: netlist variable type mismatch: {1} = {2}'.format(
                    variable, paramName, var))

    def set_attributes(self, useDefaults=True):
        for key, varname in self.varDict.items():
            setattr(self, key, self._get_variable_value(key, varname))
        for key, value in self.valueDict.items():
            setattr(self, key, value)
        if useDefaults:
            for key, info in self.paramDict.items():
                if not hasattr(self, key):
                    setattr(self, key, info[3])


class Model(ParamSet):
    """``.model`` statement: a named parameter set shared by elements."""

    def __init__(self, name, modelType, paramDict):
        self.name = name
        self.modelType = modelType
        ParamSet.__init__(self, paramDict)

    def __str__(self):
        return 'Model: {0}, Type: {1}\n\n'.format(
            self.name, self.modelType) + ParamSet.__str__(self)

    def netlist_string(self):
        return '.model {0} {1} ({2})'.format(
            self.name, self.modelType, ParamSet.netlist_string(self))

    def set_missing_attributes(self, target):
        """Give ``target`` every parameter it does not yet own."""
        for key, info in self.paramDict.items():
            if hasattr(target, key):
                continue
            if key in self.varDict:
                value = self._get_variable_value(key, self.varDict[key])
            elif key in self.valueDict:
                value = self.valueDict[key]
            else:
                value = info[3]
            setattr(target, key, value)
        if not self.valueDict:
            print('\nWarning: empty model, name: {0}, type: {1}\n'.format(
                self.name, self.modelType))
