"""
Internal circuit representation: terminals, elements, sub-circuits.

Host-side mirror of the reference's cardoon/circuit.py (GraphNode :121,
Terminal :146, InternalTerminal :171, Element :205, Xsubckt :425,
Circuit :478, SubCircuit :971).  This layer is formulation independent and
not on the hot path; the nodal back-end only reads ``elem.connection`` and
the port lists the device plugins publish.

Differences that do not change results: dictionaries keep insertion order
(the reference relied on Python-2 hash order, so its node numbering was
arbitrary; compare by terminal *name*).

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import copy

from .paramset import ParamSet, Model
from .globalVars import glVar


class CircuitError(Exception):
    pass


class GraphNode(object):
    """Graph vertex with an ordered adjacency list."""

    def __init__(self, instanceName):
        self.instanceName = instanceName
        self.connection = []

    def __str__(self):
        return '"{0}"\nLinked nodes: {1}'.format(
            self.instanceName,
            ' '.join(n.instanceName for n in self.connection))


class Terminal(GraphNode):
    """External terminal (a netlist node)."""
    unit = 'V'

    def __str__(self):
        return 'Terminal: {1}\nUnit: {0}'.format(self.unit,
                                                 GraphNode.__str__(self))

    def get_label(self):
        return 'T: "{0}"'.format(self.instanceName)


class InternalTerminal(Terminal):
    """Terminal owned by one element (its only neighbour)."""

    def __init__(self, element, name):
        Terminal.__init__(self, name)
        self.connection.append(element)
        element.connection.append(self)

    def __str__(self):
        return 'Internal Terminal: "{1}:{2}", Unit: {0}'.format(
            self.unit, self.connection[0].instanceName, self.instanceName)

    def get_label(self):
        return 'IT: "{0}:{1}"'.format(self.connection[0].instanceName,
                                      self.instanceName)


class Element(GraphNode, ParamSet):
    """Base class of every device plugin; holds the default flags."""
    numTerms = 0
    tempItem = (('temp', ('Device temperature (None: use global temp.)',
                          'C', float, None)), )
    makeAutoThermal = False
    isNonlinear = False
    nDelays = 0
    isFreqDefined = False
    isDCSource = False
    isTDSource = False
    isFDSource = False
    linearVCCS = []
    linearVCQS = []
    noisePorts = ()
    localReference = 0

    def __init__(self, instanceName):
        GraphNode.__init__(self, self.devType + ':' + instanceName)
        ParamSet.__init__(self, self.paramDict)
        self.dotModel = False

    def __str__(self):
        desc = 'Element: ' + GraphNode.__str__(self)
        desc += '\nDevice type: ' + self.devType + '\n'
        if self.dotModel:
            desc += 'Model: {0}\n'.format(self.dotModel.name)
        return desc + 'Overridden parameters: ' \
            + ParamSet.netlist_string(self)

    def netlist_string(self):
        desc = self.instanceName + ' '
        for term in self.connection:
            if isinstance(term, InternalTerminal):
                break
            desc += term.instanceName + ' '
        if self.dotModel:
            desc += 'model = {0} '.format(self.dotModel.name)
        return desc + ParamSet.netlist_string(self)

    def print_vars(self):
        print('Parameter values:\n')
        print(ParamSet.__str__(self))

    def init(self):
        """Set attributes, check terminal count, process parameters."""
        self.set_attributes()
        self.check_terms()
        self.process_params()

    def is_set(self, paramName):
        if self.dotModel and self.dotModel.is_set(paramName):
            return True
        return ParamSet.is_set(self, paramName)

    def set_attributes(self):
        # ambient temperature unless overridden by model / element line
        self.temp = glVar.temp
        if self.dotModel:
            ParamSet.set_attributes(self, useDefaults=False)
            self.dotModel.set_missing_attributes(self)
        else:
            ParamSet.set_attributes(self, useDefaults=True)

    def check_terms(self):
        if self.numTerms:
            if len(self.connection) != self.numTerms:
                raise CircuitError('{0}: must have {1} terminals.'.format(
                    self.instanceName, self.numTerms))
        else:
            self.numTerms = len(self.connection)

    def disconnect(self, terminal):
        try:
            self.connection.remove(terminal)
            terminal.connection.remove(self)
        except ValueError:
            raise CircuitError('Nodes not linked')

    def add_internal_term(self, name, unit):
        term = InternalTerminal(self, name)
        term.unit = unit
        return len(self.connection) - 1

    def add_reference_term(self):
        assert not self.localReference
        term = InternalTerminal(self, 'lref')
        term.unit = '-'
        self.localReference = len(self.connection) - 1
        return self.localReference

    def get_internal_terms(self):
        intTerms = self.connection[self.numTerms:]
        if self.localReference:
            intTerms.pop(self.localReference - self.numTerms)
        return intTerms

    def clean_internal_terms(self):
        for term in list(self.connection[self.numTerms:]):
            self.disconnect(term)
        self.connection = self.connection[:self.numTerms]
        self.localReference = 0


class Xsubckt(GraphNode):
    """Sub-circuit instance (``x...`` line)."""

    def __init__(self, instanceName, cktName):
        assert instanceName[0] in 'xX'
        GraphNode.__init__(self, instanceName)
        self.cktName = cktName

    def netlist_string(self):
        return ' '.join([self.instanceName]
                        + [t.instanceName for t in self.connection]
                        + [self.cktName])

    def check_terms(self):
        try:
            cktDef = Circuit.cktDict[self.cktName]
        except KeyError:
            raise CircuitError('{0}: subcircuit definition "{1}" not found'
                               .format(self.instanceName, self.cktName))
        if len(self.connection) != len(cktDef.extConnectionList):
            raise CircuitError(
                '{0}: xsubckt connections do not match definition in "{1}"'
                .format(self.instanceName, self.cktName))


class Circuit(object):
    """A circuit: dictionaries of terminals, elements, subckt instances.

    ``cktDict`` and ``modelDict`` are global (class level), as in the
    reference: ``.model`` and ``.subckt`` have global scope.
    """
    cktDict = dict()
    modelDict = dict()

    def __init__(self, name):
        self.name = name
        self._initialized = False
        self._flattened = False
        if name in self.cktDict:
            raise CircuitError('Circuit "' + name + '" already exists')
        self.cktDict[name] = self
        self.termDict = dict()
        self.elemDict = dict()
        self.subcktDict = dict()
        self.plotReqList = list()
        self.saveReqList = list()

    def __str__(self):
        if hasattr(self, 'title'):
            return self.title + '\n'
        return 'Circuit: {0}\n'.format(self.name)

    def netlist_string(self):
        desc = self.title + '\n' if hasattr(self, 'title') else ''
        desc += '# Circuit: ' + self.name + '\n'
        desc += '#                         *** Elements ***\n'
        for elem in self.elemDict.values():
            desc += elem.netlist_string() + '\n'
        if not self._flattened and self.subcktDict:
            used = []
            desc += '#                     *** Subcircuit instances ***\n'
            for xs in self.subcktDict.values():
                desc += xs.netlist_string() + '\n'
                if xs.cktName not in used:
                    used.append(xs.cktName)
            desc += '#                     *** Subcircuit Definitions ***\n'
            for cktName in used:
                desc += Circuit.cktDict[cktName].netlist_string()
        for tag, reqs in (('.plot ', self.plotReqList),
                          ('.save ', self.saveReqList)):
            for outreq in reqs:
                desc += '\n' + tag + outreq.type
                for name in outreq.varlist:
                    desc += ' ' + str(name)
        return desc + '\n'

    def globals_to_str(self):
        desc = '.vars ' + ''.join('{0} = {1} '.format(k, v)
                                  for k, v in ParamSet.netVar.items())
        desc += '\n\n.options ' + glVar.netlist_string() + '\n\n'
        desc += '#                  *** Models *** \n'
        for model in Circuit.modelDict.values():
            desc += model.netlist_string() + '\n\n'
        return desc

    # ---- whole-circuit actions ----------------------------------------
    def _copy_into(self, cktCopy):
        for elem in self.elemDict.values():
            elemCopy = copy.copy(elem)
            elemCopy.connection = []
            cktCopy.add_elem(elemCopy)
            cktCopy.connect(elemCopy,
                            [t.instanceName for t in elem.connection])
        if not self._flattened:
            for xs in self.subcktDict.values():
                xsCopy = copy.copy(xs)
                xsCopy.connection = []
                cktCopy.add_subckt(xsCopy)
                cktCopy.connect(xsCopy,
                                [t.instanceName for t in xs.connection])
        return cktCopy

    def copy(self, newName):
        """Copy of an uninitialised circuit (elements shallow-copied)."""
        assert not self._initialized
        return self._copy_into(Circuit(newName))

    def init(self):
        """Initialise all elements and resolve output requests (once)."""
        if self._initialized:
            return
        for elem in list(self.elemDict.values()):
            elem.init()
        if not self._flattened:
            for xs in self.subcktDict.values():
                xs.check_terms()
                Circuit.cktDict[xs.cktName].init()
        for outreq in self.plotReqList + self.saveReqList:
            outreq.termlist = []
            for termname in outreq.varlist:
                if isinstance(termname, str):
                    if termname == '0':
                        termname = 'gnd'
                    try:
                        outreq.termlist.append(self.termDict[termname])
                    except KeyError:
                        raise CircuitError(
                            'Output request for nonexistent terminal: '
                            + termname)
                else:
                    try:
                        elem = self.elemDict[termname[0]]
                    except KeyError:
                        raise CircuitError(
                            'Output request for nonexistent element: '
                            + termname[0] + ':' + termname[1])
                    for term in elem.get_internal_terms():
                        if term.instanceName == termname[1]:
                            outreq.termlist.append(term)
                            break
                    else:
                        raise CircuitError(
                            'Output request for nonexistent terminal: '
                            + termname[0] + ':' + termname[1])
        self._initialized = True

    def get_internal_terms(self):
        assert self._initialized
        out = []
        for elem in self.elemDict.values():
            out += elem.get_internal_terms()
        return out

    def flatten(self):
        """Expand sub-circuit instances in place (before init())."""
        assert not self._initialized
        if self._flattened:
            return
        for xs in self.subcktDict.values():
            cktDef = self.cktDict[xs.cktName]
            if not cktDef._flattened:
                cktDef.flatten()
            cktDef.get_circuit(xs, self)
        self._flattened = True

    def check_sanity(self):
        for elem in self.elemDict.values():
            for term in elem.connection:
                assert self.termDict[term.instanceName] == term
        for term in self.termDict.values():
            assert term.connection

    # ---- individual elements / terminals ------------------------------
    def connect(self, element, termList):
        assert not element.connection
        for termName in termList:
            terminal = self.get_term(termName)
            terminal.connection.append(element)
            element.connection.append(terminal)

    def add_elem(self, elem, modelName=None):
        if elem.instanceName in self.elemDict:
            raise CircuitError(elem.instanceName + ': Element already exists')
        self.elemDict[elem.instanceName] = elem
        if modelName:
            if elem.dotModel:
                raise CircuitError('{0} already has an assigned model'
                                   .format(elem.instanceName))
            if modelName in Circuit.modelDict:
                model = Circuit.modelDict[modelName]
                if model.modelType != elem.devType:
                    raise CircuitError(
                        'Incorrect model type "{0}" for element "{1}"'
                        .format(model.modelType, elem.instanceName))
                elem.dotModel = model
            else:
                elem.dotModel = Model(modelName, elem.devType,
                                      elem.paramDict)
                Circuit.modelDict[modelName] = elem.dotModel

    def remove_elem(self, elemName):
        try:
            elem = self.elemDict.pop(elemName)
        except KeyError:
            raise CircuitError(elemName + ': Element not found')
        elem.clean_internal_terms()
        for term in elem.connection:
            term.connection.remove(elem)

    def add_subckt(self, xsubckt):
        if xsubckt.instanceName in self.subcktDict:
            raise CircuitError('add_subckt: Subcircuit instance name "'
                               + xsubckt.instanceName + '" already in use')
        self.subcktDict[xsubckt.instanceName] = xsubckt

    def get_term(self, termName):
        if termName == '0':
            termName = 'gnd'
        if termName not in self.termDict:
            self.termDict[termName] = Terminal(termName)
        return self.termDict[termName]

    def has_term(self, termName):
        return ('gnd' if termName == '0' else termName) in self.termDict

    def find_term(self, termName):
        """Find an external ('name') or internal ('dev:inst:name') terminal."""
        head, _, tail = termName.rpartition(':')
        if head:
            if not head.rpartition(':')[0]:
                raise CircuitError('Terminal name syntax is wrong: "'
                                   + termName + '"')
            try:
                terms = self.elemDict[head].get_internal_terms()
            except KeyError:
                raise CircuitError('Element "' + head + '" not found')
            for t in terms:
                if t.instanceName == tail:
                    return t
            raise CircuitError('Terminal "{0}" not found in element "{1}"'
                               .format(tail, head))
        try:
            return self.termDict['gnd' if termName == '0' else termName]
        except KeyError:
            raise CircuitError('Terminal "' + termName + '" not found')

    def remove_term(self, termName):
        try:
            self.termDict.pop(termName)
        except KeyError:
            raise CircuitError(termName + ': Terminal not found')

    def add_plot_request(self, outreq):
        self.plotReqList.append(outreq)

    def add_save_request(self, outreq):
        self.saveReqList.append(outreq)

    def get_requested_terms(self, reqtype):
        """Terminals named by .plot/.save requests of one type (no repeats,
        in request order)."""
        out = []
        for outreq in self.plotReqList + self.saveReqList:
            if outreq.type == reqtype:
                for term in outreq.termlist:
                    if term not in out:
                        out.append(term)
        return out


class SubCircuit(Circuit):
    """Circuit with an external connection list (``.subckt`` definition)."""

    def __init__(self, name, termList):
        Circuit.__init__(self, name)
        self.extConnectionList = list(termList)
        for i, termName in enumerate(termList):
            terminal = self.get_term(termName)
            if terminal.instanceName == 'gnd':
                raise CircuitError(
                    name + ': gnd is the global reference. It is implicitly '
                    'connected and can not be included in the external '
                    'terminal list in subckt')
            terminal.subCKTconnection = i

    def copy(self, newName):
        assert not self._initialized
        return self._copy_into(SubCircuit(newName, self.extConnectionList))

    def get_connections(self):
        return [self.get_term(name) for name in self.extConnectionList]

    def get_circuit(self, xsubckt, target):
        """Dump a renamed copy of the elements into ``target`` (flatten)."""
        assert self._flattened
        prefix = xsubckt.instanceName + ':'
        for elem in self.elemDict.values():
            elemCopy = copy.copy(elem)
            elemCopy.connection = []
            elemCopy.instanceName = prefix + elem.instanceName
            target.add_elem(elemCopy)
            termList = []
            for term in elem.connection:
                if hasattr(term, 'subCKTconnection'):
                    name = xsubckt.connection[
                        term.subCKTconnection].instanceName
                elif term.instanceName == 'gnd':
                    name = 'gnd'
                else:
                    name = prefix + term.instanceName
                termList.append(name)
            target.connect(elemCopy, termList)

    def netlist_string(self):
        desc = '#-----------------------------------------------------\n'
        desc += '.subckt ' + self.name + ''.join(
            ' ' + n for n in self.extConnectionList) + '\n'
        return desc + Circuit.netlist_string(self) + '.ends\n\n'


def get_mainckt():
    if 'main' not in Circuit.cktDict:
        return Circuit('main')
    return Circuit.cktDict['main']


def reset_allckt():
    Circuit.cktDict = dict()
    Circuit.modelDict = dict()


def add_model(model):
    if model.name in Circuit.modelDict:
        raise CircuitError(model.name + ': Model already exists')
    Circuit.modelDict[model.name] = model


def get_model(modelName):
    return Circuit.modelDict.get(modelName)
