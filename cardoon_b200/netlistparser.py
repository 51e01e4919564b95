"""
Netlist reader.

Same line-oriented, SPICE-like language as the reference's
cardoon/netlistparser.py (grammar :385-474, number syntax :72-116, line
handlers :211-371), re-expressed as a small hand-written tokenizer (the
reference used pyparsing, which is not a dependency here).  Not on the hot
path.

Supported lines: comments (``*``, ``#``, ``//``), ``\\`` continuation,
``<dev>:<name> nodes.. [par=val ..]``, ``x<name> nodes.. <subckt>``,
``.model``, ``.options``, ``.vars``, ``.subckt``/``.ends``, ``.include``,
``.analysis``, ``.plot``, ``.save``, ``.end``.

Provenance and licence: this file follows the structure, messages and
behaviour of the corresponding module of cardoon (cechrist/cardoon,
Copyright Carlos Christoffersen <c.christoffersen@ieee.org>), which is free
software under the GNU General Public License version 3 or later
(http://www.gnu.org/licenses/gpl.html); this Python-3 restatement is
distributed under the same terms.
"""
import os
import re

from . import circuit as cir
from .paramset import ParamSet, ParamError
from .globalVars import glVar


class ParseError(Exception):
    pass


class NetVarException(Exception):
    pass


# circuit being filled (main circuit at the bottom, .subckt bodies on top)
cktStack = []


class Queue:
    def __init__(self):
        self.q = []

    def reset(self):
        self.q = []


analysisQueue = Queue()

allowedChars = '._+-*:'

_NUM = re.compile(r'([+-]?\d*(?:\.\d*)?(?:[Ee][+-]?\d+)?)'
                  r'(meg|mil|[tgkmunpf])?', re.IGNORECASE)
_MULT = dict(t=1e12, g=1e9, meg=1e6, k=1e3, mil=25.4e-6, m=1e-3, u=1e-6,
             n=1e-9, p=1e-12, f=1e-15)


def string_to_number(valueString):
    """SPICE number: float, optional scale factor, trailing letters ignored.

    ``10``, ``10V``, ``1e-14``, ``1KHz``, ``100MEGHz``, ``20fF``, ``2mil``.
    Raises ValueError when the string does not start with a number.
    """
    m = _NUM.match(valueString)
    value = float(m.group(1))           # ValueError if empty / just '.'
    if m.group(2):
        value *= _MULT[m.group(2).lower()]
    return value


def convert_type(name, value, ptype, isVector):
    """Convert the text of a parameter value to ``ptype``.

    Raises NetVarException when the text is not a literal of that type (it
    is then assumed to name a ``.vars`` variable).
    """
    if ptype == str:
        if isVector:
            raise ParseError('"{0}" must be a string: only the following '
                             'special characters allowed: {1}'.format(
                                 name, allowedChars))
        return value
    if isVector:
        if ptype != list:
            raise ParseError('Can not convert "{0}" value to {1}'.format(
                name, ptype))
        out = []
        for item in value:
            try:
                out.append(string_to_number(item))
            except ValueError:
                out.append(item)
        return out
    if ptype == float:
        try:
            return string_to_number(value)
        except ValueError:
            raise NetVarException('"{0}" must be numeric: {1}'.format(
                name, value))
    if ptype == int:
        try:
            return int(value)
        except ValueError:
            raise NetVarException('"{0}" must be an int: {1}'.format(
                name, value))
    if ptype == bool:
        if value.lower() == 'false' or value == '0':
            return False
        if value.lower() == 'true' or value == '1':
            return True
        raise NetVarException('Can not convert "{0}" value: {1} to bool'
                              .format(name, value))
    if ptype == list:
        raise NetVarException('Can not convert "{0}" value to list (vector)'
                              .format(name))
    raise AssertionError(name + ': Parameter has unknown type')


def match_parameter(paramset, name, value, isVector):
    """Type-check one ``name=value`` against a parameter set."""
    try:
        ptype = paramset.get_type(name)
    except ParamError:
        raise ParseError('Unrecognized parameter: {0}\n'.format(name)
                         + 'Valid parameters: ' + paramset.list_parameters())
    try:
        return convert_type(name, value, ptype, isVector)
    except NetVarException:
        # not a literal: leave the text, ParamSet treats it as a variable
        return value


# ---- tokenizer -----------------------------------------------------------
_PAR = re.compile(
    r"([A-Za-z][A-Za-z0-9_]*)\s*=\s*"
    r"(?:'([A-Za-z0-9._+\-*: ()]*)'|\[([^\]]*)\]|([A-Za-z0-9._+\-*:]+))")
_FIRSTPAR = re.compile(r"[A-Za-z][A-Za-z0-9_]*\s*=")
_IDENT = re.compile(r'[A-Za-z0-9_\-]+$')


def _split_params(text):
    """'a b c p1=v1 p2 = v2' -> (['a','b','c'], [(name, value, isVector)])"""
    m = _FIRSTPAR.search(text)
    if not m:
        return text.split(), []
    head, tail = text[:m.start()], text[m.start():]
    # model lines wrap the parameters in parentheses
    head = head.replace('(', ' ')
    tail = tail.strip()
    pars = []
    pos = 0
    while pos < len(tail):
        while pos < len(tail) and tail[pos] in ' \t)(':
            pos += 1
        if pos >= len(tail):
            break
        pm = _PAR.match(tail, pos)
        if not pm:
            raise ParseError('Syntax error near: "{0}"'.format(tail[pos:]))
        if pm.group(3) is not None:
            vec = [s.strip().strip("'") for s in pm.group(3).split(',')]
            pars.append((pm.group(1), vec, True))
        elif pm.group(2) is not None:
            pars.append((pm.group(1), pm.group(2), False))
        else:
            pars.append((pm.group(1), pm.group(4), False))
        pos = pm.end()
    return head.split(), pars


def _check_idents(tokens, what):
    for tok in tokens:
        if not _IDENT.match(tok):
            raise ParseError('Invalid {0}: "{1}"'.format(what, tok))


# ---- line handlers (reference netlistparser.py:211-371) -------------------
def parse_element(devType, instanceName, nodes, pars):
    from .devices import devClass
    try:
        dev = devClass[devType](instanceName)
    except KeyError:
        raise ParseError('Unknown device: ' + devType)
    modelname = None
    for name, value, isVector in pars:
        if name == 'model':
            modelname = value
        else:
            dev.set_param(name, match_parameter(dev, name, value, isVector))
    cktStack[-1].add_elem(dev, modelname)
    cktStack[-1].connect(dev, nodes)
    dev.check_terms()


def parse_model(modName, devType, pars):
    from .devices import devClass
    model = cir.get_model(modName)
    if not model:
        try:
            model = cir.Model(modName, devType, devClass[devType].paramDict)
        except KeyError:
            raise ParseError('Could not create model "{0}" type: {1}'.format(
                modName, devType))
        cir.add_model(model)
    if model.modelType != devType:
        raise ParseError('Model "{0}" type {1} conflicts with another model '
                         'of type {2}'.format(modName, devType,
                                              model.modelType))
    for name, value, isVector in pars:
        model.set_param(name, match_parameter(model, name, value, isVector))


def parse_options(pars):
    for name, value, isVector in pars:
        glVar.set_param(name, match_parameter(glVar, name, value, isVector))
    glVar.set_attributes(useDefaults=False)


def parse_vars(pars):
    for name, value, isVector in pars:
        try:
            if isVector:
                value = convert_type(name, value, list, True)
            elif value in ('True', 'False'):
                value = convert_type(name, value, bool, False)
            else:
                value = convert_type(name, value, float, False)
        except NetVarException as e:
            raise ParseError(str(e))
        ParamSet.netVar[name] = value


def parse_analysis(anType, pars):
    from .analyses import anClass
    try:
        an = anClass[anType]()
    except KeyError:
        raise ParseError('Unknown analysis: ' + anType)
    for name, value, isVector in pars:
        an.set_param(name, match_parameter(an, name, value, isVector))
    an.set_attributes()
    analysisQueue.q.append(an)


def _outvars(tokens):
    out = []
    for tok in tokens:
        parts = tok.split(':')
        if len(parts) == 3:
            out.append([parts[0] + ':' + parts[1], parts[2]])
        elif len(parts) == 1:
            out.append(tok)
        else:
            raise ParseError('Invalid output variable: ' + tok)
    return out


def _parse_line(line):
    if line[0] in '*#' or line.startswith('//'):
        return
    tokens, pars = _split_params(line)
    if not tokens:
        raise ParseError('Syntax error')
    key = tokens[0].lower()
    if key == '.model':
        if len(tokens) != 3:
            raise ParseError('Syntax error in .model')
        parse_model(tokens[1], tokens[2], pars)
    elif key == '.options':
        parse_options(pars)
    elif key == '.vars':
        parse_vars(pars)
    elif key == '.subckt':
        _check_idents(tokens[1:], 'name')
        cktStack.append(cir.SubCircuit(tokens[1], tokens[2:]))
    elif key == '.include':
        parse_file(tokens[1], cktStack[-1], _nested=True)
    elif key == '.analysis':
        parse_analysis(tokens[1], pars)
    elif key in ('.plot', '.save'):
        from .analyses import OutRequest
        outreq = OutRequest(tokens[1], _outvars(tokens[2:]))
        if key == '.plot':
            cktStack[0].add_plot_request(outreq)
        else:
            cktStack[0].add_save_request(outreq)
    elif key == '.ends':
        if len(cktStack) > 1 and hasattr(cktStack[-1], 'extConnectionList'):
            cktStack.pop()
        else:
            raise ParseError('.ends found but not inside a subcircuit')
    elif key == '.end':
        del cktStack[:]
    elif ':' in tokens[0]:
        devType, _, instanceName = tokens[0].partition(':')
        _check_idents([devType, instanceName] + tokens[1:], 'name')
        if len(tokens) < 2:
            raise ParseError('Element without connections')
        parse_element(devType, instanceName, tokens[1:], pars)
    elif key[0] == 'x' and not pars:
        _check_idents(tokens, 'name')
        if len(tokens) < 3:
            raise ParseError('Incomplete subcircuit instance')
        xckt = cir.Xsubckt(tokens[0], tokens[-1])
        cktStack[-1].add_subckt(xckt)
        cktStack[-1].connect(xckt, tokens[1:-1])
    else:
        raise ParseError('Syntax error')


def parse_file(filename, ckt, _nested=False):
    """Parse ``filename`` into ``ckt``; returns the list of analyses."""
    cktStack.append(ckt)
    if not _nested:
        analysisQueue.reset()
        ckt.filename = filename
    try:
        with open(filename, 'r') as f:
            lineAcc = ''
            for lineNumber, line in enumerate(f, 1):
                line = line.strip()
                if lineAcc:
                    line = lineAcc + ' ' + line
                if not line:
                    continue
                if ckt.name == 'main' and lineNumber == 1 \
                        and len(cktStack) == 1:
                    ckt.title = line
                    continue
                if line[-1] == '\\':
                    lineAcc = line[:-1]
                    continue
                lineAcc = ''
                try:
                    _parse_line(line)
                except (ParseError, NetVarException, ParamError) as pe:
                    raise ParseError('Parse error in file {0}, line {1}:\n'
                                     '{2}\n"{3}"'.format(filename,
                                                         lineNumber, pe, line))
                except cir.CircuitError as ce:
                    raise ParseError('Circuit error in file {0}, line {1}:\n'
                                     '{2}\n"{3}"'.format(filename,
                                                         lineNumber, ce, line))
                if not cktStack:
                    break
            if cktStack:
                cktStack.pop()
    except IOError as ioe:
        raise ParseError('Parse error -> ' + str(ioe))
    return analysisQueue.q
