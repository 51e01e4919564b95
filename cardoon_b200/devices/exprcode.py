"""
Tiny expression compiler for devices whose law is a netlist string
(memristor ``m``, reference devices/memristor.py:143,189; nonlinear ``vccs``
``f``, vccs.py:76-90).

The reference ``eval()``s the string inside the AD-taped function.  The
netlist grammar restricts such strings to alphanumerics, ``._+-*:``, blanks
and parentheses (netlistparser.py:56, :388), so expressions are made of
numbers, one variable, ``abs``, ``np.<fn>`` calls and ``+ - * **``.  They are
compiled (via ``ast``, nothing is executed) to a postfix program that the
CUDA kernel interprets in dual-number arithmetic
(csrc/models/exprvm.cuh).  Encoding: a flat list of doubles
``[n, op0, arg0, op1, arg1, ...]``.
"""
import ast

OP_CONST, OP_VAR, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG, OP_ABS, \
    OP_EXP, OP_LOG, OP_SQRT, OP_TANH, OP_SIN, OP_COS, OP_TAN, OP_COSH = range(17)

_FUNCS = dict(abs=OP_ABS, exp=OP_EXP, log=OP_LOG, sqrt=OP_SQRT,
              tanh=OP_TANH, sin=OP_SIN, cos=OP_COS, tan=OP_TAN, cosh=OP_COSH)
_BIN = {ast.Add: OP_ADD, ast.Sub: OP_SUB, ast.Mult: OP_MUL, ast.Div: OP_DIV,
        ast.Pow: OP_POW}

MAXCODE = 64


class ExprError(Exception):
    pass


def compile_expr(text, varname):
    """Postfix program of ``text`` (a function of ``varname``)."""
    try:
        tree = ast.parse(text.strip(), mode='eval')
    except SyntaxError as e:
        raise ExprError(str(e))
    code = []

    def emit(node):
        if isinstance(node, ast.Expression):
            emit(node.body)
        elif isinstance(node, ast.Constant) and \
                isinstance(node.value, (int, float)):
            code.extend((OP_CONST, float(node.value)))
        elif isinstance(node, ast.Name):
            if node.id != varname:
                raise ExprError('unknown name: ' + node.id)
            code.extend((OP_VAR, 0.))
        elif isinstance(node, ast.UnaryOp) and \
                isinstance(node.op, (ast.USub, ast.UAdd)):
            emit(node.operand)
            if isinstance(node.op, ast.USub):
                code.extend((OP_NEG, 0.))
        elif isinstance(node, ast.BinOp) and type(node.op) in _BIN:
            emit(node.left)
            emit(node.right)
            code.extend((_BIN[type(node.op)], 0.))
        elif isinstance(node, ast.Call) and len(node.args) == 1 \
                and not node.keywords:
            f = node.func
            if isinstance(f, ast.Attribute) and \
                    isinstance(f.value, ast.Name) and f.value.id == 'np':
                name = f.attr
            elif isinstance(f, ast.Name):
                name = f.id
            else:
                raise ExprError('unsupported call')
            if name not in _FUNCS:
                raise ExprError('unsupported function: ' + name)
            emit(node.args[0])
            code.extend((_FUNCS[name], 0.))
        else:
            raise ExprError('unsupported syntax: ' + ast.dump(node))

    emit(tree)
    if len(code) // 2 > MAXCODE:
        raise ExprError('expression too long')
    return [float(len(code) // 2)] + [float(c) for c in code]
