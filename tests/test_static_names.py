"""A poor man's pyflakes (no linter in the image): every global name a function refers to must exist at module level or
be a builtin.  Large parts of the product only execute on a GPU box — a typo in such a branch would otherwise surface
as a NameError in the middle of a GPU run."""
import builtins
import symtable
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
FILES = sorted(p for pat in ('dmlcloud_b200/**/*.py', 'oracle/*.py', 'profiles/*.py', 'examples/*.py', 'tests/*.py', '*.py')
               for p in ROOT.glob(pat) if '_ref' not in p.parts and 'shims' not in p.parts)


def _undefined(path):
    top = symtable.symtable(path.read_text(), str(path), 'exec')
    defined = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    defined |= set(dir(builtins)) | {'__file__', '__name__', '__doc__', '__package__', '__spec__', '__builtins__'}
    missing = []

    def walk(table):
        for sym in table.get_symbols():
            if sym.is_referenced() and sym.is_global() and not sym.is_assigned() and sym.get_name() not in defined:
                # `global x` inside a function that assigns x counts as assigned; anything left is unresolvable
                missing.append(f'{table.get_name()}:{table.get_lineno()}: {sym.get_name()}')
        for child in table.get_children():
            walk(child)

    for child in top.get_children():
        walk(child)
    for sym in top.get_symbols():  # module-level references
        if sym.is_referenced() and not (sym.is_assigned() or sym.is_imported() or sym.is_namespace()) \
                and sym.get_name() not in defined:
            missing.append(f'<module>: {sym.get_name()}')
    # names assigned through `global` declarations inside functions
    declared = set()

    def collect(table):
        for sym in table.get_symbols():
            if sym.is_declared_global() and sym.is_assigned():
                declared.add(sym.get_name())
        for child in table.get_children():
            collect(child)

    collect(top)
    return [m for m in missing if m.rsplit(' ', 1)[-1] not in declared]


@pytest.mark.parametrize('path', FILES, ids=lambda p: str(p.relative_to(ROOT)))
def test_no_undefined_global_names(path):
    assert _undefined(path) == []


def _dmlb_calls(path):
    import ast

    tree = ast.parse(path.read_text(), str(path))
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith('dmlb_'):
            yield node


@pytest.mark.parametrize('path', FILES, ids=lambda p: str(p.relative_to(ROOT)))
def test_c_abi_calls_pass_the_declared_number_of_arguments(path):
    """ctypes only complains about a wrong argument count when the call executes — on a GPU box for most call sites.
    Every `<lib>.dmlb_xxx(...)` call in the tree is checked against dmlcloud_b200/_native.py SIGNATURES here."""
    import ast
    import sys

    sys.path.insert(0, str(ROOT))
    from dmlcloud_b200 import _native as N

    for call in _dmlb_calls(path):
        name = call.func.attr
        assert name in N.SIGNATURES, f'{path.name}:{call.lineno}: {name} is not declared'
        assert not call.keywords, f'{path.name}:{call.lineno}: keyword arguments in a C call'
        want = len(N.SIGNATURES[name][1])
        fixed = [a for a in call.args if not isinstance(a, ast.Starred)]
        if len(fixed) == len(call.args):
            assert len(call.args) == want, f'{path.name}:{call.lineno}: {name} takes {want} arguments, {len(call.args)} given'
        else:  # a *tuple in the call: at least the explicit ones must fit
            assert len(fixed) < want, f'{path.name}:{call.lineno}: {name} takes {want} arguments'
