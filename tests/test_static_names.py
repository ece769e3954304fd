"""CPU guard for the scripts that only run on the GPU box (bench.py, __graft_entry__.py, tools/, the package): every
name a function reads must be bound somewhere it can see -- a typo or a half-applied edit in a GPU-only code path would
otherwise surface as a NameError at round end (it happened once: bench.py, round 2)."""
import builtins
import glob
import os
import symtable

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ([os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
         + sorted(glob.glob(os.path.join(ROOT, "pdb2reaction_b200", "*.py")))
         + sorted(glob.glob(os.path.join(ROOT, "pdb2reaction_b200", "csrc", "*.py")))
         + sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))))


def _module_names(table):
    """Names bound at MODULE level (assignments, imports, defs, classes) plus the builtins."""
    names = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for sym in table.get_symbols():
        if sym.is_assigned() or sym.is_imported() or sym.is_namespace():
            names.add(sym.get_name())
    return names


def _undefined(table, known, path, out):
    for child in table.get_children():
        if child.get_type() in ("function", "class"):
            for sym in child.get_symbols():
                # a name the scope only READS and resolves globally must exist at module level (or be a builtin)
                if sym.is_global() and sym.is_referenced() and not sym.is_assigned() and sym.get_name() not in known:
                    out.append(f"{os.path.relpath(path, ROOT)}: {child.get_name()}() reads undefined name {sym.get_name()!r}")
        _undefined(child, known, path, out)


@pytest.mark.parametrize("path", FILES, ids=[os.path.relpath(p, ROOT) for p in FILES])
def test_no_function_reads_an_unbound_global(path):
    src = open(path).read()
    table = symtable.symtable(src, path, "exec")
    known = _module_names(table)
    out = []
    _undefined(table, known, path, out)
    assert not out, "\n".join(out)
