"""The C-ABI library: builds for sm_100a, loads, exports every symbol include/umab.h declares and
refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "umab.h")).read()
    return sorted(set(re.findall(r"UMAB_API[^;]*?\b(umab_\w+)\s*\(", text)))


def test_header_and_binding_agree(built_lib):
    from pdb2reaction_b200 import engine
    syms = _header_symbols()
    assert len(syms) >= 15
    assert sorted(engine.EXPORTS) == syms
    lib = ctypes.CDLL(built_lib)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/umab.h but not exported"


def test_library_is_sm100a_only(built_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_abi_version_and_config_struct_layout(built_lib):
    from pdb2reaction_b200 import engine
    lib = engine.load_library()
    assert lib.umab_abi_version() == engine.ABI_VERSION
    assert ctypes.sizeof(engine.UmabConfig) == 8 * 4 + 2 * 4 + 8 + 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built_lib, state4, arch4):
    from pdb2reaction_b200 import engine
    lib = engine.load_library()
    cfg = engine.UmabConfig(128, 128, 64, 4, 300, 0, 0, 0, 6.0, 5.0, 0, 0)
    h = ctypes.c_void_p()
    assert lib.umab_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CPU fallback" in lib.umab_last_error()
    with pytest.raises(RuntimeError, match="no CPU"):
        engine.UmabEngine({}, [1, 1], arch4)
    from pdb2reaction_b200 import uma_pysis
    calc = uma_pysis(device="cpu")
    with pytest.raises(RuntimeError, match="no CPU"):
        calc.get_energy(["H", "H"], [0, 0, 0, 0, 0, 1.4])
    calc = uma_pysis()                      # device="auto" must not silently fall back either
    with pytest.raises(RuntimeError, match="no CPU|no CUDA"):
        calc.get_forces(["H", "H"], [0, 0, 0, 0, 0, 1.4])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pdb2reaction_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f


def test_header_is_plain_c_and_links_against_the_library(built_lib, tmp_path):
    """include/umab.h is the drop-in boundary: it must compile as C99 (no C++ / torch types in the signatures) and a C
    program must link against libumab.so and read the ABI version (no compute call: there is no GPU here)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include "umab.h"\n#include <stdio.h>\n'
                   'int main(void){ printf("%d %d\\n", (int)umab_abi_version(), (int)UMAB_ABI_VERSION);\n'
                   ' return umab_abi_version() == UMAB_ABI_VERSION ? 0 : 1; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    libdir = os.path.dirname(built_lib)
    exe = tmp_path / "t"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{inc}", str(src), "-o", str(exe),
                        f"-L{libdir}", "-lumab", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and len(set(r.stdout.split())) == 1, (r.stdout, r.stderr)
