"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/tppmx.h declares, and fails loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import pytest

from treatppmx_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tppmx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tppmx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    if not os.path.exists(_lib.SO_PATH):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(_lib.SO_PATH)
    decl = _declared_symbols()
    assert len(decl) >= 13
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/tppmx.h but not exported"
    assert sorted(_lib.SYMBOLS) == decl
    L.tppmx_version.restype = C.c_int
    assert L.tppmx_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from treatppmx_b200.engine import Context
    with pytest.raises(_lib.TppmxError):
        Context(0)


def test_struct_sizes_match_header():
    """ctypes mirrors must agree with the C structs (compiled probe)."""
    import subprocess, tempfile
    from treatppmx_b200 import _abi
    code = r'''
#include <stdio.h>
#include "tppmx.h"
int main(){printf("%zu %zu %zu %zu %zu\n", sizeof(tppmx_model), sizeof(tppmx_dims),
 sizeof(tppmx_shared), sizeof(tppmx_dataset), sizeof(tppmx_state));return 0;}
'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        open(src, "w").write(code)
        exe = os.path.join(td, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = list(map(int, subprocess.check_output([exe]).split()))
    assert sizes == [C.sizeof(_abi.Model), C.sizeof(_abi.Dims), C.sizeof(_abi.Shared),
                     C.sizeof(_abi.Dataset), C.sizeof(_abi.State)]
