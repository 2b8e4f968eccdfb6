"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/florence_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "florence_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from florence_b200 import _lib, build
    build.build()
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libflorence_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared


def test_library_is_built_for_sm_100a():
    from florence_b200 import _lib, build
    import subprocess
    build.build()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from florence_b200 import _lib, backend
    import numpy as np
    with pytest.raises(_lib.FlorenceB200Error):
        backend.AssemblyHandle(np.zeros((4, 3)), np.zeros((1, 4), np.uint64), np.zeros((3, 4, 1)), np.ones(1))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "florence_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), "%s mentions the oracle" % f
