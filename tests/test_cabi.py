"""The C-ABI shared library loads and exports every symbol include/gplum_b200.h declares.
CPU only: no compute call succeeds without a GPU, and the library says so loudly."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "gplum_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(gplum_b200_\w+)\s*\(", h)))


def test_header_and_binding_agree():
    from gplum_b200 import _lib
    assert _declared() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from gplum_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    h = C.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(h, name), name
    assert _lib.lib().gplum_b200_abi_version() == 3
    e, s = C.c_int(0), C.c_int(0)
    _lib.lib().gplum_b200_packed_sizes(C.byref(e), C.byref(s))
    assert (e.value, s.value) == (48, 64)


def test_no_cpu_fallback():
    """Without a CUDA device a compute call must fail with an error, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from gplum_b200 import _lib, functors, structs as S
    epi = np.zeros(1, S.EPI); epj = np.zeros(1, S.EPJ); f = S.cleared_force(1)
    with pytest.raises(_lib.GplumB200Error):
        functors.calcForceEPEPWithSearch(0.0)(epi, 1, epj, 1, f)
    assert f.tobytes() == S.cleared_force(1).tobytes()


def test_product_never_touches_oracle():
    """Only tests/, smoke() and bench.py's CPU-baseline legs may reference oracle/."""
    pkg = os.path.join(ROOT, "gplum_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(d, f)).read()
                for bad in ("liboracle", "oracle_api", "pikg_oracle", "oracle/", "import oracle", "libgplum_ref"):
                    assert bad not in txt.replace("tests/test_oracle_vs_ref.py", ""), (os.path.join(d, f), bad)
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "liboracle" not in open(p).read()
