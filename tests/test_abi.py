"""The C-ABI library loads on a machine without a GPU and exports every symbol include/msm_b200.h
declares; compute entry points fail with a code (never crash, never fall back to the CPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from msm_zprize_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.lib()


def _declared():
    src = open(os.path.join(ROOT, "include", "msm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msm_b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from msm_zprize_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names


def test_struct_sizes_match_header(lib):
    from msm_zprize_b200 import _lib
    assert ctypes.sizeof(_lib.Point) == 100
    # 9 floats + 5 ints + 8-byte counter, 8-byte aligned
    assert ctypes.sizeof(_lib.Timing) == 64


def test_no_silent_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from msm_zprize_b200 import MsmEngine, MsmError
    with pytest.raises(MsmError):
        MsmEngine("bls12-377", device=0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "msm_zprize_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and f != "hosttest.cpp":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|bigint_oracle|libmsm_port|oracle/", txt, flags=re.M), f
