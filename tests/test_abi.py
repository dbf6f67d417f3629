"""The C-ABI shared library: loads without a GPU, exports every symbol include/cmoop_b200.h
declares, and fails loudly (no CPU fallback) when no device is usable.  CPU only."""
import os
import re

import numpy as np
import pytest

from cmoop_audio_processing_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cmoop_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmoop_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run python -m cmoop_audio_processing_b200.build"
    lib = _lib.load()
    assert lib.cmoop_abi_version() >= 1


def test_every_header_symbol_is_exported_and_bound():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    for name in _lib.SIGNATURES:
        assert name in syms, f"{name} bound in Python but not declared in include/cmoop_b200.h"


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "cmoop_b200.h")).read()
    for cite in ("nsga_penalty.py:448-524", "sa_nsga_local.py:212-223", "mobo_penalty.py:265-273", "compare.ipynb"):
        assert cite in text


def test_no_silent_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.cmoop_device_count() > 0:
        pytest.skip("a GPU is present")
    from cmoop_audio_processing_b200.nsga import fast_non_dominated_sort
    with pytest.raises(_lib.CmoopError, match="no CPU fallback"):
        fast_non_dominated_sort([{"hparams": {}, "objs": [0.0, 1.0], "CV": 0.0}], 1.0)
    from cmoop_audio_processing_b200.features import MfccFrontEnd
    with pytest.raises(_lib.CmoopError):
        MfccFrontEnd()


def test_argument_validation_needs_no_device():
    lib = _lib.load()
    objs = np.zeros((4, 3))
    st = lib.cmoop_nds_crowding_host(_lib.ptr(objs), None, 4, 99, 1, 1.0, 1e-6, 0, None, None, None, None, None)
    assert st == -1 and b"m=99" in lib.cmoop_last_error()
    st = lib.cmoop_nds_crowding_host(_lib.ptr(objs), None, 100000, 3, 1, 1.0, 1e-6, 0, None, None, None, None, None)
    assert st == -1
    out = np.zeros(1)
    assert lib.cmoop_hypervolume_host(_lib.ptr(objs), 4, 5, _lib.ptr(objs), _lib.ptr(out)) == -1
    # empty problems are answered on the host
    nf = np.full(1, -7, np.int32)
    assert lib.cmoop_nds_crowding_host(None, None, 0, 3, 1, 1.0, 1e-6, 0, None, None, None, _lib.ptr(nf), None) == 0
    assert nf[0] == 0
    assert lib.cmoop_nds_workspace_bytes(512, 3, 1) == 0
    assert lib.cmoop_nds_workspace_bytes(4096, 3, 2) > 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cmoop_audio_processing_b200")
    for root, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
