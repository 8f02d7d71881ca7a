"""The C-ABI library loads without a GPU and exports exactly what include/mqcb200.h declares."""
import os
import re

import pytest

from metalquicha_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "mqcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mqcb200_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mqcb200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes prototypes drifted from the header"


def test_version_and_error_buffer():
    lib = _lib.load()
    assert lib.mqcb200_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_bad_handles_are_refused_not_dereferenced():
    import ctypes
    lib = _lib.load()
    assert lib.mqcb200_destroy(None) == _lib.MQCB200_BAD_HANDLE
    assert "names no engine" in _lib.last_error()
    n = ctypes.c_int(0)
    assert lib.mqcb200_last_launches(None, ctypes.byref(n)) == _lib.MQCB200_BAD_HANDLE
    assert lib.mqcb200_queue_destroy(None) == _lib.MQCB200_BAD_HANDLE


def test_create_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without CUDA the engine refuses to exist."""
    import ctypes
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    lib = _lib.load()
    h = ctypes.c_void_p(None)
    assert lib.mqcb200_create(0, ctypes.byref(h)) == _lib.MQCB200_FAIL
    assert h.value is None
    assert "no CPU fallback" in _lib.last_error()
    assert lib.mqcb200_create(0, None) == _lib.MQCB200_FAIL


def test_library_is_sm100a_with_tma_and_dmma():
    """The shipped cubin is sm_100a and its hot kernels use bulk-TMA + DMMA."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "DMMA.8x8x4" in sass
    assert "UBLKCP" in sass
