"""CPU-side checks of the drop-in boundary: the .so loads and exports exactly what include/lbt.h declares."""
import ctypes
import os
import re

from lbt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'lbt.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return set(re.findall(r'\b(lbt_[a-z0-9_]+)\s*\(', src))


def test_library_builds_and_loads():
    from lbt_b200 import build
    path = build.build()
    assert os.path.exists(path)
    h = _lib.lib()
    assert h.lbt_version() == 100
    assert h.lbt_strerror(0) == b'ok'
    assert b'invalid' in h.lbt_strerror(-1)


def test_every_declared_symbol_is_exported_and_bound():
    h = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared()
    assert declared, 'no declarations parsed from include/lbt.h'
    for name in declared:
        assert hasattr(h, name), 'include/lbt.h declares %s but liblbt_b200.so does not export it' % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))


def test_argument_validation_without_gpu():
    h = _lib.lib()
    # null tensor / bad bits are rejected before any CUDA call
    assert h.lbt_quantize(None, 1, 1, 8, None, 0.0, 0, None, 0, 0, None, None, None, 0, None, 0, None) == -1
    buf = (ctypes.c_float * 4)()
    ib = (ctypes.c_int32 * 1)(2)
    p = ctypes.cast(buf, ctypes.c_void_p)
    q = ctypes.cast(ib, ctypes.c_void_p)
    assert h.lbt_quantize(p, 1, 4, 1, q, 0.0, 0, None, 0, 0, None, p, None, 0, None, 0, None) == -1     # bits < 2
    assert h.lbt_quantize(p, 1, 4, 33, q, 0.0, 0, None, 0, 0, None, p, None, 0, None, 0, None) == -1    # bits > 24
    assert h.lbt_quantize(p, 1, 4, 8, q, 0.0, 1, None, 0, 0, None, p, None, 0, None, 0, None) == -1     # noise mode, no noise
    assert h.lbt_quantize(p, 1, 4, 9, q, 0.0, 0, None, 0, 0, None, None, p, 1, None, 0, None) == -1     # 9 bits into s8
    assert h.lbt_quantize(p, 1, 4, 8, q, 0.0, 0, None, 0, 0, None, p, None, 0, None, 1, None) == -1     # update w/o counters
    assert h.lbt_quantize(p, 0, 4, 8, q, 0.0, 0, None, 0, 0, None, p, None, 0, None, 0, None) == 0      # empty tensor


def test_cpu_tensor_is_an_error():
    import pytest
    import torch
    from lbt_b200 import quantizer
    with pytest.raises(_lib.LbtError):
        quantizer.quantize(torch.zeros(4, 4), 8, torch.tensor(2, dtype=torch.int32))
