"""CPU check: libvecvad.so loads and exports every symbol include/vecvad.h declares (no compute calls)."""
import ctypes
import os
import re

from vec_vad_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(REPO, 'include', 'vecvad.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(vecvad_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), 'libvecvad.so does not export %s' % n
    assert sorted(_lib.SYMBOLS) == names, 'vec_vad_b200/_lib.py SYMBOLS out of sync with include/vecvad.h'


def test_abi_version_and_error_string():
    L = _lib.lib()
    assert L.vecvad_abi_version() == _lib.ABI_VERSION
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert L.vecvad_correlation_out_shape(48, 64, 20, 1, 20, 1, 2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oc.value, oh.value, ow.value) == (441, 48, 64)
    # bad arguments return a negative status and set the message instead of aborting (correlation_cuda.c:87-89 aborts)
    assert L.vecvad_correlation_out_shape(4, 4, 0, 1, 20, 1, 2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) < 0
    assert b'empty output' in L.vecvad_last_error()


def test_net_create_validates_configuration():
    L = _lib.lib()
    cfg = _lib.NetConfig()
    h = ctypes.c_void_p()
    assert L.vecvad_net_create(ctypes.byref(cfg), ctypes.byref(h)) < 0          # n_unets == 0
    assert b'n_unets' in L.vecvad_last_error()
