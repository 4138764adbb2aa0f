"""The C-ABI library loads without a GPU and exports every symbol include/chap_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "chap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chap_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from chap_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libchap_b200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "chap_b200/_lib.py does not bind %s" % s
    assert set(_lib.SIGNATURES) == set(syms)


def test_library_loads_and_reports_version():
    from chap_b200 import _lib
    lib = _lib.load()
    assert lib.chap_abi_version() == 1
    assert lib.chap_get_force_simt() in (0, 1)
    lib.chap_set_conv_precision(32)
    assert lib.chap_get_conv_precision() == 32
    lib.chap_set_conv_precision(0)
    assert isinstance(_lib.launch_count(), int)


def test_bad_arguments_fail_loudly_without_touching_the_gpu():
    from chap_b200 import _lib
    lib = _lib.load()
    d = _lib.ConvDesc(0, 4, 1, 1, 8, 8, 4, 4)          # nd = 4 is invalid
    rc = lib.chap_conv_fwd(ctypes.byref(d), None, None, None, None, None, None)
    assert rc == -1
    assert b"nd must be 2 or 3" in lib.chap_last_error()
    assert lib.chap_conv_packed_elems(ctypes.byref(_lib.ConvDesc(0, 3, 2, 4, 4, 4, 16, 32))) == 2 * 27 * 16 * 32      # TF32 hi half + lo half
    # channel counts below 16 are zero-padded to 16 in the packed operand (tensor-core path of the 4- / 8-channel heads)
    assert lib.chap_conv_packed_elems(ctypes.byref(_lib.ConvDesc(3, 2, 2, 1, 4, 4, 16, 8))) == 2 * 4 * 16 * 16


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from chap_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.softmax(torch.zeros(1, 4, 8, 8))
