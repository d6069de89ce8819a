"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every symbol
include/fse_b200.h declares; host-only helpers answer; compute calls fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from entropy_coders_b200 import _capi, build
    build.build()
    return _capi.lib()


def test_header_symbols_exported(L):
    from entropy_coders_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "fse_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(fse_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(_capi.SYMBOLS) == declared
    for s in declared:
        assert hasattr(L, s), s


def test_sizing_helpers(L):
    from entropy_coders_b200 import _capi
    assert L.fse_b200_compress_bound(65536) == 66572          # fse.rs:191-193
    assert L.fse_b200_num_blocks(256 << 20, 65536) == 4096
    assert L.fse_b200_num_blocks(65537, 65536) == 2
    p = _capi.Params(65536, 0, 32, 0, 0, 0)
    assert L.fse_b200_compress_blocks_bound(1 << 20, C.byref(p)) >= 16 * 66572
    assert b"sm_100a" in L.fse_b200_version()


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert L.fse_b200_create(0, None, C.byref(h)) == -10      # FSE_B200_ERR_CUDA
    import entropy_coders_b200 as E
    with pytest.raises(RuntimeError):
        E.Context(0)
    with pytest.raises(RuntimeError):
        E.fse_compress2(b"hello world, hello world", bytearray())


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "entropy_coders_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "fse_oracle" not in txt and "oracle_lib" not in txt and "pymodel" not in txt, f
