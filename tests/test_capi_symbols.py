"""CPU-only checks of the drop-in boundary: the library loads and exports every entry point that
include/abcsmc_b200.h declares; creating a context without a GPU fails loudly (no CPU fallback)."""
import ctypes as C

import pytest

from abcsmc_b200 import _capi


def test_library_exports_every_declared_symbol():
    _capi.build()
    lib = _capi.lib()
    declared = _capi.declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_capi._SIGS) == set(declared)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from abcsmc_b200 import api
    with pytest.raises(_capi.Abcb200Error):
        api.Context(0)
    h = C.c_void_p()
    assert _capi.lib().abcb200_create(0, C.byref(h)) == -2   # ABCB200_ENODEV


def test_header_is_plain_c():
    """The boundary is a C ABI: include/abcsmc_b200.h compiles as C99 and as C++17 with -Wpedantic (no torch / CUDA / C++ types)."""
    import subprocess
    for lang, std in (("c", "-std=c99"), ("c++", "-std=c++17")):
        subprocess.check_call(["gcc", "-x", lang, std, "-Wall", "-Wpedantic", "-Werror", "-fsyntax-only", _capi.HEADER_PATH])
