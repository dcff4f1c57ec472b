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
