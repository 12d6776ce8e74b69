"""CPU-side checks of the C-ABI library: it loads and exports every symbol that
include/*.h declares.  No compute call is made (there is no GPU here)."""
import ctypes

from parelag_b200 import capi


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    names = capi.declared_symbols()
    assert len(names) > 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/*.h but not exported: %s" % missing


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        return
    try:
        capi.Ctx()
    except capi.PEError as e:
        assert "no CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("context creation must fail loudly without a CUDA device")
