"""parelag_b200 -- B200-native AMGe solve-and-coarsen path behind ParElag's API.

The product is the C-ABI shared library (parelag_b200/lib/libparelag_b200.so, built
from csrc/*.cu and src/*.cpp); `capi` is its ctypes binding.  Nothing here falls
back to the CPU: without the built library or without a CUDA device, calls raise.
"""
from . import capi  # noqa: F401
