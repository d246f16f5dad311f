"""resynthesizer_b200: B200-native implementation of the resynthesizer synthesis engine.

The product is the shared library resynthesizer_b200/lib/libresynthesizer_b200.so
(C-ABI in include/resynthesizer.h, include/rs_cuda.h); this package holds its
sources (csrc/), the build recipe and a ctypes mirror of the reference API.
"""
from . import abi  # noqa: F401
