"""pmclib_b200 — B200-native Population Monte Carlo iteration behind pmclib's C API.

The product is the C-ABI shared library ``libpmcb200.so`` (CUDA, sm_100a; sources in ``csrc/``, interface in
``include/pmcb200.h`` and ``include/pmclib_abi.h``).  This package only holds the in-tree build script, a ctypes
loader and thin numpy wrappers used by the tests and by bench.py.
"""
from .lib import load_library, LIB_PATH, PmcError  # noqa: F401
from .api import PmcGpu, StepResult  # noqa: F401
