"""libxsmm-1_b200: B200-native (sm_100a) implementation of LIBXSMM's sparse-A x dense-B hot path.

The product is the C-ABI shared library ``lib/libxsmm_b200.so`` (sources in ``csrc/``, headers in
``/include``).  This package is only its Python-side mirror -- the same function names, argument
order and semantics as the reference's C interface (include/libxsmm_spmdm.h:74-132,
include/libxsmm_fsspmdm.h:41-57) -- used by the tests and bench.py.  There is no CPU fallback: if the
library is missing, or no GPU is present when a compute entry is called, it fails loudly.

Import with ``importlib.import_module("libxsmm-1_b200")`` (the directory name is not an identifier).
"""
from .api import *          # noqa: F401,F403
from . import workloads     # noqa: F401
from . import sharding      # noqa: F401
