"""tenncor_b200 — B200-native evaluation back end for TEQ functor graphs.

Mirrors the reference's python module (`import tenncor as tc`,
tenncor/python/*.cpp): `tc.EVariable`, `tc.api.*`, `tc.derive`, `tc.apply_update`, ...
The native pieces are libtcr_b200.so (CUDA kernels behind the C-ABI of
include/tcr_b200.h) and the `_tenncor` extension (C++ host: teq / eteq / layr mirror).
There is no CPU fallback: evaluation raises when the CUDA library or a device is missing.
"""
from . import cabi  # noqa: F401

try:  # the host extension is optional at import time so that build() can import the package
    from ._tenncor import *  # noqa: F401,F403
    from . import _tenncor
    HAVE_HOST = True
except ImportError as _e:  # pragma: no cover
    HAVE_HOST = False
    _HOST_IMPORT_ERROR = _e

if HAVE_HOST:
    layer_rnn = _tenncor.api.layer.rnn_on


def require_host():
    if not HAVE_HOST:
        raise ImportError("tenncor_b200._tenncor is not built: run __graft_entry__.build() (%s)" % _HOST_IMPORT_ERROR)
