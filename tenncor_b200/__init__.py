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


if HAVE_HOST:
    _optimize_graphs = _tenncor.optimize

    def optimize(roots, fold_constants=True, ctx=None):
        """hone rewrites (duplicate merging, constant folding) of the graphs under `roots`: returns (new roots, stats).

        The reference's python entry point is `tc.optimize(rule_file, ctx)`, which rewrites every graph registered in the context in
        place, driven by a json rule file. There is no registry here to rewrite in place and the rule file is not part of the repository,
        so a rule-file argument is accepted and does nothing (returns None): the launch planner merges and inlines at lowering time what
        those rules would, and scripts written for the reference run unchanged. Pass the roots to get the explicit passes."""
        if isinstance(roots, str):
            return None
        return _optimize_graphs(list(roots), fold_constants)


class Shape(list):
    """`tc.Shape([...])` of the reference's python module (tenncor/python/eteq_ext.cpp): dimensions in numpy order. The API here takes
    plain lists wherever the reference takes a Shape, so this is a list that prints like one."""

    def __repr__(self):
        return "Shape(%s)" % list.__repr__(self)


class Context:
    """`tc.Context` / `tc.global_context`: the reference keeps its tensor registry, evaluator and RNG in a context object and its python
    calls take an optional `ctx`. This back end has ONE process-wide context (registry-free: versions come from a counter, the
    evaluator from teq::set_eval), so the object only exists for scripts that pass it around."""


global_context = Context()


def TenncorAPI(ctx=None):
    """`tc.TenncorAPI(ctx)`: the API object bound to a context — here always the module-level `tc.api`"""
    return _tenncor.api


def require_host():
    if not HAVE_HOST:
        raise ImportError("tenncor_b200._tenncor is not built: run __graft_entry__.build() (%s)" % _HOST_IMPORT_ERROR)
