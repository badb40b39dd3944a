"""Compare two graphs node by node (reference: dbg/compare/src/equal.cpp; python module dbg/python/compare.cpp). Nodes are paired
by their post-order index (teq::GraphIndex); leaves in the same position count as the same leaf, whatever they hold."""
import numpy as np

import tenncor_b200 as tc


def _ordered(root):
    return [t for t, _ in sorted(tc.teq.graph_index(root), key=lambda pair: pair[1])]


def _pairs(lroot, rroot):
    left, right = _ordered(lroot), _ordered(rroot)
    return list(zip(left, right)) if len(left) == len(right) else None


def is_equal(lroot, rroot):
    """Return true if lroot and rroot graphs are structurally equal: same opcodes, attributes, shapes and wiring"""
    if lroot == rroot:
        return True
    ldump, rdump = tc.dump_graph([lroot]), tc.dump_graph([rroot])
    if len(ldump) != len(rdump):
        return False
    for ln, rn in zip(ldump, rdump):
        if ln["kind"] != rn["kind"] or list(ln["shape"]) != list(rn["shape"]) or ln["dtype"] != rn["dtype"]:
            return False
        if ln["kind"] == "func" and (ln["op"] != rn["op"] or list(ln["args"]) != list(rn["args"]) or dict(ln["attrs"]) != dict(rn["attrs"])):
            return False
    return True


def _data(t):
    if t.is_leaf() or tc.testing.holder_ptr(t) != 0:
        return np.asarray(t.data())
    return None  # a functor that has not been evaluated (or whose result was consumed) holds nothing to compare


def _same(l, r):
    return l.teq_shape() == r.teq_shape() and l.dtype() == r.dtype() and np.array_equal(_data(l), _data(r))


def is_dataeq(lroot, rroot):
    """Return true if lroot and rroot graphs have the same data in every node"""
    if lroot == rroot:
        return True
    pairs = _pairs(lroot, rroot)
    if pairs is None:
        return False
    return all(_data(l) is not None and _data(r) is not None and _same(l, r) for l, r in pairs)


def percent_dataeq(lroot, rroot):
    """Return the fraction of nodes holding data on both sides that are data equivalent"""
    if lroot == rroot:
        return 1.
    pairs = _pairs(lroot, rroot)
    if pairs is None:
        return 0.
    valid = [(l, r) for l, r in pairs if _data(l) is not None and _data(r) is not None]
    return sum(_same(l, r) for l, r in valid) / len(valid) if valid else 0.
