"""teq::derive's traversal (internal/teq/src/derive.cpp:13-161) against internal/teq/test/test_grad.cpp: the
reference checks, with a mock builder, WHICH local rules are applied with WHICH upstream gradient and in what ORDER the
contributions of a shared node are summed. Here the real builder is used and the same facts are read off the structure of the
graph that comes out (the accumulation order is the order of the ADD's arguments)."""
import numpy as np
import pytest

import tenncor_b200 as tc


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def leaf(label):
    return tc.variable(np.linspace(0.2, 1.4, 6).reshape(2, 3), label)


def ops(t):
    """pre-order opcode / label list of a graph"""
    out = [t.label() if t.is_leaf() else t.opname()]
    if not t.is_leaf():
        for a in t.args():
            out += ops(a)
    return out


def test_one_and_zero():  # GRAD.OneZero :8-47
    x, y, z = leaf("leaf"), leaf("leaf2"), leaf("leaf3")
    f = x * y
    one = tc.derive(f, [f])[0]
    assert one.is_leaf() and one.usage() == "constant" and one.label() == "1" and one.shape() == f.shape()
    assert tc.derive(x, [x])[0].label() == "1"
    for root, target in [(x, z), (z, x), (f, z)]:  # no path: a constant zero shaped like the target
        zero = tc.derive(root, [target])[0]
        assert zero.is_leaf() and zero.label() == "0" and zero.shape() == target.shape()


def test_standard_v():  # GRAD.BuilderStandardV :48-79 — one local rule per argument, seeded with the constant one
    x, y = leaf("leaf"), leaf("leaf2")
    f = x * y
    gx, gy = tc.derive(f, [x, y])
    assert ops(gx) == ["MUL", "leaf2", "1"] and ops(gy) == ["MUL", "leaf", "1"]  # d(xy)/dx = y * 1


def test_diamond_accumulation_order():  # GRAD.BuilderDiamond :80-128
    """functors of equal height are visited by descending NAME, then descending post-order index (derive.cpp:110-134): in the
    reference's test FUNC2 precedes FUNC, so the sum is add({via FUNC2, via FUNC}). Same rule here with COS / SIN as the names."""
    x = leaf("leaf")
    f, f2 = tc.api.cos(x), tc.api.sin(x)              # "SIN" > "COS": f2 is visited first
    g = tc.derive(f * f2, [x])[0]
    assert g.opname() == "ADD" and len(g.args()) == 2
    via_f2, via_f = g.args()
    assert ops(via_f2)[:2] == ["MUL", "COS"]           # sin' = cos, reached through the product's argument 1
    assert ops(via_f)[:3] == ["MUL", "NEG", "SIN"]     # cos' = -sin, reached through argument 0
    # the upstream gradient each local rule received is the product's rule for that argument: d(f*f2)/df2 = f, d(f*f2)/df = f2
    assert ops(via_f2.args()[1]) == ["MUL", "COS", "leaf", "1"] and ops(via_f.args()[1]) == ["MUL", "SIN", "leaf", "1"]
    # equal names: the node created later (larger post-order index) goes first
    a, b = tc.api.square(x), tc.api.square(x * 1.0)
    g2 = tc.derive(a * b, [x])[0]
    first, second = g2.args()
    assert "MUL" in ops(first)[3:] and ops(second)[:4] == ["MUL", "MUL", "EXTEND", "2"]  # b's chain (through x * 1.0) is summed first


def test_symmetrical_diamond():  # GRAD.SymmetricalDiamond :129-165 — both arguments are the SAME node: summed first, one rule below
    x = leaf("leaf")
    f = tc.api.sin(x)
    f2 = f * f
    g = tc.derive(f2, [x])[0]
    assert ops(g)[:3] == ["MUL", "COS", "leaf"]         # exactly one application of sin's rule ...
    upstream = g.args()[1]
    assert upstream.opname() == "ADD" and len(upstream.args()) == 2  # ... on add({d/darg0, d/darg1})
    assert [ops(a) for a in upstream.args()] == [["MUL", "SIN", "leaf", "1"]] * 2


def test_tadpole():  # GRAD.TadPole :166-210 — diamond with a tail: contributions meet at FUNC, then one rule for the tail
    x = leaf("leaf")
    f = tc.api.exp(x)
    f2, f3 = tc.api.cos(f), tc.api.sin(f)               # FUNC2 = COS, FUNC3 = SIN: FUNC3 is visited first, as in the reference
    g = tc.derive(f2 * f3, [x])[0]
    assert ops(g)[:3] == ["MUL", "EXP", "leaf"]          # exp' = exp, applied once
    meet = g.args()[1]
    assert meet.opname() == "ADD"
    assert ops(meet.args()[0])[:2] == ["MUL", "COS"] and ops(meet.args()[1])[:3] == ["MUL", "NEG", "SIN"]  # add({via FUNC3, via FUNC2})
