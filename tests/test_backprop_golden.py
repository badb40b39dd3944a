"""The per-opcode gradient rules (DerivativeFuncs::lderive, tenncor/eteq/backprop.hpp:58-576) against the
reference's own structural goldens: tenncor/eteq/test/test_backprop.cpp asserts, for 33 opcodes, the exact derivative GRAPH
(EXPECT_GRAPHEQ on the PrettyEquation rendering). tests/golden/backprop_goldens.json holds those strings; here each operand
set is rebuilt through our host, `lderive` is called, and our graph is rendered in the same format and compared verbatim.

Two normalisations, both forced by the reference's use of mocks: its leaves are MockLeaf objects whose usage prints as
"constant" (ours are variables), and in the unary / binary / select cases its functor is a mock named "op" (ours is the real
functor): leaves are rendered as `constant:<label>`, and the differentiated functor itself as `op` where the golden does."""
import json
import os

import numpy as np
import pytest

import tenncor_b200 as tc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "backprop_goldens.json")))


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def render(root, op=None):
    """dbg/print/teq.hpp PrettyEquation with the test harness's '_' for spaces (testutil/tutil.hpp EXPECT_GRAPHEQ)"""
    def text(t):
        shape = "\\".join(str(d) for d in t.teq_shape())
        dtype = {"float64": "DOUBLE", "float32": "FLOAT", "int32": "INT32"}[str(t.dtype())]
        if t.is_leaf():
            name = "constant:" + t.label()
        else:
            name = "op" if op is not None and t is op else t.opname()
        return "(%s<%s>[%s])" % (name, dtype, shape)

    out = []

    def rec(t, ancestors_last):
        depth = len(ancestors_last)
        if depth == 0:
            out.append(text(t))
        else:
            out.append("_" + "".join("____" if last else "|___" for last in ancestors_last[:-1]) + "`--" + text(t))
        if t.is_leaf():
            return
        kids = t.args()
        for i, k in enumerate(kids):
            rec(k, ancestors_last + [i == len(kids) - 1])

    rec(root, [])
    return "\n".join(out) + "\n"


def var(shape, label, dtype=np.float64):
    return tc.variable(np.arange(1, int(np.prod(shape)) + 1, dtype=dtype).reshape(shape[::-1]), label)  # teq shape -> numpy shape


def F(opname, args, attrs=None):
    return tc.egen.make_functor(opname, args, attrs or {})


def std_operands():
    return var([3, 2], "super"), var([3, 2], "arg1"), var([3, 2], "arg2")


UNARY = {"Neg": "NEG", "Tan": "TAN", "Log": "LOG", "Sqrt": "SQRT", "Abs": "ABS", "Sin": "SIN", "Cos": "COS", "Exp": "EXP",
         "Square": "SQUARE", "Cube": "CUBE", "Sigmoid": "SIGMOID", "Tanh": "TANH"}
BINARY = {"Pow": "POW", "Mul": "MUL", "MinMax": "MAX", "Sub": "SUB", "Div": "DIV"}


@pytest.mark.parametrize("case", sorted(UNARY))
def test_unary_rules(case):  # unary_derivative(): lderive(op(arg1), super, 0)
    sup, a1, _ = std_operands()
    op = F(UNARY[case], [a1])
    assert render(tc.egen.lderive(op, sup, 0), op) == GOLD[case]["graphs"][0], GOLD[case]["cite"]


@pytest.mark.parametrize("case", sorted(BINARY))
def test_binary_rules(case):  # binary_derivative(): lderive(op(arg1, arg2), super, 1)
    sup, a1, a2 = std_operands()
    op = F(BINARY[case], [a1, a2])
    assert render(tc.egen.lderive(op, sup, 1), op) == GOLD[case]["graphs"][0], GOLD[case]["cite"]


def test_passthrough():  # :68-89 — IDENTITY hands the upstream gradient through untouched
    sup, a1, a2 = std_operands()
    assert tc.egen.lderive(F("IDENTITY", [a1, a2]), sup, 1) is sup


STRUCTURAL = {
    # name: (super shape, [arg shapes], opcode, attrs, arg index)
    "ReduceSum": ([3], [[3, 2]], "REDUCE_SUM", {"rank_set": {1}}, 0),
    "ReduceProd": ([3], [[3, 2]], "REDUCE_PROD", {"rank_set": {1}}, 0),
    "ReduceMinMax": ([3], [[3, 2]], "REDUCE_MAX", {"rank_set": {1}}, 0),
    "Extend": ([3, 2, 4], [[3, 2]], "EXTEND", {"dimensions": [1, 1, 4]}, 0),
    "Permute": ([3, 2], [[3, 2]], "PERMUTE", {"ranks": [1, 2, 0]}, 0),
    "Reshape": ([2, 2, 2], [[4, 2]], "RESHAPE", {"shape": [2, 2, 2]}, 0),
    "Matmul": ([2, 2], [[3, 2], [2, 3]], "CONTRACT", {"rank_pairs": [(0, 1)]}, 1),
    "Conv": ([2], [[3, 2], [2, 2]], "CONV", {"ranks": [0, 1]}, 1),
    "Slice": ([2], [[3, 2]], "SLICE", {"dimension_pairs": [(1, 2), (0, 1)]}, 0),
    "Pad": ([4, 3], [[3, 2]], "PAD", {"dimension_pairs": [(1, 1), (0, 1)]}, 0),
    "Concat": ([3, 4], [[3, 2], [3, 2]], "CONCAT", {"rank": 1}, 1),
    "Stride": ([1, 2], [[3, 2]], "STRIDE", {"dimensions": [2, 1]}, 0),
    "Scatter": ([3, 4], [[3, 2]], "SCATTER", {"shape": [3, 4], "dimensions": [1, 2]}, 0),
    "Reverse": ([3, 2], [[3, 2]], "REVERSE", {"rank_set": {1}}, 0),
}


@pytest.mark.parametrize("case", sorted(STRUCTURAL))
def test_structural_rules(case):
    sshape, ashapes, opname, attrs, idx = STRUCTURAL[case]
    sup = var(sshape, "super")
    args = [var(s, "arg%d" % (i + 1)) for i, s in enumerate(ashapes)]
    op = F(opname, args, attrs)
    assert render(tc.egen.lderive(op, sup, idx)) == GOLD[case]["graphs"][0], GOLD[case]["cite"]


def test_select_and_zero_gradients():
    sup, a1, a2 = std_operands()
    a3 = var([3, 2], "arg3")
    op = F("SELECT", [a1, a2, a3])
    assert render(tc.egen.lderive(op, sup, 1), op) == GOLD["Select"]["graphs"][0]
    op = F("RAND_UNIF", [a1, a2])  # :657-681 — no gradient flows into the bounds of a random draw
    assert render(tc.egen.lderive(op, sup, 1), op) == GOLD["Zeros"]["graphs"][0]


def test_const_helpers_and_add():
    t1, t2 = var([1, 2, 3], "t1", np.float32), var([3, 2, 4], "t2", np.float32)
    assert [render(tc.egen.const_zero(t1)), render(tc.egen.const_one(t2))] == GOLD["ZeroOnes"]["graphs"]
    sup, a1, a2 = std_operands()
    assert render(tc.egen.grad_add([a1, a2, var([3, 2], "arg3")])) == GOLD["AddHelper"]["graphs"][0]


def test_assign_has_no_derivative():  # BACKPROP.Fatals :684-721
    sup, a1, a2 = std_operands()
    with pytest.raises(Exception, match="cannot derive"):
        tc.egen.lderive(F("ASSIGN", [a1, a2]), sup, 1)
