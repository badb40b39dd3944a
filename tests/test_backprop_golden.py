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


# ---------------------------------------------------------------- optimizer update graphs (tenncor/test/test_approx.cpp)
def render_typed(root):
    """like render(), with each leaf's real usage (variable / constant), as the reference prints real leaves"""
    out = []

    def text(t):
        shape = "\\".join(str(d) for d in t.teq_shape())
        dtype = {"float64": "DOUBLE", "float32": "FLOAT", "int32": "INT32"}[str(t.dtype())]
        name = (t.usage() + ":" + t.label()) if t.is_leaf() else t.opname()
        return "(%s<%s>[%s])" % (name, dtype, shape)

    def rec(t, ancestors_last):
        if not ancestors_last:
            out.append(text(t))
        else:
            out.append("_" + "".join("____" if last else "|___" for last in ancestors_last[:-1]) + "`--" + text(t))
        if not t.is_leaf():
            kids = t.args()
            for i, k in enumerate(kids):
                rec(k, ancestors_last + [i == len(kids) - 1])

    rec(root, [])
    return "\n".join(out)


def _leaf(shape):
    return tc.EVariable(shape[::-1], 0, "leaf")  # FLOAT by default, like make_variable_scalar<float>


def _golden(name):
    return GOLD[name]["graphs"][0].rstrip("\n")


def same_graph(got, want):
    """tutil::compare_graph (testutil/src/graph_comp.cpp:14-56): lines are compared after stripping '_', blanks and newlines from
    both ends — the reference's hand-written goldens are not exact about leading indentation (Adadelta's has a line one short)"""
    strip = lambda text: [l.strip("_ \t") for l in text.split("\n") if l.strip("_ \t")]  # noqa: E731
    return strip(got) == strip(want)


def test_sgd_graph():  # APPROX.StochasticGD :21-45
    leaf = _leaf([18, 9, 3])
    groups = tc.api.approx.sgd(tc.api.abs(leaf), [leaf], learning_rate=0.67)
    assert len(groups) == 1
    assert render_typed(groups[0][1]) == _golden("ApproxStochasticGD")


def test_adagrad_graph():  # APPROX.Adagrad :48-86
    leaf = _leaf([18, 9, 3])
    groups = tc.api.approx.adagrad(tc.api.abs(leaf), [leaf], learning_rate=0.67)
    assert len(groups) == 1
    assert render_typed(groups[0][1]) == _golden("ApproxAdagrad")


def test_adadelta_graph():  # APPROX.Adadelta :89-195 (the golden is a format string: %s = fmts::to_string(float epsilon))
    leaf = _leaf([18, 9, 3])
    groups = tc.api.approx.adadelta(tc.api.sin(leaf), [leaf], 1.0, 0.91, 0.16)
    assert len(groups) == 1
    assert same_graph(render_typed(groups[0][1]), _golden("ApproxAdadelta").replace("%s", "1.19209e-07"))


def test_rms_momentum_graph():  # APPROX.RmsMomentum :289-341
    leaf = _leaf([5])
    groups = tc.api.approx.rms_momentum(tc.api.sin(leaf) / 2.0, [leaf], 1.0, 0.52, float(np.finfo(np.float32).eps))
    assert len(groups) == 1
    assert render_typed(groups[0][1]) == _golden("ApproxRmsMomentum")


def test_adam_graph():  # APPROX.Adam :388-482
    x = tc.EVariable([], 0, "x")
    loss = tc.api.square(x) - 2.0 * x + 1.0
    step = tc.api.approx.adam(loss, [x], 0.01, 0.9, 0.999, 1e-8)[0][1]
    assert render_typed(step) == _golden("ApproxAdam")


# ---------------------------------------------------------------- layer connection graphs (tenncor/test/test_layer.cpp:36-262)
def _x(teq_shape, label):
    return tc.EVariable(teq_shape[::-1], 0, label)


def _graphs(name):
    return [g.rstrip("\n").replace("%s", "1.19209e-07") for g in GOLD[name]["graphs"]]


def test_dense_connection():
    xu = tc.api.init.xavier_uniform
    biased = tc.api.layer.dense([6], [5], xu(2), xu(4))
    plain = tc.api.layer.dense([7], [6], xu(3), None, with_bias=False)
    want = _graphs("LayerDenseConnection")
    assert same_graph(render_typed(biased.connect(_x([6, 2], "x"))), want[0])
    assert same_graph(render_typed(plain.connect(_x([7, 2], "x2"))), want[1])


def test_conv_connection():
    conv = tc.api.layer.conv2d((6, 5), 4, 3, tc.api.init.xavier_uniform(1), tc.api.init.zeros())
    assert same_graph(render_typed(conv.connect(_x([4, 10, 9, 2], "x"))), _graphs("LayerConvConnection")[0])


def test_rbm_connections():
    xu = tc.api.init.xavier_uniform
    rbm, nobias = tc.api.layer.rbm(6, 5, xu(2), xu(4)), tc.api.layer.rbm(7, 6, xu(3), None, with_bias=False)
    fwd = _graphs("LayerRbmConnection")
    assert same_graph(render_typed(rbm.connect(_x([6, 2], "x"))), fwd[0])
    assert same_graph(render_typed(nobias.connect(_x([7, 2], "x2"))), fwd[1])
    bwd = _graphs("LayerRbmBackwardConnection")
    assert same_graph(render_typed(rbm.backward_connect(_x([5, 2], "y"))), bwd[0])
    assert same_graph(render_typed(nobias.backward_connect(_x([6, 2], "y2"))), bwd[1])


def test_bind_layers():
    x = _x([6, 2], "x")
    assert same_graph(render_typed(tc.api.layer.bind(tc.api.sigmoid).connect(x)), _graphs("LayerBindSigmoid")[0])
    soft = _graphs("LayerBindSoftmax")
    assert same_graph(render_typed(tc.api.layer.bind(lambda e: tc.api.softmax(e, 0, 1)).connect(x)), soft[0])
    assert same_graph(render_typed(tc.api.layer.bind(lambda e: tc.api.softmax(e, 1, 1)).connect(x)), soft[1])
