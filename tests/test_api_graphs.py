"""Graph-shape goldens of tenncor/test/test_api.cpp that need no evaluation: what `assign` does with a higher-precision source
(API.AssignHighToLowPrecision :647-735), that chained casts are not doubled (API.Cast :746-772), and that ASSIGN has no derivative
(API.Assign :636-640). Rendering and comparison follow the reference's PrettyEquation / tutil::compare_graph."""
import numpy as np
import pytest

import tenncor_b200 as tc
from tests.test_backprop_golden import render_typed, same_graph

DATA = np.array([59, 10, 28, 10, 67, 62, 23, 4, 55, 77, 28, 16, 82, 52, 47, 16, 7, 85, 37, 2, 8, 52, 62, 43], dtype=np.float64).reshape(4, 3, 2)
DATA2 = np.array([22, 15, 74, 38, 61, 95, 62, 81, 99, 76, 7, 22, 56, 50, 19, 13, 12, 10, 31, 40, 60, 54, 6, 83], dtype=np.float64).reshape(4, 3, 2)


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def test_assign_high_to_low_precision():
    target1 = tc.variable(DATA.astype(np.float32), "target1")
    target2 = tc.variable(DATA.astype(np.float32), "target2")
    src = tc.constant(DATA2)
    ass1 = tc.api.assign(target1, src)
    ass2 = tc.api.assign(target2, -src)
    assert same_graph(render_typed(ass1),
                      "(ASSIGN<FLOAT>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_`--(variable:target1<FLOAT>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_`--(CAST<FLOAT>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_____`--(constant:[22\\15\\74\\38\\61\\...]<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"), render_typed(ass1)
    assert same_graph(render_typed(ass2),
                      "(ASSIGN<FLOAT>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_`--(variable:target2<FLOAT>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_`--(CAST<FLOAT>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_____`--(NEG<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_________`--(constant:[22\\15\\74\\38\\61\\...]<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"), render_typed(ass2)
    # assignments take the type of the target even if the source has higher precision
    assert ass1.type_label() == ass2.type_label() == "FLOAT"
    for ass in (ass1, ass2):
        with pytest.raises(Exception, match="cannot derive ASSIGN"):
            tc.derive(ass, [src])


def test_cast_is_not_doubled():
    a = tc.constant(DATA)
    out = tc.api.cast(tc.api.cast(tc.api.cos(a) * 2.1 + 4.5, "INT32") / 2, "DOUBLE")
    want = ("(CAST<DOUBLE>)\n"
            "_`--(DIV<INT32>)\n"
            "_____`--(CAST<INT32>)\n"
            "_____|___`--(ADD<DOUBLE>)\n"
            "_____|_______`--(MUL<DOUBLE>)\n"
            "_____|_______|___`--(COS<DOUBLE>)\n"
            "_____|_______|___|___`--(constant:[59\\10\\28\\10\\67\\...]<DOUBLE>)\n"
            "_____|_______|___`--(EXTEND<DOUBLE>)\n"
            "_____|_______|_______`--(constant:2.1<DOUBLE>)\n"
            "_____|_______`--(EXTEND<DOUBLE>)\n"
            "_____|___________`--(constant:4.5<DOUBLE>)\n"
            "_____`--(EXTEND<INT32>)\n"
            "_________`--(constant:2<INT32>)\n")
    got = "\n".join(line[:line.rindex("[")] + ")" for line in render_typed(out).split("\n"))  # EXPECT_GRAPH_STRUCTEQ prints no shapes
    assert same_graph(got, want), got


def test_nn_dropout_graph():  # NN.Dropout (tenncor/test/test_nn.cpp:15-49): the graph, verbatim; the values are a GPU matter
    x = tc.scalar_constant(1, [5, 2], "FLOAT")
    f = tc.api.nn.dropout(x, 0.1)
    mask = ("(LT<FLOAT>[2\\5\\1\\1\\1\\1\\1\\1])\n"
            "`--(RAND_UNIF<FLOAT>[2\\5\\1\\1\\1\\1\\1\\1])\n"
            "|___`--(variable:0<FLOAT>[2\\5\\1\\1\\1\\1\\1\\1])\n"
            "|___`--(variable:1<FLOAT>[2\\5\\1\\1\\1\\1\\1\\1])\n"
            "`--(EXTEND<FLOAT>[2\\5\\1\\1\\1\\1\\1\\1])\n"
            "____`--(SUB<FLOAT>[1\\1\\1\\1\\1\\1\\1\\1])\n"
            "________`--(EXTEND<FLOAT>[1\\1\\1\\1\\1\\1\\1\\1])\n"
            "________|___`--(constant:1<FLOAT>[1\\1\\1\\1\\1\\1\\1\\1])\n"
            "________`--(variable:drop_rate<FLOAT>[1\\1\\1\\1\\1\\1\\1\\1])\n")
    indent = lambda text, pre: "".join(pre + line + "\n" for line in text.rstrip("\n").split("\n"))  # noqa: E731
    want = ("(MUL<FLOAT>[2\\5\\1\\1\\1\\1\\1\\1])\n"
            "_`--(constant:1<FLOAT>[2\\5\\1\\1\\1\\1\\1\\1])\n"
            "_`--(DIV<FLOAT>[2\\5\\1\\1\\1\\1\\1\\1])\n"
            + "_____`--" + indent(mask, "_____|___")[len("_____|___"):]
            + "_____`--(EXTEND<FLOAT>[2\\5\\1\\1\\1\\1\\1\\1])\n"
            "_________`--(DIV<FLOAT>[1\\1\\1\\1\\1\\1\\1\\1])\n"
            "_____________`--(REDUCE_SUM<FLOAT>[1\\1\\1\\1\\1\\1\\1\\1])\n"
            + "_____________|___`--" + indent(mask, "_____________|_______")[len("_____________|_______"):]
            + "_____________`--(constant:10<FLOAT>[1\\1\\1\\1\\1\\1\\1\\1])")
    assert same_graph(render_typed(f), want), render_typed(f)
    # the mask is ONE node read twice (one draw per evaluation), not two independent draws
    div = f.args()[1]
    assert div.args()[0] == div.args()[1].args()[0].args()[0].args()[0]
    assert f.teq_shape() == [2, 5, 1, 1, 1, 1, 1, 1]


# ---------------------------------------------------------------- initialisers (tenncor/test/test_init.cpp)
def test_init_zero():  # INIT.Zero :12-28
    z = tc.variable_from_init(tc.api.init.zeros(), [9, 18], "abc")
    assert z.teq_shape()[:2] == [18, 9] and str(z) == "abc"
    assert not z.data().any()


def test_init_variance_scaling():  # INIT.VarianceScaling :31-75 — truncated at two standard deviations
    factor = 0.425
    v1 = tc.variable_from_init(tc.api.init.variance_scaling(factor), [3, 9, 18], "def")
    v2 = tc.variable_from_init(tc.api.init.variance_scaling(factor, shape_factor=lambda shape: float(shape[0])), [3, 9, 18], "def")
    assert v1.teq_shape()[:3] == [18, 9, 3] and str(v1) == "def"
    bound1 = 2 * np.sqrt(factor / ((18 + 9) / 2))            # default fan: mean of the two fastest dimensions
    assert np.abs(v1.data()).max() < bound1 and v1.data().std() > bound1 / 8
    bound2 = 2 * np.sqrt(factor / 3)                         # shape.at(2): numpy lists it first
    assert np.abs(v2.data()).max() < bound2 and v2.data().std() > bound2 / 8


def test_init_xavier_uniform():  # INIT.UniformXavier :78-96
    factor = 0.712
    x = tc.variable_from_init(tc.api.init.xavier_uniform(factor), [3, 9, 18], "ghi")
    bound = factor * np.sqrt(6.0 / (18 + 9))
    assert x.teq_shape()[:3] == [18, 9, 3] and str(x) == "ghi"
    assert np.abs(x.data()).max() < bound and np.abs(x.data()).max() > 0.8 * bound


# ---------------------------------------------------------------- tenncor/test/test_layer.py
def test_conv_conv_gru_link_shape():  # LAYRTest.test_gru :63-82 — two zero-padded conv2d layers feeding a gru split along rank 2
    brain = tc.api.layer.link([
        tc.api.layer.conv2d([5, 5], 1, 16, kernel_init=tc.api.init.xavier_normal(0.5), zero_padding=((2, 2), (2, 2))),
        tc.api.layer.conv2d([5, 5], 16, 20, kernel_init=tc.api.init.xavier_normal(0.5), zero_padding=((2, 2), (2, 2))),
        tc.api.layer.gru([2647, 20], 20, 128, seq_dim=2, kernel_init=tc.api.init.zeros(), bias_init=tc.api.init.zeros()),
    ], tc.EVariable([128, 2647, 1], label="input"))
    assert list(brain.shape()) == [128, 2647, 20]


def test_conv2d_on_an_image():  # layer.yml:159-251 — conv2d(input, out_ncol, kernel_hw, ..., padding | zero_padding)
    x = tc.EVariable([2, 9, 10, 4], label="x")                   # teq [4, 10, 9, 2], the image of API.Conv (test_api.cpp:2505-2531)
    y = tc.api.layer.conv2d(x, 3, (6, 5))
    assert y.teq_shape() == [3, 6, 4, 2, 1, 1, 1, 1]
    kernel, bias = sorted(y.get_storage(), key=lambda v: -len(v.shape()))
    assert kernel.teq_shape()[:4] == [3, 4, 5, 6] and bias.teq_shape()[0] == 3 and str(kernel) == "weight" and str(bias) == "bias"
    assert tc.api.layer.conv2d(x, 3, (5, 5), padding="SAME").teq_shape() == [3, 10, 9, 2, 1, 1, 1, 1]
    assert tc.api.layer.conv2d(x, 3, (5, 5), padding=((2, 2), (0, 0))).teq_shape() == [3, 10, 5, 2, 1, 1, 1, 1]
    assert len(tc.api.layer.conv2d(x, 3, (5, 5), with_bias=False).get_storage()) == 1
    with pytest.raises(Exception, match="unsupported padding type full"):
        tc.api.layer.conv2d(x, 3, (5, 5), padding="full")


def test_training_graph_survives_save_and_load(tmp_path):  # LAYRTest.test_context_save :26-61 — the whole apply_update graph, by its rendering
    nunits, ninput, noutput, nbatch = 9, 10, 5, 10
    train_err = tc.apply_update(
        [tc.api.layer.link([
            tc.api.layer.dense([ninput], [nunits], kernel_init=tc.api.init.xavier_uniform(), bias_init=tc.api.init.zeros()),
            tc.api.layer.bind(tc.api.sigmoid),
            tc.api.layer.dense([nunits], [noutput], kernel_init=tc.api.init.xavier_uniform(), bias_init=tc.api.init.zeros()),
            tc.api.layer.bind(tc.api.sigmoid)])],
        lambda err, leaves: tc.api.approx.sgd(err, leaves, learning_rate=0.9),
        lambda models: tc.api.loss.mean_squared(tc.EVariable([nbatch, noutput]), models[0].connect(tc.EVariable([nbatch, ninput]))))
    path = str(tmp_path / "layr_test.onnx")
    assert tc.save_to_file(path, [train_err])
    roots = tc.load_from_file(path)
    assert len(roots) == 1
    assert render_typed(roots[0]) == render_typed(train_err)
    assert "ASSIGN_SUB" in render_typed(train_err)


def test_python_module_shims():  # names the reference's scripts use beyond the API proper (tenncor/python/eteq_ext.cpp)
    assert tc.Shape([2647, 20]) == [2647, 20] and repr(tc.Shape([3])) == "Shape([3])"
    assert tc.TenncorAPI(tc.global_context) is tc.api and isinstance(tc.global_context, tc.Context)
    assert tc.optimize("cfg/optimizations.json") is None        # the rule-file form: accepted, nothing to rewrite in place
    a = tc.variable(np.ones((2, 3)), "a")
    twice = tc.api.sin(a) * tc.api.sin(a)                        # two structurally equal functors
    (merged,), stats = tc.optimize([twice], fold_constants=False)
    assert stats["merged"] == 1 and merged.args()[0] == merged.args()[1]
    import tenncor_b200.compat as compat
    import sys
    keep = {name: sys.modules.get(name) for name in ("tenncor", "extenncor", "dbg")}
    try:
        names = compat.install()
        assert {"tenncor", "extenncor.dqn_trainer", "dbg.compare", "dbg.print"} <= set(names)
        import tenncor
        assert tenncor is tc
    finally:
        for name in list(sys.modules):
            if name.split(".")[0] in keep and keep[name.split(".")[0]] is None:
                del sys.modules[name]
