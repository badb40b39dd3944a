"""SERIALIZE.SaveGraph / LoadGraph (tenncor/test/test_serialize.cpp:72-198) against the reference's own fixtures
models/test/eteq.onnx + eteq.txt (committed in tests/golden/reference_models.json by make_model_goldens.py).

The fixture is the serialized form of four derivatives (dw0, db0, dw1, db1) of a hand-written sigmoid MLP: 90 functors produced by
tcr::derive. Three things are pinned, all on CPU:
  * our loader reads the reference's file into the graph the reference's PrettyEquation prints (eteq.txt, 575 lines);
  * our DerivativeFuncs + teq::derive, fed the same model through our API, build that same graph node for node;
  * our serializer writes it as the same ModelProto, message for message (generated node ids renamed in order of appearance —
    the reference test injects a counting id generator — and graph outputs compared by name: their order in the file is the
    iteration order of an unordered set unless the reference is built with ORDERED_SAVE, internal/onnx/save.hpp:293-300)."""
import base64
import json
import os

import numpy as np
import pytest

import tenncor_b200 as tc

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_models.json")))["test_models"]["eteq"]
NAMES = ("dw0", "db0", "dw1", "db1")


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


@pytest.fixture
def eteq_file(tmp_path):
    path = str(tmp_path / "eteq.onnx")
    data = base64.b64decode(GOLDEN["base64"])
    assert len(data) == GOLDEN["bytes"]
    with open(path, "wb") as f:
        f.write(data)
    return path


def pretty(root):
    """dbg/print/teq.hpp PrettyEquation without types and shapes, as eteq.txt was written"""
    out = []

    def text(t):
        return "(%s:%s)" % (t.usage(), t.label()) if t.is_leaf() else "(%s)" % t.opname()

    def rec(t, last):
        out.append(("".join("____" if flag else "|___" for flag in last[:-1]) + "`--" if last else "") + text(t))
        if not t.is_leaf():
            kids = t.args()
            for i, k in enumerate(kids):
                rec(k, last + [i == len(kids) - 1])

    rec(root, [])
    return out


def stripped(lines):  # fmts::strip(line, {' ', '\t', '\n', default_indent}) and drop empty lines (test_serialize.cpp:165-171)
    return [s for s in (line.strip(" \t\n_") for line in lines) if s]


WANT = stripped(GOLDEN["txt"].split("\n"))


def mock_model():
    """mock_model (test_serialize.cpp:30-69)"""
    def zeros(teq_shape, label):
        return tc.variable(np.zeros(teq_shape[::-1], dtype=np.float64), label)
    inp, w0, b0 = zeros([10, 3], "in"), zeros([9, 10], "weight0"), zeros([9], "bias0")
    w1, b1, out = zeros([5, 9], "weight1"), zeros([5], "bias1"), zeros([5, 3], "out")
    api = tc.api
    layer0 = api.matmul(inp, w0) + api.extend(b0, 1, [3])
    sig0 = 1. / (1. + api.exp(-layer0))
    layer1 = api.matmul(sig0, w1) + api.extend(b1, 1, [3])
    sig1 = 1. / (1. + api.exp(-layer1))
    err = api.pow(out - sig1, 2.)
    return [tc.derive(err, [v])[0] for v in (w0, b0, w1, b1)]


def test_load_graph(eteq_file):  # SERIALIZE.LoadGraph :132-198
    roots, ids = tc.load_model_ids(eteq_file)
    assert len(roots) == 4 and all(name in ids for name in NAMES)
    got = stripped(line for name in NAMES for line in pretty(ids[name]))
    assert len(got) == len(WANT) == 575
    assert got == WANT


def test_derivative_graph_matches_the_reference_file():
    got = stripped(line for root in mock_model() for line in pretty(root))
    assert got == WANT


def canonical(model):
    rename = {}

    def r(name):
        return name if name in NAMES else rename.setdefault(name, "#%d" % len(rename))

    def tensor(t):
        return dict(t, name=r(t["name"])) if t["name"] else t

    def attr(a):
        return dict(a, t=tensor(a["t"]), tensors=[tensor(t) for t in a["tensors"]])

    g = model["graph"]
    return {
        "model": {k: model[k] for k in ("ir_version", "model_version", "producer_name", "producer_version", "domain")},
        "name": g["name"],
        "node": [dict(n, input=[r(x) for x in n["input"]], output=[r(x) for x in n["output"]], name=r(n["name"]),
                      attribute=[attr(a) for a in n["attribute"]]) for n in g["node"]],
        "initializer": [tensor(t) for t in g["initializer"]],
        "input": [dict(v, name=r(v["name"])) for v in g["input"]],
        "output": sorted((dict(v, name=r(v["name"])) for v in g["output"]), key=lambda v: v["name"]),
        "annotation": [dict(q, tensor_name=r(q["tensor_name"])) for q in g["quantization_annotation"]],
    }


def test_save_graph(eteq_file, tmp_path):  # SERIALIZE.SaveGraph :72-129
    ders = mock_model()
    mine = str(tmp_path / "got_eteq.onnx")
    assert tc.save_to_file(mine, ders, dict(zip(NAMES, ders)))
    want = canonical(tc.onnx_describe(open(eteq_file, "rb").read()))
    got = canonical(tc.onnx_describe(open(mine, "rb").read()))
    assert len(want["node"]) == 90 and len(want["initializer"]) == 19
    for section in want:
        assert got[section] == want[section], section
    # and what we wrote loads back into the same graph
    _, ids = tc.load_model_ids(mine)
    assert stripped(line for name in NAMES for line in pretty(ids[name])) == WANT


# ---------------------------------------------------------------- tenncor/serial/test/test_serialize.cpp: mixed types, placeholder, constants
SERIAL = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_models.json")))["test_models"]["serial"]


def serial_roots():
    """the two sub-trees of SERIALIZE.SaveGraph :29-93; make_functor inserts the CASTs the mixed leaf types need"""
    F = tc.egen.make_functor
    zeros = lambda dtype, label: tc.variable(np.zeros((7, 3), dtype=dtype), label)  # noqa: E731  teq::Shape({3, 7})
    osrc, osrc2 = zeros(np.float32, "osrc"), zeros(np.float64, "osrc2")
    src = zeros(np.int32, "src")
    src2 = tc.scalar_constant(23, [7, 3], "DOUBLE")
    root1 = F("SUB", [src2, F("POW", [F("DIV", [F("NEG", [osrc]), F("ADD", [F("SIN", [src]), src])]), osrc2])])
    s2src = tc.placeholder(np.zeros((7, 3), dtype=np.float32), "s2src")
    s2src2, s2src3 = zeros(np.int32, "s2src2"), zeros(np.float64, "s2src3")
    root2 = F("SUB", [s2src, F("MUL", [F("ABS", [s2src]), F("EXP", [s2src2]), F("NEG", [s2src3])])])
    return root1, root2


def typed(root):
    from tests.test_backprop_golden import render_typed
    return render_typed(root).split("\n")


def test_serial_load_graph(tmp_path):  # SERIALIZE.LoadGraph, tenncor/serial/test/test_serialize.cpp:126-180
    path = str(tmp_path / "serial.onnx")
    with open(path, "wb") as f:
        f.write(base64.b64decode(SERIAL["base64"]))
    roots, ids = tc.load_model_ids(path)
    assert len(roots) == 2 and "root1" in ids and "root2" in ids
    got = stripped(line for name in ("root1", "root2") for line in typed(ids[name]))
    assert got == stripped(SERIAL["txt"].split("\n"))
    assert got == stripped(line for root in serial_roots() for line in typed(root))  # and the graph built here is that graph


def test_serial_save_graph(tmp_path):  # SERIALIZE.SaveGraph :29-123
    root1, root2 = serial_roots()
    mine = str(tmp_path / "got_serial.onnx")
    assert tc.save_to_file(mine, [root1, root2], {"root1": root1, "root2": root2})
    global NAMES
    keep, NAMES = NAMES, ("root1", "root2")
    try:
        want = canonical(tc.onnx_describe(base64.b64decode(SERIAL["base64"])))
        got = canonical(tc.onnx_describe(open(mine, "rb").read()))
    finally:
        NAMES = keep
    for section in ("name", "node", "initializer", "input", "output", "annotation"):  # serial::save_graph fills the graph only, no model header
        assert got[section] == want[section], section
