"""The reference's python debugging helpers (dbg/python/print.cpp, compare.cpp) as tenncor_b200.dbg: tree rendering in the reference's
exact format (checked against its own eteq.txt fixture) and node-by-node graph comparison. Host only."""
import base64
import json
import os

import numpy as np
import pytest

import tenncor_b200 as tc
from tenncor_b200.dbg import compare as cmp
from tenncor_b200.dbg import print as dprint


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def test_graph_to_str_format():
    a = tc.variable(np.ones((3, 2)), "a")
    b = tc.constant(np.full((3, 2), 2.0))
    f = tc.api.sin(a) * (a + b)
    assert dprint.graph_to_str(f) == ("(MUL)\n"
                                      " `--(SIN)\n"
                                      " |   `--(variable:a)\n"
                                      " `--(ADD)\n"
                                      "     `--(variable:a)\n"
                                      "     `--(constant:2)\n")
    assert dprint.graph_to_str(a + b, showshape=True) == ("(ADD[2\\3\\1\\1\\1\\1\\1\\1])\n"
                                                          " `--(variable:a[2\\3\\1\\1\\1\\1\\1\\1])\n"
                                                          " `--(constant:2[2\\3\\1\\1\\1\\1\\1\\1])\n")
    assert "(ADD<DOUBLE>)" in dprint.graph_to_str(a + b, showtype=True)


def test_graph_to_str_reproduces_the_reference_fixture(tmp_path):
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_models.json")))["test_models"]["eteq"]
    path = str(tmp_path / "eteq.onnx")
    with open(path, "wb") as f:
        f.write(base64.b64decode(golden["base64"]))
    _, ids = tc.load_model_ids(path)
    got = "".join(dprint.graph_to_str(ids[name], indent="_") for name in ("dw0", "db0", "dw1", "db1"))
    strip = lambda text: [s for s in (line.strip(" \t\n_") for line in text.split("\n")) if s]  # noqa: E731
    assert strip(got) == strip(golden["txt"])


def test_compare():
    b = tc.constant(np.full((3, 2), 2.0))

    def build(data, op=tc.api.sin):
        v = tc.variable(data, "v")
        return op(v) * (v + (b if data.shape == (3, 2) else tc.constant(np.full(data.shape, 2.0))))

    f, g = build(np.ones((3, 2))), build(np.zeros((3, 2)))
    assert cmp.is_equal(f, f) and cmp.is_equal(f, g)                 # leaves are matched by position, not by content
    assert not cmp.is_equal(f, build(np.ones((3, 2)), tc.api.cos))   # another opcode
    assert not cmp.is_equal(f, build(np.ones((2, 3))))               # another shape
    two = tc.api.sin(tc.variable(np.ones((3, 2)), "v")) * (tc.variable(np.ones((3, 2)), "v") + b)
    assert not cmp.is_equal(f, two)                                  # same picture, different wiring (one leaf read twice vs two leaves)
    assert not cmp.is_equal(tc.api.reduce_sum(f, 0, 1), tc.api.reduce_sum(f, 1, 1))   # attributes count
    # data: only the leaves hold data before an evaluation — the constant matches, the variable decides
    assert cmp.percent_dataeq(f, f) == 1.
    assert cmp.percent_dataeq(f, g) == 0.5 and not cmp.is_dataeq(f, g)
    assert cmp.percent_dataeq(f, build(np.ones((3, 2)))) == 1.
    assert cmp.percent_dataeq(f, two) == 0.
