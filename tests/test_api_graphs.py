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
