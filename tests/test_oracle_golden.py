"""Pins the CPU oracle against the reference's own golden vectors (CPU only)."""
import json
import os

import numpy as np
import pytest

from oracle import tcr_oracle as orc
from tests import opcheck

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "operator_goldens.json")) as f:
    GOLD = json.load(f)["cases"]


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_oracle_matches_reference_golden(case):
    arrs, shapes = opcheck.case_arrays(case)
    out, oshape = opcheck.oracle_run(case["op"], arrs, shapes, case["attrs"], opcheck.NP.get(case["out_dtype"]))
    expect = opcheck.expected_of(case)
    assert orc.n_elems(oshape) == orc.n_elems(case["out_shape"])
    assert [d for d in oshape if d != 1] == [d for d in case["out_shape"] if d != 1]
    # reference asserts EXPECT_DOUBLE_EQ (4 ulp) / EXPECT_VECEQ on these
    np.testing.assert_allclose(np.asarray(out, dtype=np.float64), expect, rtol=1e-14, atol=0)
    if case["out_dtype"] != case["dtype"]:
        assert np.asarray(out).dtype == opcheck.NP[case["out_dtype"]]
