"""Lowering regressions of the planned evaluator that need no device (`tc.describe_plan`).

The reference evaluates node by node (internal/teq/evaluator.hpp:34-43), so every graph that is
valid there must lower to a plan in which each step's operands are produced by an earlier step."""
import numpy as np

import tenncor_b200 as tc


def _names(steps):
    return [s.split(" ")[0] for s in steps]


def test_extend_feeding_a_fused_gemm_is_materialised():
    # derive(reduce_sum(contract(x, w)), [w]): the upstream gradient of the product is EXTEND(1) and the
    # weight gradient PERMUTE(CONTRACT(sup, x)) is fused to one GEMM^T step that reads that EXTEND
    x = tc.EVariable([5, 4], 1.0, "x")
    w = tc.EVariable([4, 3], 1.0, "w")
    y = tc.api.reduce_sum(tc.api.contract(x, w, [(0, 1)]))
    g = tc.derive(y, [w])[0]
    steps = tc.describe_plan([g])
    assert any(s.startswith("GEMM^T") for s in steps), steps
    assert "EXTEND" in _names(steps), steps
    assert _names(steps).index("EXTEND") < [i for i, s in enumerate(steps) if s.startswith("GEMM^T")][0]


def test_extend_operand_of_a_dense_layer_is_materialised():
    # matmul(extend(v), w) + extend(b): the product's operand is an EXTEND that only the fused GEMM+bias step reads
    v = tc.EVariable([1, 4], 1.0, "v")
    w = tc.EVariable([4, 3], 1.0, "w")
    b = tc.EVariable([3], 0.5, "b")
    xe = tc.api.extend(v, 1, [5])
    out = tc.api.matmul(xe, w) + tc.api.extend(b, 1, [5])
    steps = tc.describe_plan([out])
    assert any(s.startswith("GEMM+bias") for s in steps), steps
    assert "EXTEND" in _names(steps), steps
