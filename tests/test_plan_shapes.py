"""The launch plans of the two BASELINE training steps, pinned on the host (tc.describe_plan lowers without a device): a change to a
fusion pass that silently un-fuses the C3 / C4 step shows up here, on CPU, before it shows up as milliseconds on the GPU.
What the reference runs instead: one Eigen assignment per functor (internal/eigen/device.hpp:555-570) — 68 functors for C3, 12 187 for C4."""
import collections

import tenncor_b200 as tc
from tenncor_b200 import configs


def _kind(line):
    return line.split(" ")[0]


def test_c3_training_step_is_seventeen_launches():
    cfg = configs.mlp(784, 1024, 10, 8192, pixels=True, name="C3")
    plan = tc.describe_plan([cfg.train])
    assert len(plan) == 17, plan
    kinds = collections.Counter(_kind(s) for s in plan)
    # two forward passes (apply_update returns the error AFTER the update, tenncor/trainer/apply_update.hpp:13-41) = 4 fused
    # GEMM + bias + activation launches; the hidden layer's gradient with the SIGMOID' factor in its epilogue; both weight gradients
    # as transposed products; one in-place update per variable; the loss as one map-reduce launch; the pixel cast
    assert kinds["GEMM+bias+act"] == 4 and kinds["GEMM*dsigmoid"] == 1 and kinds["GEMM^T"] == 2
    assert kinds["ASSIGN_SUB"] == 4 and kinds["REDUCE_SUM"] == 2 and kinds["SUM/c"] == 1
    assert plan[0].startswith("MUL fused(3 instr, 1 in) [784\\8192")  # uint8 pixels -> float, scaled
    assert not any(_kind(s) in ("EXTEND", "PERMUTE", "CONTRACT", "MATMUL", "SIGMOID", "CAST") for s in plan), plan


def test_c4_training_step_is_one_product_and_one_cell_launch_per_time_step():
    cfg = configs.recurrent("lstm", vocab=128, hidden=1024, seq=128, batch=64, learning_rate=0.001)
    plan = tc.describe_plan([cfg.train])
    kinds = collections.Counter(_kind(s) for s in plan)
    assert len(plan) <= 580, len(plan)
    # forward twice (before and after the update): 255 gate launches with the cell update in the epilogue over the two passes;
    # backward: 127 K-segmented gradient products and 127 cell steps (126 over four gates, one at the end of the chain)
    assert kinds["GEMM-GROUP+cell"] == 255 and kinds["GEMM-SUM"] == 127
    assert sum(v for k, v in kinds.items() if k.startswith("CELL-BACKWARD")) == 127
    assert kinds["ADD-STACK(128)"] == 4          # the four stacked bias-gradient reductions
    assert kinds["GEMM^T"] == 5                  # four gate weight gradients with K = seq x batch, one for the output layer
    assert kinds["CONCAT"] + kinds.get("CONCAT-BATCH(127)", 0) + kinds.get("CONCAT-BATCH(128)", 0) <= 2
    assert not any(_kind(s) in ("PERMUTE", "SLICE", "EXTEND") for s in plan)


def test_gru_training_step_groups_its_gate_products():
    """cfg/tenncor/layer.yml:770-813: update and reset gates share their operands (one grouped launch per time step and pass), the
    candidate reads the reset hidden state and gets its own; the per-step CONCATs leave the chain as batched copies."""
    cfg = configs.recurrent("gru", vocab=128, hidden=1024, seq=128, batch=64, learning_rate=0.001)
    plan = tc.describe_plan([cfg.train])
    kinds = collections.Counter(_kind(s) for s in plan)
    assert len(plan) <= 1750, len(plan)
    assert kinds["GEMM-GROUP"] == 512 and kinds["GEMM-SUM"] == 127 and kinds["ADD-STACK(128)"] == 3
    assert sum(v for k, v in kinds.items() if k.startswith("CONCAT-BATCH")) == 2 and kinds["CONCAT"] == 0
    assert kinds["PERMUTE"] == 0 and kinds["EXTEND"] == 0
