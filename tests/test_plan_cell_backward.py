"""Host-side lowering checks (no device): the backward pass of an unrolled LSTM lowers to ONE elementwise launch per time step
(merge_elementwise_steps + match_cell_backward in host/planner.cpp) next to the K-segmented gradient product."""
import collections

import tenncor_b200 as tc
from tenncor_b200 import configs


def _plan(kind, seq):
    cfg = configs.recurrent(kind, vocab=32, hidden=64, seq=seq, batch=64, learning_rate=0.01)
    return [line.split(" [")[0] for line in tc.describe_plan([cfg.train])]


def test_lstm_backward_step_is_one_cell_backward_launch():
    plan = _plan("lstm", 24)
    count = collections.Counter(plan)
    cell = [k for k in count if k.startswith("CELL-BACKWARD(4 gates")]
    assert cell and count[cell[0]] >= 24 - 3, count.most_common(8)
    # nothing elementwise is left on the per-step chain besides it
    per_step = [k for k, v in count.items() if v >= 20 and "fused(" in k]
    assert not per_step, per_step


def test_gru_backward_merges_elementwise_chains():
    count = collections.Counter(_plan("gru", 24))
    assert any("multi(" in k and v >= 20 for k, v in count.items()), count.most_common(12)


def test_merging_can_be_switched_off(monkeypatch):
    monkeypatch.setenv("TCR_NO_EW_MERGE", "1")
    plan = _plan("lstm", 8)
    assert not any(k.startswith("CELL-BACKWARD") or "multi(" in k for k in plan)
