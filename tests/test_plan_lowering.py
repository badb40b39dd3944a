"""What the planner lowers the demo training graphs to (host logic, no device: `tc.describe_plan`). These counts are the
launch budget of a training step: a regression here is a performance regression on the GPU."""
import collections
import re

import pytest

import tenncor_b200 as tc
from tenncor_b200 import configs


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def kinds(steps):
    return collections.Counter(re.match(r"[A-Z_0-9a-z^+\-]+(?: im2col\+GEMM| GEMM\+col2im)?(?:\+bias)?(?:\+act)?", s).group(0) for s in steps)


def test_mlp_step_is_nineteen_launches():
    """~45 functor evaluations in the reference; two dense layers forward (bias + sigmoid in the GEMM
    epilogue), loss gradient, three gradient products (weight gradients written transposed, no PERMUTE pass), two bias
    reductions, four SGD updates, the post-update forward and the loss"""
    steps = tc.describe_plan([configs.mlp(784, 1024, 10, 8192).train])
    k = kinds(steps)
    assert len(steps) == 19, steps
    assert k["GEMM+bias+act"] == 4 and k["GEMM^T"] == 2 and k["CONTRACT"] == 1 and k["ASSIGN_SUB"] == 4 and k["REDUCE_SUM"] == 3
    assert not any(s.startswith(("PERMUTE", "EXTEND", "IDENTITY", "SIGMOID")) for s in steps)


def test_lstm_step_composition():
    """seq 8 instead of 128: the counts are per time step. Per step and pass: ONE launch per gate (bias + activation fused), the
    weight gradients as transposed-output GEMMs, no materialised EXTEND / PERMUTE, slices of the input as views"""
    seq = 8
    steps = tc.describe_plan([configs.recurrent("lstm", vocab=128, hidden=256, seq=seq, batch=64).train])
    k = kinds(steps)
    assert k["GEMM+bias+act"] == 2 * seq * 4          # two forward passes (before / after the update) x 4 gates
    assert k["GEMM^T"] == seq * 4 + 1                  # every gate's weight gradient per step + the dense layer's
    assert k["CONCAT"] == 2 * seq + 1                  # [x_t, h_{t-1}] per step and pass; the final CONCAT of states is n-ary (x2 passes, one fused away)
    assert not any(s.startswith(("PERMUTE", "EXTEND", "IDENTITY", "PAD")) for s in steps), [s for s in steps if s.startswith(("PERMUTE", "EXTEND", "PAD"))]
    assert len(steps) <= 45 * seq, len(steps)           # 320 at seq 8: ~40 launches per time step over three passes


def test_rbm_and_dqn_step_sizes():
    assert len(tc.describe_plan([configs.rbm(784, 64, 4096).train])) == 29
    assert len(tc.describe_plan([configs.dqn(nbatch=4096).train])) == 54
