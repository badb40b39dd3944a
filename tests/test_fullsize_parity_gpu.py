"""BASELINE's configs at their FULL size against the CPU oracle, step for step.

The oracle evaluates the same dumped functor graph node by node in the reference's order (one unfused numpy op per
functor, in-place ASSIGNs); the GPU runs the planned evaluator (fused launches, CUDA graph). The loss of every step and
every variable after the last step are compared, the way tenncor/test/test_equation.cpp:330-470 compares golden dW / db
after a training step. Tolerance 1e-4 relative to the tensor's magnitude (~10 chained fp32 ops + 3xTF32 GEMMs).

  C3  MLP 784-1024-10, batch 8192 (BASELINE config 3, one GPU's shard)      oracle ~0.3 s / step
  C4  LSTM 128-1024 seq 128, batch 64 -> dense -> softmax NLL, adagrad        oracle ~10-25 s / step
"""
import numpy as np
import pytest

import tenncor_b200 as tc
from tenncor_b200 import configs
from tests.test_train_gpu import OracleSession, rel_err

pytestmark = pytest.mark.gpu


def _train_and_compare(cfg, gen, steps, tol_loss, tol_var):
    sess = OracleSession([cfg.train])
    rng = np.random.default_rng(11)
    for step in range(steps):
        batch = gen(rng)
        for feed, arr in zip(cfg.feeds.values(), batch):
            feed.assign(arr)
            sess.assign(feed, arr)
        got = cfg.train.get()
        want = sess.run()[0]
        assert np.isfinite(want).all()
        assert rel_err(got, want) < tol_loss, (step, got, want)
    worst = {}
    for v in cfg.variables:
        worst[str(v)] = rel_err(v.data(), sess.leaf_value(v))
    assert max(worst.values()) < tol_var, worst


@pytest.mark.parametrize("pixels", [False, True])
def test_c3_full_size_three_steps_match_oracle(gpu, pixels):
    """pixels: the bench's form — a UINT8 input variable, CAST + scale on the device (tenncor/eteq/caster.hpp:10-44)."""
    tc.set_evaluator("plan")
    tc.set_matmul_precision("3xtf32")
    cfg = configs.mlp(784, 1024, 10, 8192, name="c3", pixels=pixels)
    _train_and_compare(cfg, lambda rng: configs.mlp_batch(rng, cfg.feeds, one_hot=True), steps=3, tol_loss=1e-4, tol_var=1e-4)


@pytest.mark.parametrize("kind", ["lstm", "gru"])
def test_c4_full_size_two_steps_match_oracle(gpu, kind):
    """adagrad's first update is lr * g / (|g| + eps): for a gradient element near zero the quotient is ill-conditioned, so the
    variables are compared after the accumulated squared gradient has a second term as well (two steps), with the demo's
    eps = 1e-8; learning rate 0.01 keeps the summed NLL finite at this size."""
    tc.set_evaluator("plan")
    tc.set_matmul_precision("3xtf32")
    cfg = configs.recurrent(kind, vocab=128, hidden=1024, seq=128, batch=64, learning_rate=0.01, name="c4")
    _train_and_compare(cfg, lambda rng: configs.recurrent_batch(rng, cfg.feeds, 128), steps=2, tol_loss=1e-4, tol_var=1e-3)
