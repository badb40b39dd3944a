"""Whole training steps (forward + derivative graph + optimiser ASSIGNs) on the GPU against
the CPU oracle evaluating the SAME dumped graph node by node in the reference's order.

fp32 tolerance: the north star asks 1e-5 relative per elementwise / reduce op; a training
step chains ~10 of them plus GEMMs (3xTF32), so step-level quantities are
compared at 1e-4 relative to the tensor's magnitude after several steps.
"""
import numpy as np
import pytest

import tenncor_b200 as tc
from oracle import tcr_oracle as orc
from tenncor_b200 import configs

pytestmark = pytest.mark.gpu


class OracleSession:
    """Keeps the oracle's own copy of every leaf; ASSIGN* nodes mutate it in place."""

    def __init__(self, roots):
        self.roots = roots
        self.tape = tc.dump_graph(roots)
        self.ids = tc.dump_ids(roots, self.tape)
        for node in self.tape:
            if node["kind"] == "leaf":
                node["data"] = np.array(node["data"], copy=True)

    def _find(self, tensor):
        # post-order ids: visiting the roots first reproduces dump_graph's numbering
        return tc.dump_ids(self.roots + [tensor], None)[tensor]

    def assign(self, tensor, array):
        nid = self._find(tensor)
        node = self.tape[nid]
        assert node["kind"] == "leaf"
        node["data"][...] = np.asarray(array, dtype=node["data"].dtype).reshape(-1)

    def run(self):
        vals = orc.eval_tape(self.tape)
        return [vals[self.ids[r]] for r in self.roots]

    def leaf_value(self, tensor):
        return self.tape[self._find(tensor)]["data"]


def rel_err(got, want):
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    return np.max(np.abs(got - want)) / (np.max(np.abs(want)) + 1e-30)


@pytest.mark.parametrize("evaluator", ["node", "plan"])
@pytest.mark.parametrize("dims", [(10, 9, 5, 3), (64, 48, 16, 33), (200, 300, 10, 257)])
def test_mlp_training_matches_oracle(gpu, dims, evaluator):
    tc.set_evaluator(evaluator)
    tc.set_matmul_precision("3xtf32")
    try:
        cfg = configs.mlp(*dims)
        sess = OracleSession([cfg.train])
        rng = np.random.default_rng(0)
        for step in range(5):
            x, y = configs.mlp_batch(rng, cfg.feeds)
            cfg.feeds["x"].assign(x)
            cfg.feeds["y"].assign(y)
            sess.assign(cfg.feeds["x"], x)
            sess.assign(cfg.feeds["y"], y)
            err = cfg.train.get()
            want = sess.run()[0]
            assert rel_err(err, want) < 1e-4, (step, err, want)
        for v in cfg.variables:
            assert rel_err(v.data(), sess.leaf_value(v)) < 1e-4
    finally:
        tc.set_evaluator("plan")


@pytest.mark.parametrize("evaluator", ["node", "plan"])
@pytest.mark.parametrize("kind,batch", [("lstm", None), ("lstm", 5), ("gru", None), ("gru", 4)])
def test_recurrent_training_matches_oracle(gpu, kind, batch, evaluator):
    tc.set_evaluator(evaluator)
    try:
        vocab = 12
        cfg = configs.recurrent(kind, vocab=vocab, hidden=16, seq=6, batch=batch)
        sess = OracleSession([cfg.train])
        rng = np.random.default_rng(3)
        for step in range(3):
            x, y = configs.recurrent_batch(rng, cfg.feeds, vocab)
            cfg.feeds["x"].assign(x)
            cfg.feeds["y"].assign(y)
            sess.assign(cfg.feeds["x"], x)
            sess.assign(cfg.feeds["y"], y)
            loss = cfg.train.get()
            want = sess.run()[0]
            assert rel_err(loss, want) < 1e-4, (step, loss, want)
        for v in cfg.variables:
            assert rel_err(v.data(), sess.leaf_value(v)) < 2e-4
    finally:
        tc.set_evaluator("plan")


@pytest.mark.parametrize("kind,batch,seq", [("lstm", 8, 10), ("lstm", 70, 3), ("gru", 8, 10), ("lstm", None, 10)])
def test_recurrent_training_with_step_fusion_matches_oracle(gpu, kind, batch, seq):
    """sizes at which the planner's recurrent-step fusion applies (grouped gate products, K-segmented sums, stacked weight
    gradients: units >= 32, more than 8 time steps for the n-ary ADDs): same oracle, same tolerance as the unfused plans"""
    tc.set_evaluator("plan")
    vocab, hidden = 32, 64
    cfg = configs.recurrent(kind, vocab=vocab, hidden=hidden, seq=seq, batch=batch, learning_rate=0.05)
    if kind == "lstm" and batch is not None:
        plan = tc.describe_plan([cfg.train])
        assert any(s.startswith("GEMM-GROUP x4") for s in plan) and any(s.startswith("GEMM-SUM") for s in plan), plan[:40]
    sess = OracleSession([cfg.train])
    rng = np.random.default_rng(5)
    for step in range(3):
        x, y = configs.recurrent_batch(rng, cfg.feeds, vocab)
        cfg.feeds["x"].assign(x)
        cfg.feeds["y"].assign(y)
        sess.assign(cfg.feeds["x"], x)
        sess.assign(cfg.feeds["y"], y)
        loss = cfg.train.get()
        want = sess.run()[0]
        assert rel_err(loss, want) < 1e-4, (step, loss, want)
    for v in cfg.variables:
        assert rel_err(v.data(), sess.leaf_value(v)) < 5e-4


def test_rbm_training_statistics(gpu):
    """RAND_UNIF streams differ from the reference's std::default_random_engine by design
    (SURVEY.md §2a): CD-1 is checked statistically — the reconstruction error falls."""
    cfg = configs.rbm(nvisible=64, nhidden=16, nbatch=256, learning_rate=0.1)
    rng = np.random.default_rng(1)
    proto = (rng.random((4, 64)) < 0.5).astype(np.float32)
    errs = []
    for step in range(60):
        v = proto[rng.integers(0, 4, 256)]
        cfg.feeds["x"].assign(v)
        errs.append(float(cfg.train.get()))
    assert np.isfinite(errs).all()
    assert np.mean(errs[-10:]) < 0.8 * np.mean(errs[:5]), (errs[:5], errs[-10:])


def test_version_gating_and_repeat_get(gpu):
    """`get()` twice without new inputs re-applies non-idempotent ASSIGN_SUB exactly like the
    reference (functor.hpp:246-269): the weights keep moving, the result stays finite."""
    cfg = configs.mlp(10, 9, 5, 3)
    rng = np.random.default_rng(0)
    x, y = configs.mlp_batch(rng, cfg.feeds)
    cfg.feeds["x"].assign(x)
    cfg.feeds["y"].assign(y)
    e0 = float(cfg.train.get())
    w0 = cfg.variables[0].data().copy()
    e1 = float(cfg.train.get())
    w1 = cfg.variables[0].data().copy()
    assert np.isfinite([e0, e1]).all() and e1 < e0  # second SGD step on the same batch lowers the error
    assert not np.array_equal(w0, w1)


@pytest.mark.parametrize("evaluator", ["node", "plan"])
@pytest.mark.parametrize("nbatch", [1, 32, 257])
def test_dqn_training_matches_oracle(gpu, nbatch, evaluator):
    """C5: source/target nets, masked TD error, rms_momentum + clip_by_l2norm, soft target update
    (extenncor/dqn_trainer.py:15-48)."""
    tc.set_evaluator(evaluator)
    try:
        cfg = configs.dqn(nbatch=nbatch)
        sess = OracleSession([cfg.train])
        rng = np.random.default_rng(4)
        for step in range(4):
            batch = configs.dqn_batch(rng, cfg.feeds)
            for k, f in cfg.feeds.items():
                f.assign(batch[k])
                sess.assign(f, batch[k])
            err = cfg.train.get()
            want = sess.run()[0]
            assert rel_err(err, want) < 1e-4, (step, err, want)
        for v in cfg.variables:
            assert rel_err(v.data(), sess.leaf_value(v)) < 1e-4
    finally:
        tc.set_evaluator("plan")


def test_prefetch_commit_matches_assign(gpu):
    """EVariable.prefetch / commit (copy-stream H2D one step ahead) trains exactly like assign."""
    rng = np.random.default_rng(5)
    batches = [configs.mlp_batch(rng, configs.mlp(64, 48, 16, 33).feeds) for _ in range(5)]

    def losses(pipelined):
        cfg = configs.mlp(64, 48, 16, 33)
        out = []
        if pipelined:
            for f, arr in zip(cfg.feeds.values(), batches[0]):
                f.prefetch(arr)
        for i in range(len(batches)):
            if pipelined:
                for f in cfg.feeds.values():
                    f.commit()
                if i + 1 < len(batches):
                    for f, arr in zip(cfg.feeds.values(), batches[i + 1]):
                        f.prefetch(arr)
            else:
                for f, arr in zip(cfg.feeds.values(), batches[i]):
                    f.assign(arr)
            out.append(float(cfg.train.get()))
        tc.sync_prefetch()
        return out, [v.data().copy() for v in cfg.variables]

    want, wvars = losses(False)
    got, gvars = losses(True)
    assert got == want
    for g, w in zip(gvars, wvars):
        np.testing.assert_array_equal(g, w)
    x = cfgx = configs.mlp(10, 9, 5, 3).feeds["x"]
    arr = rng.random(x.shape(), dtype=np.float32)
    x.prefetch(arr)
    x.commit()
    np.testing.assert_array_equal(x.data(), arr)
    with pytest.raises(Exception):
        x.commit()  # nothing staged
    with pytest.raises(Exception):
        x.prefetch(arr.astype(np.float64))  # asynchronous copies do not convert


@pytest.mark.parametrize("evaluator", ["node", "plan"])
def test_pretrained_onnx_model_runs_and_keeps_training(gpu, evaluator):
    """BASELINE config 1 names models/gd.onnx: the file the reference's serializer wrote loads, evaluates on the device
    like the oracle, solves the gd_demo task, and trains on from its weights (demo/gd_demo.py:72-96)."""
    import os
    tc.set_evaluator(evaluator)
    try:
        from tests.test_onnx import golden_model
        model = tc.load_from_file(golden_model("gd"))[0]
        rng = np.random.default_rng(0)
        x = rng.random((200, 10)).astype(np.float32)
        y = (x[:, 0::2] + x[:, 1::2]) / 2
        testin = tc.variable(x, "testin")
        out = model.connect(testin)
        got = out.get()
        want = OracleSession([out]).run()[0]
        assert rel_err(got, want) < 1e-5
        assert float(np.mean(np.abs(got.reshape(200, 5) - y))) < 0.05
        train_input = tc.EVariable([3, 10], 0, "train_input")
        train_exout = tc.EVariable([3, 5], 0, "train_exout")
        train = tc.apply_update([model], lambda err, leaves: tc.api.approx.sgd(err, leaves, learning_rate=0.9),
                                lambda models: tc.api.loss.mean_squared(train_exout, models[0].connect(train_input)))
        sess = OracleSession([train])
        for step in range(4):
            bx = rng.random((3, 10)).astype(np.float32)
            by = ((bx[:, 0::2] + bx[:, 1::2]) / 2).astype(np.float32)
            train_input.assign(bx)
            train_exout.assign(by)
            sess.assign(train_input, bx)
            sess.assign(train_exout, by)
            assert rel_err(train.get(), sess.run()[0]) < 1e-4
    finally:
        tc.set_evaluator("plan")
