"""Evaluator plugins (dbg/peval/plugin_eval.hpp, dbg/peval/stats/inspect.hpp) and the per-opcode profiler, installed
through the reference's evaluator slot (teq::set_eval)."""
import collections

import numpy as np
import pytest

import tenncor_b200 as tc
from tenncor_b200 import configs
from tests.test_train_gpu import OracleSession

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_evaluator():
    yield
    tc.set_evaluator("plan")


def test_plugable_evaluator_runs_plugins_after_the_traversal(gpu):
    rng = np.random.default_rng(0)
    a = tc.variable(rng.uniform(-2, 2, (5, 7)).astype(np.float32), "a")
    hidden = tc.api.tanh(a) * 3.0
    out = tc.api.reduce_sum(hidden, 0, 1)
    insp = tc.Inspector()
    insp.add(hidden, "hidden")
    insp.add(a, "a leaf is not a functor: ignored like the reference")
    ev = tc.PlugableEvaluator()
    ev.add_plugin(insp)
    tc.set_eval(ev)
    # an intermediate's buffer expires with its last consumer (device.hpp:555-570): to be inspected it must be a target,
    # exactly as with the reference's Inspector
    got = tc.run([out, hidden])[0]
    want = OracleSession([out, hidden]).run()
    np.testing.assert_allclose(got.reshape(-1), want[0], rtol=1e-5)
    last = insp.last()
    assert list(last) == ["hidden"]
    np.testing.assert_allclose(last["hidden"], (want[1].min(), want[1].max()), rtol=1e-5)


def test_op_profiler_accounts_every_functor(gpu):
    cfg = configs.mlp(64, 48, 16, 33)
    rng = np.random.default_rng(1)
    x, y = configs.mlp_batch(rng, cfg.feeds)
    cfg.feeds["x"].assign(x)
    cfg.feeds["y"].assign(y)
    prof = tc.OpProfiler()
    tc.set_eval(prof)
    loss = float(cfg.train.get())
    assert np.isfinite(loss)
    report = {r["opcode"]: r for r in prof.report()}
    want = collections.Counter(n["op"] for n in tc.dump_graph([cfg.train]) if n["kind"] != "leaf")
    assert {k: v["calls"] for k, v in report.items()} == dict(want)
    assert all(r["ms"] >= 0 and r["bytes"] > 0 for r in report.values()) and report["CONTRACT"]["ms"] > 0
    assert abs(sum(r["share"] for r in report.values()) - 1) < 1e-9
    assert report["CONTRACT"]["calls"] >= 4  # two forwards (before / after the update) plus the backward products
    prof.reset()
    assert prof.report() == []
