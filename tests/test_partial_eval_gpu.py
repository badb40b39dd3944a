"""Version-gated re-evaluation inside a plan: only the steps downstream of what changed run again, like the reference's per-functor
gate (tenncor/eteq/functor.hpp:246-269, internal/eigen/device.hpp:555-570); values must equal a full evaluation every time."""
import numpy as np
import pytest

import tenncor_b200 as tc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def plan_evaluator():
    tc.set_evaluator("plan")
    yield
    tc.set_evaluator("plan")


def _net(rng, n=48, h=64):
    x1 = tc.EVariable([n, 40], 0, "x1")
    x2 = tc.EVariable([n, 24], 0, "x2")
    w1 = tc.variable(rng.uniform(-0.3, 0.3, (40, h)).astype(np.float32), "w1")
    w2 = tc.variable(rng.uniform(-0.3, 0.3, (24, h)).astype(np.float32), "w2")
    w3 = tc.variable(rng.uniform(-0.3, 0.3, (h, h)).astype(np.float32), "w3")
    a = tc.api.tanh(tc.api.matmul(tc.api.sigmoid(tc.api.matmul(x1, w1)), w3))     # two products deep
    b = tc.api.exp(tc.api.matmul(x2, w2) * 0.1)
    mix = a * b + tc.api.square(a)
    y = tc.api.reduce_sum(mix, 0, 1)  # per-row sums (teq rank 0 = the fastest = numpy's last axis)
    return dict(x1=x1, x2=x2, w1=w1, w2=w2, w3=w3, a=a, b=b, y=y)


def _want(v):
    f = {k: np.asarray(v[k].data(), np.float64) for k in ("x1", "x2", "w1", "w2", "w3")}
    a = np.tanh((1 / (1 + np.exp(-(f["x1"] @ f["w1"])))) @ f["w3"])
    b = np.exp((f["x2"] @ f["w2"]) * 0.1)
    return (a * b + a * a).sum(axis=1)


def test_only_the_stale_steps_run_and_values_match_a_full_evaluation(gpu):
    rng = np.random.default_rng(7)
    v = _net(rng)
    new = lambda var: rng.uniform(-1, 1, var.shape()).astype(np.float32)  # noqa: E731
    v["x1"].assign(new(v["x1"]))
    v["x2"].assign(new(v["x2"]))
    for _ in range(3):  # eager run, captured run, replay: everything stale each time
        v["x1"].touch()
        v["x2"].touch()
        got = v["y"].get()
        st = tc.plan_stats()
        assert st["steps_run"] == st["steps"] > 3
        np.testing.assert_allclose(np.asarray(got).reshape(-1), _want(v), rtol=2e-5)
    base = tc.plan_stats()["partial_runs"]
    v["y"].get()  # nothing changed
    assert tc.plan_stats()["steps_run"] == 0
    seen = set()
    for name in ("x2", "x1", "w3", "x2", "w2", "w1", "x1"):
        v[name].assign((new(v[name]) * (0.3 if name[0] == "w" else 1.0)).astype(np.float32))
        got = v["y"].get()
        st = tc.plan_stats()
        assert 0 < st["steps_run"] < st["steps"], (name, st)
        seen.add(st["steps_run"])
        np.testing.assert_allclose(np.asarray(got).reshape(-1), _want(v), rtol=2e-5, err_msg=name)
    assert tc.plan_stats()["partial_runs"] == base + 7
    assert len(seen) > 1  # different inputs reach different parts of the plan
    # both inputs at once: the whole plan (graph replay) again
    v["x1"].assign(new(v["x1"]))
    v["x2"].assign(new(v["x2"]))
    v["w1"].touch(); v["w2"].touch(); v["w3"].touch()
    got = v["y"].get()
    st = tc.plan_stats()
    assert st["steps_run"] == st["steps"]
    np.testing.assert_allclose(np.asarray(got).reshape(-1), _want(v), rtol=2e-5)


def test_a_plan_notices_versions_bumped_by_another_plan(gpu):
    """Versions are global, buffers are per plan: after plan A re-evaluated a shared functor for a new input, plan B must not
    hand out what its own buffers held before."""
    rng = np.random.default_rng(8)
    v = _net(rng)
    new = lambda var: rng.uniform(-1, 1, var.shape()).astype(np.float32)  # noqa: E731
    v["x1"].assign(new(v["x1"]))
    v["x2"].assign(new(v["x2"]))
    for _ in range(2):
        v["x1"].touch()
        a0 = np.asarray(v["a"].get()).copy()   # plan B: target a
        y0 = np.asarray(v["y"].get()).copy()   # plan A: target y, contains a as an interior value
    v["x1"].assign(new(v["x1"]))
    y1 = np.asarray(v["y"].get())               # plan A bumps the versions of everything under a
    a1 = np.asarray(v["a"].get())               # plan B: nothing "changes" during its own version walk
    f = {k: np.asarray(v[k].data(), np.float64) for k in ("x1", "w1", "w3")}
    want_a = np.tanh((1 / (1 + np.exp(-(f["x1"] @ f["w1"])))) @ f["w3"])
    np.testing.assert_allclose(a1, want_a, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(y1.reshape(-1), _want(v), rtol=2e-5)
    assert not np.allclose(a0, a1) and not np.allclose(y0, y1)


def test_assign_steps_are_still_applied_on_every_evaluation(gpu):
    """an ASSIGN is stale by construction after it ran (its variable carries the newer version): repeated get() keeps applying it"""
    x = tc.variable(np.full((4, 5), 2.0, np.float32), "x")
    step = tc.api.assign_add(x, tc.api.square(tc.variable(np.full((4, 5), 0.5, np.float32), "d")))
    vals = [float(np.asarray(step.get()).reshape(-1)[0]) for _ in range(4)]
    assert vals == [2.25, 2.5, 2.75, 3.0]
