"""The evaluator's traversal (teq::Evaluator / TravEvaluator::visit_func, internal/teq/evaluator.hpp:11-62)
mirrored from internal/teq/test/test_evaluator.cpp. A recording device (tc.RecordingDevice, the reference's MockDevice) stands in
for the GPU: what is checked is WHICH functors reach device.calc, in what ORDER, with which is-target flag, and how the
`ignored` set prunes the walk. The planned evaluator must fall back to exactly this traversal on a device that is not its own."""
import numpy as np
import pytest

import tenncor_b200 as tc


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def leaf(name):
    return tc.variable(np.ones((2, 2), dtype=np.float32), name)


def walk(targets, ignored=(), evaluator=None):
    dev = tc.RecordingDevice()
    for ig in ignored:
        tc.testing.mock_data(ig)  # the reference's tests give ignored nodes a MockDeviceRef with data: ignoring needs existing data
    tc.evaluate_on(evaluator or tc.Evaluator(), dev, list(targets), list(ignored))
    return dev.calls()


EVALUATORS = [tc.Evaluator, tc.PlanEvaluator] if hasattr(tc, "PlanEvaluator") else [tc.Evaluator]


@pytest.mark.parametrize("make", EVALUATORS, ids=lambda m: m.__name__)
def test_update(make):  # EVALUATOR.Update :20-52 — post-order, leaves never reach the device, only the target carries the flag
    a, b, c = leaf("a"), leaf("b"), leaf("c")
    x = a + b
    target = x * c
    calls = walk([target], evaluator=make())
    assert [t for t, _ in calls] == [x, target]
    assert [ttl for _, ttl in calls] == [0, 1]


@pytest.mark.parametrize("make", EVALUATORS, ids=lambda m: m.__name__)
def test_update_ignore(make):  # EVALUATOR.UpdateIgnore :53-108 — an ignored functor is a leaf of the walk: nothing below it is visited
    a, b, c, d = (leaf(n) for n in "abcd")
    x = a + b
    y = x * c
    target = y - d
    assert [t for t, _ in walk([target], [y], make())] == [target]
    assert [t for t, _ in walk([target], [x], make())] == [y, target]


@pytest.mark.parametrize("make", EVALUATORS, ids=lambda m: m.__name__)
def test_update_ignore_common_descendant(make):  # EVALUATOR.UpdateIgnoreCommonDesc :109-162
    """u sits under the ignored y AND under x: it is still evaluated (once), through x"""
    a, b, c = leaf("a"), leaf("b"), leaf("c")
    u = tc.api.neg(a)
    x = u + b
    y = c * u
    target = y - x
    assert [t for t, _ in walk([target], [y], make())] == [u, x, target]
    # without the ignore, u is evaluated once, before its first reader
    assert [t for t, _ in walk([target], [], make())] == [u, y, x, target]


@pytest.mark.parametrize("make", EVALUATORS, ids=lambda m: m.__name__)
def test_targeted_update(make):  # EVALUATOR.TargetedUpdate / TargetedUpdateIgnore :163-230 — a target below the root: nothing above it runs
    a, b, c, d = (leaf(n) for n in "abcd")
    x = a + b
    y = x * c
    target = y - d
    with pytest.raises(Exception, match="cannot ignore tensor .* without existing data"):
        tc.evaluate_on(make(), tc.RecordingDevice(), [target], [y])  # evaluator.hpp:17-24
    assert [t for t, _ in walk([x], [], make())] == [x]
    assert [t for t, _ in walk([y], [x], make())] == [y]
    calls = walk([x, target], [], make())  # two targets: each flagged, shared work once
    assert sorted(t.opname() for t, ttl in calls if ttl == 1) == ["ADD", "SUB"] and len(calls) == 3
