"""API composites that complete the reference's python surface (cfg/tenncor/nn.yml, layer.yml, init.yml): pooling, batch
normalization, dropout, and the remaining initialisers. They are graphs over existing opcodes, so they are checked on CPU:
the functor graph each builds is evaluated by the oracle and compared with a direct numpy statement of the definition."""
import numpy as np
import pytest

import tenncor_b200 as tc
from oracle import tcr_oracle as orc


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def _eval(roots):
    tape = tc.dump_graph(roots)
    ids = tc.dump_ids(roots, tape)
    for node in tape:
        if node["kind"] == "leaf":
            node["data"] = np.array(node["data"], copy=True)
    vals = orc.eval_tape(tape)
    return [np.asarray(vals[ids[r]], np.float64).reshape(r.shape()) for r in roots], tape, ids


@pytest.mark.parametrize("shape", [(6, 8), (3, 5, 7), (2, 4, 4, 3)])
def test_pool2d(shape):
    """2x2 windows, stride 2, over the two fastest teq ranks = the last two numpy axes (nn.yml:209-268); odd extents keep
    the reference's behaviour: STRIDE rounds up, the shifted slices are one shorter, and ADD / MAX need equal shapes — so
    only even extents are valid inputs, as in the reference."""
    rng = np.random.default_rng(0)
    shape = tuple(s + (s % 2) if i >= len(shape) - 2 else s for i, s in enumerate(shape))
    x_np = rng.standard_normal(shape)
    x = tc.variable(x_np, "x")
    (mean, mx), _, _ = _eval([tc.api.nn.mean_pool2d(x), tc.api.nn.max_pool2d(x)])
    h, w = shape[-2] // 2, shape[-1] // 2
    win = x_np.reshape(shape[:-2] + (h, 2, w, 2))
    np.testing.assert_allclose(mean, win.mean(axis=(-3, -1)), rtol=1e-12)
    np.testing.assert_array_equal(mx, win.max(axis=(-3, -1)))
    gm = tc.derive(tc.api.reduce_sum(tc.api.nn.mean_pool2d(x)), [x])[0]
    (g,), _, _ = _eval([gm])
    np.testing.assert_allclose(g, np.full(shape, 0.25), rtol=1e-12)


def test_batch_normalization_whole_tensor_and_per_axis():
    rng = np.random.default_rng(1)
    x_np = rng.standard_normal((5, 4, 6)) * 3 + 2
    x = tc.variable(x_np, "x")
    eps = np.finfo(np.float64).eps
    (whole,), _, _ = _eval([tc.api.nn.batch_normalization(x, 0.5, 2.0)])
    np.testing.assert_allclose(whole, (x_np - x_np.mean()) / np.sqrt(x_np.var() + eps) * 2.0 + 0.5, rtol=1e-10)
    assert abs(whole.mean() - 0.5) < 1e-9 and abs(whole.std() - 2.0) < 1e-6
    # layer form, statistics along the slowest rank (teq rank 2 = numpy axis 0). reduce_*_1d drops the rank, so like the
    # reference's own composite only the last non-singular rank can be extended back (extend_like of [6,4] to [6,4,5])
    (per_axis,), _, _ = _eval([tc.api.layer.batch_normalization(x, axis=2)])
    want = (x_np - x_np.mean(axis=0, keepdims=True)) / np.sqrt(x_np.var(axis=0, keepdims=True) + eps)
    np.testing.assert_allclose(per_axis, want, rtol=1e-9, atol=1e-12)
    with pytest.raises(Exception, match="extend"):
        tc.api.layer.batch_normalization(x, axis=0)
    # tensor-valued offset / scale are broadcast like the reference's extend_like
    off, sc = tc.variable(rng.standard_normal(6), "offset"), tc.variable(rng.random(6) + 0.5, "scale")
    (affine,), _, _ = _eval([tc.api.nn.batch_normalization(x, off, sc, tc.scalar_constant(1e-3, [5, 4, 6], "DOUBLE"))])
    base = (x_np - x_np.mean()) / np.sqrt(x_np.var() + 1e-3)
    np.testing.assert_allclose(affine, base * np.asarray(sc.data()) + np.asarray(off.data()), rtol=1e-10)


def test_batch_normalization_moving_statistics():
    """with `training`, the statistics are the batch's where training != 0 and the momentum-updated moving ones elsewhere;
    the moving variables are ASSIGNed in place on every evaluation (layer.yml:581-632)."""
    rng = np.random.default_rng(2)
    x_np = rng.standard_normal((4, 3)) + 1.0
    x = tc.variable(x_np, "x")
    training = tc.variable(np.array(0.0), "training")  # inference: use (and update) the moving statistics
    out = tc.api.layer.batch_normalization(x, training=training, momentum=0.9)
    (got,), tape, ids = _eval([out])
    eps = np.finfo(np.float64).eps
    mmean = 0.0 * 0.9 + x_np.mean() * 0.1
    mvar = 1.0 * 0.9 + x_np.var() * 0.1
    np.testing.assert_allclose(got, (x_np - mmean) / np.sqrt(mvar + eps), rtol=1e-10)
    leaves = {n["label"]: n["data"] for n in tape if n["kind"] == "leaf" and n.get("label") in ("moving_mean", "moving_var")}
    np.testing.assert_allclose(leaves["moving_mean"], np.full(12, mmean), rtol=1e-12)
    np.testing.assert_allclose(leaves["moving_var"], np.full(12, mvar), rtol=1e-12)
    training.assign(np.array(1.0))
    (got,), _, _ = _eval([out])
    np.testing.assert_allclose(got, (x_np - x_np.mean()) / np.sqrt(x_np.var() + eps), rtol=1e-10)


def test_dropout_graph_and_training_switch():
    x = tc.variable(np.arange(12, dtype=np.float64).reshape(3, 4) + 1, "x")
    off = tc.variable(np.array(0.0), "training")
    out = tc.api.layer.dropout(x, 0.25, off)
    ops = {n["op"] for n in tc.dump_graph([out]) if n["kind"] != "leaf"}
    assert {"RAND_UNIF", "SELECT", "REDUCE_SUM", "DIV", "MUL"} <= ops  # nn.dropout's mask / renormalisation + if_then_else
    assert out.shape() == [3, 4]
    rate = [n for n in tc.dump_graph([out]) if n["kind"] == "leaf" and n.get("label") == "drop_rate"]
    assert len(rate) == 1 and float(rate[0]["data"][0]) == 0.25


def test_initialisers():
    tc.seed(7)
    eye = tc.api.init.identity(gain=2.5)([3, 5], "eye")  # numpy shape [3, 5] = teq [5, 3]
    want = np.zeros((3, 5))
    want[np.arange(3), np.arange(3)] = 2.5
    np.testing.assert_array_equal(eye.data(), want.astype(np.float32))
    with pytest.raises(Exception, match="2D"):
        tc.api.init.identity()([2, 3, 4], "bad")
    tn = tc.api.init.truncated_normal(mean=1.0, stddev=0.5)([200, 100], "tn").data()
    assert tn.min() >= 0.0 - 1e-6 and tn.max() <= 2.0 + 1e-6 and abs(tn.mean() - 1.0) < 0.02 and 0.35 < tn.std() < 0.5
    vs = tc.api.init.variance_scaling(2.0)([300, 100], "vs").data()  # stddev = sqrt(2 / fanavg) = sqrt(2 / 200) = 0.1
    assert abs(vs.std() - 0.1 * 0.88) < 0.01 and np.abs(vs).max() <= 0.2 + 1e-6  # truncation at 2 sigma shrinks the spread to ~0.88 sigma
    vs2 = tc.api.init.variance_scaling(1.0, shape_factor=lambda shape: float(shape[-1]))([300, 100], "vs2").data()  # fan-in only
    assert abs(vs2.std() - 0.1 * 0.88) < 0.01
    tc.seed(7)  # the host generator is seeded with the graph's: the same seed reproduces the same draws (identity draws none)
    tc.api.init.identity(gain=2.5)([3, 5], "eye")
    np.testing.assert_array_equal(tc.api.init.truncated_normal(mean=1.0, stddev=0.5)([200, 100], "tn").data(), tn)


def test_min_max_over_lists_and_clip_build_real_nodes():
    """regression: inside namespace tenncor an unqualified min(ETensor, ETensor) resolved to std::min on the shared_ptrs
    (ADL), so min / max over a list and clip_by_range returned one of their ARGUMENTS instead of a MIN / MAX functor"""
    rng = np.random.default_rng(3)
    arrs = [rng.standard_normal((3, 4)) for _ in range(3)]
    vs = [tc.variable(a, "v%d" % i) for i, a in enumerate(arrs)]
    lo = tc.api.min(vs) if callable(getattr(tc.api, "min", None)) and _accepts_list(tc.api.min, vs) else None
    x = tc.variable(arrs[0], "x")
    clipped = tc.api.clip_by_range(x, -0.25, 0.5)
    assert clipped.opname() == "MAX" and clipped.args()[0].opname() == "MIN"
    (got,), _, _ = _eval([clipped])
    np.testing.assert_array_equal(got, np.clip(arrs[0], -0.25, 0.5))
    if lo is not None:
        (got,), _, _ = _eval([lo])
        np.testing.assert_array_equal(got, np.minimum(np.minimum(arrs[0], arrs[1]), arrs[2]))


def _accepts_list(fn, vs):
    try:
        fn(vs)
        return True
    except TypeError:
        return False


def test_module_level_helpers():
    """tc.unif_gen / norm_gen / variable_from_init / set_log_level / Evaluator (tenncor/python/eteq_ext.cpp:205-405, layr_ext.cpp:74-80)"""
    tc.seed(11)
    gen = tc.unif_gen(2.0, 3.0)
    draws = [gen() for _ in range(200)]
    assert min(draws) >= 2.0 and max(draws) < 3.0 and abs(np.mean(draws) - 2.5) < 0.1
    norm = tc.norm_gen(1.0, 0.1)
    assert abs(np.mean([norm() for _ in range(500)]) - 1.0) < 0.02
    tc.seed(11)
    assert [tc.unif_gen(2.0, 3.0)() for _ in range(1)] == draws[:1]  # seeded with the rest of the host state
    v = tc.variable_from_init(tc.api.init.constants(3.5), [2, 3], "w")
    assert v.shape() == [2, 3] and np.all(v.data() == 3.5)
    w = tc.variable_from_init(lambda shape, label: tc.variable(np.ones(shape, dtype=np.float32) * 2, label), [4], "mine")
    assert np.all(w.data() == 2)
    tc.set_log_level("warn")
    assert tc.get_log_level() == "warn"
    assert isinstance(tc.Evaluator(), tc.iEvaluator) and isinstance(tc.PlanEvaluator(), tc.iEvaluator)
