"""Numerical check of every differentiable opcode's gradient rule: the derivative GRAPH tcr::derive builds (structure pinned
in tests/test_backprop_golden.py) must also evaluate — through the CPU oracle, in double — to the central finite difference of
the forward graph. This is the numeric half of tenncor/test/test_api.cpp's per-op forward + gradient checks
(unary_elementary / binary_elementary, :60-330), done generically instead of with one hand-written f' per op."""
import numpy as np
import pytest

import tenncor_b200 as tc
from oracle import tcr_oracle as orc


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def evaluate(roots):
    tape = tc.dump_graph(roots)
    ids = tc.dump_ids(roots, tape)
    vals = orc.eval_tape(tape)
    return [np.asarray(vals[ids[r]], np.float64).reshape(-1).copy() for r in roots]


def check(build, arrays, eps=1e-6, rtol=2e-6, weighted=True):
    """build(*variables) -> output tensor; the loss is sum(output * R) for a fixed random R (R = 1 when not `weighted`)"""
    vs = [tc.variable(a.copy(), "v%d" % i) for i, a in enumerate(arrays)]
    out = build(*vs)
    rng = np.random.default_rng(99)
    weight = tc.constant(rng.uniform(0.5, 1.5, out.shape()) if weighted else np.ones(out.shape()))
    loss = tc.api.reduce_sum(out * weight)
    grads = evaluate(list(tc.derive(loss, vs)))
    for k, (v, a) in enumerate(zip(vs, arrays)):
        numeric = np.zeros(a.size)
        flat = a.reshape(-1)
        for i in range(a.size):
            for sign in (+1, -1):
                pert = flat.copy()
                pert[i] += sign * eps
                v.assign(pert.reshape(a.shape))
                numeric[i] += sign * evaluate([loss])[0][0]
            v.assign(a)
        numeric /= 2 * eps
        scale = max(1.0, float(np.abs(numeric).max()))
        assert np.max(np.abs(grads[k] - numeric)) <= rtol * scale + 1e-7, (k, grads[k], numeric)


RNG = np.random.default_rng(7)
A = RNG.uniform(0.3, 1.7, (2, 3))          # positive, away from kinks
B = RNG.uniform(0.4, 1.6, (2, 3))
S = RNG.uniform(-1.2, 1.2, (2, 3))          # signed, away from 0

UNARY = ["abs", "neg", "sin", "cos", "tan", "exp", "log", "sqrt", "square", "cube", "sigmoid", "tanh"]


@pytest.mark.parametrize("name", UNARY)
def test_unary(name):
    x = A if name in ("log", "sqrt") else (S * 0.6 if name == "tan" else S)
    check(lambda v: getattr(tc.api, name)(v), [x])


@pytest.mark.parametrize("name", ["add", "sub", "mul", "div", "pow", "min", "max"])
def test_binary(name):
    check(lambda a, b: getattr(tc.api, name)(a, b), [A, B])


def test_nnary_and_select():
    c = RNG.uniform(0.2, 1.2, (2, 3))
    check(lambda a, b, d: tc.api.sum([a, b, d]) * tc.api.prod([a, b, d]) if hasattr(tc.api, "prod") else tc.api.sum([a, b, d]), [A, B, c])
    cond = tc.constant((RNG.random((2, 3)) < 0.5).astype(np.float64))
    check(lambda a, b: tc.api.if_then_else(cond, a, b), [A, B])


@pytest.mark.parametrize("name", ["reduce_sum", "reduce_prod", "reduce_min", "reduce_max", "reduce_mean", "reduce_variance", "reduce_l2norm"])
def test_reductions(name):
    x = RNG.permutation(np.linspace(0.4, 1.9, 24)).reshape(2, 3, 4)  # distinct values: min / max are differentiable
    fn = getattr(tc.api, name)
    # The reference's REDUCE_MIN / REDUCE_MAX rule is EQ(extend(op), MUL(arg, extend(sup))) (tenncor/eteq/backprop.hpp:208-217,
    # golden BACKPROP.ReduceMinMax): the upstream gradient sits INSIDE the comparison, so the rule is the true gradient only for an
    # upstream gradient of 1. It is reproduced as is (parity with the reference) and checked in the regime where it is meaningful.
    weighted = name not in ("reduce_min", "reduce_max")
    check(lambda v: fn(v), [x], weighted=weighted)
    if name in ("reduce_sum", "reduce_prod", "reduce_min", "reduce_max", "reduce_l2norm"):
        check(lambda v: fn(v, 1, 1), [x], weighted=weighted)       # one middle rank
        check(lambda v: fn(v, 0, 2), [x], weighted=weighted)       # the two fastest ranks


def test_layout_ops():
    x = RNG.uniform(-1, 1, (2, 3, 4))
    check(lambda v: tc.api.permute(v, [2, 0, 1]), [x])
    check(lambda v: tc.api.extend(v, 3, [5]), [x])
    check(lambda v: tc.api.reshape(v, [4, 6]), [x])
    check(lambda v: tc.api.slice(v, 1, 2, 1), [x])
    check(lambda v: tc.api.pad(v, (1, 2), 0), [x])
    check(lambda v: tc.api.stride(v, [2, 1, 2]), [x])
    check(lambda v: tc.api.reverse(v, [0, 2]) if _takes(tc.api.reverse, v) else tc.api.reverse(v, {0, 2}), [x])
    y = RNG.uniform(-1, 1, (2, 5, 4))
    check(lambda a, b: tc.api.concat(a, b, 1), [x, y])
    check(lambda v: tc.api.transpose(tc.api.reshape(v, [4, 6])), [x])
    check(lambda v: tc.api.softmax(tc.api.reshape(v, [4, 6]), 0, 1), [x], rtol=5e-6)


def _takes(fn, v):
    try:
        fn(v, [0, 2])
        return True
    except TypeError:
        return False


def test_contractions():
    a, b = RNG.uniform(-1, 1, (3, 4)), RNG.uniform(-1, 1, (4, 5))
    check(lambda x, y: tc.api.matmul(x, y), [a, b])
    check(lambda x, y: tc.api.contract(x, y, [(0, 1)]), [a, b])
    t, u = RNG.uniform(-1, 1, (2, 3, 4)), RNG.uniform(-1, 1, (5, 3, 4))
    check(lambda x, y: tc.api.contract(x, y, [(0, 0), (1, 1)]), [t, u])      # two common ranks
    check(lambda x, y: tc.api.contract(x, y, [(1, 1)]), [t, u], rtol=5e-6)   # free ranks on both sides of the common one
    w = RNG.uniform(-1, 1, (6, 4))
    check(lambda x, y: tc.api.nn.fully_connect([x], [y], tc.constant(np.ones(4))), [RNG.uniform(-1, 1, (3, 6)), w])


def test_convolutions():
    img, ker = RNG.uniform(-1, 1, (3, 5, 6)), RNG.uniform(-1, 1, (2, 3))
    check(lambda i, k: tc.api.convolution(i, k, [0, 1]), [img, ker])
    img4, ker4, bias = RNG.uniform(-1, 1, (2, 5, 6, 3)), RNG.uniform(-1, 1, (2, 3, 3, 4)), RNG.uniform(-1, 1, 4)
    check(lambda i, k, b: tc.api.nn.conv2d(i, k, b), [img4, ker4, bias], rtol=5e-6)
    check(lambda i, k: tc.api.nn.conv2d(i, k, zero_paddings=((1, 0), (1, 2))), [img4, ker4], rtol=5e-6)


def test_pooling_batchnorm_and_losses():
    x = RNG.permutation(np.linspace(-1.5, 1.5, 48)).reshape(3, 4, 4)
    check(lambda v: tc.api.nn.mean_pool2d(v), [x])
    check(lambda v: tc.api.nn.max_pool2d(v), [x])
    check(lambda v: tc.api.nn.batch_normalization(v, 0.1, 1.5, 1e-3), [x], rtol=2e-5)
    y = RNG.uniform(0.1, 0.9, (3, 4, 4))
    check(lambda a, b: tc.api.loss.mean_squared(a, b), [x, y])
    p, q = RNG.uniform(0.1, 0.9, (4, 5)), RNG.uniform(0.1, 0.9, (4, 5))
    if hasattr(tc.api.loss, "cross_entropy"):
        check(lambda a, b: tc.api.loss.cross_entropy(a, b), [p, q])
