"""The host generator behind the initialisers and tc.unif_gen / tc.norm_gen (global::Randomizer, internal/global/random.hpp:78-146),
mirrored from internal/global/test/test_random.cpp. The device-side RAND_UNIF stream (Philox, statistical parity by design) is
checked under -m gpu in tests/test_ops_gpu.py."""
import numpy as np
import pytest

import tenncor_b200 as tc


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def test_uniform():  # RANDOM.RandomizerUniform :34-56
    tc.seed(0)
    gen = tc.unif_gen(2, 7)
    draws = np.array([gen() for _ in range(2000)])
    assert draws.min() >= 2 and draws.max() <= 7
    assert abs(draws.mean() - 4.5) < 0.15 and draws.min() < 2.1 and draws.max() > 6.9


def norm_check(draws, mean, stdev):  # NormChecker :58-97
    d = np.abs(np.asarray(draws) - mean) / stdev
    prob68, prob95, prob99 = (100.0 * np.mean(d < k) for k in (1, 2, 3))
    assert 63 <= prob68 <= 73
    assert 92 <= prob95 <= 98
    assert 96 <= prob99


def test_normal():  # RANDOM.RandomizerNorm :100-120
    tc.seed(0)
    mean, stdev = 2, 3
    gen, gen2 = tc.norm_gen(mean, stdev), tc.norm_gen(mean, stdev)
    a, b = [], []
    for _ in range(1000):
        a.append(gen())
        b.append(gen2())
    norm_check(a, mean, stdev)
    norm_check(b, mean, stdev)


def test_seed_reproduces_the_stream():  # global::seed: same seed, same draws — initialisers included
    tc.seed(7)
    first = [tc.unif_gen(0, 1)() for _ in range(5)]
    w1 = tc.variable_from_init(tc.api.init.xavier_uniform(), [4, 6], "w").data()
    tc.seed(7)
    assert [tc.unif_gen(0, 1)() for _ in range(5)] == first
    np.testing.assert_array_equal(tc.variable_from_init(tc.api.init.xavier_uniform(), [4, 6], "w").data(), w1)
    tc.seed(8)
    assert [tc.unif_gen(0, 1)() for _ in range(5)] != first
