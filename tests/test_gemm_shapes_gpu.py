"""Skinny (HBM-bound) products and split-K: the output-layer shapes of the demo MLPs."""
import numpy as np
import pytest

from tests.test_gemm_tc_gpu import run_gemm
from tests.test_ops_gpu import _gemm

pytestmark = pytest.mark.gpu

SKINNY = [(8192, 10, 1024), (10, 1024, 8192), (8192, 1024, 10), (4096, 5, 96), (3, 2000, 4100), (700, 16, 33), (1000, 1000, 3), (64, 1, 5000), (1, 300, 70000)]


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32], ids=["f32", "f64", "i32"])
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)], ids=["NN", "TN", "NT", "TT"])
@pytest.mark.parametrize("M,N,K", SKINNY, ids=[str(s) for s in SKINNY])
def test_skinny_products_all_layouts(gpu, M, N, K, ta, tb, dt):
    rng = np.random.default_rng(M + 3 * N + 7 * K)
    if np.issubdtype(dt, np.integer):
        A, B = rng.integers(-3, 4, (1, M, K)).astype(dt), rng.integers(-3, 4, (1, K, N)).astype(dt)
    else:
        A, B = rng.uniform(-1, 1, (1, M, K)).astype(dt), rng.uniform(-1, 1, (1, K, N)).astype(dt)
    got = _gemm(gpu, A, B, ta, tb, precision=2 if dt == np.float32 else 0)
    want = np.matmul(A.astype(np.float64), B.astype(np.float64))
    if np.issubdtype(dt, np.integer):
        np.testing.assert_array_equal(got, want.astype(dt))
    else:
        bound = np.matmul(np.abs(A).astype(np.float64), np.abs(B).astype(np.float64)) * (K * np.finfo(dt).eps + 2.0 ** -19)
        assert np.all(np.abs(got - want) <= bound + 1e-300)


def test_skinny_epilogue_and_perm(gpu):
    rng = np.random.default_rng(3)
    M, N, K = 2048, 10, 512
    A, B = rng.uniform(-1, 1, (1, M, K)).astype(np.float32), rng.uniform(-1, 1, (1, K, N)).astype(np.float32)
    bias = rng.uniform(-1, 1, N).astype(np.float32)
    got = _gemm(gpu, A, B, 0, 0, precision=2, bias=bias, epi=gpu.EPI_BIAS_N, act=gpu.OP["SIGMOID"])
    want = 1 / (1 + np.exp(-(np.matmul(A.astype(np.float64), B.astype(np.float64)) + bias)))
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    got = _gemm(gpu, A, B, 1, 1, precision=2, perm_c=True)
    np.testing.assert_allclose(got, np.matmul(A.astype(np.float64), B.astype(np.float64)), rtol=0, atol=K * 1e-6)


@pytest.mark.parametrize("precision", [1, 2], ids=["tf32", "3xtf32"])
@pytest.mark.parametrize("M,N,K,ta,tb", [(1024, 784, 8192, 1, 0), (256, 128, 4096, 0, 0), (128, 128, 32768, 0, 1), (384, 640, 2048, 1, 1),
                                         (28, 64, 65536, 1, 0), (48, 40, 20000, 0, 0), (288, 64, 65536, 1, 0)])  # few outputs, long reduction: conv kernel gradients
def test_split_k_is_deterministic_and_accurate(gpu, M, N, K, ta, tb, precision):
    rng = np.random.default_rng(K)
    A, B = rng.uniform(-1, 1, (M, K)).astype(np.float32), rng.uniform(-1, 1, (K, N)).astype(np.float32)
    got = run_gemm(gpu, A, B, ta, tb, precision)
    again = run_gemm(gpu, A, B, ta, tb, precision)
    np.testing.assert_array_equal(got, again)  # two-pass reduction over the splits, no atomics
    want = A.astype(np.float64) @ B.astype(np.float64)
    S = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    bound = S * (2.0 ** -10 if precision == 1 else (2.0 ** -19 + K * 2.0 ** -23))
    assert np.all(np.abs(got - want) <= bound)
