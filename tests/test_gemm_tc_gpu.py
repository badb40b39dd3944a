"""tcgen05 fp32 GEMM (TF32 / 3xTF32) against a double-precision product.

Tolerances (stated per the north star): with |A|,|B| <= 1 and S = sum_k |a||b|,
  TF32    : |err| <= 2^-10 * S   (operands truncated to 10 mantissa bits, products exact, fp32 accumulate)
  3xTF32  : |err| <= 2^-19 * S + K * 2^-23 * S   (hi/lo split drops lo*lo; fp32 accumulation)
Bit-exact parity with the reference's Eigen GEMM is not defined (summation order differs).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_gemm(gpu, A, B, ta, tb, precision, perm_c=False, bias=None, epi=0, act=0, accumulate=None):
    lib = gpu.lib()
    M, K = A.shape
    N = B.shape[1]
    As = np.ascontiguousarray(A.T) if ta else np.ascontiguousarray(A)
    Bs = np.ascontiguousarray(B.T) if tb else np.ascontiguousarray(B)
    da, db = gpu.to_device(As), gpu.to_device(Bs)
    dc = gpu.to_device(accumulate) if accumulate is not None else gpu.empty(M * N, np.float32)
    d = gpu.GemmDesc(m=M, n=N, k=K, batch=1, a_sm=1 if ta else K, a_sk=M if ta else 1, b_sk=1 if tb else N, b_sn=K if tb else 1,
                     c_sm=1 if perm_c else N, c_sn=M if perm_c else 1, dtype=gpu.FLOAT, precision=precision, epilogue=epi,
                     activation=act, accumulate=1 if accumulate is not None else 0)
    keep = None
    if bias is not None:
        keep = gpu.to_device(bias)
        d.bias = keep.ptr
    gpu.check(lib.tcr_gemm(C.c_void_p(da.ptr), C.c_void_p(db.ptr), C.c_void_p(dc.ptr), C.byref(d)))
    out = gpu.to_host(dc, M * N, np.float32)
    return out.reshape(N, M).T if perm_c else out.reshape(M, N)


SHAPES = [(128, 128, 32), (128, 128, 256), (256, 384, 96), (300, 200, 100), (1000, 72, 520), (64, 4096, 64), (8192, 1024, 784), (129, 257, 36),
          (40000, 64, 288)]  # many tiles x few k-blocks (conv2d patch product): the shallow-ring TF32 variant


@pytest.mark.parametrize("precision", [1, 2], ids=["tf32", "3xtf32"])
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)], ids=["NN", "TN", "NT", "TT"])
@pytest.mark.parametrize("M,N,K", SHAPES, ids=[str(s) for s in SHAPES])
def test_tc_gemm_all_layouts(gpu, M, N, K, ta, tb, precision):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    got = run_gemm(gpu, A, B, ta, tb, precision)
    want = A.astype(np.float64) @ B.astype(np.float64)
    S = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    bound = S * (2.0 ** -10 if precision == 1 else (2.0 ** -19 + K * 2.0 ** -23))
    err = np.abs(got - want)
    assert np.all(err <= bound + 1e-30), (float(err.max()), float((err / (bound + 1e-30)).max()))
    if precision == 2:  # 3xTF32 must be far more accurate than one TF32 pass
        assert err.max() / (np.abs(want).max() + 1e-30) < 1e-5


def test_tc_gemm_epilogue_perm_and_accumulate(gpu):
    rng = np.random.default_rng(5)
    M, N, K = 256, 192, 128
    A, B = rng.uniform(-1, 1, (M, K)).astype(np.float32), rng.uniform(-1, 1, (K, N)).astype(np.float32)
    bias = rng.uniform(-1, 1, N).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    got = run_gemm(gpu, A, B, 0, 0, 2, bias=bias, epi=gpu.EPI_BIAS_N, act=gpu.OP["SIGMOID"])
    np.testing.assert_allclose(got, 1 / (1 + np.exp(-(ref + bias))), rtol=1e-5, atol=1e-6)
    got = run_gemm(gpu, A, B, 1, 0, 2, perm_c=True)  # trailing PERMUTE{1,0} absorbed into the output strides
    np.testing.assert_allclose(got, ref, rtol=0, atol=K * 1e-6)
    base = rng.uniform(-1, 1, (M, N)).astype(np.float32)
    got = run_gemm(gpu, A, B, 0, 1, 2, accumulate=base.reshape(-1).copy())
    np.testing.assert_allclose(got, ref + base, rtol=0, atol=K * 1e-6)


def test_unaligned_operands_take_exact_path(gpu):
    """pitch not a multiple of 16 bytes (e.g. the 1024x10 output layer of the MNIST MLP): TMA cannot address it;
    tcr_gemm silently uses the SIMT kernel (still on device) and stays within the 3xTF32 bound."""
    rng = np.random.default_rng(9)
    M, N, K = 512, 10, 1024
    A, B = rng.uniform(-1, 1, (M, K)).astype(np.float32), rng.uniform(-1, 1, (K, N)).astype(np.float32)
    got = run_gemm(gpu, A, B, 0, 0, 2)
    want = A.astype(np.float64) @ B.astype(np.float64)
    assert np.abs(got - want).max() < K * 1e-6


def test_single_cta_kernel_still_correct_in_subprocess(gpu):
    """TCR_GEMM_2CTA=0 forces the single-CTA 128x128 tcgen05 kernel (the dispatch reads the variable once,
    so the check runs in a fresh interpreter)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from tenncor_b200 import cabi\n"
        "from tests.test_gemm_tc_gpu import run_gemm\n"
        "cabi.init(0)\n"
        "rng = np.random.default_rng(0)\n"
        "for (M, N, K, ta, tb) in [(256, 384, 96, 0, 0), (300, 200, 100, 1, 0), (512, 256, 4096, 1, 1), (1024, 784, 8192, 1, 0)]:\n"
        "    A = rng.uniform(-1, 1, (M, K)).astype(np.float32); B = rng.uniform(-1, 1, (K, N)).astype(np.float32)\n"
        "    want = A.astype(np.float64) @ B.astype(np.float64)\n"
        "    S = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)\n"
        "    for prec, tol in ((1, 2.0 ** -10), (2, 2.0 ** -19 + K * 2.0 ** -23)):\n"
        "        got = run_gemm(cabi, A, B, ta, tb, prec)\n"
        "        assert np.all(np.abs(got - want) <= S * tol), (M, N, K, prec)\n"
        "print('OK')\n" % root)
    env = dict(os.environ, TCR_GEMM_2CTA="0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("precision", [1, 2], ids=["tf32", "3xtf32"])
@pytest.mark.parametrize("M,N,K,ta,tb", [(5120, 1024, 784, 0, 0), (4864, 1000, 520, 1, 0), (19200, 260, 300, 0, 1), (2560, 7680, 256, 1, 1)],
                         ids=["80-tiles", "ragged-n", "ragged-mnk", "300-tiles"])
def test_stream_k_tiles_shared_by_two_pairs(gpu, M, N, K, ta, tb, precision):
    """More 256 x 256 tiles than CTA pairs and not a multiple of them: the pairs walk equal (tile, k-block) ranges and a tile's
    accumulator is the sum of two pairs' partials (csrc/gemm_tc2.cu). Same bounds as the un-split product, bit-identical on
    repeats (fixed order own + partial, flags reset for the next launch), bias + activation applied once by the owner."""
    rng = np.random.default_rng(M + N + K)
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64)
    S = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    bound = S * (2.0 ** -10 if precision == 1 else (2.0 ** -19 + K * 2.0 ** -23))
    got = run_gemm(gpu, A, B, ta, tb, precision)
    assert np.all(np.abs(got - want) <= bound + 1e-30)
    for _ in range(2):
        np.testing.assert_array_equal(run_gemm(gpu, A, B, ta, tb, precision), got)
    bias = rng.uniform(-1, 1, N).astype(np.float32)
    got = run_gemm(gpu, A, B, ta, tb, 2, bias=bias, epi=gpu.EPI_BIAS_N, act=gpu.OP["SIGMOID"])
    assert np.all(np.abs(got - 1 / (1 + np.exp(-(want + bias)))) <= 0.25 * S * (2.0 ** -19 + K * 2.0 ** -23) + 2e-6)  # |sigmoid'| <= 1/4
