"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle is too slow
there): linearity, involutions, checksums of checksums, batch-permutation invariance, determinism.
Sizes: SURVEY.md §8(d) — M-ew 2^26, M-red [4096,16384], M-lay 8192^2, C3 784-1024-10 at batch 8192,
C4 LSTM N=128 H=1024 seq=128 batch 64. Bit-exact where the ops are data movement / integer, stated
fp32 tolerances elsewhere."""
import ctypes as C

import numpy as np
import pytest

import tenncor_b200 as tc
from tenncor_b200 import configs

pytestmark = pytest.mark.gpu

P = lambda b: C.c_void_p(b.ptr)  # noqa: E731


def test_elementwise_identities_at_2_pow_26(gpu):
    lib, F, n = gpu.lib(), gpu.FLOAT, 1 << 26
    rng = np.random.default_rng(5)
    x = rng.uniform(-4, 4, n).astype(np.float32)
    dx, a, b = gpu.to_device(x), gpu.empty(n, np.float32), gpu.empty(n, np.float32)
    # sigmoid(x) + sigmoid(-x) == 1
    gpu.check(lib.tcr_unary(gpu.OP["SIGMOID"], P(dx), P(a), C.c_int64(n), F))
    gpu.check(lib.tcr_unary(gpu.OP["NEG"], P(dx), P(b), C.c_int64(n), F))
    gpu.check(lib.tcr_unary(gpu.OP["SIGMOID"], P(b), P(b), C.c_int64(n), F))
    gpu.check(lib.tcr_binary(gpu.OP["ADD"], P(a), P(b), P(a), C.c_int64(n), F))
    s = gpu.to_host(a, n, np.float32)
    assert np.max(np.abs(s - 1.0)) <= 1e-6
    # log(exp(x)) == x within 1e-5 relative to max(|x|, 1)
    gpu.check(lib.tcr_unary(gpu.OP["EXP"], P(dx), P(a), C.c_int64(n), F))
    gpu.check(lib.tcr_unary(gpu.OP["LOG"], P(a), P(a), C.c_int64(n), F))
    back = gpu.to_host(a, n, np.float32)
    assert np.max(np.abs(back - x) / np.maximum(np.abs(x), 1.0)) <= 1e-5
    # the fused chain equals the same three launches, bit for bit (same per-element arithmetic)
    y = rng.uniform(-1, 1, n).astype(np.float32)
    z = rng.uniform(-1, 1, n).astype(np.float32)
    dy, dz = gpu.to_device(y), gpu.to_device(z)
    prog = gpu.make_program(F, (n, 1, 1), [(dx.ptr, F, (0, 0, 0)), (dy.ptr, F, (0, 0, 0)), (dz.ptr, F, (0, 0, 0))], [(a.ptr, F, 0)],
                            [(gpu.OP["MUL"], 0, 0, 1), (gpu.OP["ADD"], 0, 0, 2), (gpu.OP["TANH"], 0, 0)])
    gpu.check(lib.tcr_elementwise(C.byref(prog)))
    gpu.check(lib.tcr_binary(gpu.OP["MUL"], P(dx), P(dy), P(b), C.c_int64(n), F))
    gpu.check(lib.tcr_binary(gpu.OP["ADD"], P(b), P(dz), P(b), C.c_int64(n), F))
    gpu.check(lib.tcr_unary(gpu.OP["TANH"], P(b), P(b), C.c_int64(n), F))
    fused, unfused = gpu.to_host(a, n, np.float32), gpu.to_host(b, n, np.float32)
    # (a*b)+c may contract to one FMA in either kernel: allow the last-bit difference that implies
    np.testing.assert_allclose(fused, unfused, rtol=2e-6, atol=1e-7)


def test_reduce_checksum_of_checksums(gpu):
    """sum over dim0 then over the result == sum over dim1 then over the result == full sum."""
    lib, F = gpu.lib(), gpu.FLOAT
    R0, R1 = 4096, 16384
    rng = np.random.default_rng(6)
    x = rng.uniform(0, 1, R0 * R1).astype(np.float32)
    dx = gpu.to_device(x)
    part0, part1, tot = gpu.empty(R1, np.float32), gpu.empty(R0, np.float32), gpu.empty(1, np.float32)
    shp = gpu.shape8([R0, R1])
    op = gpu.OP["REDUCE_SUM"]
    gpu.check(lib.tcr_reduce(op, P(dx), P(part0), shp, C.c_uint32(1), F))
    gpu.check(lib.tcr_reduce(op, P(dx), P(part1), shp, C.c_uint32(2), F))
    gpu.check(lib.tcr_reduce(op, P(dx), P(tot), shp, C.c_uint32(3), F))
    full = float(gpu.to_host(tot, 1, np.float32)[0])
    want = float(x.astype(np.float64).sum())
    assert abs(full - want) <= 1e-5 * want
    for part, n in ((part0, R1), (part1, R0)):
        gpu.check(lib.tcr_reduce(op, P(part), P(tot), gpu.shape8([n]), C.c_uint32(1), F))
        assert abs(float(gpu.to_host(tot, 1, np.float32)[0]) - want) <= 1e-5 * want
    # max over either order is exact
    opm = gpu.OP["REDUCE_MAX"]
    gpu.check(lib.tcr_reduce(opm, P(dx), P(part0), shp, C.c_uint32(1), F))
    gpu.check(lib.tcr_reduce(opm, P(part0), P(tot), gpu.shape8([R1]), C.c_uint32(1), F))
    assert gpu.to_host(tot, 1, np.float32)[0] == x.max()
    # argmax along dim0 points at the row maximum (bit-exact index)
    idx = gpu.empty(R1, np.float32)
    gpu.check(lib.tcr_argmax(P(dx), P(idx), shp, 0, F))
    got = gpu.to_host(idx, R1, np.float32).astype(np.int64)
    rows = x.reshape(R1, R0)
    np.testing.assert_array_equal(got, rows.argmax(axis=1))


def test_layout_involutions_bit_exact(gpu):
    lib = gpu.lib()
    side = 8192
    n = side * side
    x = np.arange(n, dtype=np.float32)  # iota: every element distinguishable
    dx, a, b = gpu.to_device(x), gpu.empty(n, np.float32), gpu.empty(n, np.float32)
    order = (C.c_int32 * 8)(1, 0, 2, 3, 4, 5, 6, 7)
    gpu.check(lib.tcr_permute(P(dx), P(a), gpu.shape8([side, side]), order, 4))
    gpu.check(lib.tcr_permute(P(a), P(b), gpu.shape8([side, side]), order, 4))
    np.testing.assert_array_equal(gpu.to_host(b, n, np.float32), x)
    once = gpu.to_host(a, n, np.float32).reshape(side, side)
    np.testing.assert_array_equal(once[5, :64], x.reshape(side, side)[:64, 5])
    # slice(pad(x)) == x
    s3 = [1024, 128, 64]
    m = 1024 * 128 * 64
    lo = (C.c_int64 * 8)(0, 16, 0, 0, 0, 0, 0, 0)
    gpu.check(lib.tcr_pad(P(dx), P(a), gpu.shape8(s3), lo, lo, 4))
    offs = (C.c_int64 * 8)(0, 16, 0, 0, 0, 0, 0, 0)
    exts = (C.c_int64 * 8)(1024, 128, 64, 1, 1, 1, 1, 1)
    gpu.check(lib.tcr_slice(P(a), P(b), gpu.shape8([1024, 160, 64]), offs, exts, 4))
    np.testing.assert_array_equal(gpu.to_host(b, m, np.float32), x[:m])
    # reverse twice
    gpu.check(lib.tcr_reverse(P(dx), P(a), gpu.shape8(s3), C.c_uint32(0b10), 4))  # reverse rank 1
    gpu.check(lib.tcr_reverse(P(a), P(b), gpu.shape8(s3), C.c_uint32(0b10), 4))
    np.testing.assert_array_equal(gpu.to_host(b, m, np.float32), x[:m])


def test_gemm_linearity_at_c3_shapes(gpu):
    """A (B1 + B2) == A B1 + A B2 for the layer-0 forward product (8192 x 784)(784 x 1024), 3xTF32."""
    lib, F = gpu.lib(), gpu.FLOAT
    M, K, N = 8192, 784, 1024
    rng = np.random.default_rng(7)
    A = rng.uniform(-1, 1, M * K).astype(np.float32)
    B1, B2 = rng.uniform(-1, 1, K * N).astype(np.float32), rng.uniform(-1, 1, K * N).astype(np.float32)
    dA, dB1, dB2, dBs = gpu.to_device(A), gpu.to_device(B1), gpu.to_device(B2), gpu.empty(K * N, np.float32)
    c1, c2, cs = gpu.empty(M * N, np.float32), gpu.empty(M * N, np.float32), gpu.empty(M * N, np.float32)
    gpu.check(lib.tcr_binary(gpu.OP["ADD"], P(dB1), P(dB2), P(dBs), C.c_int64(K * N), F))
    d = gpu.GemmDesc(m=M, n=N, k=K, batch=1, a_sm=K, a_sk=1, b_sk=N, b_sn=1, c_sm=N, c_sn=1, dtype=F, precision=gpu.GEMM_3XTF32)
    for b, c in ((dB1, c1), (dB2, c2), (dBs, cs)):
        gpu.check(lib.tcr_gemm(P(dA), P(b), P(c), C.byref(d)))
    gpu.check(lib.tcr_binary(gpu.OP["ADD"], P(c1), P(c2), P(c1), C.c_int64(M * N), F))
    lhs, rhs = gpu.to_host(cs, M * N, np.float32), gpu.to_host(c1, M * N, np.float32)
    # |sum| <= K; fp32 accumulation over K = 784 terms plus the rounding of B1 + B2
    assert np.max(np.abs(lhs - rhs)) <= 784 * 2.0 ** -22
    # spot rows against float64
    rows = [0, 4097, 8191]
    want = A.reshape(M, K)[rows].astype(np.float64) @ (B1.astype(np.float64) + B2.astype(np.float64)).reshape(K, N)
    assert np.max(np.abs(lhs.reshape(M, N)[rows] - want)) <= 784 * 2.0 ** -21


def _mlp_steps(batch_x, batch_y, steps=2):
    cfg = configs.mlp(784, 1024, 10, batch_x.shape[0], seed=2)
    losses = []
    for _ in range(steps):
        cfg.feeds["x"].assign(batch_x)
        cfg.feeds["y"].assign(batch_y)
        losses.append(float(cfg.train.get()))
    return losses, [np.array(v.data(), copy=True) for v in cfg.variables]


def test_c3_training_is_batch_permutation_invariant_and_deterministic(gpu):
    """The loss and every gradient sum over the batch: permuting the samples changes nothing but the
    fp32 summation order. Same batch twice -> bit-identical (split-K and reductions are deterministic)."""
    tc.set_evaluator("plan")
    tc.set_matmul_precision("3xtf32")
    rng = np.random.default_rng(2)
    B = 8192
    x = rng.random((B, 784), dtype=np.float32)
    y = np.zeros((B, 10), dtype=np.float32)
    y[np.arange(B), rng.integers(0, 10, B)] = 1
    l0, w0 = _mlp_steps(x, y)
    l1, w1 = _mlp_steps(x, y)
    assert l0 == l1
    for a, b in zip(w0, w1):
        np.testing.assert_array_equal(a, b)
    perm = rng.permutation(B)
    l2, w2 = _mlp_steps(x[perm], y[perm])
    assert max(abs(a - b) / abs(a) for a, b in zip(l0, l2)) <= 1e-5
    for a, b in zip(w0, w2):
        assert np.max(np.abs(a - b)) <= 1e-5 * (np.max(np.abs(a)) + 1e-30)
    assert l0[1] < l0[0]  # SGD on the same batch lowers the error


def test_c4_lstm_step_is_finite_and_deterministic(gpu):
    tc.set_evaluator("plan")
    rng = np.random.default_rng(3)
    runs = []
    for _ in range(2):
        cfg = configs.recurrent("lstm", vocab=128, hidden=1024, seq=128, batch=64, seed=3)
        r = np.random.default_rng(3)
        x, y = configs.recurrent_batch(r, cfg.feeds, 128)
        cfg.feeds["x"].assign(x)
        cfg.feeds["y"].assign(y)
        loss = float(cfg.train.get())
        runs.append((loss, np.array(cfg.variables[0].data(), copy=True)))
        del cfg
    assert np.isfinite(runs[0][0]) and runs[0][0] > 0  # summed NLL after the first adagrad step
    assert np.isfinite(runs[0][1]).all()
    assert runs[0][0] == runs[1][0]
    np.testing.assert_array_equal(runs[0][1], runs[1][1])
    assert rng is not None


def test_conv2d_adjoint_identities_at_bench_size(gpu):
    """conv2d 3x3, 32 -> 64 channels, 34x34, batch 64 (bench.py --workload conv): the oracle's padded-rank formulation needs
    minutes there, so forward, kernel gradient and image gradient — three different lowerings (patch gather + GEMM, gather +
    transposed split-K GEMM, GEMM + patch scatter) — are tied together by the bilinear form they all evaluate:
        <conv(x, k), g>  ==  <k, dL/dk>  ==  <x, dL/dx>      for L = sum(conv(x, k) * g),
    and the planned evaluator must agree with itself under linearity in x. fp32 / 3xTF32: 1e-4 relative."""
    tc.set_evaluator("plan")
    tc.set_matmul_precision("3xtf32")
    rng = np.random.default_rng(21)
    B, H, W, cin, cout = 64, 34, 34, 32, 64
    x_np = rng.uniform(-1, 1, (B, H, W, cin)).astype(np.float32)
    k_np = rng.uniform(-1, 1, (3, 3, cin, cout)).astype(np.float32)
    g_np = rng.uniform(-1, 1, (B, H - 2, W - 2, cout)).astype(np.float32)
    x, k, g = tc.variable(x_np, "x"), tc.variable(k_np, "k"), tc.variable(g_np, "g")
    out = tc.api.nn.conv2d(x, k)
    loss = tc.api.reduce_sum(out * g)
    dk, dx = tc.derive(loss, [k, x])
    steps = tc.describe_plan([out, dk, dx])
    assert any(s.startswith("CONV2D im2col+GEMM") for s in steps) and any(s.startswith("CONV2D-dK") for s in steps)
    assert any(s.startswith("CONV2D-dX") for s in steps) and not any(s.startswith("CONV ") for s in steps), steps
    out_v, dk_v, dx_v = (a.astype(np.float64) for a in tc.run([out, dk, dx]))
    form = float(np.sum(out_v * g_np))
    scale = float(np.sum(np.abs(out_v) * np.abs(g_np)))
    assert abs(float(np.sum(dk_v * k_np)) - form) <= 1e-4 * scale
    assert abs(float(np.sum(dx_v * x_np)) - form) <= 1e-4 * scale
    # spot values against a direct evaluation of the definition at a few output positions
    for (b, y, xx, o) in [(0, 0, 0, 0), (63, 31, 31, 63), (17, 5, 29, 40), (40, 30, 2, 7)]:
        want = float(np.sum(x_np[b, y:y + 3, xx:xx + 3, :].astype(np.float64) * k_np[:, :, :, o]))
        assert abs(out_v[b, y, xx, o] - want) <= 1e-4 * float(np.sum(np.abs(x_np[b, y:y + 3, xx:xx + 3, :]) * np.abs(k_np[:, :, :, o])))
    # linearity in the image: conv(2x, k) == 2 conv(x, k) bit for bit (a power-of-two scale commutes with every rounding)
    x.assign(2 * x_np)
    np.testing.assert_array_equal(out.get().astype(np.float64), 2 * out_v)
