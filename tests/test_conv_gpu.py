"""conv2d (cfg/tenncor/nn.yml:48-98, layer.yml:130-160) on the GPU against the CPU oracle evaluating
the SAME dumped functor graph: the planned evaluator runs it as patch gather + tcgen05 GEMM
(planner.cpp fuse_convs), the node evaluator through the generic CONV kernel like the reference.

Tolerance: the GEMM runs in 3xTF32: 1e-4 relative to the tensor's magnitude, the same
bar as the dense training tests. tcr_im2col is pure data movement: bit-exact."""
import ctypes as C

import numpy as np
import pytest

import tenncor_b200 as tc
from oracle import tcr_oracle as orc
from tenncor_b200 import configs
from tests.test_conv_plan import _col2im, _im2col
from tests.test_train_gpu import OracleSession, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("img_shape,win", [
    ([3, 10, 9, 4], [3, 3, 2, 1]),        # conv2d: k = 18 -> pitch 20 (zero fill)
    ([16, 18, 18, 8], [16, 3, 3, 1]),     # k = 144
    ([1, 7, 5, 1], [1, 2, 2, 1]),
    ([5, 6, 7, 2], [2, 1, 3, 2]),         # windows that do not cover rank 0: no run merging
    ([4, 4, 4, 4], [4, 4, 4, 4]),         # one position
    ([257, 3, 2, 3], [1, 2, 1, 1]),
])
def test_im2col_bit_exact(gpu, img_shape, win):
    rng = np.random.default_rng(12)
    s8 = list(img_shape) + [1] * (8 - len(img_shape))
    w8 = list(win) + [1] * (8 - len(win))
    img = rng.standard_normal(int(np.prod(s8))).astype(np.float32)
    want = _im2col(img, s8, w8)
    rows, k = want.shape
    pitch = (k + 3) // 4 * 4
    dimg = gpu.to_device(img)
    dcols = gpu.empty(rows * pitch, np.float32)
    gpu.check(gpu.lib().tcr_memset(C.c_void_p(dcols.ptr), 0xFF, C.c_size_t(rows * pitch * 4)))
    gpu.check(gpu.lib().tcr_im2col(C.c_void_p(dimg.ptr), C.c_void_p(dcols.ptr), gpu.shape8(s8), gpu.shape8(w8), C.c_int64(pitch), 4))
    got = gpu.to_host(dcols, rows * pitch, np.float32).reshape(rows, pitch)
    np.testing.assert_array_equal(got[:, :k].view(np.uint32), want.view(np.uint32))
    assert not got[:, k:].any()  # the tail of each row is zero-filled


@pytest.mark.parametrize("img_shape,win", [([3, 10, 9, 4], [3, 3, 2, 1]), ([16, 18, 18, 8], [16, 3, 3, 1]), ([1, 7, 5, 1], [1, 2, 2, 1]),
                                           ([5, 6, 7, 2], [2, 1, 3, 2]), ([4, 4, 4, 4], [4, 4, 4, 4]), ([8, 9, 1, 3], [8, 1, 1, 1])])
def test_col2im_is_the_adjoint_of_im2col(gpu, img_shape, win):
    rng = np.random.default_rng(14)
    s8 = list(img_shape) + [1] * (8 - len(img_shape))
    w8 = list(win) + [1] * (8 - len(win))
    rows = int(np.prod([a - b + 1 for a, b in zip(s8, w8)]))
    k = int(np.prod(w8))
    pitch = (k + 3) // 4 * 4
    cols = rng.integers(-4, 5, (rows, pitch)).astype(np.float32)  # small integers: every sum is exact in fp32
    want = _col2im(cols.astype(np.float64), s8, w8)
    dcols = gpu.to_device(cols)
    n_img = int(np.prod(s8))
    dimg = gpu.empty(n_img, np.float32)
    gpu.check(gpu.lib().tcr_memset(C.c_void_p(dimg.ptr), 0xFF, C.c_size_t(n_img * 4)))
    gpu.check(gpu.lib().tcr_col2im(C.c_void_p(dcols.ptr), C.c_void_p(dimg.ptr), gpu.shape8(s8), gpu.shape8(w8), C.c_int64(pitch), gpu.FLOAT))
    np.testing.assert_array_equal(gpu.to_host(dimg, n_img, np.float32), want.astype(np.float32))
    # <im2col(x), y> == <x, col2im(y)>
    x = rng.integers(-4, 5, n_img).astype(np.float64)
    assert np.sum(_im2col(x, s8, w8) * cols[:, :k]) == np.sum(x * want)


def test_im2col_rejects_bad_arguments(gpu):
    dimg = gpu.to_device(np.zeros(64, np.float32))
    dcols = gpu.empty(64, np.float32)
    lib = gpu.lib()
    assert lib.tcr_im2col(C.c_void_p(dimg.ptr), C.c_void_p(dcols.ptr), gpu.shape8([4, 4]), gpu.shape8([5, 1]), C.c_int64(8), 4) != 0
    assert b"does not fit" in lib.tcr_last_error()
    assert lib.tcr_im2col(C.c_void_p(dimg.ptr), C.c_void_p(dcols.ptr), gpu.shape8([4, 4]), gpu.shape8([2, 2]), C.c_int64(6), 4) != 0
    assert lib.tcr_im2col(C.c_void_p(dimg.ptr), C.c_void_p(dcols.ptr), gpu.shape8([4, 4]), gpu.shape8([2, 2]), C.c_int64(4), 8) != 0


CONV_CASES = [
    # inc, outc, W, H, B, kw, kh, bias, pads, image gradient too (its oracle is (2*outc-1) x as expensive)
    (3, 8, 10, 9, 4, 3, 2, True, None, True),
    (1, 3, 5, 5, 1, 2, 2, False, None, True),
    (5, 2, 7, 4, 3, 1, 3, True, ((1, 1), (2, 0)), True),
    (7, 5, 9, 8, 3, 2, 3, True, None, True),           # k = 42: padded row pitch, n not a multiple of 4
    (16, 20, 14, 12, 6, 3, 3, True, None, True),       # image-gradient GEMM (720 x 144 x 20) on the tcgen05 kernel
    (16, 32, 18, 18, 8, 3, 3, True, None, False),      # forward and kernel-gradient GEMMs large enough for the tcgen05 kernels
    (8, 32, 34, 20, 4, 3, 3, True, ((1, 1), (1, 1)), False),
]


@pytest.mark.parametrize("evaluator", ["node", "plan"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c[:7]) for c in CONV_CASES])
def test_conv2d_forward_and_gradients_match_oracle(gpu, case, evaluator):
    inc, outc, W, H, B, kw, kh, bias, pads, with_img = case
    tc.set_evaluator(evaluator)
    tc.set_matmul_precision("3xtf32")
    try:
        rng = np.random.default_rng(13)
        img = tc.variable(rng.uniform(-1, 1, (B, H, W, inc)).astype(np.float32), "img")
        ker = tc.variable(rng.uniform(-1, 1, (kh, kw, inc, outc)).astype(np.float32), "ker")
        b = tc.variable(rng.uniform(-1, 1, (outc,)).astype(np.float32), "bias")
        kwargs = {} if pads is None else {"zero_paddings": pads}
        out = tc.api.nn.conv2d(img, ker, b, **kwargs) if bias else tc.api.nn.conv2d(img, ker, **kwargs)
        act = tc.api.tanh(out)
        loss = tc.api.reduce_sum(tc.api.square(act))
        wrt = [ker] + ([img] if with_img else []) + ([b] if bias else [])
        roots = [out, act, loss] + list(tc.derive(loss, wrt))
        sess = OracleSession(roots)
        want = sess.run()
        got = tc.run(roots)
        for r, g, w in zip(roots, got, want):
            assert rel_err(g, w) < 1e-4, (r.opname(), r.shape(), rel_err(g, w))
    finally:
        tc.set_evaluator("plan")


@pytest.mark.parametrize("evaluator", ["node", "plan"])
@pytest.mark.parametrize("dims", [dict(), dict(in_ch=4, mid_ch=16, out_ch=16, width=12, height=12, nbatch=8)])
def test_cnn_training_matches_oracle(gpu, dims, evaluator):
    tc.set_evaluator(evaluator)
    tc.set_matmul_precision("3xtf32")
    try:
        cfg = configs.cnn(**dims)
        sess = OracleSession([cfg.train])
        rng = np.random.default_rng(6)
        for step in range(3):
            x, y = configs.cnn_batch(rng, cfg.feeds)
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


def test_conv_plan_is_fused_and_graph_replayed(gpu):
    tc.set_evaluator("plan")
    cfg = configs.cnn()
    rng = np.random.default_rng(6)
    for _ in range(3):
        x, y = configs.cnn_batch(rng, cfg.feeds)
        cfg.feeds["x"].assign(x)
        cfg.feeds["y"].assign(y)
        cfg.train.get()
    stats = tc.plan_stats()
    assert stats["graph"], stats
    names = [t["what"] for t in tc.profile_plan(1)]
    assert sum(n.startswith("CONV2D im2col+GEMM") for n in names) == 4, names
    assert sum(n.startswith("CONV2D-dK") for n in names) == 2, names
    assert sum(n.startswith("CONV2D-dX") for n in names) == 1, names


@pytest.mark.parametrize("precision", [1, 2], ids=["tf32", "3xtf32"])
@pytest.mark.parametrize("img_shape,win,cout,act", [
    ([32, 10, 9, 4], [32, 3, 3, 1], 24, "TANH"),     # 224 positions: the second row tile is zero-filled past row 96; 24 of 128 columns
    ([64, 34, 34, 8], [64, 3, 3, 1], 64, None),      # the conv bench layer at batch 8: 8192 positions, K = 576
    ([8, 12, 11, 3], [8, 4, 2, 1], 40, "SIGMOID"),   # run = 32: one k-block per window row
    ([32, 6, 5, 2, 3], [32, 1, 1, 1, 1], 16, None),  # 1 x 1 window, two batch ranks
], ids=["ragged", "bench", "run32", "pointwise"])
def test_gemm_patches_matches_the_materialised_patch_matrix(gpu, img_shape, win, cout, act, precision):
    """tcr_gemm_patches (implicit GEMM: patch tiles gathered by the kernel's producer warp) against float64 over the numpy patch matrix
    (bounds of tests/test_gemm_tc_gpu.py) and against tcr_im2col + tcr_gemm on the device."""
    rng = np.random.default_rng(int(np.prod(img_shape)) + cout)
    lib = gpu.lib()
    s8 = list(img_shape) + [1] * (8 - len(img_shape))
    w8 = list(win) + [1] * (8 - len(win))
    img = rng.uniform(-1, 1, int(np.prod(s8))).astype(np.float32)
    cols = _im2col(img, s8, w8)
    m, k = cols.shape
    kern = (rng.uniform(-1, 1, (k, cout)) * 0.2).astype(np.float32)
    bias = rng.uniform(-1, 1, cout).astype(np.float32)
    pre = cols.astype(np.float64) @ kern.astype(np.float64) + bias.astype(np.float64)
    S = np.abs(cols).astype(np.float64) @ np.abs(kern).astype(np.float64)
    want = np.tanh(pre) if act == "TANH" else 1 / (1 + np.exp(-pre)) if act == "SIGMOID" else pre
    tol = S * (2.0 ** -10 if precision == 1 else (2.0 ** -19 + k * 2.0 ** -23)) + 2e-6
    dimg, dk, db = gpu.to_device(img), gpu.to_device(kern), gpu.to_device(bias)
    out = gpu.to_device(np.full(m * cout, 9.0, np.float32))
    d = gpu.GemmDesc()
    d.m, d.n, d.k, d.batch = m, cout, k, 1
    d.a_sm, d.a_sk, d.b_sk, d.b_sn, d.c_sm, d.c_sn = k, 1, cout, 1, cout, 1
    d.dtype, d.precision, d.epilogue, d.activation, d.bias = gpu.FLOAT, precision, 1, (gpu.OP[act] if act else 0), db.ptr
    launches = lib.tcr_launch_count()
    gpu.check(lib.tcr_gemm_patches(C.c_void_p(dimg.ptr), C.c_void_p(dk.ptr), C.c_void_p(out.ptr), C.byref(d), gpu.shape8(s8), gpu.shape8(w8)))
    assert lib.tcr_launch_count() == launches + 1  # one kernel, no patch matrix
    got = gpu.to_host(out, m * cout, np.float32).reshape(m, cout).astype(np.float64)
    assert np.all(np.abs(got - want) <= tol), float(np.abs(got - want).max())
    # the two-call form on the device
    dcols = gpu.empty(m * k, np.float32)
    gpu.check(lib.tcr_im2col(C.c_void_p(dimg.ptr), C.c_void_p(dcols.ptr), gpu.shape8(s8), gpu.shape8(w8), C.c_int64(k), 4))
    out2 = gpu.empty(m * cout, np.float32)
    gpu.check(lib.tcr_gemm(C.c_void_p(dcols.ptr), C.c_void_p(dk.ptr), C.c_void_p(out2.ptr), C.byref(d)))
    two = gpu.to_host(out2, m * cout, np.float32).reshape(m, cout).astype(np.float64)
    assert np.all(np.abs(got - two) <= 2 * tol)


def test_gemm_patches_declines_views_it_cannot_gather(gpu):
    lib = gpu.lib()
    s8, w8 = [3, 10, 9, 4, 1, 1, 1, 1], [3, 3, 2, 1, 1, 1, 1, 1]  # kw * C = 9
    d = gpu.GemmDesc()
    d.m, d.n, d.k, d.batch = 8 * 8 * 4, 16, 18, 1
    d.b_sk, d.b_sn, d.c_sm, d.c_sn = 16, 1, 16, 1
    d.dtype, d.precision = gpu.FLOAT, 2
    buf = gpu.empty(1 << 16, np.float32)
    launches = lib.tcr_launch_count()
    rc = lib.tcr_gemm_patches(C.c_void_p(buf.ptr), C.c_void_p(buf.ptr), C.c_void_p(buf.ptr), C.byref(d), gpu.shape8(s8), gpu.shape8(w8))
    assert rc == 6 and lib.tcr_launch_count() == launches  # TCR_ERR_UNSUPPORTED, nothing launched
