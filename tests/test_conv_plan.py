"""Host-side lowering of the conv2d composite (cfg/tenncor/nn.yml:48-98): the planner must turn
PERMUTE(CONV(PAD(img), REVERSE(kernel))) and its kernel gradient into patch-gather + GEMM steps, and
the descriptor algebra those steps use (im2col row / window enumeration, GEMM operand strides) must
reproduce the oracle's evaluation of the same functor graph. No device needed: `describe_plan` stops
after lowering, and the descriptor algebra is replayed in numpy."""
import numpy as np
import pytest

import tenncor_b200 as tc
from oracle import tcr_oracle as orc
from tenncor_b200 import configs


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def _conv_graph(rng, inc, outc, W, H, B, kw, kh, bias=True, pads=None, dtype=np.float32):
    img = tc.variable(rng.random((B, H, W, inc)).astype(dtype), "img")
    ker = tc.variable(rng.random((kh, kw, inc, outc)).astype(dtype), "ker")
    b = tc.variable(rng.random((outc,)).astype(dtype), "bias")
    kwargs = {} if pads is None else {"zero_paddings": pads}
    out = tc.api.nn.conv2d(img, ker, b, **kwargs) if bias else tc.api.nn.conv2d(img, ker, **kwargs)
    return img, ker, b, out


def test_forward_and_kernel_gradient_are_lowered_to_gemms():
    rng = np.random.default_rng(0)
    img, ker, b, out = _conv_graph(rng, 3, 8, 10, 9, 4, 3, 2)
    loss = tc.api.reduce_sum(tc.api.square(tc.api.sigmoid(out)))
    gk, gb = tc.derive(loss, [ker, b])
    fwd = tc.describe_plan([out])
    assert len(fwd) == 1 and fwd[0].startswith("CONV2D im2col+GEMM+bias m256 n8 k18"), fwd
    steps = tc.describe_plan([loss, gk, gb])
    assert any(s.startswith("CONV2D im2col+GEMM+bias+act m256 n8 k18") for s in steps), steps
    assert any(s.startswith("CONV2D-dK im2col+GEMM m18 n8 k256") for s in steps), steps
    # nothing of the composite is left to the generic kernels
    assert not any(s.split()[0] in ("CONV", "PAD", "PERMUTE", "REVERSE") for s in steps), steps


def test_zero_padding_keeps_the_spatial_pad_as_its_own_step():
    rng = np.random.default_rng(0)
    img, ker, b, out = _conv_graph(rng, 2, 4, 6, 5, 3, 3, 3, pads=((1, 1), (2, 0)))
    steps = tc.describe_plan([out])
    assert steps[0].startswith("PAD") and steps[1].startswith("CONV2D im2col+GEMM+bias m90 n4 k18"), steps


def test_non_float_and_plain_correlations_are_left_alone():
    rng = np.random.default_rng(0)
    img, ker, b, out = _conv_graph(rng, 2, 4, 6, 5, 3, 3, 3, dtype=np.float64)
    assert any(s.startswith("CONV ") for s in tc.describe_plan([out]))
    image = tc.variable(rng.random((6, 7)).astype(np.float32), "image")
    kern = tc.variable(rng.random((2, 3)).astype(np.float32), "kern")
    plain = tc.api.convolution(image, kern, [0, 1])
    assert tc.describe_plan([plain])[0].startswith("CONV "), tc.describe_plan([plain])


def test_cnn_training_plan():
    cfg = configs.cnn()
    steps = tc.describe_plan([cfg.train])
    assert sum(s.startswith("CONV2D im2col+GEMM+bias+act") for s in steps) == 4  # two layers, before and after the update
    assert sum(s.startswith("CONV2D-dK") for s in steps) == 2
    assert sum(s.startswith("CONV2D-dX GEMM+col2im m192 n72 k4") for s in steps) == 1  # image gradient of layer 2
    # nothing of the composites is left to the generic kernels
    assert not any(s.split()[0] in ("CONV", "PAD", "PERMUTE", "REVERSE", "SLICE") for s in steps), steps


def test_image_gradient_falls_back_when_the_whole_padded_output_is_read():
    """the fused step computes only the slice the PAD rule keeps; a reader of the full tensor keeps the generic CONV"""
    rng = np.random.default_rng(0)
    img, ker, b, out = _conv_graph(rng, 3, 4, 6, 5, 2, 3, 2)
    loss = tc.api.reduce_sum(tc.api.square(out))
    gi = tc.derive(loss, [img])[0]
    steps = tc.describe_plan([gi])
    assert any(s.startswith("CONV2D-dX") for s in steps), steps
    # walk down to the CONV under the gradient's SLICE and expose it as a second target
    node = gi
    while node.opname() != "CONV":
        node = node.args()[0]
    steps = tc.describe_plan([gi, node])
    assert not any(s.startswith("CONV2D-dX") for s in steps) and any(s.startswith("CONV ") for s in steps), steps


def _im2col(flat, img_shape, win):
    """numpy replay of tcr_im2col's enumeration (tenncor_b200/csrc/im2col.cu)."""
    pos = [s - w + 1 for s, w in zip(img_shape, win)]
    strides = np.cumprod([1] + list(img_shape[:-1]))
    rows, k = int(np.prod(pos)), int(np.prod(win))
    pos_idx = np.stack(np.unravel_index(np.arange(rows), pos, order="F"), 1)
    win_idx = np.stack(np.unravel_index(np.arange(k), win, order="F"), 1)
    return flat[(pos_idx @ strides)[:, None] + (win_idx @ strides)[None, :]]


def _col2im(cols, img_shape, win):
    """numpy replay of tcr_col2im (tenncor_b200/csrc/im2col.cu): the adjoint of `_im2col`."""
    pos = [s - w + 1 for s, w in zip(img_shape, win)]
    strides = np.cumprod([1] + list(img_shape[:-1]))
    rows, k = int(np.prod(pos)), int(np.prod(win))
    pos_idx = np.stack(np.unravel_index(np.arange(rows), pos, order="F"), 1)
    win_idx = np.stack(np.unravel_index(np.arange(k), win, order="F"), 1)
    img = np.zeros(int(np.prod(img_shape)))
    np.add.at(img, (pos_idx @ strides)[:, None] + (win_idx @ strides)[None, :], cols[:, :k])
    return img


@pytest.mark.parametrize("inc,outc,W,H,B,kw,kh", [(3, 8, 10, 9, 4, 3, 2), (1, 3, 5, 5, 1, 2, 2), (5, 2, 7, 4, 3, 1, 3), (2, 5, 6, 6, 2, 6, 6)])
def test_image_gradient_algebra_matches_oracle(inc, outc, W, H, B, kw, kh):
    rng = np.random.default_rng(2)
    # small integers in float32: the fused lowering is FLOAT-only and every sum stays exact
    img_np, ker_np = (rng.integers(-3, 4, s).astype(np.float32) for s in ((B, H, W, inc), (kh, kw, inc, outc)))
    g_np = rng.integers(-3, 4, (B, H - kh + 1, W - kw + 1, outc)).astype(np.float32)
    img, ker, g = (tc.variable(a, n) for a, n in ((img_np, "img"), (ker_np, "ker"), (g_np, "g")))
    gi = tc.derive(tc.api.reduce_sum(tc.api.nn.conv2d(img, ker) * g), [img])[0]
    assert any(s.startswith("CONV2D-dX") for s in tc.describe_plan([gi]))
    tape = tc.dump_graph([gi])
    want = np.asarray(orc.eval_tape(tape)[tc.dump_ids([gi], tape)[gi]]).reshape(-1)
    k = inc * kw * kh
    cols = g_np.reshape(-1, outc).astype(np.float64) @ ker_np.reshape(k, outc).T  # a = sup [rows, out]; b(kk = o, n = (c,i,j)) = kernel[o + out * n]
    got = _col2im(cols, [inc, W, H, B, 1, 1, 1, 1], [inc, kw, kh, 1, 1, 1, 1, 1])
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("inc,outc,W,H,B,kw,kh", [(3, 8, 10, 9, 4, 3, 2), (1, 3, 5, 5, 1, 2, 2), (5, 2, 7, 4, 3, 1, 3), (2, 5, 6, 6, 2, 6, 6)])
def test_descriptor_algebra_matches_oracle(inc, outc, W, H, B, kw, kh):
    rng = np.random.default_rng(1)
    img_np = rng.random((B, H, W, inc))
    ker_np = rng.random((kh, kw, inc, outc))
    b_np = rng.random((outc,))
    g_np = rng.random((B, H - kh + 1, W - kw + 1, outc))
    img, ker, b, g = (tc.variable(a, n) for a, n in ((img_np, "img"), (ker_np, "ker"), (b_np, "bias"), (g_np, "g")))
    out = tc.api.nn.conv2d(img, ker, b)
    gk = tc.derive(tc.api.reduce_sum(out * g), [ker])[0]
    tape = tc.dump_graph([out, gk])
    ids = tc.dump_ids([out, gk], tape)
    vals = orc.eval_tape(tape)
    cols = _im2col(img_np.reshape(-1), [inc, W, H, B, 1, 1, 1, 1], [inc, kw, kh, 1, 1, 1, 1, 1])
    k = inc * kw * kh
    fwd = cols @ ker_np.reshape(k, outc) + b_np[None, :]           # b_sk = n, b_sn = 1; c row-major [m, n]
    np.testing.assert_allclose(fwd.reshape(-1), np.asarray(vals[ids[out]]).reshape(-1), rtol=1e-12)
    dk = cols.T @ g_np.reshape(-1, outc)                            # cols^T . sup in the kernel's own layout
    np.testing.assert_allclose(dk.reshape(-1), np.asarray(vals[ids[gk]]).reshape(-1), rtol=1e-12)
