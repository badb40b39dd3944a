"""tcr_elementwise_multi / tcr_elementwise_reduce through the C-ABI against numpy (float32 arithmetic, same operation order).

The multi form replaces a chain of elementwise functors whose intermediate results have several readers (one Eigen assignment
each in the reference, internal/eigen/device.hpp:555-570); the reduce form replaces elementwise + REDUCE_SUM over every rank +
scalar DIV (cfg/tenncor/loss.yml:21-39)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lstm_backward_programs(gpu, n, bufs, dims):
    """dh = a + b; dc = f * dcn + o * dh; dpre_g = (1 - g^2) * (i * dc); dpre_o = o (1 - o) * (c * dh)   (backprop.hpp:136-142)"""
    F, OP = gpu.FLOAT, gpu.OP
    P = lambda name: bufs[name].ptr  # noqa: E731
    full = (0, 0, 0)
    return [
        gpu.make_program(F, dims, [(P("a"), F, full), (P("b"), F, full)], [(P("dh"), F, 0)], [(OP["ADD"], 0, 0, 1)]),
        gpu.make_program(F, dims, [(P("f"), F, full), (P("dcn"), F, full), (P("o"), F, full), (P("dh"), F, full)], [(P("dc"), F, 0)],
                         [(OP["MUL"], 0, 0, 1), (OP["MUL"], 1, 2, 3), (OP["ADD"], 0, 0, 1)]),
        gpu.make_program(F, dims, [(P("g"), F, full), (P("i"), F, full), (P("dc"), F, full)], [(P("dg"), F, 0)],
                         [(gpu.EW_CONST, 3, 0, 0, 0, 1.0), (OP["SQUARE"], 0, 0), (OP["SUB"], 0, 3, 0), (OP["MUL"], 1, 1, 2), (OP["MUL"], 0, 0, 1)]),
        gpu.make_program(F, dims, [(P("o"), F, full), (P("c"), F, full), (P("dh"), F, full)], [(P("do"), F, 0)],
                         [(gpu.EW_CONST, 3, 0, 0, 0, 1.0), (OP["SUB"], 3, 3, 0), (OP["MUL"], 0, 0, 3), (OP["MUL"], 1, 1, 2), (OP["MUL"], 0, 0, 1)]),
    ]


@pytest.mark.parametrize("n", [65536, 4099, 3])
def test_multi_forwards_results_between_programs(gpu, n):
    rng = np.random.default_rng(n)
    host = {k: rng.uniform(-1, 1, n).astype(np.float32) for k in "abfogic"}
    host["dcn"] = rng.uniform(-1, 1, n).astype(np.float32)
    bufs = {k: gpu.to_device(v) for k, v in host.items()}
    for k in ("dh", "dc", "dg", "do"):
        bufs[k] = gpu.to_device(np.full(n, 7.0, np.float32))
    progs = _lstm_backward_programs(gpu, n, bufs, (n, 1, 1))
    arr = (gpu.EwProgram * len(progs))(*progs)
    keep = (C.c_uint8 * 4)(0, 1, 1, 1)  # dh is read only inside the launch
    gpu.check(gpu.lib().tcr_elementwise_multi(arr, len(progs), keep))
    one = np.float32(1)
    dh = host["a"] + host["b"]
    dc = host["f"] * host["dcn"] + host["o"] * dh
    dg = (one - host["g"] * host["g"]) * (host["i"] * dc)
    do = (host["o"] * (one - host["o"])) * (host["c"] * dh)
    np.testing.assert_array_equal(gpu.to_host(bufs["dc"], n, np.float32), dc)
    np.testing.assert_array_equal(gpu.to_host(bufs["dg"], n, np.float32), dg)
    np.testing.assert_array_equal(gpu.to_host(bufs["do"], n, np.float32), do)
    np.testing.assert_array_equal(gpu.to_host(bufs["dh"], n, np.float32), np.full(n, 7.0, np.float32))  # not kept: untouched
    # keep = NULL stores everything, and the result equals four separate launches
    gpu.check(gpu.lib().tcr_elementwise_multi(arr, len(progs), None))
    np.testing.assert_array_equal(gpu.to_host(bufs["dh"], n, np.float32), dh)
    for prog in progs:
        gpu.check(gpu.lib().tcr_elementwise(C.byref(prog)))
    np.testing.assert_array_equal(gpu.to_host(bufs["do"], n, np.float32), do)


def test_multi_with_broadcast_external_input_and_rejects_broadcast_of_internal(gpu):
    F, OP = gpu.FLOAT, gpu.OP
    H, B = 64, 48
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (B, H)).astype(np.float32)
    bias = rng.uniform(-1, 1, H).astype(np.float32)
    dx, db, dy, dz = gpu.to_device(x), gpu.to_device(bias), gpu.empty(B * H, np.float32), gpu.empty(B * H, np.float32)
    p0 = gpu.make_program(F, (H, B, 1), [(dx.ptr, F, (0, 0, 0)), (db.ptr, F, (0, 1, 0))], [(dy.ptr, F, 0)], [(OP["ADD"], 0, 0, 1)])
    p1 = gpu.make_program(F, (H, B, 1), [(dy.ptr, F, (0, 0, 0)), (db.ptr, F, (0, 1, 0))], [(dz.ptr, F, 0)], [(OP["MUL"], 0, 0, 1), (OP["TANH"], 0, 0)])
    arr = (gpu.EwProgram * 2)(p0, p1)
    gpu.check(gpu.lib().tcr_elementwise_multi(arr, 2, None))
    np.testing.assert_allclose(gpu.to_host(dz, B * H, np.float32).reshape(B, H), np.tanh((x + bias) * bias), rtol=1e-5, atol=1e-6)
    bad = gpu.make_program(F, (H, B, 1), [(dy.ptr, F, (0, 1, 0))], [(dz.ptr, F, 0)], [(OP["NEG"], 0, 0)])  # broadcast read of program 0's result
    arr = (gpu.EwProgram * 2)(p0, bad)
    assert gpu.lib().tcr_elementwise_multi(arr, 2, None) != 0


@pytest.mark.parametrize("n", [81920, 1000003, 5])
def test_elementwise_reduce_mean_squared(gpu, n):
    F, OP = gpu.FLOAT, gpu.OP
    rng = np.random.default_rng(n)
    a, b = rng.uniform(0, 1, n).astype(np.float32), rng.uniform(0, 1, n).astype(np.float32)
    da, db, out = gpu.to_device(a), gpu.to_device(b), gpu.empty(1, np.float32)
    prog = gpu.make_program(F, (n, 1, 1), [(da.ptr, F, (0, 0, 0)), (db.ptr, F, (0, 0, 0))], [(0, F, 0)], [(OP["SUB"], 0, 0, 1), (OP["SQUARE"], 0, 0)])
    for _ in range(3):  # the ticket counter resets itself: repeated launches
        gpu.check(gpu.lib().tcr_elementwise_reduce(C.byref(prog), C.c_void_p(out.ptr), 1, C.c_double(n)))
        got = float(gpu.to_host(out, 1, np.float32)[0])
        want = float(np.sum((a.astype(np.float64) - b) ** 2) / n)
        assert abs(got - want) <= 1e-5 * abs(want)


@pytest.mark.parametrize("n", [65536, 4099])
def test_cell_backward_is_bit_identical_to_the_separate_functors(gpu, n):
    """s = a + b; c = x*y + z*s; SIGMOID gate (g (1 - g)) * (y v); TANH gate (1 - g^2) * (y v) — every operation rounded on its own,
    as the reference's one-assignment-per-functor evaluation does (internal/eigen/device.hpp:555-570)."""
    rng = np.random.default_rng(n)
    names = ["a", "b", "cx", "cy", "cz", "g0", "y0", "g1", "y1", "g2", "y2"]
    host = {k: rng.uniform(-1, 1, n).astype(np.float32) for k in names}
    dev = {k: gpu.to_device(v) for k, v in host.items()}
    outs = {k: gpu.to_device(np.full(n, 3.0, np.float32)) for k in ("s", "c", "o0", "o1", "o2")}
    d = gpu.CellBackwardDesc()
    d.n = n
    d.s_a, d.s_b, d.c_x, d.c_y, d.c_z = dev["a"].ptr, dev["b"].ptr, dev["cx"].ptr, dev["cy"].ptr, dev["cz"].ptr
    d.s_out, d.c_out = None, outs["c"].ptr
    d.n_gates = 3
    for k, (kind, sel) in enumerate([(1, 0), (2, 1), (1, 1)]):
        d.kind[k], d.sel[k] = kind, sel
        d.x[k], d.y[k], d.out[k] = dev["g%d" % k].ptr, dev["y%d" % k].ptr, outs["o%d" % k].ptr
    gpu.check(gpu.lib().tcr_cell_backward(C.byref(d)))
    one = np.float32(1)
    s = host["a"] + host["b"]
    c = host["cx"] * host["cy"] + host["cz"] * s
    want = [(host["g0"] * (one - host["g0"])) * (host["y0"] * s), (one - host["g1"] * host["g1"]) * (host["y1"] * c),
            (host["g2"] * (one - host["g2"])) * (host["y2"] * c)]
    np.testing.assert_array_equal(gpu.to_host(outs["c"], n, np.float32), c)
    np.testing.assert_array_equal(gpu.to_host(outs["s"], n, np.float32), np.full(n, 3.0, np.float32))  # NULL s_out: not stored
    for k in range(3):
        np.testing.assert_array_equal(gpu.to_host(outs["o%d" % k], n, np.float32), want[k])


def test_copy2d_batched_lays_down_concats(gpu):
    """CONCAT(x_t, h_t) for several t as one launch of strided 2-D copies (operator.hpp:336-368 semantics per pair)."""
    rng = np.random.default_rng(2)
    T, B, KX, KH = 5, 7, 12, 20  # rows of 48 / 80 bytes: 16-byte path; second case below is unaligned
    for kx, kh in ((KX, KH), (3, 5)):
        xs = rng.uniform(-1, 1, (T, B, kx)).astype(np.float32)
        hs = rng.uniform(-1, 1, (T, B, kh)).astype(np.float32)
        dx, dh = gpu.to_device(xs), gpu.to_device(hs)
        out = gpu.to_device(np.zeros((T, B, kx + kh), np.float32))

        class Item(C.Structure):
            _fields_ = [("dst", C.c_void_p), ("src", C.c_void_p), ("row_bytes", C.c_int64), ("rows", C.c_int64), ("dst_pitch", C.c_int64), ("src_pitch", C.c_int64)]
        items = (Item * (2 * T))()
        for t in range(T):
            base = out.ptr + 4 * t * B * (kx + kh)
            items[2 * t] = Item(base, dx.ptr + 4 * t * B * kx, 4 * kx, B, 4 * (kx + kh), 4 * kx)
            items[2 * t + 1] = Item(base + 4 * kx, dh.ptr + 4 * t * B * kh, 4 * kh, B, 4 * (kx + kh), 4 * kh)
        gpu.check(gpu.lib().tcr_copy2d_batched(items, 2 * T))
        got = gpu.to_host(out, T * B * (kx + kh), np.float32).reshape(T, B, kx + kh)
        np.testing.assert_array_equal(got, np.concatenate([xs, hs], axis=2))


@pytest.mark.parametrize("n", [8192 * 784, 1000003, 17, 5])
def test_pixel_cast_and_scale(gpu, n):
    """(float)u8 * c — CAST of a UINT8 variable (tenncor/eteq/caster.hpp:10-44) folded into the MUL that scales it: bit-exact."""
    F, OP = gpu.FLOAT, gpu.OP
    rng = np.random.default_rng(n)
    px = rng.integers(0, 256, n, dtype=np.uint8)
    dpx, out = gpu.to_device(px), gpu.empty(n, np.float32)
    c = np.float32(1.0 / 255.0)
    prog = gpu.make_program(F, (n, 1, 1), [(dpx.ptr, gpu.UINT8, (0, 0, 0))], [(out.ptr, F, 0)],
                            [(gpu.EW_MOV, 1, 0), (gpu.EW_CONST, 2, 0, 0, 0, float(c)), (OP["MUL"], 0, 1, 2)])
    gpu.check(gpu.lib().tcr_elementwise(C.byref(prog)))
    np.testing.assert_array_equal(gpu.to_host(out, n, np.float32), px.astype(np.float32) * c)
    prog = gpu.make_program(F, (n, 1, 1), [(dpx.ptr, gpu.UINT8, (0, 0, 0))], [(out.ptr, F, 1)], [(gpu.EW_MOV, 1, 0)])
    gpu.check(gpu.lib().tcr_elementwise(C.byref(prog)))
    np.testing.assert_array_equal(gpu.to_host(out, n, np.float32), px.astype(np.float32))
