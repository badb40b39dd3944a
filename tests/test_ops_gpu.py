"""GPU parity of every kernel class against the oracle, through the C-ABI.

Bars (BASELINE.json north_star): bit-exact for indexing / integer / layout ops; 1e-5
relative for fp32 elementwise and reductions (tolerance written at each assert);
double-precision goldens of the reference's own tests to 1e-12.
"""
import ctypes as C
import json
import os
import zlib

import numpy as np
import pytest

from oracle import tcr_oracle as orc
from tests import opcheck

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "operator_goldens.json")) as f:
    GOLD = json.load(f)["cases"]

def _seed(*a):
    return zlib.crc32(repr(a).encode())


FP32_RTOL = 1e-5  # north_star: "within 1e-5 relative for fp32 elementwise and reduce"


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_reference_goldens_on_gpu(gpu, case):
    """The reference's own operator unit-test vectors (test_operator.cpp), in double."""
    arrs, shapes = opcheck.case_arrays(case)
    got, oshape = opcheck.gpu_run(gpu, case["op"], arrs, shapes, case["attrs"], opcheck.NP.get(case["out_dtype"]))
    expect = opcheck.expected_of(case)
    assert orc.n_elems(oshape) == orc.n_elems(case["out_shape"])
    np.testing.assert_allclose(np.asarray(got, dtype=np.float64), expect, rtol=1e-12, atol=0)


UNARY = ["ABS", "NEG", "SIN", "COS", "TAN", "EXP", "LOG", "SQRT", "ROUND", "SIGMOID", "TANH", "SQUARE", "CUBE"]
BINARY = ["POW", "ADD", "SUB", "MUL", "DIV", "MIN", "MAX", "EQ", "NEQ", "LT", "GT"]


def _rand(rng, n, dt, lo=-4, hi=4):
    if np.issubdtype(dt, np.integer):
        return rng.integers(lo, hi + 1, n).astype(dt)
    return rng.uniform(lo, hi, n).astype(dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32], ids=["f32", "f64", "i32"])
@pytest.mark.parametrize("op", UNARY)
@pytest.mark.parametrize("n", [1, 7, 1000, 4099, 1 << 18])
def test_unary(gpu, op, dt, n):
    rng = np.random.default_rng(_seed(op, n))
    x = _rand(rng, n, dt)
    if op in ("LOG", "SQRT"):
        x = np.abs(x) + dt(1)
    if op == "TAN":
        x = _rand(rng, n, dt, -1, 1)
    if op == "EXP" and np.issubdtype(dt, np.integer):
        x = np.abs(x)
    shape = orc.full_shape([n])
    got, _ = opcheck.gpu_run(gpu, op, [x], [shape], {})
    want, _ = opcheck.oracle_run(op, [x], [shape], {})
    if np.issubdtype(dt, np.integer):
        if op in ("SIN", "COS", "TAN", "TANH", "SIGMOID", "EXP", "LOG", "SQRT"):
            assert np.max(np.abs(got.astype(np.int64) - want.astype(np.int64))) <= 1  # truncation of a transcendental
        else:
            np.testing.assert_array_equal(got, want)
    elif op in ("ABS", "NEG", "ROUND", "SQUARE", "CUBE", "SQRT"):
        np.testing.assert_array_equal(got, want)  # correctly rounded ops are bit-exact
    else:
        tol = FP32_RTOL if dt == np.float32 else 1e-12
        np.testing.assert_allclose(got, want, rtol=tol, atol=tol if op in ("SIN", "COS", "TAN") else 0)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32], ids=["f32", "f64", "i32"])
@pytest.mark.parametrize("op", BINARY)
@pytest.mark.parametrize("n", [1, 5, 4097, 1 << 18])
def test_binary(gpu, op, dt, n):
    rng = np.random.default_rng(_seed(op, n, 1))
    a, b = _rand(rng, n, dt), _rand(rng, n, dt)
    if op == "POW":
        a = np.abs(a) + dt(1)
        b = _rand(rng, n, dt, 0, 3)
    if op == "DIV":
        b = np.where(b == 0, dt(1), b)
    if op in ("EQ", "NEQ", "MIN", "MAX"):
        b[::3] = a[::3]
    shape = orc.full_shape([n])
    got, _ = opcheck.gpu_run(gpu, op, [a, b], [shape, shape], {})
    want, _ = opcheck.oracle_run(op, [a, b], [shape, shape], {})
    if op == "POW" and not np.issubdtype(dt, np.integer):
        np.testing.assert_allclose(got, want, rtol=FP32_RTOL if dt == np.float32 else 1e-12)
    elif op == "POW":
        assert np.max(np.abs(got.astype(np.int64) - want.astype(np.int64))) <= 1
    else:
        np.testing.assert_array_equal(got, want)  # IEEE add/sub/mul/div/compare are exact


def test_select_and_nnary(gpu):
    rng = np.random.default_rng(3)
    n = 10007
    shape = orc.full_shape([n])
    c = (rng.random(n) < 0.5).astype(np.float32)
    a, b = _rand(rng, n, np.float32), _rand(rng, n, np.float32)
    got, _ = opcheck.gpu_run(gpu, "SELECT", [c, a, b], [shape] * 3, {})
    np.testing.assert_array_equal(got, np.where(c != 0, a, b))
    for k in (3, 8, 9, 20, 130):
        args = [_rand(rng, n, np.float64) for _ in range(k)]
        got, _ = opcheck.gpu_run(gpu, "ADD", args, [shape] * k, {})
        want, _ = opcheck.oracle_run("ADD", args, [shape] * k, {})
        np.testing.assert_allclose(got, want, rtol=1e-12)
    args = [_rand(rng, n, np.int32, -3, 3) for _ in range(5)]
    got, _ = opcheck.gpu_run(gpu, "MUL", args, [shape] * 5, {})
    want, _ = opcheck.oracle_run("MUL", args, [shape] * 5, {})
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("dims,bc", [
    ((1000, 1, 1), [(0, 0, 0)] * 3),
    ((37, 19, 1), [(0, 0, 0), (0, 1, 0), (1, 0, 0)]),     # bias [H] over [H,B]; per-column scalar [1,B]
    ((8, 5, 7), [(0, 0, 0), (0, 1, 0), (1, 1, 1)]),        # middle broadcast, full scalar
    ((64, 33, 1), [(0, 1, 0), (0, 0, 0), (1, 1, 0)]),
])
def test_fused_program_with_broadcast(gpu, dims, bc):
    """sigmoid(a*b + c) as ONE kernel with broadcast operands (EXTEND never materialised)."""
    rng = np.random.default_rng(11)
    lib = gpu.lib()
    D = dims
    n = D[0] * D[1] * D[2]

    def ext(b):
        return [1 if b[k] else D[k] for k in range(3)]

    host = [rng.uniform(-2, 2, int(np.prod(ext(b)))).astype(np.float32) for b in bc]
    dev = [gpu.to_device(h) for h in host]
    out = gpu.empty(n, np.float32)
    out2 = gpu.empty(n, np.float32)
    OPC = gpu.OP
    prog = gpu.make_program(gpu.FLOAT, D,
                            [(d.ptr, gpu.FLOAT, b) for d, b in zip(dev, bc)],
                            [(out.ptr, gpu.FLOAT, 3), (out2.ptr, gpu.FLOAT, 4)],
                            [(OPC["MUL"], 3, 0, 1), (OPC["ADD"], 3, 3, 2), (gpu.EW_CONST, 5, 0, 0, 0, 0.5),
                             (OPC["MUL"], 4, 3, 5), (OPC["SIGMOID"], 3, 3)])
    gpu.check(lib.tcr_elementwise(C.byref(prog)))
    got, got2 = gpu.to_host(out, n, np.float32), gpu.to_host(out2, n, np.float32)
    full = []
    for h, b in zip(host, bc):
        a = h.reshape(ext(b)[::-1])
        full.append(np.broadcast_to(a, D[::-1]).reshape(-1).astype(np.float64))
    t = full[0] * full[1] + full[2]
    np.testing.assert_allclose(got, 1 / (1 + np.exp(-t)), rtol=FP32_RTOL, atol=1e-7)
    np.testing.assert_allclose(got2, 0.5 * t, rtol=FP32_RTOL, atol=1e-6)


@pytest.mark.parametrize("d0,rows,binop,rev,un", [
    (1024, 67, "ADD", False, "SIGMOID"),     # dense bias + activation: one column block, rows with a ragged tail of the 4-row unroll
    (64, 1031, "ADD", True, "TANH"),         # 16 vectors per row: 16 row lanes per block
    (2048 + 512, 33, "MUL", False, None),    # three column blocks, the last one partial
    (256, 300, "SUB", True, "EXP"),          # vector - x
    (128, 512, "DIV", False, "SQUARE"),
    (4096, 9, "MAX", False, "NEG"),
])
def test_vector_broadcast_program_matches_separate_functors(gpu, d0, rows, binop, rev, un):
    """un(bin(x, EXTEND(v))) with v along segment 0 takes ew_rowvec_kernel; values are bit-identical to the functors run one by one
    over the materialised broadcast (internal/eigen/operator.hpp:159-173 materialises it), also when the result overwrites x."""
    rng = np.random.default_rng(d0 + rows)
    lib = gpu.lib()
    n = d0 * rows
    x = rng.uniform(-2, 2, n).astype(np.float32)
    v = rng.uniform(0.5, 2, d0).astype(np.float32)
    dx, dv, out = gpu.to_device(x), gpu.to_device(v), gpu.empty(n, np.float32)
    OPC = gpu.OP
    ins = [(dx.ptr, gpu.FLOAT, (0, 0, 0)), (dv.ptr, gpu.FLOAT, (0, 1, 0))]
    instrs = [(OPC[binop], 2, 1, 0) if rev else (OPC[binop], 2, 0, 1)]
    if un:
        instrs.append((OPC[un], 3, 2))
    reg = 3 if un else 2
    launches = lib.tcr_launch_count()
    prog = gpu.make_program(gpu.FLOAT, (d0, rows, 1), ins, [(out.ptr, gpu.FLOAT, reg)], instrs)
    gpu.check(lib.tcr_elementwise(C.byref(prog)))
    assert lib.tcr_launch_count() == launches + 1
    got = gpu.to_host(out, n, np.float32)
    # the separate functors over the materialised broadcast
    dfull = gpu.to_device(np.tile(v, rows))
    t = gpu.empty(n, np.float32)
    a, b = (dfull, dx) if rev else (dx, dfull)
    gpu.check(lib.tcr_binary(OPC[binop], C.c_void_p(a.ptr), C.c_void_p(b.ptr), C.c_void_p(t.ptr), C.c_int64(n), gpu.FLOAT))
    if un:
        gpu.check(lib.tcr_unary(OPC[un], C.c_void_p(t.ptr), C.c_void_p(t.ptr), C.c_int64(n), gpu.FLOAT))
    want = gpu.to_host(t, n, np.float32)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    # in place: the result replaces x
    prog = gpu.make_program(gpu.FLOAT, (d0, rows, 1), ins, [(dx.ptr, gpu.FLOAT, reg)], instrs)
    gpu.check(lib.tcr_elementwise(C.byref(prog)))
    np.testing.assert_array_equal(gpu.to_host(dx, n, np.float32).view(np.uint32), want.view(np.uint32))


def test_program_mixed_dtype_cast_and_inplace(gpu):
    rng = np.random.default_rng(5)
    lib = gpu.lib()
    n = 5001
    x = rng.uniform(-100, 100, n)
    dx = gpu.to_device(x)
    out = gpu.empty(n, np.int32)
    prog = gpu.make_program(gpu.DOUBLE, (n, 1, 1), [(dx.ptr, gpu.DOUBLE, (0, 0, 0))], [(out.ptr, gpu.INT32, 0)], [])
    gpu.check(lib.tcr_elementwise(C.byref(prog)))
    np.testing.assert_array_equal(gpu.to_host(out, n, np.int32), x.astype(np.int32))
    # w -= lr * g in place
    w, g = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32)
    dw, dg = gpu.to_device(w), gpu.to_device(g)
    prog = gpu.make_program(gpu.FLOAT, (n, 1, 1), [(dw.ptr, gpu.FLOAT, (0, 0, 0)), (dg.ptr, gpu.FLOAT, (0, 0, 0))],
                            [(dw.ptr, gpu.FLOAT, 0)],
                            [(gpu.EW_CONST, 2, 0, 0, 0, 0.9), (gpu.OP["MUL"], 1, 1, 2), (gpu.OP["SUB"], 0, 0, 1)])
    gpu.check(lib.tcr_elementwise(C.byref(prog)))
    np.testing.assert_array_equal(gpu.to_host(dw, n, np.float32), w - np.float32(0.9) * g)


@pytest.mark.parametrize("src,dst", [(np.float64, np.int32), (np.float32, np.float64), (np.int32, np.float32),
                                     (np.float32, np.uint8), (np.int64, np.int16)])
def test_cast(gpu, src, dst):
    rng = np.random.default_rng(1)
    x = _rand(rng, 3001, src, 0, 100)
    shape = orc.full_shape([3001])
    got, _ = opcheck.gpu_run(gpu, "CAST", [x], [shape], {}, out_dtype=dst)
    np.testing.assert_array_equal(got, x.astype(dst))


REDUCE_SHAPES = [
    ([3, 2], [1]), ([3, 2], [0]), ([3, 2], [0, 1]),
    ([1000], [0]), ([100003], [0]), ([257, 33], [0]), ([257, 33], [1]), ([33, 257], [1]),
    ([16, 4100], [1]), ([4100, 16], [0]), ([5, 7, 9], [1]), ([5, 7, 9], [0, 2]), ([5, 7, 9], [0, 1]),
    ([4, 3, 2, 5], [1, 3]), ([1 << 20], [0]), ([1024, 1024], list(range(8))), ([8, 1, 600], [2]),
]


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32], ids=["f32", "f64", "i32"])
@pytest.mark.parametrize("op", ["REDUCE_SUM", "REDUCE_PROD", "REDUCE_MIN", "REDUCE_MAX"])
@pytest.mark.parametrize("shape,ranks", REDUCE_SHAPES, ids=[str(s) + str(r) for s, r in REDUCE_SHAPES])
def test_reduce(gpu, op, dt, shape, ranks):
    rng = np.random.default_rng(7)
    n = orc.n_elems(shape)
    if op == "REDUCE_PROD":
        x = rng.uniform(0.9, 1.1, n).astype(dt) if not np.issubdtype(dt, np.integer) else rng.choice([1, 1, 1, -1, 2], n).astype(dt)
        if np.issubdtype(dt, np.integer) and n > 2000:
            x = rng.choice([1, -1], n).astype(dt)
    else:
        x = _rand(rng, n, dt)
    s8 = orc.full_shape(shape)
    got, oshape = opcheck.gpu_run(gpu, op, [x], [s8], {"rank_set": ranks})
    # the checker accumulates in double for fp32 inputs
    want, wshape = orc.reduce(op, x.astype(np.float64) if dt == np.float32 else x, s8, ranks)
    assert oshape == wshape
    if np.issubdtype(dt, np.integer) or op in ("REDUCE_MIN", "REDUCE_MAX"):
        np.testing.assert_array_equal(got, want.astype(dt))
    else:
        rtol = FP32_RTOL if dt == np.float32 else 1e-11
        scale = float(np.max(np.abs(want))) + 1.0
        if op == "REDUCE_SUM":  # relative to the magnitude of the summed terms
            asum, _ = orc.reduce(op, np.abs(x).astype(np.float64), s8, ranks)
            assert np.all(np.abs(got - want) <= rtol * (asum + 1e-30))
        else:
            np.testing.assert_allclose(got, want, rtol=rtol * 20, atol=rtol * scale)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32], ids=["f32", "f64", "i32"])
@pytest.mark.parametrize("shape,dim", [([3, 2], 1), ([3, 2], 0), ([3, 2], 8), ([1000, 7], 0), ([7, 1000], 1),
                                       ([40, 50, 6], 1), ([1 << 20], 8), ([513, 129], 8), ([10, 4096], 0)])
def test_argmax_with_ties(gpu, dt, shape, dim):
    rng = np.random.default_rng(9)
    n = orc.n_elems(shape)
    x = rng.integers(0, 50, n).astype(dt)  # many exact ties: lowest index must win
    s8 = orc.full_shape(shape)
    got, oshape = opcheck.gpu_run(gpu, "ARGMAX", [x], [s8], {"rank": dim})
    want, wshape = orc.argmax(x, s8, dim)
    assert oshape == wshape
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("fill", [-np.inf, float(np.finfo(np.float32).min)], ids=["-inf", "lowest"])
def test_argmax_flat_of_a_tensor_of_lowest_values(gpu, fill):
    """every element equals the reduction's identity: the answer is still the first index (Eigen argmax keeps the first maximum)"""
    n = (1 << 18) + 3
    x = np.full(n, fill, np.float32)
    s8 = orc.full_shape([n])
    got, _ = opcheck.gpu_run(gpu, "ARGMAX", [x], [s8], {"rank": 8})
    assert float(np.asarray(got).reshape(-1)[0]) == 0.0
    x[12345] = fill if np.isinf(fill) else 0.0
    x[777] = np.nan  # never wins
    got, _ = opcheck.gpu_run(gpu, "ARGMAX", [x], [s8], {"rank": 8})
    assert float(np.asarray(got).reshape(-1)[0]) == (0.0 if np.isinf(fill) else 12345.0)


LAYOUT = [
    ("EXTEND", [7, 1, 5], {"dimensions": [1, 6]}),
    ("EXTEND", [1], {"dimensions": [33, 9]}),
    ("EXTEND", [128], {"dimensions": [1, 300]}),
    ("EXTEND", [1, 300], {"dimensions": [128]}),
    ("EXTEND", [3, 1, 4, 1, 2], {"dimensions": [1, 5, 1, 6]}),
    ("PERMUTE", [65, 130], {"ranks": [1, 0]}),
    ("PERMUTE", [31, 33, 5], {"ranks": [1, 0]}),
    ("PERMUTE", [4, 5, 6], {"ranks": [2, 0, 1]}),
    ("PERMUTE", [4, 5, 6, 3], {"ranks": [0, 2, 1, 3]}),
    ("PERMUTE", [2, 3, 4, 5, 2, 3, 2, 2], {"ranks": [7, 6, 5, 4, 3, 2, 1, 0]}),
    ("SLICE", [9, 8, 7], {"dimension_pairs": [[2, 4], [0, 100], [3, 2]]}),
    ("SLICE", [64, 10, 6], {"dimension_pairs": [[0, 64], [3, 1]]}),
    ("SLICE", [5, 4], {"dimension_pairs": [[100, 1]]}),
    ("PAD", [5, 4, 3], {"dimension_pairs": [[1, 2], [0, 0], [2, 0]]}),
    ("PAD", [64, 1, 6], {"dimension_pairs": [[0, 0], [3, 6]]}),
    ("SLICE", [1024, 128, 40], {"dimension_pairs": [[0, 1024], [32, 64]]}),   # row_window_kernel over several unrolled iterations
    ("PAD", [1024, 128, 24], {"dimension_pairs": [[0, 0], [16, 16]]}),
    ("STRIDE", [10, 9], {"dimensions": [2, 3]}),
    ("SCATTER", [5, 3], {"dimensions": [2, 3], "shape": [10, 9]}),
    ("SCATTER", [5, 3], {"dimensions": [2, 3], "shape": [9, 7]}),
    ("REVERSE", [6, 5, 4], {"rank_set": [0, 2]}),
    ("REVERSE", [1000], {"rank_set": [0]}),
]


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.uint8, np.int16], ids=["4B", "8B", "1B", "2B"])
@pytest.mark.parametrize("op,shape,attrs", LAYOUT, ids=["%s%s" % (o, s) for o, s, _ in LAYOUT])
def test_layout_bit_exact(gpu, op, shape, attrs, dt):
    n = orc.n_elems(shape)
    x = (np.arange(n) % 251).astype(dt) if dt in (np.uint8,) else np.arange(1, n + 1).astype(dt)
    s8 = orc.full_shape(shape)
    got, oshape = opcheck.gpu_run(gpu, op, [x], [s8], attrs)
    want, wshape = opcheck.oracle_run(op, [x], [s8], attrs)
    assert orc.n_elems(oshape) == orc.n_elems(wshape)
    assert got.tobytes() == np.ascontiguousarray(want).tobytes()


@pytest.mark.parametrize("shapes,axis", [
    ([[2, 3], [1, 3]], 0), ([[3, 2], [3, 5]], 1), ([[4, 3, 5], [4, 2, 5]], 1), ([[7, 1, 3]] * 2, 1),
    ([[6, 1, 4]] * 5, 1), ([[5, 1]] * 3, 1), ([[1, 9]] * 4, 0), ([[16, 1, 3]] * 70, 1),
])
def test_concat_bit_exact(gpu, shapes, axis):
    rng = np.random.default_rng(2)
    arrs = [rng.integers(0, 1 << 30, orc.n_elems(s)).astype(np.int32) for s in shapes]
    s8 = [orc.full_shape(s) for s in shapes]
    got, _ = opcheck.gpu_run(gpu, "CONCAT", arrs, s8, {"rank": axis})
    want, _ = opcheck.oracle_run("CONCAT", arrs, s8, {"rank": axis})
    np.testing.assert_array_equal(got, want)


def _gemm(gpu, A, B, trans_a, trans_b, precision=0, bias=None, epi=0, act=0, batch=1, perm_c=False):
    """A: [batch, M, K] logical, B: [batch, K, N]; storage transposed on request."""
    lib = gpu.lib()
    _, M, K = A.shape
    N = B.shape[2]
    As = np.ascontiguousarray(A.transpose(0, 2, 1)) if trans_a else np.ascontiguousarray(A)
    Bs = np.ascontiguousarray(B.transpose(0, 2, 1)) if trans_b else np.ascontiguousarray(B)
    da, db = gpu.to_device(As), gpu.to_device(Bs)
    dc = gpu.empty(batch * M * N, A.dtype)
    d = gpu.GemmDesc(m=M, n=N, k=K, batch=batch,
                     a_sm=1 if trans_a else K, a_sk=M if trans_a else 1, a_sb=M * K,
                     b_sk=1 if trans_b else N, b_sn=K if trans_b else 1, b_sb=K * N,
                     c_sm=1 if perm_c else N, c_sn=M if perm_c else 1, c_sb=M * N,
                     dtype=gpu.DTYPE_OF[np.dtype(A.dtype)], precision=precision, epilogue=epi, activation=act)
    keep = None
    if bias is not None:
        keep = gpu.to_device(bias)
        d.bias = keep.ptr
    gpu.check(lib.tcr_gemm(C.c_void_p(da.ptr), C.c_void_p(db.ptr), C.c_void_p(dc.ptr), C.byref(d)))
    out = gpu.to_host(dc, batch * M * N, A.dtype)
    return out.reshape(batch, N, M).transpose(0, 2, 1) if perm_c else out.reshape(batch, M, N)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32], ids=["f32", "f64", "i32"])
@pytest.mark.parametrize("M,N,K,batch", [(3, 2, 4, 1), (64, 64, 16, 1), (65, 63, 17, 1), (130, 70, 33, 3), (1, 1, 1, 1), (257, 129, 300, 1)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_gemm_exact_all_layouts(gpu, dt, M, N, K, batch, ta, tb):
    rng = np.random.default_rng(4)
    if np.issubdtype(dt, np.integer):
        A, B = rng.integers(-5, 6, (batch, M, K)).astype(dt), rng.integers(-5, 6, (batch, K, N)).astype(dt)
    else:
        A, B = rng.uniform(-1, 1, (batch, M, K)).astype(dt), rng.uniform(-1, 1, (batch, K, N)).astype(dt)
    got = _gemm(gpu, A, B, ta, tb, batch=batch, perm_c=bool(ta and tb))
    want = np.matmul(A.astype(np.float64), B.astype(np.float64))
    if np.issubdtype(dt, np.integer):
        np.testing.assert_array_equal(got, want.astype(dt))
    else:
        # fp32 FMA accumulation over K terms vs double: |err| <= K * eps * sum|a||b|
        bound = np.matmul(np.abs(A).astype(np.float64), np.abs(B).astype(np.float64)) * K * np.finfo(dt).eps
        assert np.all(np.abs(got - want) <= bound + 1e-300)


def test_gemm_bias_activation_epilogue(gpu):
    rng = np.random.default_rng(6)
    M, N, K = 70, 50, 40
    A, B = rng.uniform(-1, 1, (1, M, K)).astype(np.float32), rng.uniform(-1, 1, (1, K, N)).astype(np.float32)
    bias = rng.uniform(-1, 1, N).astype(np.float32)
    got = _gemm(gpu, A, B, 0, 0, bias=bias, epi=gpu.EPI_BIAS_N, act=gpu.OP["SIGMOID"])
    want = 1 / (1 + np.exp(-(np.matmul(A.astype(np.float64), B.astype(np.float64)) + bias)))
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


CONTRACTS = [
    ([4, 3], [2, 4], [[0, 1]]), ([4, 3], [4, 5], [[0, 0]]), ([3, 4], [5, 4], [[1, 1]]), ([3, 4], [3, 5], [[0, 0]]),
    ([4, 2, 3], [2, 4, 2], [[0, 1], [1, 2]]), ([3, 4, 5], [5, 6], [[2, 0]]), ([6, 5], [3, 4], []),
]


@pytest.mark.parametrize("ashape,bshape,pairs", CONTRACTS, ids=[str(c) for c in CONTRACTS])
def test_contract_generic(gpu, ashape, bshape, pairs):
    rng = np.random.default_rng(8)
    a = rng.integers(-4, 5, orc.n_elems(ashape)).astype(np.float64)
    b = rng.integers(-4, 5, orc.n_elems(bshape)).astype(np.float64)
    sa, sb = orc.full_shape(ashape), orc.full_shape(bshape)
    got, gs = opcheck.gpu_run(gpu, "CONTRACT", [a, b], [sa, sb], {"rank_pairs": pairs})
    want, ws = opcheck.oracle_run("CONTRACT", [a, b], [sa, sb], {"rank_pairs": pairs})
    assert gs == ws
    np.testing.assert_array_equal(got, want)  # small integers: exact in double


@pytest.mark.parametrize("ishape,kshape,order", [([3, 3], [2], [1]), ([8, 9, 3], [3, 2], [0, 1]), ([5, 6, 7, 2], [2, 3, 2], [2, 0, 1]),
                                                 ([4, 10, 10, 3], [4, 3, 3], [0, 1, 2])])
def test_conv_valid_correlation(gpu, ishape, kshape, order):
    rng = np.random.default_rng(10)
    img = rng.integers(-3, 4, orc.n_elems(ishape)).astype(np.float64)
    ker = rng.integers(-3, 4, orc.n_elems(kshape)).astype(np.float64)
    si, sk = orc.full_shape(ishape), orc.full_shape(kshape)
    got, gs = opcheck.gpu_run(gpu, "CONV", [img, ker], [si, sk], {"ranks": order})
    want, ws = opcheck.oracle_run("CONV", [img, ker], [si, sk], {"ranks": order})
    assert gs == ws
    np.testing.assert_array_equal(got, want)


def test_rand_unif_statistics(gpu):
    lib = gpu.lib()
    n = 1 << 20
    lo, hi = np.full(n, -2.0, np.float32), np.full(n, 3.0, np.float32)
    dlo, dhi, out = gpu.to_device(lo), gpu.to_device(hi), gpu.empty(n, np.float32)
    gpu.check(lib.tcr_rand_unif(C.c_void_p(dlo.ptr), C.c_void_p(dhi.ptr), C.c_void_p(out.ptr), C.c_int64(n), gpu.FLOAT, C.c_uint64(42), C.c_uint64(0)))
    x = gpu.to_host(out, n, np.float32)
    assert x.min() >= -2.0 and x.max() < 3.0  # reference test asserts a <= e <= b (test_operator.cpp:1768-1778)
    assert abs(x.mean() - 0.5) < 0.01 and abs(x.var() - 25 / 12) < 0.02
    gpu.check(lib.tcr_rand_unif(C.c_void_p(dlo.ptr), C.c_void_p(dhi.ptr), C.c_void_p(out.ptr), C.c_int64(n), gpu.FLOAT, C.c_uint64(42), C.c_uint64(0)))
    np.testing.assert_array_equal(x, gpu.to_host(out, n, np.float32))  # counter based: reproducible
    gpu.check(lib.tcr_rand_unif(C.c_void_p(dlo.ptr), C.c_void_p(dhi.ptr), C.c_void_p(out.ptr), C.c_int64(n), gpu.FLOAT, C.c_uint64(42), C.c_uint64(n)))
    assert np.mean(x == gpu.to_host(out, n, np.float32)) < 0.01  # new offset -> new stream
    ilo, ihi = np.full(4096, 3, np.int32), np.full(4096, 6, np.int32)
    dlo, dhi, out = gpu.to_device(ilo), gpu.to_device(ihi), gpu.empty(4096, np.int32)
    gpu.check(lib.tcr_rand_unif(C.c_void_p(dlo.ptr), C.c_void_p(dhi.ptr), C.c_void_p(out.ptr), C.c_int64(4096), gpu.INT32, C.c_uint64(1), C.c_uint64(0)))
    xi = gpu.to_host(out, 4096, np.int32)
    assert set(np.unique(xi)) == {3, 4, 5, 6}


def test_graph_capture_replay_and_arena(gpu):
    """A captured launch sequence replays with stable addresses; temporaries freed inside
    the capture stay reserved for the graph."""
    lib = gpu.lib()
    n = 1 << 16
    x = np.random.default_rng(0).uniform(-1, 1, n).astype(np.float32)
    dx, dy, dz = gpu.to_device(x), gpu.empty(n, np.float32), gpu.empty(1, np.float32)
    gpu.check(lib.tcr_graph_begin())
    gpu.check(lib.tcr_unary(gpu.OP["EXP"], C.c_void_p(dx.ptr), C.c_void_p(dy.ptr), C.c_int64(n), gpu.FLOAT))
    gpu.check(lib.tcr_reduce(gpu.OP["REDUCE_SUM"], C.c_void_p(dy.ptr), C.c_void_p(dz.ptr), gpu.shape8([n]), C.c_uint32(1), gpu.FLOAT))
    exe = C.c_void_p()
    gpu.check(lib.tcr_graph_end(C.byref(exe)))
    hog = [gpu.empty(1 << 10, np.float32) for _ in range(8)]  # must not receive the graph's temporaries
    for h in hog:
        gpu.check(lib.tcr_memset(C.c_void_p(h.ptr), 0xFF, C.c_size_t(h.nbytes)))
    for scale in (1.0, 0.5):
        x2 = (x * scale).astype(np.float32)
        gpu.check(lib.tcr_h2d(C.c_void_p(dx.ptr), x2.ctypes.data_as(C.c_void_p), C.c_size_t(x2.nbytes)))
        gpu.check(lib.tcr_graph_launch(exe))
        got = gpu.to_host(dz, 1, np.float32)[0]
        assert abs(got - np.exp(x2.astype(np.float64)).sum()) <= 1e-5 * np.exp(x2.astype(np.float64)).sum()
    gpu.check(lib.tcr_graph_destroy(exe))


@pytest.mark.parametrize("op", ["REDUCE_SUM", "REDUCE_MAX"])
def test_reduce_more_rows_than_a_grid_dimension(gpu, op):
    """[1100, 66000] reduced over the fast rank: 66000 kept rows (> 65535, the SURVEY §8d
    [4096, 65536] case scaled down) must take the row kernel, not the one-thread-per-output fallback."""
    rng = np.random.default_rng(11)
    shape = [1100, 66000]
    x = rng.uniform(-1, 1, shape[0] * shape[1]).astype(np.float32)
    s8 = orc.full_shape(shape)
    got, oshape = opcheck.gpu_run(gpu, op, [x], [s8], {"rank_set": [0]})
    rows = x.reshape(shape[1], shape[0]).astype(np.float64)
    want = rows.sum(axis=1) if op == "REDUCE_SUM" else rows.max(axis=1)
    assert list(oshape)[:2] == [1, 66000]
    if op == "REDUCE_MAX":
        np.testing.assert_array_equal(got.reshape(-1), want.astype(np.float32))
    else:
        assert np.all(np.abs(got.reshape(-1) - want) <= FP32_RTOL * np.abs(rows).sum(axis=1))


def test_graph_capture_lanes(gpu):
    """Two independent chains captured on different lanes, joined by a mark: the replayed graph
    computes z = exp(x) + tanh(y) (lane 1 and lane 2 feed lane 0). Lane calls outside a capture fail."""
    lib = gpu.lib()
    n = 1 << 15
    rng = np.random.default_rng(1)
    x, y = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32)
    dx, dy = gpu.to_device(x), gpu.to_device(y)
    ex, ty, dz = gpu.empty(n, np.float32), gpu.empty(n, np.float32), gpu.empty(n, np.float32)
    assert lib.tcr_graph_lane(1) != 0  # no capture in progress
    gpu.check(lib.tcr_graph_begin())
    m1, m2 = C.c_int(-1), C.c_int(-1)
    gpu.check(lib.tcr_graph_lane(1))
    gpu.check(lib.tcr_unary(gpu.OP["EXP"], C.c_void_p(dx.ptr), C.c_void_p(ex.ptr), C.c_int64(n), gpu.FLOAT))
    gpu.check(lib.tcr_graph_record(C.byref(m1)))
    gpu.check(lib.tcr_graph_lane(2))
    gpu.check(lib.tcr_unary(gpu.OP["TANH"], C.c_void_p(dy.ptr), C.c_void_p(ty.ptr), C.c_int64(n), gpu.FLOAT))
    gpu.check(lib.tcr_graph_record(C.byref(m2)))
    gpu.check(lib.tcr_graph_lane(0))
    gpu.check(lib.tcr_graph_wait(m1.value))
    gpu.check(lib.tcr_graph_wait(m2.value))
    gpu.check(lib.tcr_binary(gpu.OP["ADD"], C.c_void_p(ex.ptr), C.c_void_p(ty.ptr), C.c_void_p(dz.ptr), C.c_int64(n), gpu.FLOAT))
    exe = C.c_void_p()
    gpu.check(lib.tcr_graph_end(C.byref(exe)))  # joins lanes 1 and 2
    assert lib.tcr_graph_wait(m1.value) != 0  # marks die with the capture
    for scale in (1.0, -0.5):
        x2, y2 = (x * scale).astype(np.float32), (y * scale).astype(np.float32)
        gpu.check(lib.tcr_h2d(C.c_void_p(dx.ptr), x2.ctypes.data_as(C.c_void_p), C.c_size_t(x2.nbytes)))
        gpu.check(lib.tcr_h2d(C.c_void_p(dy.ptr), y2.ctypes.data_as(C.c_void_p), C.c_size_t(y2.nbytes)))
        gpu.check(lib.tcr_graph_launch(exe))
        got = gpu.to_host(dz, n, np.float32)
        want = np.exp(x2.astype(np.float64)) + np.tanh(y2.astype(np.float64))
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    gpu.check(lib.tcr_graph_destroy(exe))


def test_errors_are_loud(gpu):
    lib = gpu.lib()
    buf = gpu.empty(16, np.float32)
    assert lib.tcr_unary(gpu.OP["ADD"], C.c_void_p(buf.ptr), C.c_void_p(buf.ptr), C.c_int64(4), gpu.FLOAT) != 0
    assert b"not a unary" in lib.tcr_last_error()
    assert lib.tcr_unary(gpu.OP["EXP"], C.c_void_p(buf.ptr), C.c_void_p(buf.ptr), C.c_int64(4), gpu.UINT16) != 0
    bc = (C.c_int64 * 8)(2, 1, 1, 1, 1, 1, 1, 1)
    assert lib.tcr_extend(C.c_void_p(buf.ptr), C.c_void_p(buf.ptr), gpu.shape8([3]), bc, 4) != 0
    assert b"non-singular" in lib.tcr_last_error()
