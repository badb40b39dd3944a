"""Shared helpers: run one op case through the oracle (CPU) and through the C-ABI (GPU)."""
import ctypes as C
import math

import numpy as np

from oracle import tcr_oracle as orc

NP = {"double": np.float64, "float": np.float32, "int32": np.int32, "int64": np.int64}

LIBM = {"ABS": abs, "NEG": lambda e: -e, "SIN": math.sin, "COS": math.cos, "TAN": math.tan, "EXP": math.exp,
        "LOG": math.log, "SQRT": math.sqrt, "ROUND": lambda e: float(math.floor(abs(e) + 0.5)) * (1 if e >= 0 else -1),
        "SIGMOID": lambda e: 1. / (1. + math.exp(-e)), "TANH": math.tanh, "SQUARE": lambda e: e * e,
        "CUBE": lambda e: e * e * e}


def expected_of(case):
    """Golden expectation; unary cases follow the reference test and apply libm in double."""
    if case["expect"] is not None:
        return np.array(case["expect"], dtype=np.float64)
    return np.array([LIBM[case["op"]](float(e)) for e in case["inputs"][0]["data"]], dtype=np.float64)


def case_arrays(case):
    dt = NP[case["dtype"]]
    return [np.array(i["data"], dtype=dt) for i in case["inputs"]], [orc.full_shape(i["shape"]) for i in case["inputs"]]


def oracle_run(op, arrs, shapes, attrs, out_dtype=None):
    """-> (flat ndarray, shape8)"""
    a, s = arrs, shapes
    if op in orc.UNARY:
        return orc.unary(op, a[0]), s[0]
    if op in ("ADD", "MUL"):
        return orc.nnary(op, a), s[0]
    if op in orc.BINARY:
        return orc.binary(op, a[0], a[1]), s[0]
    if op == "SELECT":
        return orc.select(a[0], a[1], a[2]), s[0]
    if op == "CAST":
        return orc.cast(a[0], orc.DTYPE_CODE[np.dtype(out_dtype)]), s[0]
    if op.startswith("REDUCE"):
        return orc.reduce(op, a[0], s[0], attrs["rank_set"])
    if op == "ARGMAX":
        return orc.argmax(a[0], s[0], attrs["rank"])
    if op == "EXTEND":
        return orc.extend(a[0], s[0], attrs["dimensions"])
    if op == "PERMUTE":
        return orc.permute(a[0], s[0], attrs["ranks"])
    if op == "SLICE":
        return orc.slice_(a[0], s[0], attrs["dimension_pairs"])
    if op == "PAD":
        return orc.pad(a[0], s[0], attrs["dimension_pairs"])
    if op == "STRIDE":
        return orc.stride(a[0], s[0], attrs["dimensions"])
    if op == "SCATTER":
        return orc.scatter(a[0], s[0], attrs["shape"], attrs["dimensions"])
    if op == "REVERSE":
        return orc.reverse(a[0], s[0], attrs["rank_set"])
    if op == "CONCAT":
        return orc.concat(a, s, attrs["rank"])
    if op == "MATMUL":
        return orc.matmul(a[0], s[0], a[1], s[1])
    if op == "CONTRACT":
        return orc.contract(a[0], s[0], a[1], s[1], attrs["rank_pairs"])
    if op == "CONV":
        return orc.conv(a[0], s[0], a[1], s[1], attrs["ranks"])
    if op.startswith("ASSIGN"):
        t = a[0].copy()
        return orc.assign(op, t, a[1]), s[0]
    raise ValueError(op)


def _i64(vals):
    return (C.c_int64 * 8)(*[int(v) for v in vals])


def _i32(vals):
    return (C.c_int32 * 8)(*[int(v) for v in vals])


def gpu_run(cabi, op, arrs, shapes, attrs, out_dtype=None, precision=0):
    """Run one reference opcode through the C-ABI on device buffers -> (flat ndarray, shape8)."""
    lib = cabi.lib()
    dt = arrs[0].dtype
    code = cabi.DTYPE_OF[np.dtype(dt)]
    es = dt.itemsize
    bufs = [cabi.to_device(a) for a in arrs]
    ptr = [C.c_void_p(b.ptr) for b in bufs]
    s = shapes
    n0 = arrs[0].size
    opc = cabi.OP[op]

    def out_buf(n, d=dt):
        return cabi.empty(n, d)

    if op in orc.UNARY:
        o = out_buf(n0)
        cabi.check(lib.tcr_unary(opc, ptr[0], C.c_void_p(o.ptr), C.c_int64(n0), code))
        return cabi.to_host(o, n0, dt), s[0]
    if op in ("ADD", "MUL") and len(arrs) != 2:
        o = out_buf(n0)
        tab = (C.c_void_p * len(arrs))(*[b.ptr for b in bufs])
        cabi.check(lib.tcr_nnary(opc, tab, len(arrs), C.c_void_p(o.ptr), C.c_int64(n0), code))
        return cabi.to_host(o, n0, dt), s[0]
    if op in orc.BINARY:
        o = out_buf(n0)
        cabi.check(lib.tcr_binary(opc, ptr[0], ptr[1], C.c_void_p(o.ptr), C.c_int64(n0), code))
        return cabi.to_host(o, n0, dt), s[0]
    if op == "SELECT":
        o = out_buf(n0)
        cabi.check(lib.tcr_select(ptr[0], ptr[1], ptr[2], C.c_void_p(o.ptr), C.c_int64(n0), code))
        return cabi.to_host(o, n0, dt), s[0]
    if op == "CAST":
        od = np.dtype(out_dtype)
        o = out_buf(n0, od)
        cabi.check(lib.tcr_cast(ptr[0], code, C.c_void_p(o.ptr), cabi.DTYPE_OF[od], C.c_int64(n0)))
        return cabi.to_host(o, n0, od), s[0]
    if op.startswith("REDUCE"):
        mask = 0
        oshape = list(s[0])
        for r in attrs["rank_set"]:
            if r < 8:
                mask |= 1 << r
                oshape[r] = 1
        n = orc.n_elems(oshape)
        o = out_buf(n)
        cabi.check(lib.tcr_reduce(opc, ptr[0], C.c_void_p(o.ptr), _i64(s[0]), C.c_uint32(mask), code))
        return cabi.to_host(o, n, dt), oshape
    if op == "ARGMAX":
        rd = attrs["rank"]
        oshape = [1] * 8 if rd >= 8 else [1 if r == rd else d for r, d in enumerate(s[0])]
        n = orc.n_elems(oshape)
        o = out_buf(n)
        cabi.check(lib.tcr_argmax(ptr[0], C.c_void_p(o.ptr), _i64(s[0]), int(rd), code))
        return cabi.to_host(o, n, dt), oshape
    if op == "EXTEND":
        bc = orc.full_shape(attrs["dimensions"])
        oshape = [a * b for a, b in zip(s[0], bc)]
        n = orc.n_elems(oshape)
        o = out_buf(n)
        cabi.check(lib.tcr_extend(ptr[0], C.c_void_p(o.ptr), _i64(s[0]), _i64(bc), es))
        return cabi.to_host(o, n, dt), oshape
    if op == "PERMUTE":
        order = orc.complete_order(attrs["ranks"])
        oshape = [s[0][order[r]] for r in range(8)]
        o = out_buf(n0)
        cabi.check(lib.tcr_permute(ptr[0], C.c_void_p(o.ptr), _i64(s[0]), _i32(order), es))
        return cabi.to_host(o, n0, dt), oshape
    if op == "SLICE":
        offs, exts = [0] * 8, list(s[0])
        for r, (off, ext) in enumerate(attrs["dimension_pairs"][:8]):
            off = min(int(off), s[0][r] - 1)
            offs[r], exts[r] = off, min(int(ext), s[0][r] - off)
        n = orc.n_elems(exts)
        o = out_buf(n)
        cabi.check(lib.tcr_slice(ptr[0], C.c_void_p(o.ptr), _i64(s[0]), _i64(offs), _i64(exts), es))
        return cabi.to_host(o, n, dt), exts
    if op == "PAD":
        lo, hi = [0] * 8, [0] * 8
        for r, (l, h) in enumerate(attrs["dimension_pairs"][:8]):
            lo[r], hi[r] = int(l), int(h)
        oshape = [d + l + h for d, l, h in zip(s[0], lo, hi)]
        n = orc.n_elems(oshape)
        o = out_buf(n)
        cabi.check(lib.tcr_pad(ptr[0], C.c_void_p(o.ptr), _i64(s[0]), _i64(lo), _i64(hi), es))
        return cabi.to_host(o, n, dt), oshape
    if op == "STRIDE":
        inc = orc.full_shape(attrs["dimensions"])
        oshape = [(d + i - 1) // i for d, i in zip(s[0], inc)]
        n = orc.n_elems(oshape)
        o = out_buf(n)
        cabi.check(lib.tcr_stride(ptr[0], C.c_void_p(o.ptr), _i64(s[0]), _i64(inc), es))
        return cabi.to_host(o, n, dt), oshape
    if op == "SCATTER":
        inc = orc.full_shape(attrs["dimensions"])
        oshape = orc.full_shape(attrs["shape"])
        n = orc.n_elems(oshape)
        o = out_buf(n)
        cabi.check(lib.tcr_scatter(ptr[0], C.c_void_p(o.ptr), _i64(s[0]), _i64(oshape), _i64(inc), es))
        return cabi.to_host(o, n, dt), oshape
    if op == "REVERSE":
        mask = 0
        for r in attrs["rank_set"]:
            mask |= 1 << r
        o = out_buf(n0)
        cabi.check(lib.tcr_reverse(ptr[0], C.c_void_p(o.ptr), _i64(s[0]), C.c_uint32(mask), es))
        return cabi.to_host(o, n0, dt), s[0]
    if op == "CONCAT":
        axis = attrs["rank"]
        oshape = list(s[0])
        oshape[axis] = sum(sh[axis] for sh in s)
        n = orc.n_elems(oshape)
        o = out_buf(n)
        tab = (C.c_void_p * len(arrs))(*[b.ptr for b in bufs])
        shp = (C.c_int64 * (8 * len(arrs)))(*[int(d) for sh in s for d in sh])
        cabi.check(lib.tcr_concat(tab, shp, len(arrs), C.c_void_p(o.ptr), int(axis), es))
        return cabi.to_host(o, n, dt), oshape
    if op == "MATMUL":
        # C[N,M,batch] = A[K,M,batch] . B[N,K,batch]: row-major (MxK)(KxN) (operator.hpp:1108-1139)
        K, M = s[0][0], s[0][1]
        N = s[1][0]
        batch = orc.n_elems(s[0][2:])
        oshape = [N, M] + list(s[0][2:])
        n = orc.n_elems(oshape)
        o = out_buf(n)
        d = cabi.GemmDesc(m=M, n=N, k=K, batch=batch, a_sm=K, a_sk=1, a_sb=M * K, b_sk=N, b_sn=1, b_sb=K * N,
                          c_sm=N, c_sn=1, c_sb=M * N, dtype=code, precision=precision)
        cabi.check(lib.tcr_gemm(ptr[0], ptr[1], C.c_void_p(o.ptr), C.byref(d)))
        return cabi.to_host(o, n, dt), orc.full_shape(oshape)
    if op == "CONTRACT":
        pairs = attrs["rank_pairs"]
        acom = {p for p, _ in pairs}
        bcom = {q for _, q in pairs}
        oshape = [s[1][r] for r in range(8) if r not in bcom and s[1][r] != 1]
        oshape += [s[0][r] for r in range(8) if r not in acom and s[0][r] != 1]
        n = orc.n_elems(oshape)
        o = out_buf(n)
        fl = (C.c_int32 * (2 * len(pairs)))(*[int(v) for p in pairs for v in p])
        cabi.check(lib.tcr_contract(ptr[0], ptr[1], C.c_void_p(o.ptr), _i64(s[0]), _i64(s[1]), fl, len(pairs), code))
        return cabi.to_host(o, n, dt), orc.full_shape(oshape)
    if op == "CONV":
        order = orc.complete_order(attrs["ranks"])
        oshape = list(s[0])
        for i in range(8):
            oshape[order[i]] = s[0][order[i]] - s[1][i] + 1
        n = orc.n_elems(oshape)
        o = out_buf(n)
        cabi.check(lib.tcr_conv(ptr[0], ptr[1], C.c_void_p(o.ptr), _i64(s[0]), _i64(s[1]), _i32(order), code))
        return cabi.to_host(o, n, dt), oshape
    if op.startswith("ASSIGN"):
        cabi.check(lib.tcr_assign(opc, ptr[0], ptr[1], C.c_int64(n0), code))
        return cabi.to_host(bufs[0], n0, dt), s[0]
    raise ValueError(op)
