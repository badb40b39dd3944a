"""Per-kernel micro-benchmarks (M-ew, M-red, M-lay, M-mm) through the C-ABI.
Prints one JSON line per case: achieved GB/s (algorithmic bytes / CUDA-event time) or TFLOP/s.
Inputs are larger than L2 (126 MB) for the HBM-bound cases so no flush is needed.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tenncor_b200 import cabi  # noqa: E402


def timeit(fn, warmup=3, iters=10):
    lib = cabi.lib()
    e0, e1 = C.c_void_p(), C.c_void_p()
    cabi.check(lib.tcr_event_create(C.byref(e0)))
    cabi.check(lib.tcr_event_create(C.byref(e1)))
    for _ in range(warmup):
        fn()
    cabi.sync()
    cabi.check(lib.tcr_event_record(e0))
    for _ in range(iters):
        fn()
    cabi.check(lib.tcr_event_record(e1))
    ms = C.c_float()
    cabi.check(lib.tcr_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=26)
    ap.add_argument("--only", default="")
    ap.add_argument("--peaks", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))
    args = ap.parse_args()
    peak = 6547.8
    if os.path.exists(args.peaks):
        peak = json.load(open(args.peaks)).get("hbm_gbs", peak)
    cabi.init(0)
    lib = cabi.lib()
    n = 1 << args.log2n
    rng = np.random.default_rng(5)
    host = rng.uniform(-4, 4, n).astype(np.float32)
    a, b, c = cabi.to_device(host), cabi.to_device(host[::-1].copy()), cabi.to_device(np.abs(host))
    out = cabi.empty(n, np.float32)
    F = cabi.FLOAT
    P = lambda x: C.c_void_p(x.ptr)  # noqa: E731
    res = []

    def rec(name, ms, nbytes=None, flops=None):
        r = {"case": name, "ms": round(ms, 4)}
        if nbytes is not None:
            r["GBps"] = round(nbytes / ms / 1e6, 1)
            r["frac_of_measured_hbm"] = round(r["GBps"] / peak, 3)
        if flops is not None:
            r["TFLOPs"] = round(flops / ms / 1e9, 2)
        res.append(r)
        print(json.dumps(r), flush=True)

    def want(name):
        return not args.only or any(name.startswith(p) for p in args.only.split(","))

    if want("ew"):
        for op in ["EXP", "SIGMOID", "TANH", "NEG"]:
            rec("ew_unary_%s_2^%d" % (op, args.log2n), timeit(lambda: cabi.check(lib.tcr_unary(cabi.OP[op], P(a), P(out), C.c_int64(n), F))), 8 * n)
        for op in ["ADD", "MUL"]:
            rec("ew_binary_%s_2^%d" % (op, args.log2n), timeit(lambda: cabi.check(lib.tcr_binary(cabi.OP[op], P(a), P(b), P(out), C.c_int64(n), F))), 12 * n)
        prog = cabi.make_program(F, (n, 1, 1), [(a.ptr, F, (0, 0, 0)), (b.ptr, F, (0, 0, 0)), (c.ptr, F, (0, 0, 0))], [(out.ptr, F, 0)],
                                 [(cabi.OP["MUL"], 0, 0, 1), (cabi.OP["ADD"], 0, 0, 2), (cabi.OP["SIGMOID"], 0, 0)])
        rec("ew_fused_sigmoid(a*b+c)_2^%d" % args.log2n, timeit(lambda: cabi.check(lib.tcr_elementwise(C.byref(prog)))), 16 * n)
        prog3 = cabi.make_program(F, (n, 1, 1), [(a.ptr, F, (0, 0, 0)), (b.ptr, F, (0, 0, 0)), (c.ptr, F, (0, 0, 0))], [(out.ptr, F, 0)],
                                  [(cabi.OP["MUL"], 0, 0, 1), (cabi.OP["ADD"], 0, 0, 2)])
        rec("ew_fused_a*b+c_2^%d" % args.log2n, timeit(lambda: cabi.check(lib.tcr_elementwise(C.byref(prog3)))), 16 * n)
        rec("ew_assign_sub_2^%d" % args.log2n, timeit(lambda: cabi.check(lib.tcr_assign(cabi.OP["ASSIGN_SUB"], P(out), P(a), C.c_int64(n), F))), 12 * n)
        H, B = 1024, n // 1024
        bias = cabi.to_device(host[:H].copy())
        prog2 = cabi.make_program(F, (H, B, 1), [(a.ptr, F, (0, 0, 0)), (bias.ptr, F, (0, 1, 0))], [(out.ptr, F, 0)],
                                  [(cabi.OP["ADD"], 0, 0, 1), (cabi.OP["SIGMOID"], 0, 0)])
        rec("ew_fused_bias_sigmoid_[1024,%d]" % B, timeit(lambda: cabi.check(lib.tcr_elementwise(C.byref(prog2)))), 8 * n + 4 * H)
    if want("red"):
        R0, R1 = 4096, n // 4096
        shape = cabi.shape8([R0, R1])
        small = cabi.empty(max(R0, R1), np.float32)
        for op in ["REDUCE_SUM", "REDUCE_MAX"]:
            for mask, nm, nout in [(1, "dim0", R1), (2, "dim1", R0), (3, "full", 1)]:
                rec("red_%s_%s_[4096,%d]" % (op, nm, R1), timeit(lambda: cabi.check(lib.tcr_reduce(cabi.OP[op], P(a), P(small), shape, C.c_uint32(mask), F))), 4 * (n + nout))
        for dim, nm, nout in [(0, "dim0", R1), (1, "dim1", R0), (8, "flat", 1)]:
            rec("argmax_%s_[4096,%d]" % (nm, R1), timeit(lambda: cabi.check(lib.tcr_argmax(P(a), P(small), shape, dim, F))), 4 * (n + nout))
    if want("lay"):
        side = 1 << (args.log2n // 2)
        m = side * side
        order = (C.c_int32 * 8)(1, 0, 2, 3, 4, 5, 6, 7)
        rec("lay_permute10_[%d,%d]" % (side, side), timeit(lambda: cabi.check(lib.tcr_permute(P(a), P(out), cabi.shape8([side, side]), order, 4))), 8 * m)
        bc = (C.c_int64 * 8)(1, n // 1024, 1, 1, 1, 1, 1, 1)
        rec("lay_extend_[1024]->[1024,%d]" % (n // 1024), timeit(lambda: cabi.check(lib.tcr_extend(P(a), P(out), cabi.shape8([1024]), bc, 4))), 4 * n + 4096)
        s3 = [1024, 128, n // (1024 * 128)]
        offs = (C.c_int64 * 8)(0, 32, 0, 0, 0, 0, 0, 0)
        exts = (C.c_int64 * 8)(1024, 64, s3[2], 1, 1, 1, 1, 1)
        rec("lay_slice_mid_[1024,128,%d]" % s3[2], timeit(lambda: cabi.check(lib.tcr_slice(P(a), P(out), cabi.shape8(s3), offs, exts, 4))), 8 * (n // 2))
        lo = (C.c_int64 * 8)(0, 16, 0, 0, 0, 0, 0, 0)
        hi = (C.c_int64 * 8)(0, 16, 0, 0, 0, 0, 0, 0)
        s3h = [1024, 96, s3[2]]
        rec("lay_pad_mid_[1024,96,%d]" % s3[2], timeit(lambda: cabi.check(lib.tcr_pad(P(a), P(out), cabi.shape8(s3h), lo, hi, 4))), 4 * (1024 * 96 * s3[2] + n))
        half = [1024, 64, s3[2]]
        tab = (C.c_void_p * 2)(a.ptr, b.ptr)
        shp = (C.c_int64 * 16)(*(cabi.shape8(half)[:] + cabi.shape8(half)[:]))
        rec("lay_concat_axis1_[1024,64,%d]x2" % s3[2], timeit(lambda: cabi.check(lib.tcr_concat(tab, shp, 2, P(out), 1, 4))), 8 * n)
    if want("survey"):
        # the exact shapes SURVEY.md §8(d) lists that the n-derived cases above do not hit
        R0, R1 = 4096, 65536
        big = cabi.to_device(rng.uniform(0.9, 1.1, R0 * R1).astype(np.float32))
        small2 = cabi.empty(R1, np.float32)
        shp = cabi.shape8([R0, R1])
        nn = R0 * R1
        for op in ["REDUCE_SUM", "REDUCE_MAX", "REDUCE_MIN", "REDUCE_PROD"]:
            for mask, nm, nout in [(1, "dim0", R1), (2, "dim1", R0), (3, "full", 1)]:
                rec("red_%s_%s_[4096,65536]" % (op, nm), timeit(lambda: cabi.check(lib.tcr_reduce(cabi.OP[op], P(big), P(small2), shp, C.c_uint32(mask), F)), iters=5), 4 * (nn + nout))
        for dim, nm, nout in [(0, "dim0", R1), (1, "dim1", R0), (8, "flat", 1)]:
            rec("argmax_%s_[4096,65536]" % nm, timeit(lambda: cabi.check(lib.tcr_argmax(P(big), P(small2), shp, dim, F)), iters=5), 4 * (nn + nout))
        big2 = cabi.empty(nn, np.float32)
        for op in ["EXP", "SIGMOID", "TANH"]:
            rec("ew_unary_%s_2^28" % op, timeit(lambda: cabi.check(lib.tcr_unary(cabi.OP[op], P(big), P(big2), C.c_int64(nn), F)), iters=5), 8 * nn)
        big3 = cabi.to_device(rng.uniform(-1, 1, nn).astype(np.float32))
        big4 = cabi.to_device(rng.uniform(-1, 1, nn).astype(np.float32))
        for op in ["ADD", "MUL"]:
            rec("ew_binary_%s_2^28" % op, timeit(lambda: cabi.check(lib.tcr_binary(cabi.OP[op], P(big), P(big3), P(big2), C.c_int64(nn), F)), iters=5), 12 * nn)
        progb = cabi.make_program(F, (nn, 1, 1), [(big.ptr, F, (0, 0, 0)), (big3.ptr, F, (0, 0, 0)), (big4.ptr, F, (0, 0, 0))], [(big2.ptr, F, 0)],
                                  [(cabi.OP["MUL"], 0, 0, 1), (cabi.OP["ADD"], 0, 0, 2), (cabi.OP["SIGMOID"], 0, 0)])
        rec("ew_fused_sigmoid(a*b+c)_2^28", timeit(lambda: cabi.check(lib.tcr_elementwise(C.byref(progb))), iters=5), 16 * nn)
        rec("ew_assign_sub_2^28", timeit(lambda: cabi.check(lib.tcr_assign(cabi.OP["ASSIGN_SUB"], P(big2), P(big), C.c_int64(nn), F)), iters=5), 12 * nn)
        bc = (C.c_int64 * 8)(1, 65536, 1, 1, 1, 1, 1, 1)
        rec("lay_extend_[1024]->[1024,65536]", timeit(lambda: cabi.check(lib.tcr_extend(P(big), P(big2), cabi.shape8([1024]), bc, 4))), 4 * 1024 * 65536 + 4096)
        s3 = [1024, 128, 64]
        m3 = 1024 * 128 * 64
        offs = (C.c_int64 * 8)(0, 32, 0, 0, 0, 0, 0, 0)
        exts = (C.c_int64 * 8)(1024, 64, 64, 1, 1, 1, 1, 1)
        rec("lay_slice_mid_[1024,128,64] (L2-resident)", timeit(lambda: cabi.check(lib.tcr_slice(P(big), P(big2), cabi.shape8(s3), offs, exts, 4))), 8 * (m3 // 2))
        lo = (C.c_int64 * 8)(0, 16, 0, 0, 0, 0, 0, 0)
        rec("lay_pad_mid_[1024,128,64] (L2-resident)", timeit(lambda: cabi.check(lib.tcr_pad(P(big), P(big2), cabi.shape8(s3), lo, lo, 4))), 4 * (m3 + 1024 * 160 * 64))
        tab = (C.c_void_p * 2)(big.ptr, big.ptr)
        shp2 = (C.c_int64 * 16)(*(cabi.shape8(s3)[:] + cabi.shape8(s3)[:]))
        rec("lay_concat_axis1_[1024,128,64]x2 (L2-resident)", timeit(lambda: cabi.check(lib.tcr_concat(tab, shp2, 2, P(big2), 1, 4))), 16 * m3)
        for prec, nm in [(1, "tf32"), (2, "3xtf32")]:
            M = N = K = 8192
            d = cabi.GemmDesc(m=M, n=N, k=K, batch=1, a_sm=K, a_sk=1, a_sb=0, b_sk=N, b_sn=1, b_sb=0, c_sm=N, c_sn=1, c_sb=0, dtype=F, precision=prec)
            rec("mm_%s_8192^3" % nm, timeit(lambda: cabi.check(lib.tcr_gemm(P(big), P(big), P(big2), C.byref(d))), iters=3), flops=2.0 * M * N * K)
    if want("mm"):
        for prec, nm in [(0, "exact_simt"), (1, "tf32"), (2, "3xtf32")]:
            for s in [1024, 4096]:
                if prec == 0 and s > 2048:
                    continue
                M = N = K = s
                d = cabi.GemmDesc(m=M, n=N, k=K, batch=1, a_sm=K, a_sk=1, a_sb=0, b_sk=N, b_sn=1, b_sb=0, c_sm=N, c_sn=1, c_sb=0,
                                  dtype=F, precision=prec)
                rec("mm_%s_%d^3" % (nm, s), timeit(lambda: cabi.check(lib.tcr_gemm(P(a), P(b), P(out), C.byref(d))), iters=5), flops=2.0 * M * N * K)
        # operand majors (the backward contractions are the TN / NT forms) and the c3 training shapes
        for prec, nm in [(1, "tf32"), (2, "3xtf32")]:
            for tag, (M, N, K), ta, tb in [("NN_4096^3", (4096, 4096, 4096), 0, 0), ("TN_4096^3", (4096, 4096, 4096), 1, 0),
                                           ("NT_4096^3", (4096, 4096, 4096), 0, 1), ("TT_4096^3", (4096, 4096, 4096), 1, 1),
                                           ("c3_fwd_NN_8192x1024x784", (8192, 1024, 784), 0, 0),
                                           ("c3_dW_TN_1024x784x8192", (1024, 784, 8192), 1, 0),
                                           ("c4_gate_NN_64x4096x1152", (64, 4096, 1152), 0, 0)]:
                d = cabi.GemmDesc(m=M, n=N, k=K, batch=1, a_sm=1 if ta else K, a_sk=M if ta else 1, b_sk=1 if tb else N, b_sn=K if tb else 1,
                                  c_sm=N, c_sn=1, dtype=F, precision=prec)
                rec("mm_%s_%s" % (nm, tag), timeit(lambda: cabi.check(lib.tcr_gemm(P(a), P(b), P(out), C.byref(d))), iters=5), flops=2.0 * M * N * K)
    if want("conv"):
        # conv2d composite as the planner lowers it (fuse_convs): patch gather (HBM-bound) + GEMM (tensor-bound), and the
        # generic CONV kernel running the reference's padded-rank formulation of the same layer for contrast (small case only)
        for tag, (inc, outc, W, Hh, B, kw, kh) in [("3x3_64->64_34x34_B64", (64, 64, 34, 34, 64, 3, 3)), ("3x3_3->64_66x66_B64", (3, 64, 66, 66, 64, 3, 3)),
                                                   ("5x5_32->128_20x20_B128", (32, 128, 20, 20, 128, 5, 5))]:
            img_shape, win = [inc, W, Hh, B], [inc, kw, kh, 1]
            rows, k = (W - kw + 1) * (Hh - kh + 1) * B, inc * kw * kh
            pitch = (k + 3) // 4 * 4
            img = cabi.to_device(rng.uniform(-1, 1, inc * W * Hh * B).astype(np.float32))
            ker = cabi.to_device(rng.uniform(-1, 1, k * outc).astype(np.float32))
            sup = cabi.to_device(rng.uniform(-1, 1, rows * outc).astype(np.float32))
            cols = cabi.empty(rows * pitch, np.float32)
            o = cabi.empty(rows * outc, np.float32)
            dk = cabi.empty(k * outc, np.float32)
            s8, w8 = cabi.shape8(img_shape), cabi.shape8(win)
            gather = lambda: cabi.check(lib.tcr_im2col(P(img), P(cols), s8, w8, C.c_int64(pitch), 4))  # noqa: E731
            rec("conv_im2col_%s" % tag, timeit(gather), 4 * (rows * pitch + inc * W * Hh * B))
            for prec, nm in [(1, "tf32"), (2, "3xtf32")]:
                d = cabi.GemmDesc(m=rows, n=outc, k=k, batch=1, a_sm=pitch, a_sk=1, b_sk=outc, b_sn=1, c_sm=outc, c_sn=1, dtype=F, precision=prec)
                fwd = lambda: cabi.check(lib.tcr_gemm(P(cols), P(ker), P(o), C.byref(d)))  # noqa: E731
                rec("conv_fwd_gemm_%s_%s" % (nm, tag), timeit(fwd), flops=2.0 * rows * outc * k)
                rec("conv_fwd_total_%s_%s" % (nm, tag), timeit(lambda: (gather(), fwd())), flops=2.0 * rows * outc * k)
                g = cabi.GemmDesc(m=k, n=outc, k=rows, batch=1, a_sm=1, a_sk=pitch, b_sk=outc, b_sn=1, c_sm=outc, c_sn=1, dtype=F, precision=prec)
                rec("conv_dK_gemm_%s_%s" % (nm, tag), timeit(lambda: cabi.check(lib.tcr_gemm(P(cols), P(sup), P(dk), C.byref(g)))), flops=2.0 * rows * outc * k)
                x = cabi.GemmDesc(m=rows, n=k, k=outc, batch=1, a_sm=outc, a_sk=1, b_sk=1, b_sn=outc, c_sm=pitch, c_sn=1, dtype=F, precision=prec)
                dx_gemm = lambda: cabi.check(lib.tcr_gemm(P(sup), P(ker), P(cols), C.byref(x)))  # noqa: E731
                rec("conv_dX_gemm_%s_%s" % (nm, tag), timeit(dx_gemm), flops=2.0 * rows * outc * k)
            scatter = lambda: cabi.check(lib.tcr_col2im(P(cols), P(img), s8, w8, C.c_int64(pitch), F))  # noqa: E731
            rec("conv_col2im_%s" % tag, timeit(scatter), 4 * (rows * pitch + inc * W * Hh * B))
        inc, outc, W, Hh, B, kw, kh = 8, 16, 18, 18, 8, 3, 3
        padded = cabi.to_device(np.zeros(inc * W * Hh * B * (2 * outc - 1), np.float32))
        kr = cabi.to_device(rng.uniform(-1, 1, outc * inc * kw * kh).astype(np.float32))
        og = cabi.empty((W - kw + 1) * (Hh - kh + 1) * B * outc, np.float32)
        order = (C.c_int32 * 8)(4, 0, 1, 2, 3, 5, 6, 7)
        rec("conv_generic_padded_rank_3x3_8->16_18x18_B8",
            timeit(lambda: cabi.check(lib.tcr_conv(P(padded), P(kr), P(og), cabi.shape8([inc, W, Hh, B, 2 * outc - 1]), cabi.shape8([outc, inc, kw, kh]), order, F))),
            flops=2.0 * (W - kw + 1) * (Hh - kh + 1) * B * outc * inc * kw * kh)
    return res


if __name__ == "__main__":
    main()
