"""The kernels of the C3 step (MLP 784-1024-10, batch 8192) launched directly through the C-ABI, a few times each, with CUDA-event
timings: the command to put under `ncu` (few launches, no CUDA graph). `--only` selects by prefix.

    python tools/c3_kernels.py [--only skinny,dact,red,gemm] [--reps 5]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tenncor_b200 import cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    cabi.init(0)
    lib = cabi.lib()
    F = cabi.FLOAT
    rng = np.random.default_rng(0)
    B, IN, H, OUT = 8192, 784, 1024, 10
    P = lambda b: C.c_void_p(b.ptr)  # noqa: E731
    e0, e1 = C.c_void_p(), C.c_void_p()
    cabi.check(lib.tcr_event_create(C.byref(e0)))
    cabi.check(lib.tcr_event_create(C.byref(e1)))

    def timed(name, fn, nbytes=None, flops=None):
        for _ in range(2):
            fn()
        cabi.sync()
        cabi.check(lib.tcr_event_record(e0))
        for _ in range(args.reps):
            fn()
        cabi.check(lib.tcr_event_record(e1))
        ms = C.c_float()
        cabi.check(lib.tcr_event_elapsed_ms(e0, e1, C.byref(ms)))
        us = ms.value * 1e3 / args.reps
        r = {"kernel": name, "us": round(us, 2)}
        if nbytes:
            r["GBps"] = round(nbytes / us / 1e3, 1)
        if flops:
            r["TFLOPs"] = round(flops / us / 1e6, 1)
        print(json.dumps(r), flush=True)

    want = lambda n: not args.only or any(n.startswith(p) for p in args.only.split(","))  # noqa: E731
    x = cabi.to_device(rng.random((B, IN), dtype=np.float32))
    w0 = cabi.to_device(rng.uniform(-0.1, 0.1, (IN, H)).astype(np.float32))
    b0 = cabi.to_device(np.zeros(H, np.float32))
    h = cabi.to_device(rng.random((B, H), dtype=np.float32))
    w1 = cabi.to_device(rng.uniform(-0.1, 0.1, (H, OUT)).astype(np.float32))
    b1 = cabi.to_device(np.zeros(OUT, np.float32))
    out = cabi.empty(B * OUT, np.float32)
    dpre1 = cabi.to_device(rng.uniform(-1, 1, (B, OUT)).astype(np.float32))
    dh = cabi.empty(B * H, np.float32)
    dw1 = cabi.empty(H * OUT, np.float32)
    dw0 = cabi.empty(IN * H, np.float32)
    small = cabi.empty(4096, np.float32)
    if want("skinny"):
        d = cabi.GemmDesc(m=B, n=OUT, k=H, batch=1, a_sm=H, a_sk=1, b_sk=OUT, b_sn=1, c_sm=OUT, c_sn=1, dtype=F, precision=2, epilogue=cabi.EPI_BIAS_N,
                          activation=cabi.OP["SIGMOID"])
        d.bias = b1.ptr
        timed("skinny_fwd_8192x10x1024 (layer-1 forward)", lambda: cabi.check(lib.tcr_gemm(P(h), P(w1), P(out), C.byref(d))), 4 * (B * H + B * OUT))
        g = cabi.GemmDesc(m=H, n=OUT, k=B, batch=1, a_sm=1, a_sk=H, b_sk=OUT, b_sn=1, c_sm=OUT, c_sn=1, dtype=F, precision=2)
        timed("skinny_dW1_1024x10x8192 (layer-1 weight gradient)", lambda: cabi.check(lib.tcr_gemm(P(h), P(dpre1), P(dw1), C.byref(g))), 4 * (B * H + B * OUT))
    if want("dact"):
        d = cabi.GemmDesc(m=B, n=H, k=OUT, batch=1, a_sm=OUT, a_sk=1, b_sk=1, b_sn=OUT, c_sm=H, c_sn=1, dtype=F, precision=2)
        timed("smallk_dX1_8192x1024x10", lambda: cabi.check(lib.tcr_gemm(P(dpre1), P(w1), P(dh), C.byref(d))), 4 * B * H)
        d.post_op = cabi.POST_MUL_DSIGMOID
        d.aux = h.ptr
        timed("smallk_dX1_x_dsigmoid_8192x1024x10", lambda: cabi.check(lib.tcr_gemm(P(dpre1), P(w1), P(dh), C.byref(d))), 8 * B * H)
    if want("red"):
        for shape, mask, name in (([H, B], 2, "db0 [1024,8192] over batch"), ([OUT, B], 2, "db1 [10,8192] over batch"), ([OUT, B], 3, "loss [10,8192] full")):
            n = int(np.prod(shape))
            timed("reduce_sum " + name, lambda: cabi.check(lib.tcr_reduce(cabi.OP["REDUCE_SUM"], P(h), P(small), cabi.shape8(shape), C.c_uint32(mask), F)), 4 * n)
    if want("gemm"):
        for prec, nm in ((1, "tf32"), (2, "3xtf32")):
            d = cabi.GemmDesc(m=B, n=H, k=IN, batch=1, a_sm=IN, a_sk=1, b_sk=H, b_sn=1, c_sm=H, c_sn=1, dtype=F, precision=prec, epilogue=cabi.EPI_BIAS_N,
                              activation=cabi.OP["SIGMOID"])
            d.bias = b0.ptr
            timed("gemm_fwd0_%s_8192x1024x784" % nm, lambda: cabi.check(lib.tcr_gemm(P(x), P(w0), P(dh), C.byref(d))), flops=2.0 * B * H * IN)
            g = cabi.GemmDesc(m=IN, n=H, k=B, batch=1, a_sm=1, a_sk=IN, b_sk=H, b_sn=1, c_sm=H, c_sn=1, dtype=F, precision=prec)
            timed("gemm_dW0_%s_784x1024x8192" % nm, lambda: cabi.check(lib.tcr_gemm(P(x), P(h), P(dw0), C.byref(g))), flops=2.0 * B * H * IN)


if __name__ == "__main__":
    main()
