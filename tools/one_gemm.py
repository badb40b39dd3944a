"""Launch ONE tcr_gemm configuration a few times (for ncu captures and quick timing).

  python tools/one_gemm.py --m 4096 --n 4096 --k 4096 --prec 1 --ta 0 --tb 0 --iters 3
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tenncor_b200 import cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=4096)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--prec", type=int, default=1, help="0 exact, 1 tf32, 2 3xtf32")
    ap.add_argument("--ta", type=int, default=0)
    ap.add_argument("--tb", type=int, default=0)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--graph", action="store_true", help="capture the launches in one CUDA graph and time its replay (GPU-side cost without CPU launch overhead)")
    args = ap.parse_args()
    cabi.init(0)
    lib = cabi.lib()
    M, N, K = args.m, args.n, args.k
    rng = np.random.default_rng(7)
    a = cabi.to_device(rng.uniform(-1, 1, M * K).astype(np.float32))
    b = cabi.to_device(rng.uniform(-1, 1, K * N).astype(np.float32))
    out = cabi.empty(M * N, np.float32)
    d = cabi.GemmDesc(m=M, n=N, k=K, batch=1, a_sm=1 if args.ta else K, a_sk=M if args.ta else 1, b_sk=1 if args.tb else N,
                      b_sn=K if args.tb else 1, c_sm=N, c_sn=1, dtype=cabi.FLOAT, precision=args.prec)
    P = lambda x: C.c_void_p(x.ptr)  # noqa: E731
    e0, e1 = C.c_void_p(), C.c_void_p()
    cabi.check(lib.tcr_event_create(C.byref(e0)))
    cabi.check(lib.tcr_event_create(C.byref(e1)))
    for _ in range(args.warmup):
        cabi.check(lib.tcr_gemm(P(a), P(b), P(out), C.byref(d)))
    cabi.sync()
    if args.graph:
        g = C.c_void_p()
        cabi.check(lib.tcr_graph_begin())
        for _ in range(args.iters):
            cabi.check(lib.tcr_gemm(P(a), P(b), P(out), C.byref(d)))
        cabi.check(lib.tcr_graph_end(C.byref(g)))
        cabi.check(lib.tcr_graph_launch(g))
        cabi.sync()
        cabi.check(lib.tcr_event_record(e0))
        cabi.check(lib.tcr_graph_launch(g))
        cabi.check(lib.tcr_event_record(e1))
    else:
        cabi.check(lib.tcr_event_record(e0))
        for _ in range(args.iters):
            cabi.check(lib.tcr_gemm(P(a), P(b), P(out), C.byref(d)))
        cabi.check(lib.tcr_event_record(e1))
    ms = C.c_float()
    cabi.check(lib.tcr_event_elapsed_ms(e0, e1, C.byref(ms)))
    per = ms.value / args.iters
    print(json.dumps({"m": M, "n": N, "k": K, "prec": args.prec, "ta": args.ta, "tb": args.tb, "ms": round(per, 4),
                      "TFLOPs": round(2.0 * M * N * K / per / 1e9, 2)}), flush=True)


if __name__ == "__main__":
    main()
