"""Where does a stream-K product differ from the float64 product? Prints the failing 256 x 256 tiles (tile row, tile col, share of bad elements)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tenncor_b200 import cabi  # noqa: E402
from tests.test_gemm_tc_gpu import run_gemm  # noqa: E402

cabi.init(0)
M, N, K = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 1024, 784))]
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rng = np.random.default_rng(1)
A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
want = A.astype(np.float64) @ B.astype(np.float64)
S = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
bound = S * (2.0 ** -10 if prec == 1 else (2.0 ** -19 + K * 2.0 ** -23))
for rep in range(3):
    got = run_gemm(cabi, A, B, 0, 0, prec)
    bad = np.abs(got - want) > bound
    print("rep", rep, "bad elements", int(bad.sum()), "max err", float(np.abs(got - want).max()))
    if bad.any():
        tiles = {}
        for tm in range((M + 255) // 256):
            for tn in range((N + 255) // 256):
                b = bad[tm * 256:(tm + 1) * 256, tn * 256:(tn + 1) * 256]
                if b.any():
                    rows = np.where(b.any(axis=1))[0]
                    cols = np.where(b.any(axis=0))[0]
                    tiles[(tm, tn)] = (float(b.mean()), int(rows.min()), int(rows.max()), int(cols.min()), int(cols.max()))
        for k, v in list(tiles.items())[:24]:
            t = k[0] * ((N + 255) // 256) + k[1]
            print(" tile", k, "index", t, "kb-unit start", t * ((K + 31) // 32), "bad share %.3f rows %d..%d cols %d..%d" % v)
        print(" failing tiles:", len(tiles))
