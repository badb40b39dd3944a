"""Diagnostic: which k-blocks / tiles of the tcgen05 GEMM are wrong (prints, no asserts)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tenncor_b200 import cabi  # noqa: E402
from tests.test_gemm_tc_gpu import run_gemm  # noqa: E402

cabi.init(0)
rng = np.random.default_rng(0)
for (M, N, K) in [(128, 128, 32), (128, 128, 64), (128, 128, 96), (128, 128, 192), (128, 128, 224), (128, 128, 256), (256, 128, 32), (128, 256, 32), (256, 256, 64)]:
    for prec in (1, 2):
        for ta, tb in ((0, 0), (1, 0), (0, 1), (1, 1)):
            A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
            B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
            got = run_gemm(cabi, A, B, ta, tb, prec)
            want = A.astype(np.float64) @ B.astype(np.float64)
            err = np.abs(got - want).max()
            # per-k-block attribution: subtract each block's exact contribution and see which one is missing / doubled
            msg = ""
            if err > 0.1:
                blocks = []
                for kb in range(K // 32):
                    part = A[:, kb * 32:(kb + 1) * 32].astype(np.float64) @ B[kb * 32:(kb + 1) * 32].astype(np.float64)
                    # least-squares coefficient of this block's contribution in the result
                    coef = float((got * part).sum() / (part * part).sum())
                    blocks.append(round(coef, 2))
                msg = " block coefficients " + str(blocks)
            print("M%d N%d K%d prec%d ta%d tb%d maxerr %.3g%s" % (M, N, K, prec, ta, tb, err, msg), flush=True)
