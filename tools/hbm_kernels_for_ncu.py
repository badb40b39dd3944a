"""A few launches of the round-2 HBM-bound kernels at the micro-benchmark sizes (for ncu)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tenncor_b200 import cabi  # noqa: E402

cabi.init(0)
lib = cabi.lib()
F = cabi.FLOAT
rng = np.random.default_rng(0)
n = 1 << 28
a = cabi.empty(n, np.float32)
seed = cabi.to_device(rng.uniform(-1, 1, 1 << 24).astype(np.float32))
for j in range(16):
    cabi.check(lib.tcr_d2d(C.c_void_p(a.ptr + 4 * j * (1 << 24)), C.c_void_p(seed.ptr), C.c_size_t(4 << 24)))
small = cabi.empty(65536, np.float32)
shp = cabi.shape8([4096, 65536])
for _ in range(2):
    cabi.check(lib.tcr_reduce(cabi.OP["REDUCE_SUM"], C.c_void_p(a.ptr), C.c_void_p(small.ptr), shp, C.c_uint32(2), F))
s3 = [1024, 128, 64]
out = cabi.empty(2 * 1024 * 128 * 64, np.float32)
lo = (C.c_int64 * 8)(0, 16, 0, 0, 0, 0, 0, 0)
for _ in range(2):
    cabi.check(lib.tcr_pad(C.c_void_p(a.ptr), C.c_void_p(out.ptr), cabi.shape8(s3), lo, lo, 4))
px = cabi.to_device(rng.integers(0, 256, 8192 * 784, dtype=np.uint8))
res = cabi.empty(8192 * 784, np.float32)
prog = cabi.make_program(F, (8192 * 784, 1, 1), [(px.ptr, cabi.UINT8, (0, 0, 0))], [(res.ptr, F, 0)],
                         [(cabi.EW_MOV, 1, 0), (cabi.EW_CONST, 2, 0, 0, 0, 1.0 / 255.0), (cabi.OP["MUL"], 0, 1, 2)])
for _ in range(2):
    cabi.check(lib.tcr_elementwise(C.byref(prog)))
# round 2, session 3: vector-broadcast elementwise, flat ARGMAX with 16-byte loads, SLICE through the row-window kernel
bias = cabi.to_device(rng.uniform(-1, 1, 1024).astype(np.float32))
big = cabi.empty(1 << 26, np.float32)
prog2 = cabi.make_program(F, (1024, (1 << 26) // 1024, 1), [(a.ptr, F, (0, 0, 0)), (bias.ptr, F, (0, 1, 0))], [(big.ptr, F, 0)],
                          [(cabi.OP["ADD"], 0, 0, 1), (cabi.OP["SIGMOID"], 0, 0)])
for _ in range(2):
    cabi.check(lib.tcr_elementwise(C.byref(prog2)))
for _ in range(2):
    cabi.check(lib.tcr_argmax(C.c_void_p(a.ptr), C.c_void_p(small.ptr), shp, 8, F))
offs = (C.c_int64 * 8)(0, 32, 0, 0, 0, 0, 0, 0)
exts = (C.c_int64 * 8)(1024, 64, 64, 1, 1, 1, 1, 1)
for _ in range(2):
    cabi.check(lib.tcr_slice(C.c_void_p(a.ptr), C.c_void_p(out.ptr), cabi.shape8(s3), offs, exts, 4))
cabi.sync()
print("ok")
