"""T chained LSTM gate steps: per-step tcr_gemm_grouped launches against ONE tcr_gemm_grouped_seq launch (bit-identical results
expected: same kernel code per step), then the time of both as CUDA-graph replays.

    python tools/rnn_seq_check.py [steps=128] [precision=2] [batch=64] [vocab=128] [hidden=1024]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tenncor_b200 import cabi  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    precision = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    N = int(sys.argv[4]) if len(sys.argv) > 4 else 128
    H = int(sys.argv[5]) if len(sys.argv) > 5 else 1024
    cabi.init(0)
    lib = cabi.lib()
    rng = np.random.default_rng(0)
    keep = []

    def dev(a):
        b = cabi.to_device(np.ascontiguousarray(a, np.float32))
        keep.append(b)
        return b

    W = [dev(rng.uniform(-0.05, 0.05, (N + H, H))) for _ in range(4)]
    bias = [dev(rng.uniform(-0.1, 0.1, H)) for _ in range(4)]
    x = dev(rng.uniform(0, 1, (T, B, N)))
    zeros = dev(np.zeros((B, H)))
    hs = cabi.empty(T * B * H, np.float32)
    cs = cabi.empty(T * B * H, np.float32)
    gates = [cabi.empty(T * B * H, np.float32) for _ in range(4)]
    slab = 4 * B * H
    descs = (cabi.GemmGroupDesc * T)()
    for t in range(T):
        d = descs[t]
        d.m, d.n, d.groups, d.segments = B, H, 4, 2
        d.seg_k[0], d.seg_k[1] = N, H
        d.a[0], d.a_pitch[0] = x.ptr + 4 * t * B * N, N
        d.a[1], d.a_pitch[1] = (zeros.ptr if t == 0 else hs.ptr + (t - 1) * slab), H
        for g in range(4):
            d.b[g][0] = W[g].ptr
            d.b[g][1] = W[g].ptr + 4 * N * H
            d.bias[g] = bias[g].ptr
            d.act[g] = cabi.OP["TANH"] if g == 0 else cabi.OP["SIGMOID"]
            d.out[g] = gates[g].ptr + t * slab if t % 2 == 0 else None  # odd steps do not keep their gate activations
        d.b_pitch, d.b_trans, d.precision, d.out_pitch = H, 0, precision, H
        d.cell, d.role_cand, d.role_in, d.role_forget, d.role_out = 1, 0, 1, 2, 3
        d.c_prev = zeros.ptr if t == 0 else cs.ptr + (t - 1) * slab
        d.c_out, d.h_out, d.state_pitch = cs.ptr + t * slab, hs.ptr + t * slab, H

    def clear():
        for buf in [hs, cs] + gates:
            cabi.check(lib.tcr_memset(C.c_void_p(buf.ptr), 0xff, C.c_size_t(T * slab)))

    def per_step():
        for t in range(T):
            cabi.check(lib.tcr_gemm_grouped(C.byref(descs[t])))

    clear()
    per_step()
    cabi.sync()
    want = [cabi.to_host(b, T * B * H, np.float32).copy() for b in [hs, cs] + gates]
    handle = C.c_void_p()
    cabi.check(lib.tcr_gemm_grouped_seq_prepare(descs, T, C.byref(handle)))
    ok = True
    for rep in range(3):  # the grid-barrier counters reset themselves: repeated launches
        clear()
        cabi.check(lib.tcr_gemm_grouped_seq_launch(handle))
        cabi.sync()
        got = [cabi.to_host(b, T * B * H, np.float32) for b in [hs, cs] + gates]
        same = [bool(np.array_equal(a.view(np.uint32), b.view(np.uint32))) for a, b in zip(want, got)]
        ok = ok and all(same)
        if not all(same):
            bad = [int((a.view(np.uint32) != b.view(np.uint32)).sum()) for a, b in zip(want, got)]
            first = [int(np.argmax(a.view(np.uint32) != b.view(np.uint32))) // (B * H) if n else -1 for a, b, n in zip(want, got, bad)]
            print("rep", rep, "mismatching elements (h, c, gates)", bad, "first bad step", first, flush=True)

    e0, e1 = C.c_void_p(), C.c_void_p()
    cabi.check(lib.tcr_event_create(C.byref(e0)))
    cabi.check(lib.tcr_event_create(C.byref(e1)))

    def timed(fn, reps=5):
        g = C.c_void_p()
        cabi.check(lib.tcr_graph_begin())
        fn()
        cabi.check(lib.tcr_graph_end(C.byref(g)))
        for _ in range(2):
            cabi.check(lib.tcr_graph_launch(g))
        cabi.sync()
        cabi.check(lib.tcr_event_record(e0))
        for _ in range(reps):
            cabi.check(lib.tcr_graph_launch(g))
        cabi.check(lib.tcr_event_record(e1))
        ms = C.c_float()
        cabi.check(lib.tcr_event_elapsed_ms(e0, e1, C.byref(ms)))
        cabi.check(lib.tcr_graph_destroy(g))
        return ms.value * 1e3 / reps / T

    us_step = timed(per_step)
    us_seq = timed(lambda: cabi.check(lib.tcr_gemm_grouped_seq_launch(handle)))
    cabi.check(lib.tcr_gemm_grouped_seq_destroy(handle))
    print(json.dumps({"steps": T, "precision": precision, "batch": B, "hidden": H, "bit_identical": ok, "us_per_step_separate_launches": round(us_step, 3),
                      "us_per_step_one_launch": round(us_seq, 3)}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
