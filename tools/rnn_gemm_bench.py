"""Latency of tcr_gemm_grouped on the C4 shapes, measured the way the step uses it: a chain of 128 dependent launches
captured in one CUDA graph (time step t reads the h_{t-1} the previous launch wrote). Prints us per launch.

    python tools/rnn_gemm_bench.py            # sweeps cluster sizes through TCR_RNN_CLUSTER in sub-processes
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(config, precision, steps=128):
    from tenncor_b200 import cabi
    cabi.init(0)
    lib = cabi.lib()
    rng = np.random.default_rng(0)
    B, N, H = 64, 128, 1024
    keep = []

    def dev(a):
        b = cabi.to_device(np.ascontiguousarray(a, np.float32))
        keep.append(b)
        return b

    descs = []
    if config == "fwd":
        W = [dev(rng.uniform(-0.05, 0.05, (N + H, H))) for _ in range(4)]
        bias = [dev(np.zeros(H)) for _ in range(4)]
        x = dev(rng.uniform(0, 1, (steps, B, N)))
        h = [cabi.empty(B * H, np.float32) for _ in range(2)]
        c = [cabi.empty(B * H, np.float32) for _ in range(2)]
        gates = [cabi.empty(B * H, np.float32) for _ in range(4)]
        for buf in h + c:
            cabi.check(lib.tcr_memset(C.c_void_p(buf.ptr), 0, C.c_size_t(B * H * 4)))
        for t in range(steps):
            d = cabi.GemmGroupDesc()
            d.m, d.n, d.groups, d.segments = B, H, 4, 2
            d.seg_k[0], d.seg_k[1] = N, H
            d.a[0], d.a_pitch[0] = x.ptr + 4 * t * B * N, N
            d.a[1], d.a_pitch[1] = h[t % 2].ptr, H
            for g in range(4):
                d.b[g][0] = W[g].ptr
                d.b[g][1] = W[g].ptr + 4 * N * H
                d.bias[g] = bias[g].ptr
                d.act[g] = cabi.OP["TANH"] if g == 0 else cabi.OP["SIGMOID"]
                d.out[g] = gates[g].ptr
            d.b_pitch, d.b_trans, d.precision, d.out_pitch = H, 0, precision, H
            d.cell, d.role_cand, d.role_in, d.role_forget, d.role_out = 1, 0, 1, 2, 3
            d.c_prev, d.c_out, d.h_out, d.state_pitch = c[t % 2].ptr, c[(t + 1) % 2].ptr, h[(t + 1) % 2].ptr, H
            descs.append(d)
    else:
        W = [dev(rng.uniform(-0.05, 0.05, (N + H, H))) for _ in range(4)]
        dpre = [dev(rng.uniform(-1, 1, (B, H))) for _ in range(4)]
        dh = [cabi.empty(B * H, np.float32) for _ in range(2)]
        for t in range(steps):
            d = cabi.GemmGroupDesc()
            d.m, d.n, d.groups, d.segments = B, H, 1, 4
            for s in range(4):
                d.seg_k[s] = H
                # the chain: segment 0 reads what the previous launch produced
                d.a[s], d.a_pitch[s] = (dh[t % 2].ptr if s == 0 else dpre[s].ptr), H
                d.b[0][s] = W[s].ptr + 4 * N * H
            d.b_pitch, d.b_trans, d.precision, d.out_pitch = H, 1, precision, H
            d.out[0] = dh[(t + 1) % 2].ptr
            descs.append(d)
        cabi.check(lib.tcr_memset(C.c_void_p(dh[0].ptr), 0, C.c_size_t(B * H * 4)))

    def launch_all():
        for d in descs:
            cabi.check(lib.tcr_gemm_grouped(C.byref(d)))

    launch_all()
    cabi.sync()
    if os.environ.get("TCR_RNN_DEBUG"):
        cabi.check(lib.tcr_gemm_grouped(C.byref(descs[5])))
        st = (C.c_longlong * 16)()
        cabi.check(lib.tcr_rnn_debug_read(st))
        t = [int(x) for x in st[:11]]
        names = ["entry", "setup", "tma0", "tmaN", "land0", "mmaN", "acc", "sent", "xchg", "stored", "exit"]
        print("stamps (SM cycles from entry): " + ", ".join("%s=%d" % (n, x - t[0]) for n, x in zip(names, t)), flush=True)
    cabi.check(lib.tcr_graph_begin())
    launch_all()
    g = C.c_void_p()
    cabi.check(lib.tcr_graph_end(C.byref(g)))
    e0, e1 = C.c_void_p(), C.c_void_p()
    cabi.check(lib.tcr_event_create(C.byref(e0)))
    cabi.check(lib.tcr_event_create(C.byref(e1)))
    for _ in range(3):
        cabi.check(lib.tcr_graph_launch(g))
    cabi.sync()
    cabi.check(lib.tcr_event_record(e0))
    reps = 10
    for _ in range(reps):
        cabi.check(lib.tcr_graph_launch(g))
    cabi.check(lib.tcr_event_record(e1))
    ms = C.c_float()
    cabi.check(lib.tcr_event_elapsed_ms(e0, e1, C.byref(ms)))
    cabi.check(lib.tcr_graph_destroy(g))
    return ms.value * 1e3 / reps / steps


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print(json.dumps({"config": sys.argv[1], "precision": int(sys.argv[2]), "cluster": os.environ.get("TCR_RNN_CLUSTER", "auto"),
                          "us_per_launch": round(one(sys.argv[1], int(sys.argv[2])), 3)}), flush=True)
    else:
        for config in ("fwd", "bwd"):
            for prec in (1, 2):
                for cl in ("auto", "1", "2", "4", "8", "16"):
                    env = dict(os.environ)
                    if cl != "auto":
                        env["TCR_RNN_CLUSTER"] = cl
                    r = subprocess.run([sys.executable, __file__, config, str(prec)], env=env, capture_output=True, text=True, timeout=120)
                    print(r.stdout.strip() or ("FAILED " + config + " " + cl + " " + r.stderr[-300:]), flush=True)
