#!/usr/bin/env python
"""bench.py — train steps/sec of the TEQ-graph evaluation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c1w|c4|c4a|c2] [--impl reference]

One JSON line on stdout (rank 0). A "step" is one evaluation of the `apply_update` root of a
demo model: forward + derivative graph + optimiser ASSIGNs over one synthetic batch.

  value      steps/s with the batch already resident in HBM (CUDA events on the library stream,
             barrier + synchronize on both sides, max over ranks)
  e2e        the same metric through the public API with HOST buffers: every step assigns the
             batch from pinned host memory (H2D) and reads the loss back (D2H)
  roofline   the dominant kernel of the step, timed live with CUDA events through the C-ABI
  cpu_baseline / --impl reference
             the CPU oracle (numpy restatement of the reference's Eigen path; the reference
             itself cannot be built here) evaluating the same dumped graph node
             by node on the host cores

Default workload c3 = BASELINE.json configs[2] "MNIST-shaped MLP ... batch 65536 sharded over
8xB200" = 8192 samples per GPU (weak scaling); configs[1] (RBM) is `--workload c2`.
Inputs per step are larger than nothing cached: weights + activations exceed L2 only for c3/c4
(activations 33.5 MB x ~10 live tensors); smaller workloads say "l2_resident" in config.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (builder kwargs, description)
    "c1": dict(kind="mlp", ninput=10, nhidden=9, noutput=5, nbatch=3),
    "c1w": dict(kind="mlp", ninput=1024, nhidden=1024, noutput=512, nbatch=8192),
    # c3 feeds UINT8 pixels (the way MNIST is stored) and casts them on the device with the reference's own CAST (tenncor/eteq/caster.hpp),
    # both arms evaluate that same graph; c3f is the float-input form of round 1
    "c3": dict(kind="mlp", ninput=784, nhidden=1024, noutput=10, nbatch=8192, one_hot=True, pixels=True),
    "c3f": dict(kind="mlp", ninput=784, nhidden=1024, noutput=10, nbatch=8192, one_hot=True),
    "c2": dict(kind="rbm", nvisible=784, nhidden=64, nbatch=4096),
    # learning rate: the demo's adagrad 0.1 (demo/lstm/latin_demo.py:120-122) on a NLL summed over 8192 tokens diverges at this size
    # — the CPU oracle of the reference path reaches inf at step 3 and NaN at step 4 as well. 0.01 survives the ~50 steps
    # of a 1-GPU run but not a 2-GPU one (the bench replays ONE batch: every step pushes the weights the same way until a softmax
    # saturates and -log(0) appears), so the batched configs train with 0.001; the arithmetic per step is identical
    "c4a": dict(kind="lstm", vocab=128, hidden=1024, seq=128, batch=None, learning_rate=0.001),
    "c4": dict(kind="lstm", vocab=128, hidden=1024, seq=128, batch=64, learning_rate=0.001),
    "c4gru": dict(kind="gru", vocab=128, hidden=1024, seq=128, batch=64, learning_rate=0.001),
    "c5": dict(kind="dqn", nobs=10, nunits=9, nactions=9, nbatch=4096),
    # SURVEY §8a row a12 (CONV; not one of BASELINE's configs): one conv2d layer + sigmoid trained like gd_demo. The CPU oracle of the
    # reference's padded-rank formulation needs minutes per step at this size: run with --cpu-seconds 0 to skip that leg.
    "conv": dict(kind="conv", in_ch=32, out_ch=64, width=34, height=34, nbatch=64),
}


def build_config(name):
    from tenncor_b200 import configs
    w = dict(WORKLOADS[name])
    kind = w.pop("kind")
    one_hot = w.pop("one_hot", False)
    if kind == "mlp":
        cfg = configs.mlp(name=name, **w)
        gen = lambda rng: configs.mlp_batch(rng, cfg.feeds, one_hot=one_hot)  # noqa: E731
    elif kind == "rbm":
        cfg = configs.rbm(name=name, **w)
        gen = lambda rng: ((rng.random(cfg.feeds["x"].shape()) < 0.5).astype(np.float32),)  # noqa: E731
    elif kind == "conv":
        cfg = configs.conv_layer(name=name, **w)
        gen = lambda rng: configs.cnn_batch(rng, cfg.feeds)  # noqa: E731
    elif kind == "dqn":
        cfg = configs.dqn(name=name, **w)
        gen = lambda rng: tuple(configs.dqn_batch(rng, cfg.feeds)[k] for k in cfg.feeds)  # noqa: E731
    else:
        cfg = configs.recurrent(kind, name=name, **w)
        gen = lambda rng: configs.recurrent_batch(rng, cfg.feeds, w["vocab"])  # noqa: E731
    return cfg, gen, w


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][2]) if self.rows[0][2].replace(".", "").isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def pinned_array(cabi, shape, dtype=np.float32):
    n = int(np.prod(shape))
    p = C.c_void_p()
    cabi.check(cabi.lib().tcr_host_alloc(C.byref(p), C.c_size_t(n * np.dtype(dtype).itemsize)))
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def event_timer(cabi):
    lib = cabi.lib()
    e0, e1 = C.c_void_p(), C.c_void_p()
    cabi.check(lib.tcr_event_create(C.byref(e0)))
    cabi.check(lib.tcr_event_create(C.byref(e1)))

    def start():
        cabi.check(lib.tcr_event_record(e0))

    def stop_ms():
        cabi.check(lib.tcr_event_record(e1))
        ms = C.c_float()
        cabi.check(lib.tcr_event_elapsed_ms(e0, e1, C.byref(ms)))
        return ms.value
    return start, stop_ms


def dominant_kernel_roofline(cabi, name, w, peaks):
    """Time the step's dominant kernel alone, on the library stream, with the workload's shapes."""
    lib = cabi.lib()
    F = cabi.FLOAT
    start, stop_ms = event_timer(cabi)
    rng = np.random.default_rng(7)
    if "ninput" in w or "vocab" in w or "in_ch" in w:
        if "in_ch" in w:
            M, K, N = (w["width"] - 2) * (w["height"] - 2) * w["nbatch"], 9 * w["in_ch"], w["out_ch"]
            label = "tcr_gemm conv2d forward (positions x patch)(patch x out) 3xTF32"
        elif "ninput" in w:
            M, K, N = w["nbatch"], w["ninput"], w["nhidden"]
            label = "tcr_gemm fwd layer0 (B x in)(in x hid) 3xTF32"
        else:
            B = w["batch"] or 1
            M, K, N = B, w["vocab"] + w["hidden"], w["hidden"]
            label = "tcr_gemm gate (B x (N+H))((N+H) x H) 3xTF32"
        # rotate over enough (A, C) buffer sets that no launch finds its operands in the 126 MB L2 (the weights B stay hot, as in the step)
        set_bytes = 4 * (M * K + M * N)
        nsets = max(3, min(8, int(np.ceil(3 * 126e6 / set_bytes))))
        b = cabi.to_device(rng.uniform(-1, 1, K * N).astype(np.float32))
        sets = [(cabi.to_device(rng.uniform(-1, 1, M * K).astype(np.float32)), cabi.empty(M * N, np.float32)) for _ in range(nsets)]
        d = cabi.GemmDesc(m=M, n=N, k=K, batch=1, a_sm=K, a_sk=1, b_sk=N, b_sn=1, c_sm=N, c_sn=1, dtype=F, precision=cabi.GEMM_3XTF32)

        def call(i):
            a, c = sets[i % nsets]
            cabi.check(lib.tcr_gemm(C.c_void_p(a.ptr), C.c_void_p(b.ptr), C.c_void_p(c.ptr), C.byref(d)))

        def timed(iters=24):
            for i in range(nsets):
                call(i)
            cabi.sync()
            start()
            for i in range(iters):
                call(i)
            return stop_ms() / iters

        ms = timed()
        d.precision = cabi.GEMM_TF32  # the plain-TF32 variant of the same launch, reported beside the 3xTF32 product path
        ms_tf32 = timed()
        d.precision = cabi.GEMM_3XTF32
        flops = 2.0 * M * N * K
        # TF32 dense peak is not in MEASURED_PEAKS.json: half of the measured bf16 burst (B200_PROFILING.md table: tf32 = bf16 / 2)
        peak = peaks.get("bf16_tflops", 1590.0) / 2
        return {"bound": "tensor", "kernel": label, "achieved": round(flops / ms / 1e9, 2), "peak": round(peak, 1), "unit": "TFLOP/s",
                "frac": round(flops / ms / 1e9 / peak, 4), "traffic": None, "ms_per_launch": round(ms, 5),
                "issued_mma_frac": round(3 * flops / ms / 1e9 / peak, 4),  # 3xTF32 issues 3 TF32 MMAs per algorithmic MMA
                "tf32_variant": {"achieved": round(flops / ms_tf32 / 1e9, 2), "frac": round(flops / ms_tf32 / 1e9 / peak, 4), "ms_per_launch": round(ms_tf32, 5)},
                "peak_source": "0.5 x measured bf16_tflops (MEASURED_PEAKS.json)" if "bf16_tflops" in peaks else "0.5 x fallback 1590",
                "buffers": "%d rotating (A, C) sets of %.0f MB: operands are not L2-resident between launches" % (nsets, set_bytes / 1e6)}
    # RBM: HBM-bound elementwise/RNG over [784, B]; DQN: the same kernel over [nobs, B] (latency-bound at that size)
    width = w.get("nvisible", w.get("nobs", 1))
    n = width * w["nbatch"]
    x = cabi.to_device(rng.uniform(-4, 4, n).astype(np.float32))
    y = cabi.empty(n, np.float32)
    call = lambda: cabi.check(lib.tcr_unary(cabi.OP["SIGMOID"], C.c_void_p(x.ptr), C.c_void_p(y.ptr), C.c_int64(n), F))  # noqa: E731
    for _ in range(3):
        call()
    cabi.sync()
    start()
    for _ in range(50):
        call()
    ms = stop_ms() / 50
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"bound": "hbm", "kernel": "tcr_unary SIGMOID [%d,B]" % width, "achieved": round(8 * n / ms / 1e6, 1), "peak": peak, "unit": "GB/s",
            "frac": round(8 * n / ms / 1e6 / peak, 4), "traffic": None, "ms_per_launch": round(ms, 5),
            "peak_source": "measured hbm_gbs (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650"}


def hbm_micro(cabi, peaks):
    """The HBM half of the metric (M-ew at 2^26 and 2^28, M-red on [4096, 65536], M-lay), each case timed with CUDA
    events on the library stream through the C-ABI: achieved = ALGORITHMIC bytes / time, frac = achieved / measured copy bandwidth.
    Every operand set is larger than the 126 MB L2, or rotates over 8 disjoint copies when the case's own shape is smaller."""
    lib = cabi.lib()
    F = cabi.FLOAT
    peak = peaks.get("hbm_gbs", 6650.0)
    start, stop_ms = event_timer(cabi)
    rng = np.random.default_rng(5)
    n28 = 1 << 28
    seed = cabi.to_device(rng.uniform(-4, 4, 1 << 24).astype(np.float32))
    bufs = [cabi.empty(n28, np.float32) for _ in range(4)]
    for k, buf in enumerate(bufs[:3]):  # fill by replication on the device (values only matter for range)
        for j in range(16):
            cabi.check(lib.tcr_d2d(C.c_void_p(buf.ptr + 4 * j * (1 << 24)), C.c_void_p(seed.ptr), C.c_size_t(4 << 24)))
    a, b, c, out = bufs
    # PROD needs values near 1: |x|/4 * 0.2 + 0.9
    P = lambda x, off=0: C.c_void_p(x.ptr + 4 * off)  # noqa: E731
    res = {}

    def timeit(fn, iters, graph=False):
        """graph=True: the launches are captured once and replayed — for kernels of a few microseconds, where the host's launch rate
        (ctypes + cudaLaunchKernel, ~5 us) and not the device would otherwise be what is measured"""
        for i in range(2):
            fn(i)
        cabi.sync()
        if graph:
            g = C.c_void_p()
            cabi.check(lib.tcr_graph_begin())
            for i in range(iters):
                fn(i)
            cabi.check(lib.tcr_graph_end(C.byref(g)))
            cabi.check(lib.tcr_graph_launch(g))
            cabi.sync()
            start()
            cabi.check(lib.tcr_graph_launch(g))
            ms = stop_ms() / iters
            cabi.check(lib.tcr_graph_destroy(g))
            return ms
        start()
        for i in range(iters):
            fn(i)
        return stop_ms() / iters

    def rec(name, ms, nbytes):
        gbps = nbytes / ms / 1e6
        res[name] = {"GBps": round(gbps, 1), "frac": round(gbps / peak, 3), "ms": round(ms, 4)}

    for log2n in (26, 28):
        n = 1 << log2n
        it = 10 if log2n == 26 else 5
        rot = (lambda i: (i % 4) * n) if log2n == 26 else (lambda i: 0)
        for op in ("EXP", "SIGMOID", "TANH"):
            rec("ew_%s_2^%d" % (op, log2n), timeit(lambda i: cabi.check(lib.tcr_unary(cabi.OP[op], P(a, rot(i)), P(out, rot(i)), C.c_int64(n), F)), it), 8 * n)
        for op in ("ADD", "MUL"):
            rec("ew_%s_2^%d" % (op, log2n), timeit(lambda i: cabi.check(lib.tcr_binary(cabi.OP[op], P(a, rot(i)), P(b, rot(i)), P(out, rot(i)), C.c_int64(n), F)), it), 12 * n)
        progs = [cabi.make_program(F, (n, 1, 1), [(a.ptr + 4 * rot(i), F, (0, 0, 0)), (b.ptr + 4 * rot(i), F, (0, 0, 0)), (c.ptr + 4 * rot(i), F, (0, 0, 0))],
                                   [(out.ptr + 4 * rot(i), F, 0)], [(cabi.OP["MUL"], 0, 0, 1), (cabi.OP["ADD"], 0, 0, 2), (cabi.OP["SIGMOID"], 0, 0)]) for i in range(4)]
        rec("ew_fused_sigmoid(a*b+c)_2^%d" % log2n, timeit(lambda i: cabi.check(lib.tcr_elementwise(C.byref(progs[i % 4]))), it), 16 * n)
        rec("ew_ASSIGN_SUB_2^%d" % log2n, timeit(lambda i: cabi.check(lib.tcr_assign(cabi.OP["ASSIGN_SUB"], P(out, rot(i)), P(a, rot(i)), C.c_int64(n), F)), it), 12 * n)
    H = 1024
    progs = [cabi.make_program(F, (H, (1 << 26) // H, 1), [(a.ptr + (i << 28), F, (0, 0, 0)), (b.ptr, F, (0, 1, 0))], [(out.ptr + (i << 28), F, 0)],
                               [(cabi.OP["ADD"], 0, 0, 1), (cabi.OP["SIGMOID"], 0, 0)]) for i in range(4)]
    rec("ew_fused_sigmoid(x+bias[1024])_[1024,65536]", timeit(lambda i: cabi.check(lib.tcr_elementwise(C.byref(progs[i % 4]))), 10), 8 * (1 << 26) + 4 * H)
    # reductions on [4096, 65536] (= 2^28 elements); PROD reads a tensor of values near 1
    R0, R1 = 4096, 65536
    shp = cabi.shape8([R0, R1])
    near1 = [cabi.make_program(F, (n28, 1, 1), [(a.ptr, F, (0, 0, 0))], [(c.ptr, F, 0)],
                               [(cabi.EW_CONST, 1, 0, 0, 0, 0.025), (cabi.OP["MUL"], 0, 0, 1), (cabi.EW_CONST, 1, 0, 0, 0, 1.0), (cabi.OP["ADD"], 0, 0, 1)])]
    cabi.check(lib.tcr_elementwise(C.byref(near1[0])))  # c = 1 + a / 40 in [0.9, 1.1]
    small = cabi.empty(R1, np.float32)
    for op in ("REDUCE_SUM", "REDUCE_MAX", "REDUCE_MIN", "REDUCE_PROD"):
        src = c if op == "REDUCE_PROD" else a
        for mask, nm, nout in ((1, "dim0", R1), (2, "dim1", R0), (3, "full", 1)):
            rec("%s_%s_[4096,65536]" % (op, nm), timeit(lambda i: cabi.check(lib.tcr_reduce(cabi.OP[op], P(src), P(small), shp, C.c_uint32(mask), F)), 5), 4 * (n28 + nout))
    for dim, nm, nout in ((0, "dim0", R1), (1, "dim1", R0), (8, "flat", 1)):
        rec("ARGMAX_%s_[4096,65536]" % nm, timeit(lambda i: cabi.check(lib.tcr_argmax(P(a), P(small), shp, dim, F)), 5), 4 * (n28 + nout))
    # layout
    side = 8192
    order = (C.c_int32 * 8)(1, 0, 2, 3, 4, 5, 6, 7)
    rec("PERMUTE10_[8192,8192]", timeit(lambda i: cabi.check(lib.tcr_permute(P(a, (i % 4) * side * side), P(out, (i % 4) * side * side), cabi.shape8([side, side]), order, 4)), 10, graph=True), 8 * side * side)
    bc = (C.c_int64 * 8)(1, 65536, 1, 1, 1, 1, 1, 1)
    rec("EXTEND_[1024]->[1024,65536]", timeit(lambda i: cabi.check(lib.tcr_extend(P(a), P(out, (i % 4) << 26), cabi.shape8([1024]), bc, 4)), 10, graph=True), 4 * 1024 * 65536 + 4096)
    s3 = [1024, 128, 64]
    m3 = 1024 * 128 * 64
    offs = (C.c_int64 * 8)(0, 32, 0, 0, 0, 0, 0, 0)
    exts = (C.c_int64 * 8)(1024, 64, 64, 1, 1, 1, 1, 1)
    rec("SLICE_mid_[1024,128,64]", timeit(lambda i: cabi.check(lib.tcr_slice(P(a, (i % 8) * m3), P(out, (i % 8) * m3), cabi.shape8(s3), offs, exts, 4)), 16, graph=True), 8 * (m3 // 2))
    lo = (C.c_int64 * 8)(0, 16, 0, 0, 0, 0, 0, 0)
    rec("PAD_mid_[1024,128,64]", timeit(lambda i: cabi.check(lib.tcr_pad(P(a, (i % 8) * m3), P(out, (i % 8) * 2 * m3), cabi.shape8(s3), lo, lo, 4)), 16, graph=True), 4 * (m3 + 1024 * 160 * 64))
    shp2 = (C.c_int64 * 16)(*(cabi.shape8(s3)[:] + cabi.shape8(s3)[:]))
    tabs = [(C.c_void_p * 2)(a.ptr + 4 * (i % 8) * m3, b.ptr + 4 * (i % 8) * m3) for i in range(8)]
    rec("CONCAT_axis1_[1024,128,64]x2", timeit(lambda i: cabi.check(lib.tcr_concat(tabs[i % 8], shp2, 2, P(out, (i % 8) * 2 * m3), 1, 4)), 16, graph=True), 16 * m3)
    res["_note"] = {"peak_GBps": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                    "bytes": "algorithmic: 4*(inputs un-broadcast + outputs) per element (SURVEY.md §8d)",
                    "cache": "operands exceed L2 (126 MB) or rotate over disjoint copies",
                    "timing": "CUDA events on the library stream; layout cases (< 0.1 ms per launch) are captured in one CUDA graph and replayed"}
    del bufs, a, b, c, out, seed, small
    return res


def host_threads():
    """Threads the CPU legs may use: every core the process is allowed to run on. torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers, which would silently throttle the BLAS inside the oracle, so the count is
    set explicitly (threadpoolctl) instead of being inherited."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_steps(cfg, gen, budget_s, steps, warmup=1, target=None):
    """The oracle port of the reference's CPU path: node-by-node numpy evaluation of the dumped
    graph (TravEvaluator order, one unfused op per functor, in-place ASSIGNs). Runs `warmup` untimed and then `steps`
    timed steps; a step that is too slow for that (c4: ~7 s) stops at `budget_s` with at least two timed steps.
    Returns (seconds per step, timed steps, warm-up steps, BLAS threads)."""
    import tenncor_b200 as tc
    from threadpoolctl import threadpool_limits, threadpool_info
    from oracle import tcr_oracle as orc  # cpu_baseline leg only: never on the product path
    orc.set_baseline_mode(True)
    target = cfg.train if target is None else target
    tape = tc.dump_graph([target])
    for node in tape:
        if node["kind"] == "leaf":
            node["data"] = np.array(node["data"], copy=True)
    feed_ids = tc.dump_ids([target] + list(cfg.feeds.values()), None)
    rng = np.random.default_rng(0)
    times = []
    nthreads = host_threads()
    with threadpool_limits(limits=nthreads):
        blas = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        t_begin = time.perf_counter()
        for it in range(warmup + steps):
            if it >= warmup + 2 and time.perf_counter() - t_begin > budget_s:
                break
            batch = gen(rng)
            t0 = time.perf_counter()
            for feed, arr in zip(cfg.feeds.values(), batch):
                if feed_ids[feed] < len(tape) and tape[feed_ids[feed]]["kind"] == "leaf":  # inference does not read the labels
                    tape[feed_ids[feed]]["data"][...] = arr.reshape(-1)
            orc.eval_tape(tape)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return float(np.mean(times)), len(times), warmup, (max(blas) if blas else 1)


def e2e_entry(world, steps, pipelined_ms, serial_ms, h2d, d2h):
    """Both loops go through the public API with the same per-step H2D + D2H; the headline is the faster one (prefetching pays
    when the copy is long enough to hide: c3 1.7x; for the latency-bound toy sizes its extra device copy costs ~10 %)."""
    pipe = {"value": round(world * steps * 1e3 / pipelined_ms, 3), "ms_per_step": round(pipelined_ms / steps, 4),
            "input": "batch prefetched one step ahead on a copy stream (EVariable.prefetch / commit); the loss of every step read back through ETensor.get_later, one step behind the launches"}
    serial = {"value": round(world * steps * 1e3 / serial_ms, 3), "ms_per_step": round(serial_ms / steps, 4),
              "input": "EVariable.assign then get(), no overlap"}
    best = pipe if pipelined_ms <= serial_ms else serial
    return {"value": best["value"], "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": best["ms_per_step"],
            "input": best["input"], "pipelined": pipe, "serial": serial}


def pick_target(cfg, mode):
    if mode == "train":
        return cfg.train
    if "x" not in cfg.feeds or not hasattr(cfg.model, "get") or cfg.name.startswith("c4"):
        raise SystemExit("--mode inference is defined for the models linked to their input variable: mlp (c1, c1w, c3) and conv")
    return cfg.model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--evaluator", default="plan", choices=["plan", "node"])
    ap.add_argument("--precision", default="3xtf32", choices=["3xtf32", "tf32", "exact"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--extras", default="auto", choices=["auto", "none"],
                    help="auto: the default c3 line also carries the LSTM sub-record (`workloads.c4`) and, at N = 1, the HBM micro-benchmarks (`micro`)")
    ap.add_argument("--mode", default="train", choices=["train", "inference"],
                    help="train: one apply_update step (the metric); inference: forward pass of the model only (BASELINE config 3 'inference+training')")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world

    import __graft_entry__ as g
    g.build()
    import tenncor_b200 as tc
    from tenncor_b200 import cabi

    metric = "train steps/sec" if args.mode == "train" else "inference steps/sec"
    wdesc = {"workload": args.workload, **{k: v for k, v in WORKLOADS[args.workload].items()}}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        cfg, gen, w = build_config(args.workload)
        target = pick_target(cfg, args.mode)
        sec, nsteps, nwarm, threads = cpu_reference_steps(cfg, gen, budget_s=max(args.cpu_seconds, 10.0) * 12, steps=args.steps, warmup=args.warmup, target=target)
        sample = "%d timed full steps (after %d warm-up) of the same graph on the host: numpy oracle of the Eigen path, one thread per elementwise op, %d BLAS threads in GEMM" % (nsteps, nwarm, threads)
        line = {"impl": "reference", "metric": metric, "value": round(1.0 / sec, 4), "unit": "steps/s", "n_gpus": args.gpus, "steps": nsteps,
                "warmup": nwarm, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": dict(wdesc, desc=cfg.desc, batch_per_gpu=w.get("nbatch", w.get("batch"))),
                "cpu_baseline": {"value": round(1.0 / sec, 4), "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": round(1.0 / sec, 4), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "one host process evaluates one local batch per step regardless of --gpus (the reference's CPU path has no multi-device mode)"}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ B200 arm
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)  # host-side rendezvous only; data path is NCCL in libtcr_b200
    os.environ["TCR_DEVICE"] = str(local_rank)
    cabi.init(local_rank)
    tc.set_evaluator(args.evaluator)
    tc.set_matmul_precision(args.precision)
    dp_parity = None
    if world > 1:
        # sharded-vs-full-batch parity on this box's GPUs, before anything is timed (tests/test_dp_nccl_gpu.py is skipped on 1-GPU boxes)
        from tools import dp_check
        dp_parity = dp_check.run_check(dist, rank, world)  # initialises the communicator
        if not dp_parity["ok"]:
            if rank == 0:
                sys.stderr.write("bench.py: data-parallel parity check failed: %s\n" % json.dumps(dp_parity))
            return 4

    def barrier():
        tc.sync()
        if dist is not None:
            dist.barrier()

    def reduce_max(values):
        if dist is None:
            return [float(v) for v in values]
        import torch
        t = torch.tensor(values, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def measure(name, steps, warmup, mode):
        """resident + end-to-end timing of one workload; every rank takes part (collectives), results are max over ranks"""
        if world > 1:
            tc.dp.set_mean_reduce(WORKLOADS[name]["kind"] in ("mlp", "rbm", "dqn", "conv"))  # reduce_mean losses; the LSTM's NLL is a sum
        cfg, gen, w = build_config(name)
        target = pick_target(cfg, mode)
        rng = np.random.default_rng(1000 + rank)
        feeds = list(cfg.feeds.values()) if mode == "train" else [cfg.feeds["x"]]  # inference reads no labels
        host = [pinned_array(cabi, f.shape(), np.dtype(f.dtype())) for f in feeds]
        for buf, arr in zip(host, gen(rng)):
            buf[...] = arr
        h2d = int(sum(b.nbytes for b in host))
        start, stop_ms = event_timer(cabi)

        # resident: batch uploaded once, timed region = graph evaluation only
        for f, buf in zip(feeds, host):
            f.assign(buf)
        for _ in range(warmup):
            target.calc()
            for f in feeds:
                f.touch()  # new input version, data stays resident: the next step recomputes everything
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = cabi.lib().tcr_launch_count()
        start()
        t_host = time.perf_counter()
        for _ in range(steps):
            for f in feeds:
                f.touch()
            target.calc()
        host_ms = (time.perf_counter() - t_host) * 1e3  # the calls have returned, the device is still working: host cost of K evaluations
        ms_total = stop_ms()
        launches = int(cabi.lib().tcr_launch_count() - launches0)
        barrier()
        plan = tc.plan_stats()

        # end to end, serial: pinned host batch -> H2D -> step -> D2H of the loss, every step, one after the other
        loss = None
        for _ in range(warmup):
            for f, buf in zip(feeds, host):
                f.assign(buf)
            loss = target.get()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            for f, buf in zip(feeds, host):
                f.assign(buf)
            loss = target.get()
        tc.sync()
        e2e_serial_s = time.perf_counter() - t0
        barrier()
        # end to end, pipelined (the headline): the same per-step H2D of the batch and D2H of the loss, but the batch of
        # step i+1 crosses PCIe on the copy stream (EVariable.prefetch) while step i computes; commit() swaps it in
        for f, buf in zip(feeds, host):
            f.prefetch(buf)
        for _ in range(warmup):
            for f in feeds:
                f.commit()
            for f, buf in zip(feeds, host):
                f.prefetch(buf)
            loss = target.get()
        tc.sync()
        tc.sync_prefetch()
        barrier()
        # the loss of EVERY step is copied to pinned host memory and read by the host (get_later / result), one step behind the
        # launches: step i+1 is queued before the host waits for the loss of step i, so the device does not idle for a host round trip
        t0 = time.perf_counter()
        pending = None
        for _ in range(steps):
            for f in feeds:
                f.commit()
            for f, buf in zip(feeds, host):
                f.prefetch(buf)
            nxt = target.get_later()
            if pending is not None:
                loss = pending.result()
            pending = nxt
        loss = pending.result()
        tc.sync()
        tc.sync_prefetch()  # K copies were issued inside the timed region: all of them must have landed
        e2e_s = time.perf_counter() - t0
        del pending, nxt
        barrier()
        clocks = sampler.summary()
        d2h = int(np.asarray(loss).nbytes)
        ms_total, e2e_ms, e2e_serial_ms = reduce_max([ms_total, e2e_s * 1e3, e2e_serial_s * 1e3])
        final_loss = float(np.asarray(loss).reshape(-1)[0])
        # replicas must hold the same weights after the same all-reduced updates
        replicas = None
        if dist is not None and mode == "train":
            sums = [float(np.asarray(v.data(), np.float64).sum()) for v in cfg.variables]
            allsums = [None] * world
            dist.all_gather_object(allsums, sums)
            dev = max(abs(a - b) / (abs(b) + 1e-30) for other in allsums for a, b in zip(other, allsums[0]))
            replicas = {"variables": len(sums), "max_rel_dev_of_checksums_vs_rank0": dev, "ok": bool(dev < 1e-6)}
        finite = bool(np.isfinite(final_loss))
        if dist is not None:
            finite = bool(reduce_max([0.0 if finite else 1.0])[0] == 0.0)
        return dict(cfg=cfg, gen=gen, w=w, target=target, ms_total=ms_total, host_ms=host_ms, launches=launches, plan=plan, e2e_ms=e2e_ms, e2e_serial_ms=e2e_serial_ms,
                    h2d=h2d, d2h=d2h, final_loss=final_loss, finite=finite, clocks=clocks, replicas=replicas)

    def sub_record(name, m, steps):
        """compact record of a secondary workload (same timing rules) for the `workloads` key"""
        w = m["w"]
        ms_step = m["ms_total"] / steps
        flops = m["cfg"].flops_per_step
        rec = {"metric": "train steps/sec", "value": round(world * 1e3 / ms_step, 3), "unit": "steps/s", "ms_per_step": round(ms_step, 4), "steps": steps,
               "config": dict({"workload": name}, **WORKLOADS[name], desc=m["cfg"].desc, batch_per_gpu=w.get("nbatch", w.get("batch")), matmul=args.precision),
               "tflops": round(flops / ms_step / 1e9, 2), "launches_per_step": round(m["launches"] / steps, 1), "plan": m["plan"],
               "host_ms_per_step": round(m["host_ms"] / steps, 4),
               "e2e": e2e_entry(world, steps, m["e2e_ms"], m["e2e_serial_ms"], m["h2d"], m["d2h"]), "final_loss": m["final_loss"]}
        if m["replicas"] is not None:
            rec["replicas"] = m["replicas"]
        return rec

    main_m = measure(args.workload, args.steps, args.warmup, args.mode)
    cfg, gen, w, target = main_m["cfg"], main_m["gen"], main_m["w"], main_m["target"]
    rc = 0
    if not main_m["finite"]:
        sys.stderr.write("bench.py: final loss of %s is not finite (%r): the run is invalid\n" % (args.workload, main_m["final_loss"]))
        rc = 5
    if main_m["replicas"] is not None and not main_m["replicas"]["ok"]:
        sys.stderr.write("bench.py: replicas diverged: %s\n" % json.dumps(main_m["replicas"]))
        rc = 6

    # the LSTM half of the metric ("gd MLP, LSTM"): a sub-record in the same line, same timing rules, fewer steps
    extra = {}
    extras = [] if args.extras == "none" or args.mode != "train" else [x for x in ("c4",) if x != args.workload and args.workload == "c3"]
    for name in extras:
        sub_steps = max(3, min(args.steps, 10))
        m = measure(name, sub_steps, 3, "train")
        extra[name] = sub_record(name, m, sub_steps)
        if not m["finite"]:
            sys.stderr.write("bench.py: final loss of %s is not finite (%r)\n" % (name, m["final_loss"]))
            rc = 5
        if m["replicas"] is not None and not m["replicas"]["ok"]:
            rc = 6
        del m

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        roof = dominant_kernel_roofline(cabi, args.workload, w, peaks)
        try:  # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this launch
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
            if tr:
                roof["traffic"] = tr["bytes"]
                roof["traffic_source"] = tr["source"]
        except Exception:
            pass
        micro = None
        if world == 1 and args.extras != "none" and args.workload == "c3":
            micro = hbm_micro(cabi, peaks)
        if args.cpu_seconds > 0 and world == 1:
            sec, nsteps, nwarm, threads = cpu_reference_steps(cfg, gen, budget_s=args.cpu_seconds, steps=20, warmup=1, target=target)
            cpu_baseline = {"value": round(1.0 / sec, 4), "unit": "steps/s", "cores": threads, "kind": "port",
                            "sample": "%d full steps of the same graph on the host (numpy oracle of the Eigen path, %d BLAS threads)" % (nsteps, threads)}
        else:
            cpu_baseline = None  # timed on rank 0 at N = 1 only (or skipped with --cpu-seconds 0)
        ms_step = main_m["ms_total"] / args.steps
        # inference = forward only: a third of forward + both gradients for the GEMM-dominated models (approximate for c3, whose
        # first layer has no input gradient: 2*B*(in*hid + hid*out) exactly)
        flops_step = cfg.flops_per_step if args.mode == "train" else (2 * w["nbatch"] * (w["ninput"] * w["nhidden"] + w["nhidden"] * w["noutput"]) if "ninput" in w else cfg.flops_per_step // 3)
        line = {
            "metric": metric, "value": round(world * 1e3 / ms_step, 3), "unit": "steps/s",
            "unit_note": "sum over GPUs of per-GPU steps/s; each step = one local batch",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(wdesc, desc=cfg.desc, batch_per_gpu=w.get("nbatch", w.get("batch")), global_batch=(w.get("nbatch") or w.get("batch") or 1) * world,
                           evaluator=args.evaluator, matmul=args.precision, parallelism="dp%d" % world, mode=args.mode,
                           cache="step working set %s L2 (126 MB); no flush" % ("exceeds" if args.workload in ("c3", "c1w", "c4", "c4gru", "c2", "conv") else "is resident in")),
            "samples_per_s": round(world * (w.get("nbatch") or w.get("batch") or 1) * 1e3 / ms_step, 1),
            "flops_per_step": flops_step, "tflops": round(flops_step / ms_step / 1e9, 2),
            "clocks": main_m["clocks"],
            "e2e": e2e_entry(world, args.steps, main_m["e2e_ms"], main_m["e2e_serial_ms"], main_m["h2d"], main_m["d2h"]),
            "gpu_launches": main_m["launches"], "launches_per_step": round(main_m["launches"] / args.steps, 1), "plan": main_m["plan"],
            "host_ms_per_step": round(main_m["host_ms"] / args.steps, 4),
            "roofline": roof,
            "cpu_baseline": cpu_baseline,
            "final_loss": main_m["final_loss"],
        }
        if extra:
            line["workloads"] = extra
        if micro is not None:
            line["micro"] = micro
        if dp_parity is not None:
            line["dp_parity"] = dp_parity
        if main_m["replicas"] is not None:
            line["replicas"] = main_m["replicas"]
        print(json.dumps(line), flush=True)
    del cfg, gen, target, main_m
    if dist is not None:
        dist.barrier()  # rank 0 may still be in its micro-benchmark leg: leave together
        tc.dp.shutdown()
        dist.destroy_process_group()
    tc.shutdown()  # plans, graphs and the arena go before the interpreter tears the CUDA context down
    return rc


def _watchdog(seconds):
    """A bench that wedges (a peer died, a collective never completes) must end, not hold the box."""
    import threading

    def fire():
        sys.stderr.write("bench.py: watchdog expired after %ds, aborting\n" % seconds)
        sys.stderr.flush()
        os._exit(3)

    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


if __name__ == "__main__":
    _watchdog(int(os.environ.get("TCR_BENCH_WATCHDOG_S", "900")))
    rc = main()
    sys.stdout.flush()
    sys.stderr.flush()
    sys.exit(rc)  # normal interpreter teardown (atexit hooks of the harness run); the watchdog above is the only hard exit
