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
    "c3": dict(kind="mlp", ninput=784, nhidden=1024, noutput=10, nbatch=8192, one_hot=True),
    "c2": dict(kind="rbm", nvisible=784, nhidden=64, nbatch=4096),
    "c4a": dict(kind="lstm", vocab=128, hidden=1024, seq=128, batch=None),
    "c4": dict(kind="lstm", vocab=128, hidden=1024, seq=128, batch=64),
    "c4gru": dict(kind="gru", vocab=128, hidden=1024, seq=128, batch=64),
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
        a = cabi.to_device(rng.uniform(-1, 1, M * K).astype(np.float32))
        b = cabi.to_device(rng.uniform(-1, 1, K * N).astype(np.float32))
        c = cabi.empty(M * N, np.float32)
        d = cabi.GemmDesc(m=M, n=N, k=K, batch=1, a_sm=K, a_sk=1, b_sk=N, b_sn=1, c_sm=N, c_sn=1, dtype=F, precision=cabi.GEMM_3XTF32)
        call = lambda: cabi.check(lib.tcr_gemm(C.c_void_p(a.ptr), C.c_void_p(b.ptr), C.c_void_p(c.ptr), C.byref(d)))  # noqa: E731

        def timed(iters=20):
            for _ in range(3):
                call()
            cabi.sync()
            start()
            for _ in range(iters):
                call()
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
                "peak_source": "0.5 x measured bf16_tflops (MEASURED_PEAKS.json)" if "bf16_tflops" in peaks else "0.5 x fallback 1590"}
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


def cpu_reference_steps(cfg, gen, budget_s, max_steps, target=None):
    """The oracle port of the reference's CPU path: node-by-node numpy evaluation of the dumped
    graph (TravEvaluator order, one unfused op per functor, in-place ASSIGNs)."""
    import tenncor_b200 as tc
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
    t_begin = time.perf_counter()
    while len(times) < max_steps and (time.perf_counter() - t_begin < budget_s or len(times) < 2):
        batch = gen(rng)
        t0 = time.perf_counter()
        for feed, arr in zip(cfg.feeds.values(), batch):
            if feed_ids[feed] < len(tape) and tape[feed_ids[feed]]["kind"] == "leaf":  # inference does not read the labels
                tape[feed_ids[feed]]["data"][...] = arr.reshape(-1)
        orc.eval_tape(tape)
        times.append(time.perf_counter() - t0)
    steady = times[1:] if len(times) > 1 else times
    return float(np.median(steady)), len(times)


def e2e_entry(world, steps, pipelined_ms, serial_ms, h2d, d2h):
    """Both loops go through the public API with the same per-step H2D + D2H; the headline is the faster one (prefetching pays
    when the copy is long enough to hide: c3 1.7x; for the latency-bound toy sizes its extra device copy costs ~10 %)."""
    pipe = {"value": round(world * steps * 1e3 / pipelined_ms, 3), "ms_per_step": round(pipelined_ms / steps, 4),
            "input": "prefetched one step ahead on a copy stream (EVariable.prefetch / commit)"}
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
        med, nsteps = cpu_reference_steps(cfg, gen, budget_s=max(args.cpu_seconds, 10.0) * 4, max_steps=args.steps + args.warmup, target=target)
        cores = os.cpu_count()
        line = {"impl": "reference", "metric": metric, "value": round(1.0 / med, 4), "unit": "steps/s", "n_gpus": 0, "steps": nsteps,
                "warmup": 1, "ms_per_step": round(med * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": dict(wdesc, desc=cfg.desc, batch_per_gpu=w.get("nbatch", w.get("batch"))),
                "cpu_baseline": {"value": round(1.0 / med, 4), "unit": "steps/s", "cores": cores, "kind": "port",
                                 "sample": "%d full steps of the same graph (numpy oracle: 1 thread per elementwise op, BLAS threads in GEMM)" % nsteps},
                "e2e": {"value": round(1.0 / med, 4), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    if world > 1:
        ids = [tc.dp.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        mean_loss = WORKLOADS[args.workload]["kind"] in ("mlp", "rbm", "dqn", "conv")  # reduce_mean losses; the LSTM's NLL is a sum
        tc.dp.init(rank, world, ids[0], mean_reduce=mean_loss)

    cfg, gen, w = build_config(args.workload)
    target = pick_target(cfg, args.mode)
    rng = np.random.default_rng(1000 + rank)
    feeds = list(cfg.feeds.values()) if args.mode == "train" else [cfg.feeds["x"]]  # inference reads no labels
    host = [pinned_array(cabi, f.shape()) for f in feeds]
    for buf, arr in zip(host, gen(rng)):
        buf[...] = arr
    h2d = int(sum(b.nbytes for b in host))

    def barrier():
        tc.sync()
        if dist is not None:
            dist.barrier()

    start, stop_ms = event_timer(cabi)

    # resident: batch uploaded once, timed region = graph evaluation only
    for f, buf in zip(feeds, host):
        f.assign(buf)
    for _ in range(args.warmup):
        target.calc()
        for f in feeds:
            f.touch()  # new input version, data stays resident: the next step recomputes everything
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = cabi.lib().tcr_launch_count()
    start()
    for _ in range(args.steps):
        for f in feeds:
            f.touch()
        target.calc()
    ms_total = stop_ms()
    launches = int(cabi.lib().tcr_launch_count() - launches0)
    barrier()
    plan = tc.plan_stats()

    # end to end, serial: pinned host batch -> H2D -> step -> D2H of the loss, every step, one after the other
    loss = None
    for _ in range(args.warmup):
        for f, buf in zip(feeds, host):
            f.assign(buf)
        loss = target.get()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
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
    for _ in range(args.warmup):
        for f in feeds:
            f.commit()
        for f, buf in zip(feeds, host):
            f.prefetch(buf)
        loss = target.get()
    tc.sync()
    tc.sync_prefetch()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for f in feeds:
            f.commit()
        for f, buf in zip(feeds, host):
            f.prefetch(buf)
        loss = target.get()
    tc.sync()
    tc.sync_prefetch()  # K copies were issued inside the timed region: all of them must have landed
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.summary()
    d2h = int(np.asarray(loss).nbytes)

    if dist is not None:
        import torch
        t = torch.tensor([ms_total, e2e_s * 1e3, e2e_serial_s * 1e3], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, e2e_serial_ms = float(t[0]), float(t[1]), float(t[2])
    else:
        e2e_ms, e2e_serial_ms = e2e_s * 1e3, e2e_serial_s * 1e3

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        roof = dominant_kernel_roofline(cabi, args.workload, w, peaks)
        try:  # DRAM traffic of the dominant kernel from the committed ncu capture (null when none was taken)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(args.workload)
            if tr:
                roof["traffic"] = tr["bytes"]
                roof["traffic_source"] = tr["source"]
        except Exception:
            pass
        if args.cpu_seconds > 0:
            med, nsteps = cpu_reference_steps(cfg, gen, budget_s=args.cpu_seconds, max_steps=20, target=target)
            cpu_baseline = {"value": round(1.0 / med, 4), "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": "%d full steps of the same graph on the host (numpy oracle of the Eigen path)" % nsteps}
        else:
            cpu_baseline = None  # --cpu-seconds 0: the CPU leg was skipped on request
        ms_step = ms_total / args.steps
        # inference = forward only: a third of forward + both gradients for the GEMM-dominated models (approximate for c3, whose
        # first layer has no input gradient: 2*B*(in*hid + hid*out) exactly)
        flops_step = cfg.flops_per_step if args.mode == "train" else (2 * w["nbatch"] * (w["ninput"] * w["nhidden"] + w["nhidden"] * w["noutput"]) if "ninput" in w else cfg.flops_per_step // 3)
        line = {
            "metric": metric, "value": round(world * 1e3 / ms_step, 3), "unit": "steps/s (sum over GPUs of per-GPU steps/s; each step = one local batch)",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(wdesc, desc=cfg.desc, batch_per_gpu=w.get("nbatch", w.get("batch")), global_batch=(w.get("nbatch") or w.get("batch") or 1) * world,
                           evaluator=args.evaluator, matmul=args.precision, parallelism="dp%d" % world, mode=args.mode,
                           cache="step working set %s L2 (126 MB); no flush" % ("exceeds" if args.workload in ("c3", "c1w", "c4", "c4gru", "c2", "conv") else "is resident in")),
            "samples_per_s": round(world * (w.get("nbatch") or w.get("batch") or 1) * 1e3 / ms_step, 1),
            "flops_per_step": flops_step, "tflops": round(flops_step / ms_step / 1e9, 2),
            "clocks": clocks,
            "e2e": e2e_entry(world, args.steps, e2e_ms, e2e_serial_ms, h2d, d2h),
            "gpu_launches": launches, "launches_per_step": round(launches / args.steps, 1), "plan": plan,
            "roofline": roof,
            "cpu_baseline": cpu_baseline,
            "final_loss": float(np.asarray(loss).reshape(-1)[0]),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()  # rank 0 may still be in its CPU-baseline leg: leave together
        tc.dp.shutdown()
        dist.destroy_process_group()
    return 0


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
    os._exit(rc)  # skip interpreter teardown: nothing after the JSON line may block the launcher
