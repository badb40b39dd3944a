"""Peer-memory all-reduce (allreduce_p2p.cu) against numpy on real GPUs; run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/p2p_allreduce_check.py

Every rank fills symmetric buffers of several sizes (one-shot and two-shot regimes, ragged lengths) with rank-dependent values;
after tcr_allreduce_sum every rank must hold sum over ranks x scale, bit-identical across ranks and across repeats, also when the
call is replayed from a CUDA graph. Prints one JSON line from rank 0 (with timings vs NCCL); exit code 1 on mismatch."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tenncor_b200 import cabi
    os.environ["TCR_DEVICE"] = str(local)
    cabi.init(local)
    lib = cabi.lib()
    ids = [None]
    if rank == 0:
        buf = C.create_string_buffer(cabi.COMM_ID_BYTES)
        cabi.check(lib.tcr_comm_unique_id(buf))
        ids = [buf.raw]
    dist.broadcast_object_list(ids, src=0)
    cabi.check(lib.tcr_comm_init(rank, world, ids[0]))
    ready = int(lib.tcr_comm_p2p_ready())
    ok = True
    report = {"world": world, "p2p_ready": ready, "cases": []}
    e0, e1 = C.c_void_p(), C.c_void_p()
    cabi.check(lib.tcr_event_create(C.byref(e0)))
    cabi.check(lib.tcr_event_create(C.byref(e1)))
    for n in (149, 189, 4096, 51024, 814090, 4860000):
        rng = np.random.default_rng(100 + rank)
        mine = rng.uniform(-1, 1, n).astype(np.float32)
        everyone = [np.random.default_rng(100 + r).uniform(-1, 1, n).astype(np.float32) for r in range(world)]
        want = everyone[0].copy()
        for r in range(1, world):
            want = want + everyone[r]  # rank order, fp32: what the kernel does
        scale = 1.0 / world
        want = (want * np.float32(scale)).astype(np.float32)
        ptr = C.c_void_p()
        rc = lib.tcr_comm_symm_alloc(C.byref(ptr), C.c_size_t(4 * ((n + 3) // 4 * 4)))
        symmetric = rc == 0
        if not symmetric:
            keep = cabi.empty((n + 3) // 4 * 4, np.float32)
            ptr = C.c_void_p(keep.ptr)
        host = np.zeros((n + 3) // 4 * 4, np.float32)

        def fill():
            host[:n] = mine
            cabi.check(lib.tcr_h2d(ptr, host.ctypes.data_as(C.c_void_p), C.c_size_t(host.nbytes)))

        fill()
        cabi.sync()
        dist.barrier()
        cabi.check(lib.tcr_allreduce_sum(ptr, C.c_int64(n), cabi.FLOAT, C.c_double(scale)))
        got = np.empty_like(host)
        cabi.check(lib.tcr_d2h(got.ctypes.data_as(C.c_void_p), ptr, C.c_size_t(host.nbytes)))
        cabi.sync()
        exact = bool(np.array_equal(got[:n], want))
        close = bool(np.allclose(got[:n], want, rtol=1e-6, atol=1e-6))
        gathered = [None] * world
        dist.all_gather_object(gathered, got[:n].tobytes())
        identical = all(g == gathered[0] for g in gathered)
        # timing: 20 back-to-back exchanges (values grow, irrelevant)
        dist.barrier()
        cabi.check(lib.tcr_event_record(e0))
        for _ in range(20):
            cabi.check(lib.tcr_allreduce_sum(ptr, C.c_int64(n), cabi.FLOAT, C.c_double(1.0)))
        cabi.check(lib.tcr_event_record(e1))
        ms = C.c_float()
        cabi.check(lib.tcr_event_elapsed_ms(e0, e1, C.byref(ms)))
        # replay from a CUDA graph
        fill()
        cabi.sync()
        dist.barrier()
        cabi.check(lib.tcr_graph_begin())
        cabi.check(lib.tcr_allreduce_sum(ptr, C.c_int64(n), cabi.FLOAT, C.c_double(scale)))
        g = C.c_void_p()
        cabi.check(lib.tcr_graph_end(C.byref(g)))
        cabi.check(lib.tcr_graph_launch(g))
        cabi.check(lib.tcr_d2h(got.ctypes.data_as(C.c_void_p), ptr, C.c_size_t(host.nbytes)))
        cabi.sync()
        graph_ok = bool(np.allclose(got[:n], want, rtol=1e-6, atol=1e-6))
        cabi.check(lib.tcr_graph_destroy(g))
        case = {"n": n, "symmetric": symmetric, "exact_rank_order_sum": exact, "close": close, "identical_on_all_ranks": identical, "graph_replay": graph_ok,
                "us_per_call": round(ms.value * 1e3 / 20, 2)}
        report["cases"].append(case)
        ok = ok and close and identical and graph_ok
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    ok = all(flags)
    report["ok"] = ok
    if rank == 0:
        print(json.dumps(report), flush=True)
    dist.barrier()
    cabi.check(lib.tcr_comm_destroy())
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
