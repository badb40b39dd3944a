"""Data-parallel host logic on CPU with world_size 2 over gloo (no GPU): batch sharding, and
the identity SUM-all-reduce(shard gradients) * scale == full-batch gradient for the two loss
normalisations the configs use (reduce_mean -> 1/N, summed NLL -> 1). Gradients come from the
CPU oracle evaluating graphs built by the host (test-only use of the oracle)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _grads(x, y, w0, b0, w1, b1, mean):
    """gradients of the gd_demo MLP loss for one shard, through host graph + oracle tape."""
    sys.path.insert(0, ROOT)
    import tenncor_b200 as tc
    from oracle import tcr_oracle as orc
    X, Y = tc.variable(x, "x"), tc.variable(y, "y")
    W0, B0, W1, B1 = tc.variable(w0, "w0"), tc.variable(b0, "b0"), tc.variable(w1, "w1"), tc.variable(b1, "b1")
    h = tc.api.sigmoid(tc.api.nn.fully_connect([X], [W0], B0))
    o = tc.api.sigmoid(tc.api.nn.fully_connect([h], [W1], B1))
    sq = tc.api.square(Y - o)
    loss = tc.api.reduce_mean(sq) if mean else tc.api.reduce_sum(sq)
    ders = tc.derive(loss, [W0, B0, W1, B1])
    tape = tc.dump_graph(ders)
    vals = orc.eval_tape(tape)
    ids = tc.dump_ids(ders, None)
    return [np.asarray(vals[ids[d]], dtype=np.float64) for d in ders]


def _worker(rank, world, port, mean, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import tenncor_b200 as tc
    rng = np.random.default_rng(0)  # same data on every rank; each takes its shard
    B = 11  # ragged on purpose: shards of 6 and 5
    x, y = rng.random((B, 7)), rng.random((B, 3))
    w0, b0, w1, b1 = rng.normal(size=(7, 5)), rng.normal(size=5), rng.normal(size=(5, 3)), rng.normal(size=3)
    off, cnt = tc.dp.shard(B, rank, world)
    gs = _grads(x[off:off + cnt], y[off:off + cnt], w0, b0, w1, b1, mean)
    # reduce_mean bakes the LOCAL element count (core.yml:883-889): weight each shard's mean by its share
    scale = (cnt / B) if mean else 1.0
    flat = torch.from_numpy(np.concatenate([g.reshape(-1) for g in gs]) * scale)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)  # one flat bucket per step
    if rank == 0:
        full = np.concatenate([g.reshape(-1) for g in _grads(x, y, w0, b0, w1, b1, mean)])
        out.put((flat.numpy().copy(), full, (off, cnt)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mean", [True, False], ids=["reduce_mean", "reduce_sum"])
def test_sharded_gradients_allreduce_to_full_batch(built, mean):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mean, out)) for r in range(2)]
    for p in procs:
        p.start()
    got, full, shard0 = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert shard0 == (0, 6)
    np.testing.assert_allclose(got, full, rtol=1e-10, atol=1e-12)


def test_shard_partition_covers_batch(built):
    import tenncor_b200 as tc
    for total, n in [(65536, 8), (11, 2), (7, 4), (3, 8)]:
        spans = [tc.dp.shard(total, r, n) for r in range(n)]
        assert sum(c for _, c in spans) == total
        pos = 0
        for off, cnt in spans:
            assert off == pos
            pos += cnt


def test_dp_marker_only_when_group_active(built):
    import tenncor_b200 as tc
    x = tc.variable(np.ones((2, 3), np.float32), "x")
    g = tc.derive(tc.api.reduce_sum(tc.api.square(x)), [x])[0]
    assert g.opname() != "IDENTITY"  # no communicator: derive returns the plain gradient
    assert tc.dp.size() == 1 and tc.dp.rank() == 0
