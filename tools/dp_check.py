"""Data-parallel parity on real GPUs (run under torchrun, one rank per GPU; also called by bench.py --gpus N before its
timed region, because the pytest NCCL test is skipped on 1-GPU boxes):

every rank trains the gd_demo-pattern MLP on its shard of a global batch with NCCL all-reduced
gradients; rank 0 also trains a single-GPU replica on the WHOLE batch. With a reduce_mean loss and
equal shards the two must agree step for step (same summation up to fp32 reassociation).
Prints one JSON line from rank 0; exit code 1 on mismatch.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_check(dist, rank, world, dims=(64, 48, 16), per_gpu=96, steps=5):
    """Must be called BEFORE tc.dp.init (the full-batch replica is built without a communicator); leaves the communicator
    initialised with mean_reduce=True. Returns {"world", "weights_rel_err", "loss_rel_err", "ok"} on every rank."""
    import tenncor_b200 as tc
    from tenncor_b200 import configs
    rng = np.random.default_rng(123)
    xs = rng.random((steps, per_gpu * world, dims[0]), dtype=np.float32)
    ys = rng.random((steps, per_gpu * world, dims[2]), dtype=np.float32)

    # reference replica: the whole batch on this GPU, no communicator
    full = configs.mlp(dims[0], dims[1], dims[2], per_gpu * world, seed=7)
    full_losses = []
    for s in range(steps):
        full.feeds["x"].assign(xs[s])
        full.feeds["y"].assign(ys[s])
        full_losses.append(float(full.train.get()))
    full_w = [np.array(v.data(), copy=True) for v in full.variables]

    ids = [tc.dp.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    tc.dp.init(rank, world, ids[0], mean_reduce=True)
    shard = configs.mlp(dims[0], dims[1], dims[2], per_gpu, seed=7)  # same seed -> same initial weights
    lo = rank * per_gpu
    losses = []
    for s in range(steps):
        shard.feeds["x"].assign(xs[s, lo:lo + per_gpu])
        shard.feeds["y"].assign(ys[s, lo:lo + per_gpu])
        losses.append(float(shard.train.get()))
    w = [np.array(v.data(), copy=True) for v in shard.variables]
    worst = 0.0
    for a, b in zip(w, full_w):
        worst = max(worst, float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30)))
    gathered = [None] * world
    dist.all_gather_object(gathered, (worst, losses))
    worst_all = max(g[0] for g in gathered)
    # the returned loss is the post-update error on the LOCAL shard; its mean over ranks is the full-batch error
    mean_losses = np.mean([g[1] for g in gathered], axis=0)
    loss_err = float(np.max(np.abs(mean_losses - np.array(full_losses)) / np.abs(full_losses)))
    ok = bool(worst_all < 1e-4 and loss_err < 1e-4)
    return {"world": world, "weights_rel_err": worst_all, "loss_rel_err": loss_err, "ok": ok,
            "what": "MLP %d-%d-%d, %d samples per rank, %d SGD steps: NCCL-sharded ranks vs one replica fed the whole batch" % (dims + (per_gpu, steps))}


def main():
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tenncor_b200 as tc
    from tenncor_b200 import cabi
    os.environ["TCR_DEVICE"] = str(local)
    cabi.init(local)
    tc.set_evaluator("plan")
    res = run_check(dist, rank, world)
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    tc.dp.shutdown()
    dist.destroy_process_group()
    sys.stdout.flush()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
