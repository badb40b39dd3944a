"""Data-parallel parity for the recurrent workload on real GPUs (torchrun, one rank per GPU): every rank trains the LSTM on its
shard of the batch with all-reduced (summed) gradients; rank 0 also trains one replica on the whole batch. Prints one JSON line."""
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
    import tenncor_b200 as tc
    from tenncor_b200 import cabi, configs
    os.environ["TCR_DEVICE"] = str(local)
    cabi.init(local)
    tc.set_evaluator("plan")
    vocab, hidden, seq, per = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 64, 10, 8))]
    steps = 3
    rng = np.random.default_rng(7)
    eye = np.eye(vocab, dtype=np.float32)
    ids = rng.integers(0, vocab, (steps, seq + 1, per * world))
    full = configs.recurrent("lstm", vocab=vocab, hidden=hidden, seq=seq, batch=per * world, learning_rate=0.01, seed=3)
    full_losses = []
    for s in range(steps):
        full.feeds["x"].assign(eye[ids[s, :-1]])
        full.feeds["y"].assign(eye[ids[s, 1:]])
        full_losses.append(float(full.train.get()))
    full_w = [np.array(v.data(), copy=True) for v in full.variables]
    uid = [tc.dp.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    tc.dp.init(rank, world, uid[0], mean_reduce=False)
    shard = configs.recurrent("lstm", vocab=vocab, hidden=hidden, seq=seq, batch=per, learning_rate=0.01, seed=3)
    lo = rank * per
    losses = []
    for s in range(steps):
        shard.feeds["x"].assign(eye[ids[s, :-1, lo:lo + per]])
        shard.feeds["y"].assign(eye[ids[s, 1:, lo:lo + per]])
        losses.append(float(shard.train.get()))
    w = [np.array(v.data(), copy=True) for v in shard.variables]
    worst = max(float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30)) for a, b in zip(w, full_w))
    gathered = [None] * world
    dist.all_gather_object(gathered, (worst, losses))
    if rank == 0:
        sums = np.sum([g[1] for g in gathered], axis=0)  # summed NLL: shard losses add up to the full-batch loss
        print(json.dumps({"world": world, "dims": [vocab, hidden, seq, per], "weights_rel_err": max(g[0] for g in gathered),
                          "loss_rel_err": float(np.max(np.abs(sums - np.array(full_losses)) / np.abs(full_losses))), "full_losses": full_losses,
                          "shard_loss_sums": [float(x) for x in sums]}), flush=True)
    dist.barrier()
    tc.dp.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
