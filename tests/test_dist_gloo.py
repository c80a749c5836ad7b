"""N > 1 host logic on CPU: two gloo ranks shard the env ids, reduce rollout statistics and take the max step time."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mujocoposelearning_b200.dist import env_shard, max_over_ranks, reduce_rollout_stats
    off, total = env_shard(4096)
    t = max_over_ranks([1.0 + rank, 5.0 - rank])
    stats = reduce_rollout_stats(10.0 * (rank + 1), 667 * 2, 2)
    q.put((rank, off, total, t, stats))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [o[1] for o in out] == [0, 4096] and all(o[2] == 8192 for o in out)      # disjoint shards of the global ids
    assert all(o[3] == [2.0, 5.0] for o in out)                                       # max over ranks
    assert all(np.allclose(o[4], (30.0 / 4, 667.0, 4)) for o in out)                 # summed statistics
