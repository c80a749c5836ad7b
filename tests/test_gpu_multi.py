"""Multi-GPU on real hardware (skipped with fewer than two visible GPUs; the world_size-2 host logic is covered on CPU by
tests/test_dist_gloo.py): two NCCL ranks, each with its own env shard, run one PPO update whose gradients are
averaged by the flat all-reduce -- and end with identical replicas that equal the update computed in ONE process from
the two ranks' minibatches (mean of the two per-rank gradients, then clip and Adam: train_sb3.py:208-231 semantics
under data parallelism)."""
import math
import os
import socket

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _worker(rank, world, port, q, allreduce):
    import torch.distributed as dist
    import torch.nn.functional as F
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.ppo import PPOTrainer
    n, T = 128, 8
    b = HumanoidBatch(n, frame_skip=3, duration=10.0, reward_type="stand", device=rank, seed=2, env_id_offset=rank * n)
    tr = PPOTrainer(b, n_steps=T, batch_size=n * T, n_epochs=1, lr=3e-4, seed=5, allreduce=allreduce)   # update_impl native: b2h_ppo_* kernels
    assert tr.kernels.p2p == (allreduce == "p2p")
    with torch.no_grad():
        tr.col.collect()
    before = [t.detach().clone() for t in tr.tensors]
    tr.update()
    tr.policy.check_error()
    c = tr.col
    local = [c.obs.reshape(n * T, -1), c.actions.reshape(n * T, -1), c.log_probs.reshape(n * T, 1), c.advantages.reshape(n * T, 1),
             c.returns.reshape(n * T, 1)]
    flat = torch.cat([t.detach().reshape(-1) for t in tr.tensors])
    gathered = []
    for t in local + [flat]:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        gathered.append(parts)
    ok, err = True, 0.0
    if rank == 0:
        assert torch.equal(gathered[5][0], gathered[5][1])                    # identical replicas after the update
        assert not torch.equal(gathered[0][0], gathered[0][1])                # ... from different env shards
        ref = [t.clone().requires_grad_(True) for t in before]
        opt = torch.optim.Adam(ref, lr=3e-4, eps=1e-5)
        net = lambda w, x: F.linear(F.relu(F.linear(F.relu(F.linear(x, w[0], w[1])), w[2], w[3])), w[4], w[5])
        opt.zero_grad()
        for r in range(world):
            obs, act, olp, adv, ret = (gathered[k][r] for k in range(5))
            olp, adv, ret = olp[:, 0], adv[:, 0], ret[:, 0]
            a = (adv - adv.mean()) / (adv.std() + 1e-8)
            mean, value, log_std = net(ref[0:6], obs), net(ref[6:12], obs).squeeze(1), ref[12]
            logp = (-0.5 * ((act - mean) / log_std.exp()) ** 2 - log_std - 0.5 * math.log(2 * math.pi)).sum(1)
            ratio = torch.exp(logp - olp)
            loss = -torch.min(a * ratio, a * torch.clamp(ratio, 0.8, 1.2)).mean() + 0.5 * F.mse_loss(ret, value)
            (loss / world).backward()                                          # accumulates the mean of the per-rank gradients
        torch.nn.utils.clip_grad_norm_(ref, 0.5)
        opt.step()
        err = max(float((g - w).abs().max()) for g, w in zip(tr.tensors, ref))
        moved = max(float((g - w).abs().max()) for g, w in zip(tr.tensors, before))
        ok = err < 4e-6 and moved > 1e-5
    q.put((rank, ok, err))
    dist.barrier()
    b.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("allreduce", ["p2p", "nccl"])
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_update_equals_single_process_mean_gradient(allreduce):
    """allreduce='p2p': the gradients are summed by peer loads over NVLink inside the update kernels (CUDA IPC buffers, flag
    barrier); 'nccl': one flat NCCL all-reduce between b2h_ppo_minibatch_grad and b2h_ppo_apply."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, allreduce)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(o[1] for o in out), out
