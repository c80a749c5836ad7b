#!/usr/bin/env python
"""PPO on the batched humanoid (BASELINE config 4 shape): rollouts on the hand-written kernels, update in PyTorch.

    python examples/train_ppo.py --iters 50                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_ppo.py --n-envs 16384

Mirrors train_sb3.py:170-236 (reference) with SubprocVecEnv + SB3 replaced by HumanoidBatch + PPOTrainer.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-envs", type=int, default=4096)
    ap.add_argument("--n-steps", type=int, default=64)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--reward", default="stand")
    ap.add_argument("--duration", type=float, default=10.0)
    ap.add_argument("--frame-skip", type=int, default=3)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--batch-size", type=int, default=16384)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--ent-coef", type=float, default=0.0)
    ap.add_argument("--update-tf32", action="store_true", help="update GEMMs in a single tf32 pass (default: fp32-faithful hi / lo split)")
    ap.add_argument("--update-impl", choices=["native", "torch"], default="native", help="native: the kernels of csrc/b2h_ppo.cu; torch: autograd + library GEMMs")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.ppo import PPOTrainer
    b = HumanoidBatch(args.n_envs, frame_skip=args.frame_skip, duration=args.duration, reward_type=args.reward, device=local, seed=0,
                      env_id_offset=rank * args.n_envs)
    tr = PPOTrainer(b, n_steps=args.n_steps, batch_size=args.batch_size, n_epochs=args.epochs, lr=args.lr, ent_coef=args.ent_coef, update_tf32=args.update_tf32,
                    update_impl=args.update_impl)
    for it in range(args.iters):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad():
            tr.col.collect()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        stats = tr.update()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if rank == 0:
            print(json.dumps({"iter": it, "rollout_s": round(t1 - t0, 4), "update_s": round(t2 - t1, 4),
                              "steps_per_s": round(world * args.n_envs * args.n_steps * args.frame_skip / (t1 - t0)),
                              "policy_loss": float(stats["policy_loss"]), "value_loss": float(stats["value_loss"]),
                              "clip_fraction": float(stats["clip_fraction"]), "mean_step_reward": float(tr.col.rewards.mean()),
                              "mean_height": float(tr.col.obs[:, :, 0].mean())}), flush=True)
    tr.policy.check_error()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
