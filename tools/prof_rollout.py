#!/usr/bin/env python
"""Profiling driver for the in-library rollout loop (b2h_rollout_collect): a few control steps of policy -> sample -> env step
-> record at n_envs, nothing else.

    python tools/prof_rollout.py [n_envs] [n_steps] [rollouts]
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_rollout.py 4096 8 3
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujocoposelearning_b200.batch import HumanoidBatch  # noqa: E402
from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams, RolloutCollector  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
R = int(sys.argv[3]) if len(sys.argv) > 3 else 3
b = HumanoidBatch(E, frame_skip=3, duration=10.0, reward_type="stand", seed=1234)
col = RolloutCollector(b, MlpPolicy(MlpPolicyParams(seed=7), precise=True, seed=11), n_steps=T, cuda_graph=False)
for _ in range(R):
    col.collect()
torch.cuda.synchronize()
col.check_error()
print(b.counters(), col.stats.tolist())
