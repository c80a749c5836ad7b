#!/usr/bin/env python
"""Profiling driver: N control steps of the random-action rollout (BASELINE config 3) and nothing else.

    python tools/prof_step.py [n_envs] [steps]
    ncu --set full --import-source on --clock-control none -k regex:step_kernel -s 700 -c 1 -o gpurun_out/prof python tools/prof_step.py 4096 40
(667 pre-roll launches stagger the episode phases like bench.py does; B2H_PROF_STAGGER=0 skips that)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujocoposelearning_b200.batch import HumanoidBatch  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 320
b = HumanoidBatch(E, frame_skip=3, duration=10.0, reward_type="stand", seed=1234)
g = torch.Generator(device="cuda").manual_seed(1234)
pool = torch.rand(16, E, b.nu, device="cuda", generator=g) * 2 - 1
b.reset()
if os.environ.get("B2H_PROF_STAGGER", "1") != "0":     # the bench's timed state: episode phases spread uniformly (untimed pre-roll)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import stagger
    stagger(b, pool)
for i in range(K):
    b.step(pool[i % 16])
torch.cuda.synchronize()
print(b.counters())
