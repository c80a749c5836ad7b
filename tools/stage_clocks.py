#!/usr/bin/env python
"""Cycle split of the step kernel by stage (tuning build only).

    B2H_LIB=/path/libb2h_clk.so B2H_NVCC_EXTRA=-DB2H_STAGE_CLOCKS python -c 'from mujocoposelearning_b200.build import build; build(force=True)'
    B2H_LIB=/path/libb2h_clk.so python tools/stage_clocks.py [n_envs] [steps] [warmup]

Every warp stamps clock64() at the stage boundaries of physics_step / env_step; the sums (warp-cycles) are printed as
fractions of the warps' total residence in the claim loop.  The bench workload (BASELINE config 3) is used.
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujocoposelearning_b200 import lib as L  # noqa: E402
from mujocoposelearning_b200.batch import HumanoidBatch  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
W = int(sys.argv[3]) if len(sys.argv) > 3 else 300
NAMES = ["pre-solver stages", "factor M + solve", "checkAcc", "euler", "sub-step barrier wait", "newton: update + J^T f + move",
         "newton: hessian build", "newton: factor + solve", "newton: line search", "env_step minus barriers", "claim barrier + atomic", "-"]
lib = L.load()
fn = lib.b2h_debug_stage_clocks
fn.restype, fn.argtypes = C.c_int, [C.POINTER(C.c_uint64), C.c_int]
b = HumanoidBatch(E, frame_skip=3, duration=10.0, reward_type="stand", seed=1234)
g = torch.Generator(device="cuda").manual_seed(1234)
pool = torch.rand(16, E, b.nu, device="cuda", generator=g) * 2 - 1
b.reset()
for i in range(W):
    b.step(pool[i % 16])
out = (C.c_uint64 * 48)()
fn(out, 1)
c0 = b.counters()
for i in range(K):
    b.step(pool[(W + i) % 16])
fn(out, 0)
c1 = b.counters()
clk = [int(x) for x in out]
total = clk[9] + clk[4] + clk[10]
psteps = c1["physics_steps"] - c0["physics_steps"]
res = {"n_envs": E, "steps": K, "launch": b.launch_info(), "warp_cycles_per_physics_step": total / psteps,
       "newton_iter_per_step": (c1["newton_iter"] - c0["newton_iter"]) / psteps,
       "frac": {NAMES[i]: round(clk[i] / total, 4) for i in range(11)}}
res["frac"]["env epilogue + io (rest)"] = round((clk[9] - sum(clk[i] for i in (0, 1, 2, 3, 5, 6, 7, 8))) / total, 4)
PRE = {12: "kinematics", 13: "frames + com + cinert + cdof", 14: "crb", 15: "collision broad phase", 16: "collision narrow phase",
       17: "constraint rows (J)", 18: "row parameters + limits", 19: "velocity stage (comvel, rne, passive, actuation)"}
res["pre-solver split"] = {v: round(clk[k] / total, 4) for k, v in PRE.items()}
res["newton exits"] = {"improvement": clk[24], "gradient": clk[25], "exact_stop": clk[26], "improvement exits that exact_stop would also take": clk[29],
                       "top-of-iteration checks with unchanged active set": clk[31]}
res["line searches"] = {"count": clk[27], "exact newton step accepted": clk[28], "done after two evaluations": clk[30]}
res["iterations at exit (histogram)"] = clk[32:48]
res["cta_exit_spread_ns"] = {"mean_after_first": clk[22], "last_after_first": clk[23]}   # tail imbalance of the last launch
print(json.dumps(res, indent=1))
