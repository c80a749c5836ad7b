#!/usr/bin/env python
"""Experiment behind DESIGN.md section 5 (CPU lane emulation of the fp32 kernel source vs the fp64 oracle, ONE mj_step from
identical float32-representable states): where does the fp32 single-step error come from?

    python tools/exp_fp32_single_step_error.py                                 # fp32 kernel arithmetic
    B2H_EMU_EXTRA=-DB2H_EXP_CHOL_F64 python tools/exp_fp32_single_step_error.py  # all five factor+solves carried in double

Columns: pre-steps, nefc, element-wise relative error |a-b| / max(|b|, 1e-3) of qpos, qvel, obs, |reward error|,
qvel error relative to |qvel|_inf.  Result (round 2): the double-precision solves change nothing systematic (max
element-wise qvel error 1.4e-4 -> 5.0e-5 over these ten states, most states unchanged): the error is the round-off already present in the solve's fp32 inputs (M, J, D,
forces) amplified by cond(H), so neither fp64 pivots nor an iterative-refinement pass can buy the element-wise 1e-5;
relative to the vector norm a single mj_step is within 2.4e-6."""
import sys, numpy as np
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import emu_harness
emu_harness.build(force=True)
from emu_harness import EmuBatch
from mujocoposelearning_b200.abi import make_config, pack_model
from mujocoposelearning_b200.mjcf import compile_mjcf
from oracle.oracle import OracleEnv
cm=compile_mjcf(); ms=pack_model(cm)
def erel(a,b,floor=1e-3): return (np.abs(a-b)/np.maximum(np.abs(b),floor)).max()
res=[]
for seed,pre in [(1,0),(2,5),(3,30),(4,80),(5,150),(6,300),(7,500),(8,40),(9,60),(10,100)]:
    e=OracleEnv(ms,cm.nq,cm.nv,cm.nu); rng=np.random.default_rng(seed)
    e.env_reset(rng.uniform(-0.01,0.01,55))
    for _ in range(pre): e.env_step(rng.uniform(-1,1,21).astype(np.float32))
    emu=EmuBatch(ms,make_config(1,frame_skip=1,reward_type="stand",dtype="f32",duration=10.0),cm.nq,cm.nv,cm.nu)
    s=e.get_state()
    emu.qpos[0],emu.qvel[0],emu.warm[0],emu.nstep[0],emu.step_count[0]=s["qpos"],s["qvel"],s["warmstart"],s["nstep"],s["step_count"]
    # oracle starts from the float32-rounded state too
    e.set_state(s["qpos"].astype(np.float32).astype(np.float64), s["qvel"].astype(np.float32).astype(np.float64), s["warmstart"].astype(np.float32).astype(np.float64), int(s["nstep"]), int(s["step_count"]))
    act=rng.uniform(-1,1,(1,21)).astype(np.float32)
    obs,rew,*_=emu.step(act)
    o,r,*_=e.env_step(act[0],frame_skip=1)
    s2=e.get_state()
    nefc=int(e.get("nefc")[0])
    res.append((pre,nefc,erel(emu.qpos[0],s2["qpos"]),erel(emu.qvel[0],s2["qvel"]),erel(obs[0],o),abs(rew[0]-r), np.abs(emu.qvel[0]-s2["qvel"]).max()/max(1,np.abs(s2["qvel"]).max())))
    print(res[-1],flush=True)
