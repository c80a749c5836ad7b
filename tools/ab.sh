#!/bin/bash
# A/B tuning runs on the GPU box: tools/ab.sh "<name>|<env assignments>" ...   -> gpurun_out/ab_<name>.log
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%|*}"; envs="${spec#*|}"
  for E in 4096 16384; do
    env $envs python bench.py --quick --steps 1400 --warmup 100 --n-envs $E > gpurun_out/ab_${name}_$E.log 2>&1
    tail -1 gpurun_out/ab_${name}_$E.log | cut -c1-300
  done
done
